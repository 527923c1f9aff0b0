"""Every operator once on small awkward shapes -- meant to be run under `compute-sanitizer --tool memcheck`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depthmapcompression_b200 as dmc
rs = np.random.RandomState(0)
pfs = dmc.PostFilterSet()
for (H, W) in [(83, 131), (1, 5), (5, 1), (16, 16), (37, 260)]:
    a = rs.randint(1, 256, size=(H, W)).astype(np.uint8)
    for (mr, gr, mmr, br, th) in [(2, 1, 3, 5, 10), (1, 0, 1, 3, 10), (3, 3, 6, 7, 40)]:
        pfs(a, None, mr, gr, mmr, br, th); pfs(a, None, mr, gr, mmr, br, th, dmc.SEPARABLE_KERNEL)
        pfs.filterDisp8U2Depth32F(a, None, 75, 575, 2.6, mr, gr, mmr, br, th * 6.5)
        pfs.filterDisp8U2Depth16U(a, None, 75, 575, 2.6, mr, gr, mmr, br, th * 6.5)
        pfs.filterDisp8U2Disp32F(a, None, mr, gr, mmr, br, float(th))
    for dt, cn in [(np.uint8, 1), (np.uint8, 3), (np.uint16, 1), (np.int16, 1), (np.float32, 1), (np.float32, 3)]:
        b = (rs.rand(H, W, cn) * 200).astype(dt) if cn > 1 else (rs.rand(H, W) * 200).astype(dt)
        for k in (1, 3, 11, 21):
            dmc.binalyWeightedRangeFilter(b, None, (k, k), 10, dmc.FULL_KERNEL)
        dmc.binalyWeightedRangeFilter(b, None, (11, 5), 10, dmc.SEPARABLE_KERNEL)
    for dt in (np.uint8, np.uint16, np.int16, np.float32, np.float64):
        b = (rs.rand(H, W) * 200).astype(dt)
        dmc.blurRemoveMinMax(b, None, 3); dmc.blurRemoveMinMax(b, None, 10)
        dmc.boundaryReconstructionFilter(b, None, (13, 13), 1, 1, 1); dmc.boundaryReconstructionFilter(b, None, (5, 9), 1, 1, 1)
        dmc.boundaryReconstructionFilter((b.astype(np.int64) % 3).astype(dt), None, (7, 7), 1, 1, 1)       # few values: every pass shape of the id table
        if dt in (np.uint8, np.uint16, np.int16):
            dmc.minmaxBoundaryReconstructionFilter(b, None, 3, (7, 7), 1, 1, 1)
        if dt != np.float64:
            dmc.maxFilter(b, None, (7, 5)); dmc.minFilter(b, None, (3, 9))
    dmc.medianBlur(a, None, 3); dmc.medianBlur(a, None, 5); dmc.medianBlur(a, None, 9)
    dmc.smallGaussianBlur(a, None, 3, 1.5); dmc.smallGaussianBlur(a, None, 5, 2.5); dmc.smallGaussianBlur(a, None, 9, 4.5)
    d32 = dmc.disp8U2depth32F(a, None, 43125.0, 2.6, 0.0); dmc.depth32F2disp8U(d32, None, 43125.0, 2.6, 0.0)
    dmc.depth16U2disp8U((a.astype(np.uint16) * 9), None, 43125.0, 2.6, 0.0); dmc.disp16S2depth16U(a.astype(np.int16), None, 43125.0, 2.6, 1.0)
    if W > 2:
        dmc.fillOcclusion(a.copy(), 7, dmc.FILL_DISPARITY)
    dmc.reprojectXYZ(d32, None, 510.0)
frames = rs.randint(1, 256, size=(5, 100, 260)).astype(np.uint8); out = np.zeros_like(frames)
from depthmapcompression_b200.filters import chain_params
from depthmapcompression_b200 import capi
dmc.default_context().chain_batch(frames, out, 5, 100, 260, chain_params(capi.CHAIN_DISP8U, 2, 1, 3, 5, 10), device=False)
# round-2 paths: pinned host images in place, JPEG decode and bitstream -> chain, point-cloud render, joint filter
p_in = dmc.pinned_empty((100, 260), np.uint8); p_out = dmc.pinned_empty((100, 260), np.float32); p_in[:] = frames[0]
pfs.filterDisp8U2Depth32F(p_in, p_out, 75, 575, 2.6, 1, 0, 1, 3, 65.0); pfs.filterDisp8U2Depth32F(p_in, p_out, 75, 575, 2.6, 1, 0, 1, 3, 65.0)
try:
    import cv2
    streams = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 70])[1].tobytes() for f in frames]
    dmc.jpegDecodeGrayBatch(streams, 100, 260)
    dmc.default_context().chain_batch_jpeg(streams, 100, 260, out, chain_params(capi.CHAIN_DISP8U, 1, 0, 1, 3, 10))
except ImportError:
    pass
dmc.jointBinalyWeightedRangeFilter(frames[0], np.stack([frames[1], frames[2], frames[3]], -1).copy(), None, (7, 7), 20)
d = dmc.disp8U2depth32F(frames[0], None, 43125.0, 2.6, 0.0); xyz = dmc.reprojectXYZ(d, None, 510.0)
K = np.array([[510., 0, 130], [0, 510., 50], [0, 0, 1]]); R = np.eye(3); t = np.array([20., -10., 30.])
img = np.stack([frames[0]] * 3, -1).copy()
dmc.projectImagefromXYZ(img, None, xyz, R, t, K, None, None, True)
print("sanitize_small: all operators ran")
