"""Runs ONE operator a few times on a 1080p fixture so that ncu can capture it:
    ncu --set full -k regex:brf_rank -c 1 ... python tools/run_op.py brf13 [kinect|noise] [u8|s16]
ops: brf13, brf7, fused13 (min-max r=3 -> BRF 13x13), bwrfXX_rR (range filter radius R; XX free text, 'c3' = 3 channels; the dtype argument picks the depth), median_kK, jpeg, depth32f, reproject"""
import os, sys
import numpy as np, cv2
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
import depthmapcompression_b200 as dmc
from oracle.oracle_py import synth_disp, degrade_blocks
op = sys.argv[1]; fixture = sys.argv[2] if len(sys.argv) > 2 else "kinect"; dt = sys.argv[3] if len(sys.argv) > 3 else "u8"
if fixture == "kinect":
    img = cv2.imread(os.path.join(root, "tests/golden/kinect_desk_q50.png"), cv2.IMREAD_UNCHANGED)
    a = np.ascontiguousarray(np.tile(img, (3, 3))[:1080, :1920])
else:
    a = degrade_blocks(synth_disp(1080, 1920, 7), 7)
if dt == "s16": a = (a.astype(np.int16) * 37).astype(np.int16)
if dt == "u16": a = (a.astype(np.uint16) * 37).astype(np.uint16)
if dt == "f32": a = a.astype(np.float32) * 3.7
ctx = dmc.default_context()
for _ in range(3):
    if op.startswith("brf"): k = int(op[3:]); out = dmc.boundaryReconstructionFilter(a, None, (k, k), 1.0, 1.0, 1.0)
    elif op.startswith("fused"): k = int(op[5:]); out = dmc.minmaxBoundaryReconstructionFilter(a, None, 3, (k, k), 1.0, 1.0, 1.0)
    elif op.startswith("bwrf"):
        r = int(op.split("_r")[1]); b = a
        if "c3" in op: b = np.ascontiguousarray(np.stack([a, a[::-1], a[:, ::-1]], -1))
        out = dmc.binalyWeightedRangeFilter(b, None, (2 * r + 1, 2 * r + 1), (10.0 if dt == "u8" and "c3" not in op else 30.0) * (37 if dt in ("s16", "u16") else 1), dmc.FULL_KERNEL)
    elif op.startswith("median_k"): out = dmc.medianBlur(a, None, int(op[8:]))
    elif op == "jpeg":
        streams = [cv2.imencode(".jpg", np.roll(a, 3 * i, axis=1), [cv2.IMWRITE_JPEG_QUALITY, 80])[1].tobytes() for i in range(64)]
        out = dmc.jpegDecodeGrayBatch(streams, 1080, 1920)
    elif op == "depth32f": out = dmc.PostFilterSet().filterDisp8U2Depth32F(a, None, 75.0, 575.0, 2.6, 1, 0, 1, 3, 65.0)
    elif op == "reproject": out = dmc.reprojectXYZ(a.astype(np.float32) * 3.0 + 500.0, None, 510.0)
    else: raise SystemExit("unknown op " + op)
print(op, fixture, dt, out.shape, out.dtype, int(out.astype(np.int64).sum() & 0xffffffff))
