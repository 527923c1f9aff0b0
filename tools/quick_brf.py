"""Times the boundary reconstruction filter (and the fused min-max -> BRF extension against the two separate calls) on the
Kinect fixture tiled to 1080p and on the benchmark's block-noise frames: 8-bit and 16-bit, 13x13 and 7x7.
    python tools/quick_brf.py [out.json]"""
import json, os, sys
import numpy as np, torch, cv2
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
import depthmapcompression_b200 as dmc
from depthmapcompression_b200 import capi
from depthmapcompression_b200.filters import _img
import ctypes as C
from oracle.oracle_py import synth_disp, degrade_blocks
lib = capi.lib
ctx = dmc.Context(0)
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
img = cv2.imread(os.path.join(root, "tests/golden/kinect_desk_q50.png"), cv2.IMREAD_UNCHANGED)
kin = np.ascontiguousarray(np.tile(img, (3, 3))[:1080, :1920])
noise = degrade_blocks(synth_disp(1080, 1920, 7), 7)
res = {}


def dimg(t):
    cvt = {torch.uint8: 0, torch.int16: 3, torch.float32: 5}[t.dtype] if t.dtype != torch.uint16 else 2
    return capi.DmcImage(t.data_ptr(), t.shape[0], t.shape[1], cvt, 0, 1)


def timeit(f, n=10):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(n): f()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, a in (("kinect_1080p", kin), ("blocknoise_1080p", noise), ("kinect_640x480", img)):
    for dt in ("u8", "u16"):
        b = a if dt == "u8" else (a.astype(np.int16) * 37).astype(np.int16)      # 16-bit: the same structure, spread values (16S device view)
        src = torch.from_numpy(b).to(dev); dst = torch.empty_like(src); tmp = torch.empty_like(src)
        s, d, t = dimg(src), dimg(dst), dimg(tmp)
        for k in (13, 7):
            ms = timeit(lambda: ctx.check(lib.dmc_boundary_reconstruction(ctx.h, C.byref(s), C.byref(d), k, k, 1.0, 1.0, 1.0)))
            ms2 = timeit(lambda: (ctx.check(lib.dmc_blur_remove_minmax(ctx.h, C.byref(s), C.byref(t), 3)), ctx.check(lib.dmc_boundary_reconstruction(ctx.h, C.byref(t), C.byref(d), k, k, 1.0, 1.0, 1.0))))
            two = dst.clone()
            msf = timeit(lambda: ctx.check(lib.dmc_minmax_boundary_reconstruction(ctx.h, C.byref(s), C.byref(d), 3, k, k, 1.0, 1.0, 1.0)))
            res["%s %s %dx%d" % (name, dt, k, k)] = {"brf_ms": round(ms, 4), "minmax_then_brf_ms": round(ms2, 4), "fused_ms": round(msf, 4), "fused_equals_two_calls": bool(torch.equal(two, dst))}
            print(name, dt, k, res["%s %s %dx%d" % (name, dt, k, k)], flush=True)
if len(sys.argv) > 1:
    json.dump({"gpu": torch.cuda.get_device_name(0), "results": res}, open(sys.argv[1], "w"), indent=1)
