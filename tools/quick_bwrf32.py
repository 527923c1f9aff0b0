"""Times the 32-bit binary-weighted range filter on one 3840x2160 view (BASELINE config 4): 16UC1 th=160, 32FC1 th=30.5 and
32FC3 th=30.5, radius 1..10.    python tools/quick_bwrf32.py [out.json]"""
import ctypes as C, json, os, sys
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, root)
import depthmapcompression_b200 as dmc
from depthmapcompression_b200 import capi
from depthmapcompression_b200.capi import DmcImage, lib
from oracle.oracle_py import synth_disp
H, W = 2160, 3840
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = dmc.Context(0); ctx.set_stream(stream.cuda_stream)
rs = np.random.RandomState(5)
base = synth_disp(H, W, 3)
d16 = (base.astype(np.uint16) * 16 + rs.randint(0, 16, size=(H, W)).astype(np.uint16))
f32 = d16.astype(np.float32) * 0.37
f3 = np.ascontiguousarray(np.stack([f32, np.roll(f32, 7, 1), np.roll(f32, 11, 0)], axis=2))
HBM = 6536.7
try: HBM = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception: pass
res = {}
for name, arr, cvt, th, bpp, radii in (("16UC1_th160", d16.view(np.int16), capi.CV_16U, 160.0, 4, range(1, 11)), ("32FC1_th30.5", f32, capi.CV_32F, 30.5, 8, range(1, 11)),
                                       ("32FC3_th30.5", f3, capi.CV_32F + (2 << 3), 30.5, 24, range(1, 8))):
    src = torch.from_numpy(arr).to(dev); dst = torch.empty_like(src)
    s, d = DmcImage(src.data_ptr(), H, W, cvt, 0, capi.MEM_DEVICE), DmcImage(dst.data_ptr(), H, W, cvt, 0, capi.MEM_DEVICE)
    res[name] = {}
    for r in radii:
        k = 2 * r + 1
        f = lambda: ctx.check(lib.dmc_bwrf(ctx.h, C.byref(s), C.byref(d), k, k, th, 0, 1))
        for _ in range(2): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(8): f()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 8
        res[name]["r%d" % r] = {"ms_per_view": round(ms, 4), "mpix_s": round(H * W / ms / 1e3, 1), "hbm_frac": round(H * W * bpp / ms / 1e6 / HBM, 4)}
        print(name, r, res[name]["r%d" % r], flush=True)
if len(sys.argv) > 1:
    json.dump({"gpu": torch.cuda.get_device_name(0), "frame": [H, W], "results": res}, open(sys.argv[1], "w"), indent=1)
