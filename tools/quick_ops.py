import sys, os, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depthmapcompression_b200 as dmc
from depthmapcompression_b200 import capi
from depthmapcompression_b200.capi import DmcImage, lib
from oracle.oracle_py import synth_disp, degrade_blocks
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = dmc.Context(0); ctx.set_stream(stream.cuda_stream)
def T(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(iters): fn()
    e1.record(stream); torch.cuda.synchronize(); return e0.elapsed_time(e1) / iters
for (H, W) in [(480, 640), (1080, 1920)]:
    img = degrade_blocks(synth_disp(H, W, 3), 3)
    d8 = torch.from_numpy(img).to(dev); o8 = torch.empty_like(d8)
    f32 = (d8.float() * 16).contiguous(); of = torch.empty_like(f32)
    s8, q8 = DmcImage(d8.data_ptr(), H, W, 0, 0, 1), DmcImage(o8.data_ptr(), H, W, 0, 0, 1)
    sf, qf = DmcImage(f32.data_ptr(), H, W, 5, 0, 1), DmcImage(of.data_ptr(), H, W, 5, 0, 1)
    print(H, W, "BRF 13x13 8u   %.3f ms" % T(lambda: lib.dmc_boundary_reconstruction(ctx.h, C.byref(s8), C.byref(q8), 13, 13, 1.0, 1.0, 1.0), 5))
    print(H, W, "BRF 7x7 8u     %.3f ms" % T(lambda: lib.dmc_boundary_reconstruction(ctx.h, C.byref(s8), C.byref(q8), 7, 7, 1.0, 1.0, 1.0), 5))
    for r in (1, 3, 5):
        k = 2 * r + 1
        print(H, W, "bwrf32f r%d     %.4f ms" % (r, T(lambda: lib.dmc_bwrf(ctx.h, C.byref(sf), C.byref(qf), k, k, 160.0, 0, 1))))
        print(H, W, "bwrf8u  r%d     %.4f ms" % (r, T(lambda: lib.dmc_bwrf(ctx.h, C.byref(s8), C.byref(q8), k, k, 10.0, 0, 1))))
    print(H, W, "median5 %.4f  gauss3 %.4f  minmax3 %.4f ms" % (T(lambda: lib.dmc_median_blur(ctx.h, C.byref(s8), C.byref(q8), 5)), T(lambda: lib.dmc_small_gaussian(ctx.h, C.byref(s8), C.byref(q8), 3, 1.5)), T(lambda: lib.dmc_blur_remove_minmax(ctx.h, C.byref(s8), C.byref(q8), 3))))
