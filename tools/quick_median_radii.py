import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
import depthmapcompression_b200 as dmc
from depthmapcompression_b200.capi import DmcImage, lib
from oracle.oracle_py import synth_disp, degrade_blocks
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = dmc.Context(0); ctx.set_stream(stream.cuda_stream)
H, W = 1080, 1920
img = degrade_blocks(synth_disp(H, W, 3), 3)
d8 = torch.from_numpy(img).to(dev); o8 = torch.empty_like(d8)
s8, q8 = DmcImage(d8.data_ptr(), H, W, 0, 0, 1), DmcImage(o8.data_ptr(), H, W, 0, 0, 1)
def T(f, it=5):
    for _ in range(2): assert f() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(it): f()
    e1.record(stream); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
for r in range(1, 11):
    print("median r%d %.3f ms" % (r, T(lambda: lib.dmc_median_blur(ctx.h, C.byref(s8), C.byref(q8), 2 * r + 1))), flush=True)
