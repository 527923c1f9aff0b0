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
g3 = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device=dev); g1 = g3[:, :, 0].contiguous()
s8, q8 = DmcImage(d8.data_ptr(), H, W, 0, 0, 1), DmcImage(o8.data_ptr(), H, W, 0, 0, 1)
G3, G1 = DmcImage(g3.data_ptr(), H, W, 16, 0, 1), DmcImage(g1.data_ptr(), H, W, 0, 0, 1)
for name, G in (("C3", G3), ("C1", G1)):
    for r in (1, 3, 5, 7):
        k = 2 * r + 1
        f = lambda: lib.dmc_joint_bwrf(ctx.h, C.byref(s8), C.byref(G), C.byref(q8), k, k, 30.0, 0)
        for _ in range(3): assert f() == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(10): f()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("joint guide %s 1080p r%d: %.3f ms  %.1f Gpix/s" % (name, r, ms, H * W / ms / 1e6))
