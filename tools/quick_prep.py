import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depthmapcompression_b200 as dmc
from depthmapcompression_b200.capi import DmcImage, lib
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = dmc.Context(0); ctx.set_stream(stream.cuda_stream)
def T(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(iters): fn()
    e1.record(stream); torch.cuda.synchronize(); return e0.elapsed_time(e1) / iters
for (H, W) in [(480, 640), (1080, 1920)]:
    rs = np.random.RandomState(1)
    d16 = (rs.rand(H, W) * 4000 + 500).astype(np.uint16); d16[rs.rand(H, W) < 0.2] = 0
    t16 = torch.from_numpy(d16.view(np.int16)).to(dev); o8 = torch.empty((H, W), dtype=torch.uint8, device=dev)
    s16, q8 = DmcImage(t16.data_ptr(), H, W, 2, 0, 1), DmcImage(o8.data_ptr(), H, W, 0, 0, 1)
    print(H, W, "depth16U2disp8U %.4f ms" % T(lambda: lib.dmc_depth16u2disp8u(ctx.h, C.byref(s16), C.byref(q8), 43125.0, 2.6, 0.0)))
    print(H, W, "fillOcclusion    %.4f ms" % T(lambda: lib.dmc_fill_occlusion(ctx.h, C.byref(q8), 0, 0)))
