"""Host-to-host throughput of dmc_chain_batch for a few chunk sizes (DMC_CHUNK_MB is read per call)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depthmapcompression_b200 as dmc
from depthmapcompression_b200 import capi
from depthmapcompression_b200.filters import chain_params
H, W, N = 1080, 1920, 1000
ctx = dmc.Context(0)
h_in = torch.randint(1, 256, (N, H, W), dtype=torch.uint8).pin_memory(); h_out = torch.empty((N, H, W), dtype=torch.uint8).pin_memory()
p = chain_params(capi.CHAIN_DISP8U, 2, 1, 3, 5, 10)
for mb in [int(a) for a in sys.argv[1:]] or (8, 16, 32, 64, 128, 256):
    os.environ["DMC_CHUNK_MB"] = str(mb)
    for _ in range(2): ctx.chain_batch(h_in.data_ptr(), h_out.data_ptr(), N, H, W, p, device=False)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); ctx.chain_batch(h_in.data_ptr(), h_out.data_ptr(), N, H, W, p, device=False); ts.append(time.perf_counter() - t0)
    print("chunk %3d MB: best %.1f  median %.1f  worst %.1f Gpix/s" % (mb, N * H * W / min(ts) / 1e9, N * H * W / np.median(ts) / 1e9, N * H * W / max(ts) / 1e9), flush=True)
