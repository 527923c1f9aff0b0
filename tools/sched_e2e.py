"""Host-to-host throughput of the in-process frame-batch scheduler (dmc_sched_*) over 1, 2, 4, ... visible GPUs.

    python tools/sched_e2e.py [--frames-per-gpu 500] [--reps 3] [--chunks 16,48,128] [--orders natural,interleaved]

One process, one pinned host batch (dmc_host_alloc), PostFilterSet::operator()(2,1,3,5,10) on 1080p frames; the batch is
cut into contiguous shards, one per device, each streamed through its device's H2D / kernels / D2H pipeline by its own
host thread.  Prints one JSON object; every entry carries the device list, the chunk size and the wall-clock rate of the
whole call (the call returns when the last output frame is in host memory).  The first output of every configuration is
compared with a single-device run of the same frames.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depthmapcompression_b200 as dmc          # noqa: E402
from depthmapcompression_b200 import capi       # noqa: E402
from depthmapcompression_b200.filters import chain_params, FrameBatchScheduler   # noqa: E402

H, W = 1080, 1920


def pinned(nbytes):
    p = capi.lib.dmc_host_alloc(nbytes)
    if not p:
        raise MemoryError("dmc_host_alloc(%d)" % nbytes)
    return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-per-gpu", type=int, default=500)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--chunks", default="48")
    ap.add_argument("--orders", default="natural")
    args = ap.parse_args()
    ndev = capi.lib.dmc_device_count()
    p = chain_params(capi.CHAIN_DISP8U, median_r=2, gaussian_r=1, minmax_r=3, brange_r=5, brange_th=10)
    nmax = ndev * args.frames_per_gpu
    pin_in, a_in = pinned(nmax * H * W); pin_out, a_out = pinned(nmax * H * W)
    rs = np.random.RandomState(5)
    base = (90 + 40 * np.sin(np.arange(W)[None, :] / 320.0) + 30 * np.cos(np.arange(H)[:, None] / 216.0))
    tile = np.stack([np.clip(base + rs.randint(-5, 6, (H, W)) + 8 * (rs.randint(0, 3, (H // 8, W // 8)).repeat(8, 0).repeat(8, 1)), 1, 255).astype(np.uint8) for _ in range(4)])
    fin = a_in.reshape(nmax, H, W); fout = a_out.reshape(nmax, H, W)
    for f in range(nmax):
        fin[f] = tile[f % 4]
    ctx = dmc.Context(0)
    want = np.empty((4, H, W), np.uint8)
    ctx.chain_batch(tile, want, 4, H, W, p, device=False)
    out = {"frames_per_gpu": args.frames_per_gpu, "visible_devices": ndev, "runs": []}
    sizes = [k for k in (1, 2, 4, 8) if k <= ndev]
    for order in args.orders.split(","):
        for k in sizes:
            devs = list(range(k))
            if order == "interleaved" and ndev >= 8:
                devs = [0, 4, 1, 5, 2, 6, 3, 7][:k]
            elif order == "interleaved":
                continue
            sched = FrameBatchScheduler(devs)
            gws, all_gbs, best_gbs = sched.routing()
            n = k * args.frames_per_gpu
            for chunk in [int(c) for c in args.chunks.split(",")]:
                os.environ["DMC_CHUNK_MB"] = str(chunk)
                fout[:n] = 0
                sched.chain_batch(pin_in, pin_out, n, H, W, p)       # warm-up (allocations)
                ok = all(np.array_equal(fout[f], want[f % 4]) for f in (0, 1, n // 2, n - 1))
                ts = []
                for _ in range(args.reps):
                    t0 = time.perf_counter(); sched.chain_batch(pin_in, pin_out, n, H, W, p); ts.append(time.perf_counter() - t0)
                best = min(ts)
                out["runs"].append({"devices": devs, "order": order, "chunk_mb": chunk, "frames": n, "ms": [round(t * 1e3, 2) for t in ts],
                                    "gpix_s_best": round(n * H * W / best / 1e9, 2), "gb_s_each_way": round(n * H * W / best / 1e9, 2), "bit_exact_vs_single_device": bool(ok),
                                    "gateways": gws, "link_all_gbs": round(all_gbs, 1), "link_best_gbs": round(best_gbs, 1)})
                print(json.dumps(out["runs"][-1]), file=sys.stderr, flush=True)
            del sched
    print(json.dumps(out))


if __name__ == "__main__":
    main()
