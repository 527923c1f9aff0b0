#!/usr/bin/env python
"""bench_multigpu.py -- BASELINE.json configs[3] and configs[4] on N GPUs of one box (NOT the contract line: that is bench.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_multigpu.py [--seconds 10]
    python bench_multigpu.py                      (one GPU)

One process per GPU, no collective on the data path (views and frames are independent); torch.distributed carries the
barriers and the max-over-ranks of the timings.  Every timed configuration is preceded by a bit-exact check against the
CPU oracle on this rank's own data.

  C4  3840x2160 16-bit depth + RGB, multi-view batch: view v lives on GPU v; binalyWeightedRangeFilter
      FULL_KERNEL r = 1..7 on the 16UC1 depth (th = 160) and on the 8UC3 colour view (th = 30), device-resident,
      CUDA events; per radius: slowest rank's ms per view and views x pixels / that time.
  C5  pointcloudTest() pipeline at 1080p (main.cpp:276-321): JPEG bitstream (pinned host) -> decode on the GPU ->
      filterDisp8U2Depth32F(1,0,1,3,65) -> reprojectXYZ(f) of every frame; SUSTAINED for --seconds per phase on every GPU
      at once: decode-bound rate (decode only), filter-bound rate (chain + reprojection on resident frames) and the
      whole pipeline, frames/s per GPU and for the box.
Rank 0 prints one JSON object.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import depthmapcompression_b200 as dmc  # noqa: E402
from depthmapcompression_b200 import capi  # noqa: E402
from depthmapcompression_b200.capi import DmcImage, lib  # noqa: E402
from depthmapcompression_b200.filters import chain_params  # noqa: E402
from oracle.oracle_py import Port, synth_disp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--skip", default="", help="comma list of c4,c5")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
    ctx = dmc.Context(local); ctx.set_stream(stream.cuda_stream)
    port = Port(); port.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    out = {"n_gpus": world, "gpu": torch.cuda.get_device_name(local)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t[0])

    def allsum(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t[0])

    def gather(v):
        if world == 1:
            return [float(v)]
        t = torch.tensor([v], device=dev, dtype=torch.float64); g = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(g, t)
        return [float(x[0]) for x in g]

    def timed(fn, iters, warm=2):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream); barrier()
        return e0.elapsed_time(e1) / iters

    # ---------------------------------------------------------------------------------------------- C4: view v on GPU v
    if "c4" not in args.skip:
        H, W = 2160, 3840
        rs = np.random.RandomState(5 + rank)                     # every view is its own image
        base = synth_disp(H, W, 3 + rank)
        d16 = (base.astype(np.uint16) * 16 + rs.randint(0, 16, size=(H, W)).astype(np.uint16))
        rgb = np.stack([base, np.roll(base, 7, 1), np.roll(base, 11, 0)], axis=2).copy()
        rgb = np.clip(rgb.astype(np.int16) + rs.randint(-4, 5, size=rgb.shape), 0, 255).astype(np.uint8)
        b16 = torch.from_numpy(d16.view(np.int16)).to(dev); o16 = torch.empty_like(b16)
        b3 = torch.from_numpy(rgb).to(dev); o3 = torch.empty_like(b3)
        s16, q16 = DmcImage(b16.data_ptr(), H, W, capi.CV_16U, 0, capi.MEM_DEVICE), DmcImage(o16.data_ptr(), H, W, capi.CV_16U, 0, capi.MEM_DEVICE)
        s3, q3 = DmcImage(b3.data_ptr(), H, W, capi.CV_8U + (2 << 3), 0, capi.MEM_DEVICE), DmcImage(o3.data_ptr(), H, W, capi.CV_8U + (2 << 3), 0, capi.MEM_DEVICE)
        sweep = {"16UC1_th160": {}, "8UC3_th30": {}}
        for r in range(1, 8):
            k = 2 * r + 1
            ctx.check(lib.dmc_bwrf(ctx.h, C.byref(s16), C.byref(q16), k, k, 160.0, 0, 1)); ctx.check(lib.dmc_bwrf(ctx.h, C.byref(s3), C.byref(q3), k, k, 30.0, 0, 1)); ctx.synchronize()
            if r in (1, 4, 7):                                    # parity on a crop of this rank's view (rows whose window stays inside the crop)
                want = port.bwrf(np.ascontiguousarray(d16[:300, -600:]), k, k, 160.0)
                got = o16.cpu().numpy().view(np.uint16)[:300, -600:]
                assert np.array_equal(got[:300 - r, r:], want[:300 - r, r:]), ("C4 16U rank %d r %d" % (rank, r))
                want = port.bwrf(np.ascontiguousarray(rgb[:300, :600]), k, k, 30.0)
                got = o3.cpu().numpy()[:300, :600]
                assert np.array_equal(got[:300 - r, :600 - r], want[:300 - r, :600 - r]), ("C4 8UC3 rank %d r %d" % (rank, r))
            for name, fn, bpp in (("16UC1_th160", lambda: lib.dmc_bwrf(ctx.h, C.byref(s16), C.byref(q16), k, k, 160.0, 0, 1), 4),
                                  ("8UC3_th30", lambda: lib.dmc_bwrf(ctx.h, C.byref(s3), C.byref(q3), k, k, 30.0, 0, 1), 6)):
                ms = timed(fn, 10)
                per = gather(ms); worst = max(per)
                sweep[name]["r%d" % r] = {"ms_per_view_slowest_gpu": round(worst, 4), "ms_per_view_by_gpu": [round(x, 4) for x in per],
                                          "mpix_s_all_views": round(world * H * W / worst / 1e3, 1)}
        out["C4_4K_multiview_view_v_on_gpu_v"] = {"views": world, "frame": [H, W], "sweep": sweep,
                                                  "note": "one view per GPU, device-resident, CUDA events bracketed by barriers; parity on crops of every rank's own view at r = 1, 4, 7"}
        del b16, o16, b3, o3

    # ---------------------------------------------------------------------------------------------- C5: sustained pipeline
    if "c5" not in args.skip:
        import cv2
        H, W, NJ = 1080, 1920, 240
        y = cv2.imread(os.path.join(ROOT, "tests", "golden", "kinect_desk_q80.png"), cv2.IMREAD_UNCHANGED)
        base = np.ascontiguousarray(np.tile(y, (3, 3))[:H, :W])
        uniq = [np.roll(base, (7 * i + rank, 13 * i), axis=(0, 1)) for i in range(16)]
        coded = [cv2.imencode(".jpg", u, [cv2.IMWRITE_JPEG_QUALITY, 80])[1].tobytes() for u in uniq]
        blob, offs = dmc.pack_streams([coded[i % 16] for i in range(NJ)])
        pblob = torch.empty(blob.size, dtype=torch.uint8).pin_memory(); pblob.numpy()[:] = blob
        d_dec = torch.empty((NJ, H, W), dtype=torch.uint8, device=dev); d_dep = torch.empty((NJ, H, W), dtype=torch.float32, device=dev)
        xyz = torch.empty((H * W, 3), dtype=torch.float32, device=dev)
        pj = chain_params(capi.CHAIN_DEPTH32F, 1, 0, 1, 3, 65, focus=75.0, baseline=575.0, amp=2.6)
        sx = DmcImage(xyz.data_ptr(), H * W, 1, capi.CV_32F + (2 << 3), 0, capi.MEM_DEVICE)

        def decode_only():
            dmc.jpegDecodeGrayBatch((pblob.numpy(), offs), H, W, dst=d_dec.data_ptr(), ctx=ctx)

        def filter_only():
            ctx.chain_batch(d_dec.data_ptr(), d_dep.data_ptr(), NJ, H, W, pj, device=True)
            for i in range(NJ):                                   # reprojectXYZ of every frame (the renderer's input, main.cpp:308)
                sd = DmcImage(d_dep[i].data_ptr(), H, W, capi.CV_32F, 0, capi.MEM_DEVICE)
                lib.dmc_reproject_xyz(ctx.h, C.byref(sd), C.byref(sx), 510.0)

        def pipeline():
            decode_only(); filter_only()

        pipeline(); ctx.synchronize()
        for i in (0, NJ - 1):                                     # parity: decode == cv2.imdecode, chain == oracle, reprojection == oracle
            ref_dec = cv2.imdecode(np.frombuffer(coded[i % 16], np.uint8), 0)
            assert np.array_equal(d_dec[i].cpu().numpy(), ref_dec), "C5 decode rank %d" % rank
            ref_dep = port.filter_disp8u_depth32f(ref_dec, 75.0, 575.0, 2.6, 1, 0, 1, 3, 65.0)
            assert np.array_equal(d_dep[i].cpu().numpy().view(np.uint32), ref_dep.view(np.uint32)), "C5 chain rank %d" % rank
        assert np.array_equal(xyz.cpu().numpy().view(np.uint32), port.reproject_xyz(d_dep[NJ - 1].cpu().numpy(), 510.0).view(np.uint32)), "C5 reproject"

        def sustained(fn):
            """runs fn (NJ frames per call) on every GPU at once until `seconds` have passed on this rank; -> frames/s of this rank"""
            fn(); barrier()
            t0 = time.perf_counter(); n = 0
            while True:
                fn(); ctx.synchronize(); n += NJ
                if time.perf_counter() - t0 >= args.seconds:
                    break
            dt = time.perf_counter() - t0
            barrier()
            return n / dt, dt

        c5 = {}
        for name, fn in (("decode_bound", decode_only), ("filter_bound", filter_only), ("pipeline", pipeline)):
            fps, dt = sustained(fn)
            per = gather(fps)
            c5[name] = {"fps_box": round(sum(per), 1), "fps_by_gpu": [round(x, 1) for x in per], "mpix_s_box": round(sum(per) * H * W / 1e6, 1), "seconds": round(allmax(dt), 2)}
        out["C5_pointcloud_pipeline_1080p_sustained"] = dict(c5, frames_per_call=NJ, bitstream_kb_per_frame=round(blob.size / NJ / 1e3, 1),
            note="JPEG q80 of the reference's Kinect frame tiled to 1080p; decode_bound = bitstream H2D + GPU decode; filter_bound = filterDisp8U2Depth32F(1,0,1,3,65) + reprojectXYZ "
                 "of every frame on resident frames; pipeline = both; every GPU runs concurrently for the stated seconds per phase; parity gate on every rank")

    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
