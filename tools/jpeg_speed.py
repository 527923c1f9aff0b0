"""Speed of the GPU JPEG row on 1080p: decode alone, bitstream -> chain on the device, and bitstream ->
chain -> host through the streamed entry point (pinned blob and pinned output).  Prints one JSON object.

    python tools/jpeg_speed.py [--frames 1000]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depthmapcompression_b200 as dmc                      # noqa: E402
from depthmapcompression_b200 import capi                   # noqa: E402
from depthmapcompression_b200.filters import chain_params   # noqa: E402
from oracle.oracle_py import synth_disp, degrade_blocks     # noqa: E402

H, W = 1080, 1920


def pinned(nbytes):
    p = capi.lib.dmc_host_alloc(nbytes)
    return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=1000); args = ap.parse_args()
    ctx = dmc.default_context()
    rs = np.random.RandomState(1)
    clean = [synth_disp(H, W, 1000 + f, shift=(2 * f, f)) for f in range(8)]
    degraded = [degrade_blocks(c, i) for i, c in enumerate(clean)]
    noisy = [np.clip(c.astype(int) + rs.randint(-12, 13, (H, W)), 0, 255).astype(np.uint8) for c in clean]
    sets = {"clean_q50": (clean, 50), "degraded_q80": (degraded, 80), "noisy_q80": (noisy, 80)}
    out = {"gpu": torch.cuda.get_device_name(0), "frame": [H, W], "sets": {}}
    p8 = chain_params(capi.CHAIN_DISP8U, 2, 1, 3, 5, 10)
    p32 = chain_params(capi.CHAIN_DEPTH32F, 1, 0, 1, 3, 65.0, focus=75.0, baseline=575.0, amp=2.6)
    for name, (imgs, q) in sets.items():
        streams = [cv2.imencode(".jpg", b, [cv2.IMWRITE_JPEG_QUALITY, q])[1] for b in imgs]
        res = {"kb_per_frame": round(sum(len(s) for s in streams) / 8 / 1e3, 1), "decode": {}}
        for n in (1, 64, 480, args.frames):
            ss = [streams[i % 8] for i in range(n)]
            blob, offs = dmc.pack_streams(ss)
            d = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
            dmc.jpegDecodeGrayBatch((blob, offs), H, W, dst=d.data_ptr())
            reps = 20 if n == 1 else 3
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(reps):
                dmc.jpegDecodeGrayBatch((blob, offs), H, W, dst=d.data_ptr())
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
            assert np.array_equal(d[n - 1].cpu().numpy(), cv2.imdecode(ss[n - 1], 0))
            res["decode"][str(n)] = {"ms": round(dt * 1e3, 3), "fps": round(n / dt)}
        # streamed: pinned blob -> decode -> chain -> pinned host output / device output
        n = args.frames
        ss = [streams[i % 8] for i in range(n)]
        blob, offs = dmc.pack_streams(ss)
        pb, ab = pinned(blob.size); ab[:] = blob
        po, ao = pinned(n * H * W)
        d32 = torch.empty((n, H, W), dtype=torch.float32, device="cuda")
        for label, params, dst, dev in (("bitstream_to_chain8u_host", p8, po, False), ("bitstream_to_depth32f_device", p32, d32.data_ptr(), True)):
            ctx.chain_batch_jpeg((pb, offs), H, W, dst, params, device=dev); ctx.synchronize()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); ctx.chain_batch_jpeg((pb, offs), H, W, dst, params, device=dev); ctx.synchronize(); ts.append(time.perf_counter() - t0)
            res[label] = {"ms": round(min(ts) * 1e3, 2), "fps": round(n / min(ts)), "gpix_s": round(n * H * W / min(ts) / 1e9, 2)}
        capi.lib.dmc_host_free(pb); capi.lib.dmc_host_free(po)
        out["sets"][name] = res
        print(name, json.dumps(res), file=sys.stderr, flush=True)
    t0 = time.perf_counter(); [cv2.imdecode(streams[i % 8], 0) for i in range(16)]
    out["cv2_imdecode_ms_per_frame_1_thread_noisy_q80"] = round((time.perf_counter() - t0) / 16 * 1e3, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
