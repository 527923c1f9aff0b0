import sys, os, time
import numpy as np, cv2, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import depthmapcompression_b200 as dmc
from oracle.oracle_py import synth_disp, degrade_blocks
H, W = 1080, 1920
base = [synth_disp(H, W, 1000 + f, shift=(2 * f, f)) for f in range(8)]
for q in (50, 80):
    streams = [cv2.imencode(".jpg", b, [cv2.IMWRITE_JPEG_QUALITY, q])[1] for b in base]
    for n in (64, 256, 1000):
        ss = [streams[i % 8] for i in range(n)]
        blob, offs = dmc.pack_streams(ss)
        d = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
        ctx = dmc.default_context()
        dmc.jpegDecodeGrayBatch((blob, offs), H, W, dst=d.data_ptr())
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): dmc.jpegDecodeGrayBatch((blob, offs), H, W, dst=d.data_ptr())
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        assert np.array_equal(d[n - 1].cpu().numpy(), cv2.imdecode(ss[n - 1], 0))
        print("q%d n=%4d  %.1f ms  -> %.0f frames/s  %.1f Mpix/s   (%.0f KB/frame)" % (q, n, dt * 1e3, n / dt, n * H * W / dt / 1e6, len(blob) / n / 1e3))
t0 = time.perf_counter(); [cv2.imdecode(streams[i % 8], 0) for i in range(32)]; print("cv2.imdecode 1 thread: %.2f ms/frame" % ((time.perf_counter() - t0) / 32 * 1e3))
