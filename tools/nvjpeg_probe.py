"""Probe: does nvJPEG (GPU backends) decode baseline grayscale JPEG bit-identically to libjpeg-turbo (cv2.imdecode)?"""
import ctypes as C, os, sys
import numpy as np, cv2, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle_py import synth_disp
nv = C.CDLL("libnvjpeg.so.12")
class Img(C.Structure):
    _fields_ = [("channel", C.c_void_p * 4), ("pitch", C.c_size_t * 4)]
def chk(rc, what):
    if rc != 0: raise RuntimeError("%s -> %d" % (what, rc))
for backend, name in [(0, "DEFAULT"), (1, "HYBRID"), (2, "GPU_HYBRID"), (3, "HARDWARE")]:
    h = C.c_void_p(); 
    rc = nv.nvjpegCreateEx(backend, None, None, 0, C.byref(h))
    if rc != 0: print(name, "create failed", rc); continue
    st = C.c_void_p(); chk(nv.nvjpegJpegStateCreate(h, C.byref(st)), "state")
    tot = bad = 0; maxd = 0
    for (H, W, q, seed) in [(480, 640, 50, 1), (480, 640, 80, 2), (1080, 1920, 50, 3), (1080, 1920, 80, 4), (131, 150, 50, 5)]:
        img = synth_disp(H, W, seed)
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
        ref = cv2.imdecode(buf, 0)
        out = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
        im = Img(); im.channel[0] = out.data_ptr(); im.pitch[0] = W
        data = buf.tobytes()
        rc = nv.nvjpegDecode(h, st, data, C.c_size_t(len(data)), 0, C.byref(im), None)   # NVJPEG_OUTPUT_UNCHANGED = 0 (gray stays gray)
        torch.cuda.synchronize()
        if rc != 0: print(name, "decode rc", rc); break
        got = out.cpu().numpy()
        d = np.abs(got.astype(int) - ref.astype(int)); tot += d.size; bad += int((d > 0).sum()); maxd = max(maxd, int(d.max()))
    print("%-10s pixels %d differing %d (%.4f%%) max|diff| %d" % (name, tot, bad, 100.0 * bad / max(tot, 1), maxd))
