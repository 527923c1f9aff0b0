"""Times dmc_boundary_reconstruction (13x13 and 7x7) on the Kinect fixture for one or more builds of the library."""
import ctypes as C, sys, os, numpy as np, torch, cv2
class Img(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("cvtype", C.c_int), ("step", C.c_size_t), ("mem", C.c_int)]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
img = cv2.imread(os.path.join(root, "tests/golden/kinect_desk_q50.png"), cv2.IMREAD_UNCHANGED)
big = np.ascontiguousarray(np.tile(img, (3, 3))[:1080, :1920])
ref = {}
for path in sys.argv[1:]:
    lib = C.CDLL(path); ctx = C.c_void_p()
    assert lib.dmc_create(0, C.byref(ctx)) == 0
    lib.dmc_set_stream(ctx, C.c_void_p(stream.cuda_stream))
    lib.dmc_boundary_reconstruction.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    out = []
    for name, a in (("640x480", img), ("1080p", big)):
        src = torch.from_numpy(a).to(dev); dst = torch.empty_like(src)
        H, W = a.shape
        for k in (13, 7):
            s, d = Img(src.data_ptr(), H, W, 0, 0, 1), Img(dst.data_ptr(), H, W, 0, 0, 1)
            f = lambda: lib.dmc_boundary_reconstruction(ctx, C.byref(s), C.byref(d), k, k, 1.0, 1.0, 1.0)
            for _ in range(3): assert f() == 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(stream)
            for _ in range(10): f()
            e1.record(stream); torch.cuda.synchronize()
            key = (name, k); same = ""
            if key in ref: same = " same" if torch.equal(ref[key], dst) else " DIFFERENT"
            else: ref[key] = dst.clone()
            out.append("%s %dx%d %.3f ms%s" % (name, k, k, e0.elapsed_time(e1) / 10, same))
    print(os.path.basename(path), " | ".join(out), flush=True)
    lib.dmc_destroy(ctx)
