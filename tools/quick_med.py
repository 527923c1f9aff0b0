"""Times dmc_median_blur of several builds of the library (tuning aid): python tools/quick_med.py lib1.so lib2.so ..."""
import ctypes as C, sys, torch
class Img(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("cvtype", C.c_int), ("step", C.c_size_t), ("mem", C.c_int)]
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
H, W, NF = 1080, 1920, 100
g = torch.Generator(device=dev); g.manual_seed(1)
src = torch.randint(0, 256, (NF * H, W), dtype=torch.uint8, device=dev, generator=g); dst = torch.empty_like(src)
ref = None
for path in sys.argv[1:]:
    lib = C.CDLL(path); ctx = C.c_void_p()
    assert lib.dmc_create(0, C.byref(ctx)) == 0
    lib.dmc_set_stream(ctx, C.c_void_p(stream.cuda_stream))
    out = []
    for k in (5, 3):
        s, d = Img(src.data_ptr(), NF * H, W, 0, 0, 1), Img(dst.data_ptr(), NF * H, W, 0, 0, 1)
        f = lambda: lib.dmc_median_blur(ctx, C.byref(s), C.byref(d), k)
        for _ in range(3): assert f() == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(10): f()
        e1.record(stream); torch.cuda.synchronize()
        out.append("k=%d %.3f ms/100f" % (k, e0.elapsed_time(e1) / 10))
        if k == 5:
            if ref is None: ref = dst.clone()
            else: out.append("same" if torch.equal(ref, dst) else "DIFFERENT")
    print(path.split("/")[-1], " ".join(out), flush=True)
    lib.dmc_destroy(ctx)
