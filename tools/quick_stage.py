"""Times the three front stages (median 5x5, Gaussian 3x3, min-max r=3) of several builds of the library on a 100-frame
1080p strip (tuning aid): python tools/quick_stage.py lib1.so lib2.so ..."""
import ctypes as C, sys, torch
class Img(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("cvtype", C.c_int), ("step", C.c_size_t), ("mem", C.c_int)]
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
H, W, NF = 1080, 1920, 100
g = torch.Generator(device=dev); g.manual_seed(1)
src = torch.randint(0, 256, (NF * H, W), dtype=torch.uint8, device=dev, generator=g); dst = torch.empty_like(src)
ref = {}
for path in sys.argv[1:]:
    lib = C.CDLL(path); ctx = C.c_void_p()
    assert lib.dmc_create(0, C.byref(ctx)) == 0
    lib.dmc_set_stream(ctx, C.c_void_p(stream.cuda_stream))
    lib.dmc_small_gaussian.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    s, d = Img(src.data_ptr(), NF * H, W, 0, 0, 1), Img(dst.data_ptr(), NF * H, W, 0, 0, 1)
    ops = {"median5": lambda: lib.dmc_median_blur(ctx, C.byref(s), C.byref(d), 5),
           "gauss3": lambda: lib.dmc_small_gaussian(ctx, C.byref(s), C.byref(d), 3, 1.5),
           "minmax3": lambda: lib.dmc_blur_remove_minmax(ctx, C.byref(s), C.byref(d), 3)}
    out = []
    for name, f in ops.items():
        for _ in range(3): assert f() == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(10): f()
        e1.record(stream); torch.cuda.synchronize()
        same = ""
        if name in ref: same = " same" if torch.equal(ref[name], dst) else " DIFFERENT"
        else: ref[name] = dst.clone()
        out.append("%s %.3f ms%s" % (name, e0.elapsed_time(e1) / 10, same))
    print(path.split("/")[-1], " | ".join(out), flush=True)
    lib.dmc_destroy(ctx)
