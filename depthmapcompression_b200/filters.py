"""Host-side mirror of the reference's filter.h / util.h operators over numpy arrays (the role cv::Mat plays in the
reference).  Same names, argument order, defaults and error behaviour:

  * a type/size violation that trips CV_Assert in the reference raises DmcError here;
  * the reference's silent no-ops ((type, method) pairs its dispatcher ignores) leave `dst` untouched here too;
  * `dst=None` plays the role of an empty Mat: it is allocated like the reference allocates it;
  * in-place calls (dst is src) are legal for every operator.

Every call is one C-ABI call into libdmc_b200.so; nothing is computed in Python.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import lib, DmcImage, DmcChainParams, DmcError, MEM_HOST, MEM_DEVICE, FULL_KERNEL, BORDER_REPLICATE

_DEPTH = {np.dtype(np.uint8): capi.CV_8U, np.dtype(np.uint16): capi.CV_16U, np.dtype(np.int16): capi.CV_16S,
          np.dtype(np.float32): capi.CV_32F, np.dtype(np.float64): capi.CV_64F}


def _cvtype(a):
    if a.dtype not in _DEPTH:
        raise DmcError(capi.DMC_ERR_TYPE, "unsupported dtype %s" % a.dtype)
    cn = 1 if a.ndim == 2 else a.shape[2]
    return _DEPTH[a.dtype] + ((cn - 1) << 3)


def _img(a):
    """numpy array (rows x cols [x cn], last dims contiguous, row stride free) -> dmc_image"""
    if a.ndim not in (2, 3):
        raise DmcError(capi.DMC_ERR_SIZE, "image must be 2-D or 3-D")
    es = a.dtype.itemsize * (1 if a.ndim == 2 else a.shape[2])
    inner_ok = (a.strides[-1] == a.dtype.itemsize) and (a.ndim == 2 or a.strides[1] == es)
    if not inner_ok:
        raise DmcError(capi.DMC_ERR_SIZE, "pixels of a row must be contiguous")
    return DmcImage(a.ctypes.data, a.shape[0], a.shape[1], _cvtype(a), a.strides[0] if a.shape[0] > 1 else 0, MEM_HOST)


class Context:
    """One dmc_ctx: a device, a stream and scratch memory.  Not thread-safe (like a PostFilterSet instance)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib.dmc_create(int(device), C.byref(h))
        if rc != capi.DMC_OK:
            raise DmcError(rc, (lib.dmc_last_error(None) or b"").decode())
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            lib.dmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc < 0:
            raise DmcError(rc, (lib.dmc_last_error(self.h) or b"").decode())
        return rc

    def synchronize(self):
        self.check(lib.dmc_synchronize(self.h))

    def set_stream(self, cuda_stream):
        self.check(lib.dmc_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_gateway(self, device):
        """Route this context's host traffic over another device's link (kernels reach its HBM through NVLink); -1 = own link."""
        self.check(lib.dmc_set_gateway(self.h, int(device)))

    @property
    def gateway(self):
        return int(lib.dmc_get_gateway(self.h))

    def set_lanes(self, lanes):
        self.check(lib.dmc_set_lanes(self.h, int(lanes)))

    def profile_enable(self, stage_mask):
        self.check(lib.dmc_profile_enable(self.h, int(stage_mask)))

    def profile_read(self, stage, reset=True):
        """-> (total_ms, launches, pixels) of the event-bracketed launches of `stage` since the last reset"""
        ms, n, px = C.c_double(), C.c_uint64(), C.c_uint64()
        self.check(lib.dmc_profile_read(self.h, stage, C.byref(ms), C.byref(n), C.byref(px), int(reset)))
        return ms.value, n.value, px.value

    @property
    def kernel_launches(self):
        return int(lib.dmc_kernel_launches(self.h))

    # frame batches: src/dst are device pointers (ints) or C-contiguous numpy arrays [n, rows, cols]
    def chain_batch(self, src, dst, n_frames, rows, cols, params, device=False):
        sp = C.c_void_p(src if isinstance(src, int) else src.ctypes.data)
        dp = C.c_void_p(dst if isinstance(dst, int) else dst.ctypes.data)
        return self.check(lib.dmc_chain_batch(self.h, sp, dp, n_frames, rows, cols, C.byref(params), MEM_DEVICE if device else MEM_HOST))

    def chain_batch_jpeg(self, streams, rows, cols, dst, params, device=False):
        """dmc_chain_batch_jpeg: JPEG bitstreams (list of bytes, or a (blob, offsets) pair of numpy arrays / (pointer, offsets)
        for pinned memory) -> decode -> chain, streamed; dst is a host array [n, rows, cols] or a pointer (device=True:
        device memory)."""
        blob, offsets = streams if isinstance(streams, tuple) else pack_streams(streams)
        n = len(offsets) - 1
        bp = C.c_void_p(blob if isinstance(blob, int) else blob.ctypes.data)
        dp = C.c_void_p(dst if isinstance(dst, int) else dst.ctypes.data)
        return self.check(lib.dmc_chain_batch_jpeg(self.h, bp, C.c_void_p(offsets.ctypes.data), n, rows, cols, dp, MEM_DEVICE if device else MEM_HOST, C.byref(params)))

    def chain_batch_images(self, srcs, dsts, params):
        """dmc_chain_batch_images: lists of 2-D numpy arrays (any row stride), one size; dsts are written in place."""
        n = len(srcs)
        S = (capi.DmcImage * n)(*[_img(a) for a in srcs]); D = (capi.DmcImage * n)(*[_img(a) for a in dsts])
        return self.check(lib.dmc_chain_batch_images(self.h, S, D, n, C.byref(params)))


class FrameBatchScheduler:
    """dmc_sched: persistent contexts on several GPUs of one box; chain_batch() shards a host batch over them."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = lib.dmc_sched_create(devs, len(devices), C.byref(h))
        if rc != capi.DMC_OK:
            raise DmcError(rc, (lib.dmc_last_error(None) or b"").decode())
        self.h = h

    def chain_batch(self, src, dst, n_frames, rows, cols, params):
        sp = C.c_void_p(src if isinstance(src, int) else src.ctypes.data)
        dp = C.c_void_p(dst if isinstance(dst, int) else dst.ctypes.data)
        rc = lib.dmc_sched_chain_batch(self.h, sp, dp, n_frames, rows, cols, C.byref(params))
        if rc < 0:
            raise DmcError(rc, (lib.dmc_sched_last_error(self.h) or b"").decode())
        return rc

    def routing(self):
        """-> (gateways per device (-1 = own link), GB/s each way with every device on its own link, GB/s each way as routed)"""
        n = lib.dmc_sched_device_count(self.h)
        g = (C.c_int * max(n, 1))(); a, b = C.c_double(), C.c_double()
        lib.dmc_sched_get_routing(self.h, g, C.byref(a), C.byref(b))
        return list(g)[:n], a.value, b.value

    def close(self):
        if getattr(self, "h", None):
            lib.dmc_sched_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hostlink_probe(devices):
    """dmc_hostlink_probe: measures the host link of `devices` (all copying both ways at once) and proposes which devices'
    links should carry the traffic.  Call it while the devices are idle.  -> dict"""
    devs = (C.c_int * len(devices))(*devices)
    info = capi.DmcHostlinkInfo()
    rc = lib.dmc_hostlink_probe(devs, len(devices), C.byref(info))
    if rc != capi.DMC_OK:
        raise DmcError(rc, "dmc_hostlink_probe failed")
    n = info.n_devices
    return {"devices": list(info.device)[:n], "gateway": list(info.gateway)[:n], "loaded_gbs": [round(x, 2) for x in list(info.loaded_gbs)[:n]],
            "all_gbs": round(info.all_gbs, 2), "best_gbs": round(info.best_gbs, 2), "n_link": info.n_link}


def multi_chain_batch(devices, src, dst, params):
    """Frame-batch scheduler across GPUs in one process: src/dst are host arrays [n, rows, cols]; frames are sharded
    contiguously over `devices` (one host thread and context per device, no collective)."""
    n, rows, cols = src.shape
    devs = (C.c_int * len(devices))(*devices)
    err = C.create_string_buffer(512)
    rc = lib.dmc_multi_chain_batch(devs, len(devices), C.c_void_p(src.ctypes.data), C.c_void_p(dst.ctypes.data), n, rows, cols, C.byref(params), err, 512)
    if rc < 0:
        raise DmcError(rc, err.value.decode())
    return dst


def pack_streams(streams):
    """list of bytes-like JPEG streams -> (blob uint8 array, offsets uint64 array of len n+1)"""
    sizes = [len(x) for x in streams]
    offsets = np.zeros(len(streams) + 1, np.uint64); offsets[1:] = np.cumsum(sizes)
    blob = np.empty(int(offsets[-1]) + 16, np.uint8)
    for i, x in enumerate(streams):
        blob[int(offsets[i]):int(offsets[i + 1])] = np.frombuffer(x, np.uint8) if not isinstance(x, np.ndarray) else x.ravel()
    return blob, offsets


def jpegProbe(stream):
    """dmc_jpeg_probe (host only): -> (rows, cols) of a stream this library can decode; raises DmcError with the reason
    otherwise (malformed / truncated / progressive / colour ...)."""
    a = np.frombuffer(stream, np.uint8) if not isinstance(stream, np.ndarray) else np.ascontiguousarray(stream).ravel()
    r, c = C.c_int(), C.c_int(); err = C.create_string_buffer(256)
    rc = lib.dmc_jpeg_probe(C.c_void_p(a.ctypes.data if a.size else None), a.size, C.byref(r), C.byref(c), err, 256)
    if rc != capi.DMC_OK:
        raise DmcError(rc, err.value.decode())
    return r.value, c.value


def jpegDecodeGrayBatch(streams, rows, cols, dst=None, ctx=None):
    """JPEG bitstreams (list of bytes, or a (blob, offsets) pair) -> [n, rows, cols] uint8, bit-identical to
    cv::imdecode(buf, 0) as the reference calls it (main.cpp:284, :521).  dst: None (new host array), a host array,
    or a device pointer (int)."""
    ctx = ctx or default_context()
    blob, offsets = streams if isinstance(streams, tuple) else pack_streams(streams)
    n = len(offsets) - 1
    device = isinstance(dst, int)
    if dst is None:
        dst = np.empty((n, rows, cols), np.uint8)
    dp = C.c_void_p(dst if device else dst.ctypes.data)
    ctx.check(lib.dmc_jpeg_decode_gray_batch(ctx.h, C.c_void_p(blob.ctypes.data), C.c_void_p(offsets.ctypes.data), n, rows, cols, dp,
                                             MEM_DEVICE if device else MEM_HOST))
    return dst


_default = {}


def default_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


def chain_params(chain, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method=FULL_KERNEL, focus=0.0, baseline=0.0, amp=0.0):
    return DmcChainParams(chain, median_r, gaussian_r, minmax_r, brange_r, float(brange_th), brange_method, focus, baseline, amp)


def _out(dst, src, dtype=None, shape=None):
    dtype = np.dtype(dtype or src.dtype); shape = shape or src.shape
    if dst is None or dst.shape != tuple(shape) or dst.dtype != dtype:
        return np.empty(shape, dtype)          # Mat::create on an empty / mismatching Mat
    return dst


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            lib.dmc_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint8):
    """numpy array in pinned, device-mapped host memory (dmc_host_alloc): single frames of up to 2 MB are then processed in place
    over the host link, and the streaming entry points overlap their copies.  Freed with the array."""
    dtype = np.dtype(dtype); n = int(np.prod(shape)) * dtype.itemsize
    ptr = lib.dmc_host_alloc(max(n, 1))
    if not ptr:
        raise DmcError(capi.DMC_ERR_CUDA, "dmc_host_alloc failed")
    owner = _PinnedOwner(ptr)
    buf = (C.c_uint8 * max(n, 1)).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[ptr] = owner            # numpy keeps `buf` alive, not `owner`: tie the allocation to the interpreter unless released
    return a


_PINNED = {}


def pinned_free(a):
    """releases an array made by pinned_empty (optional: otherwise it lives until the interpreter exits)"""
    _PINNED.pop(a.ctypes.data, None)


def host_register(a):
    """pins and device-maps an existing contiguous numpy array (cudaHostRegister) -- see dmc_host_register"""
    rc = lib.dmc_host_register(C.c_void_p(a.ctypes.data), a.nbytes)
    if rc < 0:
        raise DmcError(rc, "dmc_host_register failed")


def host_unregister(a):
    rc = lib.dmc_host_unregister(C.c_void_p(a.ctypes.data))
    if rc < 0:
        raise DmcError(rc, "dmc_host_unregister failed")


class PostFilterSet:
    """filter.h:32-42.  Owns nothing in Python: the scratch Mats `buff`, `bufff` of the reference live in the dmc_ctx."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()

    def __call__(self, src, dest, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method=FULL_KERNEL):
        dest = _out(dest, src, np.uint8)
        s, d = _img(src), _img(dest)
        self.ctx.check(lib.dmc_post_filter_set(self.ctx.h, C.byref(s), C.byref(d), median_r, gaussian_r, minmax_r, brange_r, int(brange_th), brange_method))
        return dest

    def filterDisp8U2Depth32F(self, src, dest, focus, baseline, amp, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method=FULL_KERNEL):
        dest = _out(dest, src, np.float32)
        s, d = _img(src), _img(dest)
        self.ctx.check(lib.dmc_filter_disp8u_depth32f(self.ctx.h, C.byref(s), C.byref(d), focus, baseline, amp, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method))
        return dest

    def filterDisp8U2Depth16U(self, src, dest, focus, baseline, amp, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method=FULL_KERNEL):
        dest = _out(dest, src, np.uint16)
        s, d = _img(src), _img(dest)
        self.ctx.check(lib.dmc_filter_disp8u_depth16u(self.ctx.h, C.byref(s), C.byref(d), focus, baseline, amp, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method))
        return dest

    def filterDisp8U2Disp32F(self, src, dest, median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method=FULL_KERNEL):
        dest = _out(dest, src, np.uint16)
        s, d = _img(src), _img(dest)
        self.ctx.check(lib.dmc_filter_disp8u_disp32f(self.ctx.h, C.byref(s), C.byref(d), median_r, gaussian_r, minmax_r, brange_r, brange_th, brange_method))
        return dest


def _ksize(k):
    return (k, k) if isinstance(k, int) else (int(k[0]), int(k[1]))     # Size(width, height)


def binalyWeightedRangeFilter(src, dst, kernelSize, threshold, method, borderType=BORDER_REPLICATE, ctx=None):
    """filter.h:29.  `dst=None`: created like `dst.create(src.size(), src.type())` (binalyWeightedRangeFilter.cpp:1108);
    for the reference's no-op (type, method) pairs a fresh dst therefore holds unspecified bytes, as in the reference."""
    ctx = ctx or default_context()
    kw, kh = _ksize(kernelSize)
    dst = _out(dst, src)
    s, d = _img(src), _img(dst)
    ctx.check(lib.dmc_bwrf(ctx.h, C.byref(s), C.byref(d), kw, kh, threshold, method, borderType))
    return dst


def jointBinalyWeightedRangeFilter(src, guide, dst, kernelSize, threshold, method=FULL_KERNEL, ctx=None):
    """Extension (no reference counterpart): binalyWeightedRangeFilter on the 8UC1 image `src` with the binary
    weights taken from `guide` (8UC3 colour or 8UC1) instead of from src.  guide == src gives binalyWeightedRangeFilter."""
    ctx = ctx or default_context()
    kw, kh = _ksize(kernelSize)
    dst = _out(dst, src)
    s, g, d = _img(src), _img(guide), _img(dst)
    ctx.check(lib.dmc_joint_bwrf(ctx.h, C.byref(s), C.byref(g), C.byref(d), kw, kh, threshold, method))
    return dst


def blurRemoveMinMax(src, dest, r, ctx=None):
    """filter.h:19"""
    ctx = ctx or default_context()
    dest = _out(dest, src)
    s, d = _img(src), _img(dest)
    ctx.check(lib.dmc_blur_remove_minmax(ctx.h, C.byref(s), C.byref(d), r))
    return dest


blurRemoveMinMaxBase = blurRemoveMinMax      # filter.h:20 -- scalar twin of the same computation (minmaxFilter.cpp:7-46)


def maxFilter(src, dest, ksize, borderType=BORDER_REPLICATE, ctx=None):
    """filter.h:17"""
    ctx = ctx or default_context()
    kw, kh = _ksize(ksize)
    dest = _out(dest, src)
    s, d = _img(src), _img(dest)
    ctx.check(lib.dmc_max_filter(ctx.h, C.byref(s), C.byref(d), kw, kh, borderType))
    return dest


def minFilter(src, dest, ksize, borderType=BORDER_REPLICATE, ctx=None):
    """filter.h:18"""
    ctx = ctx or default_context()
    kw, kh = _ksize(ksize)
    dest = _out(dest, src)
    s, d = _img(src), _img(dest)
    ctx.check(lib.dmc_min_filter(ctx.h, C.byref(s), C.byref(d), kw, kh, borderType))
    return dest


def boundaryReconstructionFilter(src, dest, ksize, frec, color, space, ctx=None):
    """filter.h:45"""
    ctx = ctx or default_context()
    kw, kh = _ksize(ksize)
    dest = _out(dest, src)
    s, d = _img(src), _img(dest)
    ctx.check(lib.dmc_boundary_reconstruction(ctx.h, C.byref(s), C.byref(d), kw, kh, frec, color, space))
    return dest


def minmaxBoundaryReconstructionFilter(src, dest, r, ksize, frec, color, space, ctx=None):
    """Extension (filter_ext.h): blurRemoveMinMax(r) fused into boundaryReconstructionFilter, one kernel, src read once;
    bit-identical to the two reference calls.  8U / 16U / 16S, single channel."""
    ctx = ctx or default_context()
    kw, kh = _ksize(ksize)
    dest = _out(dest, src)
    s, d = _img(src), _img(dest)
    ctx.check(lib.dmc_minmax_boundary_reconstruction(ctx.h, C.byref(s), C.byref(d), r, kw, kh, frec, color, space))
    return dest


def smallGaussianBlur(src, dest, d, sigma, ctx=None):
    """filter.h:14"""
    ctx = ctx or default_context()
    dest = _out(dest, src)
    s, o = _img(src), _img(dest)
    ctx.check(lib.dmc_small_gaussian(ctx.h, C.byref(s), C.byref(o), d, sigma))
    return dest


def medianBlur(src, dst, ksize, ctx=None):
    """cv::medianBlur as the chain calls it (postFilterSet.cpp:23)"""
    ctx = ctx or default_context()
    dst = _out(dst, src)
    s, d = _img(src), _img(dst)
    ctx.check(lib.dmc_median_blur(ctx.h, C.byref(s), C.byref(d), ksize))
    return dst


def _convert(fn, src, dest, ddtype, fb, a, b, ctx, zero_new=True):
    ctx = ctx or default_context()
    if dest is None or dest.shape != src.shape or dest.dtype != np.dtype(ddtype):
        dest = np.zeros(src.shape, ddtype)     # Mat::zeros(src.size(), type) (depthmapUtil.cpp:925-926)
    s, d = _img(src), _img(dest)
    ctx.check(fn(ctx.h, C.byref(s), C.byref(d), fb, a, b))
    return dest


def disp8U2depth32F(src, dest, focal_baseline, a=1.0, b=0.0, ctx=None):
    """util.h:28"""
    return _convert(lib.dmc_disp8u2depth32f, src, dest, np.float32, focal_baseline, a, b, ctx)


def depth32F2disp8U(src, dest, focal_baseline, a=1.0, b=0.0, ctx=None):
    """util.h:25"""
    return _convert(lib.dmc_depth32f2disp8u, src, dest, np.uint8, focal_baseline, a, b, ctx)


def depth16U2disp8U(src, dest, focal_baseline, a=1.0, b=0.0, ctx=None):
    """util.h:27"""
    return _convert(lib.dmc_depth16u2disp8u, src, dest, np.uint8, focal_baseline, a, b, ctx)


def disp16S2depth16U(src, dest, focal_baseline, a=1.0, b=0.0, ctx=None):
    """util.h:26"""
    return _convert(lib.dmc_disp16s2depth16u, src, dest, np.uint16, focal_baseline, a, b, ctx)


def fillOcclusion(src, invalidvalue, disp_or_depth=capi.FILL_DEPTH, ctx=None):
    """util.h:24 -- in place"""
    ctx = ctx or default_context()
    s = _img(src)
    ctx.check(lib.dmc_fill_occlusion(ctx.h, C.byref(s), int(invalidvalue), disp_or_depth))
    return src


def transpose(src, dst=None, ctx=None):
    """cv::transpose (main.cpp:258, :260), single channel"""
    ctx = ctx or default_context()
    if dst is None or dst.shape != (src.shape[1], src.shape[0]) or dst.dtype != src.dtype:
        dst = np.empty((src.shape[1], src.shape[0]), src.dtype)
    s, d = _img(src), _img(dst)
    ctx.check(lib.dmc_transpose(ctx.h, C.byref(s), C.byref(d)))
    return dst


def reprojectXYZ(depth, xyz, f, ctx=None):
    """util.h:11 -- xyz is (rows*cols) x 1 x 3 float32"""
    ctx = ctx or default_context()
    n = depth.shape[0] * depth.shape[1]
    if xyz is None or xyz.size != n * 3 or xyz.dtype != np.float32:
        xyz = np.zeros((n, 1, 3), np.float32)          # Mat::zeros(area, 1, CV_32FC3) depthmapUtil.cpp:453
    s = _img(depth)
    x3 = xyz.reshape(n, 1, 3)
    d = _img(x3)
    ctx.check(lib.dmc_reproject_xyz(ctx.h, C.byref(s), C.byref(d), f))
    return xyz


# ---- point-cloud render (util.h:12-13, :25, :33) ------------------------------------------------------------------------
def _cam(R, t, K):
    R = np.ascontiguousarray(R, np.float64).reshape(9); t = np.ascontiguousarray(t, np.float64).reshape(3); K = np.ascontiguousarray(K, np.float64).reshape(9)
    DP = C.POINTER(C.c_double)
    return (R, t, K), (R.ctypes.data_as(DP), t.ctypes.data_as(DP), K.ctypes.data_as(DP))


def projectPointsSimple(xyz, R, t, K, dest=None, exact_divide=False, ctx=None):
    """util.h:33: xyz [n, 3] float32 -> [n, 2] float32 (the reference fills a vector<Point2f>)."""
    ctx = ctx or default_context()
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 1, 3); n = xyz.shape[0]
    dest = _out(dest, xyz, np.float32, (n, 1, 2))
    keep, cam = _cam(R, t, K)
    a, b = _img(xyz), _img(dest)
    ctx.check(lib.dmc_project_points(ctx.h, C.byref(a), cam[0], cam[1], cam[2], C.byref(b), capi.RENDER_EXACT_DIVIDE if exact_divide else 0))
    return dest.reshape(n, 2)


def projectImagefromXYZ(image, destimage, xyz, R, t, K, dist=None, mask=None, isSub=False, want_depth=False, exact_divide=False, ctx=None):
    """util.h:12-13.  `dist` and `mask` are accepted and ignored, as in the reference.  Returns destimage, or
    (destimage, depth, pt) with want_depth (the second overload's outputs)."""
    ctx = ctx or default_context()
    if image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:
        raise DmcError(capi.DMC_ERR_TYPE, "projectImagefromXYZ: image must be CV_8UC3")
    H, W = image.shape[:2]
    destimage = _out(destimage, image)
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 1, 3)
    keep, cam = _cam(R, t, K)
    a, d, x = _img(image), _img(destimage), _img(xyz)
    depth = np.empty((H, W), np.float32) if want_depth else None
    pt = np.empty((H * W, 1, 2), np.float32) if want_depth else None
    zi = C.byref(_img(depth)) if want_depth else None
    pi = C.byref(_img(pt)) if want_depth else None
    ctx.check(lib.dmc_project_image_from_xyz(ctx.h, C.byref(a), C.byref(d), C.byref(x), cam[0], cam[1], cam[2], int(bool(isSub)), zi, pi,
                                             capi.RENDER_EXACT_DIVIDE if exact_divide else 0))
    return (destimage, depth, pt.reshape(-1, 2)) if want_depth else destimage


def fillSmallHole(src, dest=None, ctx=None):
    """util.h:25.  dest=None or dest is src: in place, as main.cpp:355 calls it; otherwise dest keeps its content outside the holes."""
    ctx = ctx or default_context()
    if dest is None:
        dest = src
    a, b = _img(src), _img(dest)
    ctx.check(lib.dmc_fill_small_hole(ctx.h, C.byref(a), C.byref(b)))
    return dest
