"""Boundary reconstruction filter on the tile-rank kernel (csrc/dmc_brf.cu) and the fused min-max -> BRF extension,
bit-exact against the oracle port (ref: boundaryReconstructionFilter.cpp:12-131, minmaxFilter.cpp:48-174).  The cases
are chosen to reach every path of the kernel: flat tiles, tiles with <= 32 / <= 256 / > 256 distinct values (one pass,
several passes, per-thread-list fallback), the hashed id table of the 16-bit / float / double types, float tiles with
NaN / -0 (list fallback), non-square windows, images smaller than a tile and ragged edges."""
import numpy as np
import pytest

from _util import assert_bits_equal, make_image, load_png

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dmc():
    import depthmapcompression_b200 as m
    return m


@pytest.fixture(scope="module")
def port():
    from oracle.oracle_py import Port
    return Port()


def graded(rs, H, W, dt, levels):
    """`levels` distinct values per 40x40 neighbourhood: smooth ramp + bounded noise"""
    yy, xx = np.mgrid[0:H, 0:W]
    a = (xx // 3 + yy // 5) % 200 + rs.randint(0, max(1, levels), size=(H, W))
    if dt == np.uint8:
        return np.clip(a, 0, 255).astype(dt)
    if dt == np.uint16:
        return (a * 37 + 1000).astype(dt)
    if dt == np.int16:
        return (a * 37 - 3000).astype(dt)
    return (a * 0.37 - 11.0).astype(dt)


@pytest.mark.parametrize("dt", [np.uint8, np.uint16, np.int16, np.float32, np.float64])
def test_brf_rank_paths(dmc, port, dt):
    rs = np.random.RandomState(5)
    for (H, W) in [(37, 70), (64, 96), (9, 33)]:
        for levels in (1, 3, 20, 90):
            a = graded(rs, H, W, dt, levels)
            for k in ((3, 3), (7, 7), (13, 13), (5, 9), (21, 21)):
                if levels >= 20 and k[0] > 13:
                    continue
                assert_bits_equal(dmc.boundaryReconstructionFilter(a, None, k, 1.0, 1.0, 1.0), port.brf(a, k[0], k[1], 1.0, 1.0, 1.0),
                                  "brf %s %dx%d levels %d k%s" % (np.dtype(dt).name, H, W, levels, k))
    # > 256 distinct values in a tile: the per-thread list inside the same kernel
    if dt != np.uint8:
        a = (rs.randint(0, 30000, size=(40, 70))).astype(dt)
        assert_bits_equal(dmc.boundaryReconstructionFilter(a, None, (7, 7), 1.0, 0.5, 2.0), port.brf(a, 7, 7, 1.0, 0.5, 2.0), "brf many values")
    a = rs.randint(0, 256, size=(50, 90)).astype(dt)       # up to 256 values: several passes per tile
    assert_bits_equal(dmc.boundaryReconstructionFilter(a, None, (9, 9), 1.0, 1.0, 1.0), port.brf(a, 9, 9, 1.0, 1.0, 1.0), "brf 256 values")


def test_brf_float_special_values(dmc, port):
    rs = np.random.RandomState(6)
    a = graded(rs, 48, 80, np.float32, 4)
    a[5, 7] = -0.0; a[5, 8] = 0.0; a[30, 60] = np.nan; a[31, 61] = np.inf; a[10, 40] = -np.inf
    a.view(np.uint32)[40, 20] = 0xFFFFFFFF                      # the NaN whose bits are the hash table's "empty" marker
    assert_bits_equal(dmc.boundaryReconstructionFilter(a, None, (7, 7), 1.0, 1.0, 1.0), port.brf(a, 7, 7, 1.0, 1.0, 1.0), "brf float specials")
    d = a.astype(np.float64); d.view(np.uint64)[41, 21] = 0xFFFFFFFFFFFFFFFF
    assert_bits_equal(dmc.boundaryReconstructionFilter(d, None, (5, 5), 1.0, 1.0, 1.0), port.brf(d, 5, 5, 1.0, 1.0, 1.0), "brf double specials")


def test_brf_fixture_1080p(dmc, port):
    img = load_png("kinect_desk_q50.png")
    big = np.ascontiguousarray(np.tile(img, (3, 3))[:1080, :1920])
    port.set_num_threads(0)
    assert_bits_equal(dmc.boundaryReconstructionFilter(big, None, (13, 13), 1.0, 1.0, 1.0), port.brf(big, 13, 13, 1.0, 1.0, 1.0), "brf 1080p fixture")


@pytest.mark.parametrize("dt", [np.uint8, np.uint16, np.int16])
def test_fused_minmax_brf(dmc, port, dt):
    """extension: one kernel == blurRemoveMinMax then boundaryReconstructionFilter (oracle = composition of the two ports)"""
    rs = np.random.RandomState(7)
    for (H, W) in [(37, 70), (64, 96), (9, 33), (100, 131)]:
        for kind in ("pw", "noise", "const"):
            a = make_image(rs, H, W, dt, kind=kind)
            for (r, k) in [(1, (3, 3)), (3, (7, 7)), (2, (13, 13)), (3, (5, 9)), (0, (7, 7)), (5, (9, 9))]:
                if kind == "noise" and k[0] > 9:
                    continue
                want = port.brf(port.blur_remove_minmax(a, r), k[0], k[1], 1.0, 1.0, 1.0)
                got = dmc.minmaxBoundaryReconstructionFilter(a, None, r, k, 1.0, 1.0, 1.0)
                assert_bits_equal(got, want, "minmax+brf %s %dx%d %s r%d k%s" % (np.dtype(dt).name, H, W, kind, r, k))
    a = make_image(rs, 60, 90, dt); c = a.copy()
    dmc.minmaxBoundaryReconstructionFilter(c, c, 3, (7, 7), 1.0, 1.0, 1.0)
    assert_bits_equal(c, port.brf(port.blur_remove_minmax(a, 3), 7, 7, 1.0, 1.0, 1.0), "minmax+brf in place")
    with pytest.raises(Exception):
        dmc.minmaxBoundaryReconstructionFilter(a.astype(np.float32), None, 3, (7, 7), 1.0, 1.0, 1.0)


def test_fused_minmax_brf_fixture(dmc, port):
    img = load_png("kinect_desk_q50.png")
    want = port.brf(port.blur_remove_minmax(img, 3), 13, 13, 1.0, 1.0, 1.0)
    assert_bits_equal(dmc.minmaxBoundaryReconstructionFilter(img, None, 3, (13, 13), 1.0, 1.0, 1.0), want, "minmax+brf fixture")


def test_large_shared_memory_kernels_on_every_device(dmc, port):
    """The kernels that ask for more than 48 KB of dynamic shared memory (boundary reconstruction, JPEG frame decode) set a
    per-function, per-DEVICE attribute: they must work on every visible GPU, not only on the first one that ran them."""
    import torch
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(9)
    a = graded(rs, 64, 96, np.uint8, 5)
    want = port.brf(a, 7, 7, 1.0, 1.0, 1.0)
    enc = cv2.imencode(".jpg", a, [cv2.IMWRITE_JPEG_QUALITY, 80])[1].tobytes()
    dec = cv2.imdecode(np.frombuffer(enc, np.uint8), 0)
    for dev in range(torch.cuda.device_count()):
        ctx = dmc.Context(dev)
        assert_bits_equal(dmc.boundaryReconstructionFilter(a, None, (7, 7), 1.0, 1.0, 1.0, ctx=ctx), want, "brf on device %d" % dev)
        assert_bits_equal(dmc.jpegDecodeGrayBatch([enc], 64, 96, ctx=ctx)[0], dec, "jpeg decode on device %d" % dev)


def test_contexts_release_their_device_memory(dmc, port):
    """dmc_destroy gives back everything a context allocated (scratch of every operator family, JPEG and render buffers,
    pinned staging, graphs): create / use / destroy in a loop and watch the device's free memory."""
    import torch
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(3)
    a = np.maximum(make_image(rs, 480, 640), 1)
    enc = cv2.imencode(".jpg", a, [cv2.IMWRITE_JPEG_QUALITY, 80])[1].tobytes()

    def use(ctx):
        pfs = dmc.PostFilterSet(ctx)
        pfs(a, None, 2, 1, 3, 5, 10); pfs.filterDisp8U2Depth32F(a, None, 75.0, 575.0, 2.6, 1, 0, 1, 3, 65.0)
        dmc.boundaryReconstructionFilter(a, None, (7, 7), 1.0, 1.0, 1.0, ctx=ctx)
        dmc.jpegDecodeGrayBatch([enc] * 4, 480, 640, ctx=ctx)
        ctx.synchronize()

    c = dmc.Context(0); use(c); c.close(); torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info(0)[0]
    for _ in range(12):
        c = dmc.Context(0); use(c); c.close()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info(0)[0]
    assert free0 - free1 < 32 << 20, "device memory leaked across contexts: %.1f MB" % ((free0 - free1) / 1e6)
