"""The joint (colour-guided) range filter is an extension with no reference counterpart: its oracle is pinned
to the reference on the slices where the two coincide (guide == src), on CPU; the CUDA kernels are then checked against
that oracle bit for bit, on the GPU."""
import numpy as np
import pytest

from _util import assert_bits_equal, make_image


def _guide(rs, base, cn):
    """A colour guide correlated with the depth image: same edges, different levels, some noise."""
    H, W = base.shape
    chans = []
    for c in range(cn):
        lut = rs.permutation(256).astype(np.uint8)
        g = lut[(base // 16) * 16].astype(np.int32) + rs.randint(-3, 4, size=(H, W))
        chans.append(np.clip(g, 0, 255).astype(np.uint8))
    return chans[0] if cn == 1 else np.ascontiguousarray(np.stack(chans, axis=2))


def test_oracle_coincides_with_reference_when_guide_is_src(port, ref):
    rs = np.random.RandomState(5)
    for shape in [(37, 53), (60, 90), (5, 1), (1, 5)]:
        img = make_image(rs, *shape)
        for (kw, kh), th in [((3, 3), 10), ((7, 7), 5), ((11, 11), 10), ((5, 5), 255), ((9, 9), 0), ((7, 3), 20), ((1, 5), 8), ((0, 3), 8), ((4, 4), 12)]:
            want = ref.bwrf(img, kw, kh, th)
            assert_bits_equal(port.joint_bwrf(img, img, kw, kh, th), want, "joint(guide=src) C1 %dx%d th%d" % (kw, kh, th))
            # three identical guide channels: L1 distance = 3|d|, i.e. the C1 filter with threshold th // 3 (255 stays 255)
            g3 = np.ascontiguousarray(np.repeat(img[:, :, None], 3, axis=2))
            want3 = ref.bwrf(img, kw, kh, 255 if th == 255 else th // 3)
            assert_bits_equal(port.joint_bwrf(img, g3, kw, kh, th), want3, "joint(guide=src x3) %dx%d th%d" % (kw, kh, th))


def test_constant_guide_is_the_window_mean(port):
    rs = np.random.RandomState(6)
    img = make_image(rs, 40, 60, kind="noise")
    guide = np.full((40, 60, 3), 99, np.uint8)
    got = port.joint_bwrf(img, guide, 5, 5, 0)
    # every tap accepted: plain mean over the 21-tap disc with replicated borders, RNE
    pad = np.pad(img.astype(np.int64), 2, mode="edge")
    acc = np.zeros(img.shape, np.int64); n = 0
    for i in range(-2, 3):
        for j in range(-2, 3):
            if i * i + j * j <= 4:
                acc += pad[2 + i:2 + i + 40, 2 + j:2 + j + 60]; n += 1
    want = np.rint(acc.astype(np.float32) / np.float32(n)).astype(np.uint8)
    assert_bits_equal(got, want, "constant guide")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(37, 53), (1, 5), (5, 1), (83, 131), (100, 1023), (64, 256)])
def test_gpu_joint_bwrf(shape, port):
    import depthmapcompression_b200 as dmc
    dmc.default_context(0)
    rs = np.random.RandomState(7); H, W = shape
    base = make_image(rs, H, W)
    for cn in (1, 3):
        guide = _guide(rs, base, cn)
        for (kw, kh) in [(1, 1), (3, 3), (5, 5), (7, 7), (9, 9), (11, 11), (13, 13), (21, 21), (7, 3), (1, 5), (4, 4), (0, 3)]:
            for th in (0, 10, 30, 254, 255, 10.9):
                if kw > 11 and th not in (10, 30):
                    continue
                want = port.joint_bwrf(base, guide, kw, kh, th)
                got = dmc.jointBinalyWeightedRangeFilter(base, guide, None, (kw, kh), th)
                assert_bits_equal(got, want, "joint %dx%d th%s guide C%d" % (kw, kh, th, cn))
    # guide == src reproduces binalyWeightedRangeFilter; in place; dst aliasing the guide
    assert_bits_equal(dmc.jointBinalyWeightedRangeFilter(base, base, None, (11, 11), 10), dmc.binalyWeightedRangeFilter(base, None, (11, 11), 10, dmc.FULL_KERNEL), "guide == src")
    g1 = _guide(rs, base, 1)
    c = base.copy(); dmc.jointBinalyWeightedRangeFilter(c, g1, c, (7, 7), 12)
    assert_bits_equal(c, port.joint_bwrf(base, g1, 7, 7, 12), "in place")
    g = g1.copy(); dmc.jointBinalyWeightedRangeFilter(base, g, g, (7, 7), 12)
    assert_bits_equal(g, port.joint_bwrf(base, g1, 7, 7, 12), "dst is the guide")
    with pytest.raises(dmc.DmcError):
        dmc.jointBinalyWeightedRangeFilter(base, g1[:-1] if H > 1 else g1[:, :-1], None, (3, 3), 10)
    with pytest.raises(dmc.DmcError):
        dmc.jointBinalyWeightedRangeFilter(base, g1, None, (3, 3), 10, dmc.SEPARABLE_KERNEL)


@pytest.mark.gpu
def test_gpu_joint_bwrf_full_size(port):
    import depthmapcompression_b200 as dmc
    from oracle.oracle_py import synth_disp, degrade_blocks
    rs = np.random.RandomState(8)
    base = degrade_blocks(synth_disp(1080, 1920, 4), 4)
    guide = _guide(rs, base, 3)
    assert_bits_equal(dmc.jointBinalyWeightedRangeFilter(base, guide, None, (11, 11), 30), port.joint_bwrf(base, guide, 11, 11, 30), "1080p joint r5")
