"""SURVEY.md 8f-3, the point-cloud render that consumes reprojectXYZ's output: projectPointsSimple
(ref:depthmapUtil.cpp:10-156), projectImagefromXYZ (:285-448), fillSmallHole (:187-283), call sites ref:main.cpp:341-373.

CPU part (no GPU): the plain-C port equals the UNMODIFIED reference build bit for bit -- the projected points (the
reference's _mm_rcp_ps is executed as the same instruction), the serial z-buffer splat with and without isSub, and
fillSmallHole -- on several views of the two Kinect fixtures.  That pins the port.

GPU part: libdmc_b200.so equals the port bit for bit.  The GPU carries Intel's RCPPS table, so the rcp comparison needs an
Intel host (tools/gen_rcp_table.c verifies the table against the instruction); the exact-division mode is host independent.
"""
import numpy as np
import pytest

from _util import assert_bits_equal, load_png

cv2 = pytest.importorskip("cv2")
FOCUS, BASELINE, AMP = 75.0, 575.0, 2.6


def rot(yaw, pitch):
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    return np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])


VIEWS = [("identity", np.eye(3), np.zeros(3)), ("shift", np.eye(3), np.array([300., -200., 400.])),
         ("rot", rot(0.15, -0.1), np.array([-400., 100., -300.])), ("zoom_in", rot(-0.3, 0.2), np.array([0., 0., -1500.])),
         ("zoom_out", rot(0.05, 0.0), np.array([50., 20., 9000.]))]


def scene(port, name):
    """pointcloudTest's data flow (main.cpp:303, :322, :341): decoded disparity -> filterDisp8U2Depth32F -> reprojectXYZ;
    the image that is splatted is the grey disparity map itself ("depth map view", main.cpp:345-349)."""
    img = load_png(name); H, W = img.shape
    d32 = port.filter_disp8u_depth32f(img, FOCUS, BASELINE, AMP, 1, 0, 1, 3, 65.0)
    xyz = port.reproject_xyz(d32, 510.0)
    K = np.eye(3) * 510.0; K[0, 2] = (W - 1) * 0.5; K[1, 2] = (H - 1) * 0.5; K[2, 2] = 1.0      # main.cpp:132-136
    return cv2.cvtColor(img, cv2.COLOR_GRAY2BGR), xyz, K


@pytest.mark.parametrize("name", ["kinect_meeting_q80.png", "kinect_desk_q80.png"])
def test_port_equals_reference_build(port, ref, name):
    col, xyz, K = scene(port, name)
    for vname, R, t in VIEWS:
        assert_bits_equal(port.project_points(xyz, R, t, K), ref.project_points(xyz, R, t, K), "%s projectPointsSimple" % vname)
        for sub in (False, True):
            d0, z0 = ref.project_image(col, xyz, R, t, K, sub)
            d1, z1 = port.project_image(col, xyz, R, t, K, sub)
            assert_bits_equal(d1, d0, "%s isSub=%d image" % (vname, sub))
            assert_bits_equal(z1, z0, "%s isSub=%d depth" % (vname, sub))
            assert_bits_equal(port.fill_small_hole(d0), ref.fill_small_hole(d0), "%s fillSmallHole in place" % vname)
            canvas = np.full_like(d0, 37)
            assert_bits_equal(port.fill_small_hole(d0, canvas), ref.fill_small_hole(d0, canvas), "%s fillSmallHole into a canvas" % vname)


def test_port_small_and_degenerate(port, ref):
    rs = np.random.RandomState(3)
    for (H, W) in [(3, 3), (4, 7), (2, 9), (9, 2), (17, 23)]:
        col = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        z = rs.randint(0, 4, (H, W)).astype(np.float32) * 500.0          # zeros -> 10000 (never drawn)
        xyz = port.reproject_xyz(z, 20.0)
        K = np.eye(3) * 20.0; K[0, 2] = (W - 1) * 0.5; K[1, 2] = (H - 1) * 0.5; K[2, 2] = 1.0
        for R, t in ((np.eye(3), np.zeros(3)), (rot(0.2, 0.1), np.array([30., -10., -200.])), (np.eye(3), np.array([0., 0., -500.]))):
            for sub in (False, True):
                d0, z0 = ref.project_image(col, xyz, R, t, K, sub)
                d1, z1 = port.project_image(col, xyz, R, t, K, sub)
                assert_bits_equal(d1, d0, "%dx%d image" % (H, W)); assert_bits_equal(z1, z0, "%dx%d depth" % (H, W))
            assert_bits_equal(port.fill_small_hole(col), ref.fill_small_hole(col), "%dx%d fillSmallHole" % (H, W))


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dmc():
    import depthmapcompression_b200 as m
    m.default_context(0)
    return m


def intel_host():
    try:
        return "GenuineIntel" in open("/proc/cpuinfo").read()
    except Exception:
        return False


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kinect_meeting_q80.png", "kinect_desk_q80.png"])
def test_gpu_render_equals_port(dmc, port, name):
    col, xyz, K = scene(port, name)
    modes = [False] + ([True] if intel_host() else [])          # rcp emulation (Intel table) / exact division
    for vname, R, t in VIEWS:
        for rcp in modes:
            want_pt = port.project_points(xyz, R, t, K, rcp=rcp)
            assert_bits_equal(dmc.projectPointsSimple(xyz, R, t, K, exact_divide=not rcp), want_pt, "%s projectPointsSimple rcp=%d" % (vname, rcp))
            for sub in (False, True):
                want, wz = port.project_image(col, xyz, R, t, K, sub, rcp=rcp)
                got, gz, gpt = dmc.projectImagefromXYZ(col, None, xyz, R, t, K, isSub=sub, want_depth=True, exact_divide=not rcp)
                assert_bits_equal(gpt, want_pt, "%s pt rcp=%d" % (vname, rcp))
                assert_bits_equal(gz, wz, "%s isSub=%d rcp=%d depth" % (vname, sub, rcp))
                assert_bits_equal(got, want, "%s isSub=%d rcp=%d image" % (vname, sub, rcp))
                assert_bits_equal(dmc.projectImagefromXYZ(col, None, xyz, R, t, K, isSub=sub, exact_divide=not rcp), want, "%s first overload" % vname)
        hole_src = port.project_image(col, xyz, R, t, K, True, rcp=False)[0]
        assert_bits_equal(dmc.fillSmallHole(hole_src.copy()), port.fill_small_hole(hole_src), "%s fillSmallHole in place" % vname)
        canvas = np.full_like(hole_src, 37)
        assert_bits_equal(dmc.fillSmallHole(hole_src, canvas.copy()), port.fill_small_hole(hole_src, canvas), "%s fillSmallHole into a canvas" % vname)


@pytest.mark.gpu
def test_gpu_render_small_degenerate_and_1080p(dmc, port):
    rs = np.random.RandomState(3)
    for (H, W) in [(3, 3), (4, 7), (2, 9), (9, 2), (17, 23)]:
        col = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        z = rs.randint(0, 4, (H, W)).astype(np.float32) * 500.0
        xyz = port.reproject_xyz(z, 20.0)
        K = np.eye(3) * 20.0; K[0, 2] = (W - 1) * 0.5; K[1, 2] = (H - 1) * 0.5; K[2, 2] = 1.0
        for R, t in ((np.eye(3), np.zeros(3)), (rot(0.2, 0.1), np.array([30., -10., -200.])), (np.eye(3), np.array([0., 0., -500.]))):
            for sub in (False, True):
                want, wz = port.project_image(col, xyz, R, t, K, sub, rcp=False)
                got, gz, _ = dmc.projectImagefromXYZ(col, None, xyz, R, t, K, isSub=sub, want_depth=True, exact_divide=True)
                assert_bits_equal(got, want, "%dx%d image" % (H, W)); assert_bits_equal(gz, wz, "%dx%d depth" % (H, W))
        assert_bits_equal(dmc.fillSmallHole(col.copy()), port.fill_small_hole(col), "%dx%d fillSmallHole" % (H, W))
    # every point on one pixel (far away camera): long per-pixel lists, the heap-sort path
    H, W = 120, 160
    col = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    z = (500 + rs.randint(0, 3000, (H, W))).astype(np.float32)
    xyz = port.reproject_xyz(z, 100.0)
    K = np.eye(3) * 100.0; K[0, 2] = (W - 1) * 0.5; K[1, 2] = (H - 1) * 0.5; K[2, 2] = 1.0
    for tz in (3e5, 3e7):
        t = np.array([0., 0., tz])
        for sub in (False, True):
            want, wz = port.project_image(col, xyz, np.eye(3), t, K, sub, rcp=False)
            got, gz, _ = dmc.projectImagefromXYZ(col, None, xyz, np.eye(3), t, K, isSub=sub, want_depth=True, exact_divide=True)
            assert_bits_equal(got, want, "collapsed view tz=%g image" % tz); assert_bits_equal(gz, wz, "collapsed view depth")
    # 1080p synthetic
    from oracle.oracle_py import synth_disp
    H, W = 1080, 1920
    disp = synth_disp(H, W, 5)
    d32 = port.filter_disp8u_depth32f(disp, FOCUS, BASELINE, AMP, 1, 0, 1, 3, 65.0)
    xyz = port.reproject_xyz(d32, 1100.0)
    K = np.eye(3) * 1100.0; K[0, 2] = (W - 1) * 0.5; K[1, 2] = (H - 1) * 0.5; K[2, 2] = 1.0
    col = cv2.cvtColor(disp, cv2.COLOR_GRAY2BGR)
    R, t = rot(0.1, -0.05), np.array([-300., 80., -200.])
    want, wz = port.project_image(col, xyz, R, t, K, True, rcp=False)
    got, gz, _ = dmc.projectImagefromXYZ(col, None, xyz, R, t, K, isSub=True, want_depth=True, exact_divide=True)
    assert_bits_equal(got, want, "1080p image"); assert_bits_equal(gz, wz, "1080p depth")
