"""Shared helpers for the parity tests (bit-exact, NaN-aware comparison; golden fixture loading)."""
import json
import os
import zlib

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return "%08x" % (zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF)


def bits_equal(a, b):
    """Bit-exact equality; for floats 'both NaN' counts as equal (x86 emits 0xFFC00000, CUDA 0x7FFFFFFF)."""
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        if not np.array_equal(na, nb):
            return False
        iv = np.uint32 if a.dtype == np.float32 else np.uint64
        return np.array_equal(a[~na].view(iv), b[~nb].view(iv))
    return np.array_equal(a, b)


def assert_bits_equal(a, b, msg=""):
    if not bits_equal(a, b):
        a = np.asarray(a); b = np.asarray(b)
        if a.shape == b.shape and a.dtype == b.dtype:
            with np.errstate(invalid="ignore"):
                bad = ~((a == b) | (np.isnan(a) & np.isnan(b))) if a.dtype.kind == "f" else (a != b)
            idx = np.argwhere(bad)
            first = tuple(idx[0]) if len(idx) else None
            raise AssertionError("%s: %d of %d elements differ; first at %s: got %r expected %r"
                                 % (msg, int(bad.sum()), a.size, first, a[first] if first else None, b[first] if first else None))
        raise AssertionError("%s: shape/dtype mismatch %s %s vs %s %s" % (msg, a.shape, a.dtype, b.shape, b.dtype))


def canon_crc(a):
    """CRC with NaNs canonicalised to the x86 default NaN (0xFFC00000) so GPU outputs hash like the reference's."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        a = a.copy(); v = a.view(np.uint32); v[np.isnan(a)] = 0xFFC00000
    return crc(a)


def load_golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


def load_png(name):
    import cv2
    img = cv2.imread(os.path.join(GOLDEN_DIR, name), cv2.IMREAD_UNCHANGED)
    assert img is not None, name
    return img


def make_image(rs, H, W, dtype=np.uint8, cn=1, kind="pw"):
    """Piecewise-constant + noise (exercises both weight outcomes) or pure noise."""
    shape = (H, W) if cn == 1 else (H, W, cn)
    if kind == "noise":
        a = rs.randint(0, 256, size=shape)
    elif kind == "const":
        a = np.full(shape, 77)
    else:
        blk = rs.randint(0, 256, size=((H + 7) // 8, (W + 7) // 8) + (() if cn == 1 else (cn,)))
        a = np.kron(blk, np.ones((8, 8) + (() if cn == 1 else (1,))))[:H, :W] + rs.randint(-6, 7, size=shape)
        a = np.clip(a, 0, 255)
    if dtype == np.uint16:
        a = a * 200 + rs.randint(0, 50, size=shape)
    elif dtype == np.int16:
        a = a * 200 - 20000
    elif dtype in (np.float32, np.float64):
        a = a * 3.7 + 0.25 * rs.randint(0, 4, size=shape)
    return np.ascontiguousarray(a.astype(dtype))
