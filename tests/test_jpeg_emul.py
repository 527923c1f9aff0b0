"""CPU checks of the JPEG row that need no GPU:

  * the self-synchronising parallel Huffman decode of csrc/dmc_jpeg.cu, emulated lane by lane on the CPU from the very same
    host/device primitives (tests/cpp/jpeg_emul.cpp over csrc/dmc_jpeg_core.h), must reproduce cv2.imdecode bit for bit for
    any number of lanes -- including lane counts small and large enough to force several synchronisation rounds;
  * the host-side marker parser must refuse malformed, truncated and hostile streams instead of reading out of bounds
    (over-subscribed DHT, segment length < 2, truncated DQT / DHT / DRI / SOS).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from _util import assert_bits_equal, make_image

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "jpeg_emul.cpp")
LIB = os.path.join(HERE, "cpp", "libjpeg_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "depthmapcompression_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    deps = [SRC, os.path.join(CSRC, "dmc_jpeg_core.h"), os.path.join(CSRC, "dmc_jpeg_parse.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fwrapv", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.jpeg_emul_decode.restype = C.c_int
    lib.jpeg_emul_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    lib.jpeg_emul_probe.restype = C.c_int
    lib.jpeg_emul_probe.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    return lib


def enc(img, q, *extra):
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q] + list(extra))
    assert ok
    return np.ascontiguousarray(buf).ravel()


def decode(lib, stream, H, W, lanes):
    out = np.full((H, W), 0xAB, np.uint8); rounds = C.c_int(); err = C.create_string_buffer(256)
    rc = lib.jpeg_emul_decode(stream.ctypes.data, stream.size, H, W, out.ctypes.data, lanes, C.byref(rounds), err, 256)
    assert rc == 0, err.value
    return out, rounds.value


def probe(lib, data):
    a = np.frombuffer(bytes(data), np.uint8).copy() if len(data) else np.zeros(0, np.uint8)
    buf = np.concatenate([a, np.zeros(0, np.uint8)])
    err = C.create_string_buffer(256); r, c = C.c_int(), C.c_int()
    rc = lib.jpeg_emul_probe(buf.ctypes.data if buf.size else None, buf.size, C.byref(r), C.byref(c), err, 256)
    return rc, err.value.decode(), r.value, c.value


@pytest.mark.parametrize("shape", [(480, 640), (131, 150), (8, 8), (1, 1), (7, 9), (64, 641)])
def test_parallel_decode_emulation_matches_libjpeg(emul, shape):
    from oracle.oracle_py import synth_disp
    rs = np.random.RandomState(31)
    H, W = shape
    worst = 0
    for kind in ("synth", "pw", "noise"):
        img = synth_disp(H, W, 3) if kind == "synth" and H > 16 else make_image(rs, H, W, kind="noise" if kind == "noise" else "pw")
        for q in (5, 50, 80, 95, 100):
            for s in (enc(img, q), enc(img, q, cv2.IMWRITE_JPEG_OPTIMIZE, 1)):
                want = cv2.imdecode(s, 0)
                for lanes in (1, 7, 64, 768, 4096):
                    got, rounds = decode(emul, s, H, W, lanes)
                    assert_bits_equal(got, want, "%dx%d %s q%d lanes %d" % (H, W, kind, q, lanes))
                    worst = max(worst, rounds)
    assert worst >= 1 or H * W <= 64        # the synchronisation rounds were actually exercised


def test_parallel_decode_emulation_1080p(emul):
    from oracle.oracle_py import synth_disp, degrade_blocks
    img = degrade_blocks(synth_disp(1080, 1920, 11), 11)
    for q in (50, 80):
        s = enc(img, q)
        got, rounds = decode(emul, s, 1080, 1920, 768)
        assert_bits_equal(got, cv2.imdecode(s, 0), "1080p q%d" % q)
        assert rounds <= 16, "synchronisation took %d rounds" % rounds


def test_parser_refuses_malformed_streams(emul):
    img = make_image(np.random.RandomState(5), 24, 40)
    good = bytes(enc(img, 75))
    rc, why, r, c = probe(emul, good)
    assert rc == 0 and (r, c) == (24, 40), why
    # every truncation of a valid stream is refused or accepted, never crashes (run under -fsanitize=undefined; bounds are
    # additionally covered by the explicit cases below)
    for n in range(0, len(good)):
        probe(emul, good[:n])
    assert probe(emul, b"")[0] != 0 and probe(emul, b"not a jpeg at all")[0] != 0
    # segment length below 2 (used to wrap around to 2^64)
    assert probe(emul, b"\xff\xd8\xff\xdb\x00\x00")[0] != 0
    assert probe(emul, b"\xff\xd8\xff\xdb\x00\x01" + b"\x00" * 8)[0] != 0
    # DQT with a one-byte payload at the end of the stream
    assert "DQT" in probe(emul, b"\xff\xd8\xff\xdb\x00\x03\x00")[1]
    # DHT shorter than its 17-byte header, and with fewer values than BITS announces
    assert "DHT" in probe(emul, b"\xff\xd8\xff\xc4\x00\x05\x00\x01\x02")[1]
    assert "DHT" in probe(emul, b"\xff\xd8\xff\xc4\x00\x14\x00" + bytes([0, 4] + [0] * 14) + b"\x01")[1]
    # over-subscribed code lengths: three codes of length 1 / 200 codes of length 2 (used to overflow look[])
    i = good.index(b"\xff\xc4")
    for l, cnt in ((1, 3), (2, 200)):
        bits = [0] * 16; bits[l - 1] = cnt
        seg = bytes([0x00] + bits) + bytes([0] * cnt)
        bad = good[:i] + b"\xff\xc4" + (len(seg) + 2).to_bytes(2, "big") + seg + good[i:]
        rc, why, _, _ = probe(emul, bad)
        assert rc != 0 and "prefix code" in why, why
    # truncated DRI and SOS
    assert "DRI" in probe(emul, good[:2] + b"\xff\xdd\x00\x02" + good[2:])[1]
    j = good.index(b"\xff\xda")
    assert probe(emul, good[:j] + b"\xff\xda\x00\x03\x01")[0] != 0
    # progressive and colour streams
    assert "progressive" in probe(emul, bytes(enc(img, 75, cv2.IMWRITE_JPEG_PROGRESSIVE, 1)))[1]
    assert "grayscale" in probe(emul, bytes(enc(np.dstack([img, img, img]), 75)))[1]


def test_garbage_scans_terminate(emul):
    """A valid header followed by a random scan: the decode must terminate with SOME deterministic picture (libjpeg would
    warn about corrupt data); this guards the loop bounds of the parallel decoder."""
    rs = np.random.RandomState(9)
    img = make_image(rs, 64, 96, kind="noise")
    good = enc(img, 90)
    j = bytes(good).index(b"\xff\xda")
    sos_len = int.from_bytes(bytes(good[j + 2:j + 4]), "big")
    head = bytes(good[:j + 2 + sos_len])
    for trial in range(20):
        n = int(rs.randint(0, 3000))
        body = bytes(rs.randint(0, 256, n).astype(np.uint8)).replace(b"\xff", b"\xfe")
        s = np.frombuffer(head + body + b"\xff\xd9", np.uint8).copy()
        a, _ = decode(emul, s, 64, 96, 768)
        b, _ = decode(emul, s, 64, 96, 5)
        assert_bits_equal(a, b, "garbage scan %d: result depends on the number of lanes" % trial)
