"""No-GPU checks of the boundary: the C-ABI library loads, exports every symbol include/dmc_c.h declares, and
refuses to run without a CUDA device (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from depthmapcompression_b200 import capi
    return capi


def test_header_symbols_are_exported(capi):
    hdr = open(os.path.join(ROOT, "include", "dmc_c.h")).read()
    declared = sorted(set(re.findall(r"\b(dmc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(capi.lib, name), "libdmc_b200.so does not export %s" % name
    assert sorted(capi.EXPORTS) == declared


def test_struct_layouts_match_header(capi):
    assert C.sizeof(capi.DmcImage) == 40          # void*, 3 x int, size_t, int (+pad)
    assert C.sizeof(capi.DmcChainParams) == 56


def test_shard_frames_host_logic(capi):
    for n in (0, 1, 7, 999, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            sizes = []
            for r in range(world):
                b, c = capi.shard_frames(n, r, world)
                seen += list(range(b, b + c)); sizes.append(c)
            assert seen == list(range(n)) and max(sizes) - min(sizes) <= 1
    with pytest.raises(capi.DmcError):
        capi.shard_frames(10, 4, 4)


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert capi.lib.dmc_device_count() == 0
    h = C.c_void_p()
    assert capi.lib.dmc_create(0, C.byref(h)) == capi.DMC_ERR_CUDA
    assert b"no CPU fallback" in capi.lib.dmc_last_error(None)
    import depthmapcompression_b200 as m
    with pytest.raises(m.DmcError):
        m.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "depthmapcompression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle of record", ""), "%s mentions the oracle" % f
