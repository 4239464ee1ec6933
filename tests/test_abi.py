"""CPU checks of the C ABI library: it loads, exports every symbol the header
declares, keeps the pure-arithmetic entry points exact, and refuses to compute
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def snap():
    import __graft_entry__ as g
    g.build_cuda()
    return g.load_package()


def test_header_symbols_exported(snap):
    hdr = open(os.path.join(ROOT, "include", "snapb200.h")).read()
    declared = set(re.findall(r"\b((?:sb|snappy)_[a-z0-9_]+)\s*\(", hdr))
    lib = snap._lib.lib()
    assert declared == set(snap._lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_pure_arithmetic_entry_points(snap, oracle):
    for n in [0, 1, 16, 17, 65535, 65536, 65537, 1 << 20, (1 << 32) - 1, 1 << 32, 3681400534]:
        assert snap.raw.max_compress_len(n) == oracle.max_compress_len(n)
    assert snap.raw.max_compress_len(65536) == 76490
    assert snap.raw.decompress_len(b"") == 0
    assert snap.raw.decompress_len(b"\x80\x80\x04") == 65536
    for bad, want in [(b"\xff", ("Header", 0, 0, 0)), (b"\x80\x80\x80\x80\x10", ("TooBig", 4294967296, 4294967295, 0))]:
        with pytest.raises(snap.Error) as ei:
            snap.raw.decompress_len(bad)
        assert ei.value.as_tuple() == want


def test_no_cpu_fallback(snap):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the refusal path is for GPU-less hosts")
    with pytest.raises(snap.NoDevice):
        snap.raw.Encoder().compress_vec(b"hello hello hello hello")
    with pytest.raises(snap.NoDevice):
        snap.raw.Decoder().decompress_vec(b"\x05\x10hello")


def test_product_does_not_touch_oracle():
    """Nothing under rust-snappy_b200/ or include/ may reference oracle/ or the emulator."""
    bad = []
    for base in ("rust-snappy_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"import\s+oracle|from\s+oracle|oracle/|liboracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
