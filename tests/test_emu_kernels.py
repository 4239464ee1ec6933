"""Kernel LOGIC checks on CPU: the CUDA kernel bodies (rust-snappy_b200/csrc/*.cuh)
compiled by g++ against the fiber warp emulator in tests/emu, compared with the
oracle. This is test tooling for GPU-less development -- the product library is
never built this way."""
import random

import pytest

import emu_helpers as emu
from conftest import corpus
from kats import (COPY_CLOSE_TO_END, DECODE_ERRORS, RANDOM, adversarial_blocks, small_copy_inputs,
                  small_regular_inputs)


def blocks_of(data):
    return [data[i:i + 65536] for i in range(0, len(data), 65536)]


@pytest.mark.parametrize("name", ["html", "urls.10K", "fireworks.jpeg", "paper-100k.pdf", "alice29.txt",
                                  "geo.protodata", "kppkn.gtb", "Mark.Twain-Tom.Sawyer.txt"])
def test_k1_k2_corpus_blocks(oracle, name):
    blocks = blocks_of(corpus(name))[:3]
    want = [oracle.compress(b) for b in blocks]
    assert emu.compress_units(blocks, grid=2, multi=True) == want          # product default: 7 pairs per CTA
    assert emu.compress_units(blocks[:1], global_window=True) == want[:1]  # one-pair global-window variant
    for (st, out, guard), b in zip(emu.decompress_units(want, [len(b) for b in blocks], grid=2, block=64), blocks):
        assert st[0] == "Ok" and out == b and guard == b"\xee" * 16


def test_k1_small_inputs(oracle):
    units = [b"", b"\x00"] + RANDOM + small_copy_inputs() + small_regular_inputs()[::9]
    assert emu.compress_units(units, multi=True) == [oracle.compress(u) for u in units]
    assert emu.compress_units(units, hybrid=True) == [oracle.compress(u) for u in units]


def test_k2_error_kats():
    kats = [k for k in DECODE_ERRORS]
    res = emu.decompress_units([k[1] for k in kats], [1024 if k[3] else 64 for k in kats])
    for k, (st, _, guard) in zip(kats, res):
        if k[3] or k[0] == "err_empty":
            assert st == k[2]
        assert guard == b"\xee" * 16
    # non-header KATs decode into exactly decompress_len bytes
    from oracle import oracle as o
    body = [k for k in kats if not k[3] and k[1]]
    res = emu.decompress_units([k[1] for k in body], [o.decompress_len(k[1]) for k in body])
    for k, (st, _, guard) in zip(body, res):
        assert st == k[2], k[0]
        assert guard == b"\xee" * 16
    for stream, want in COPY_CLOSE_TO_END:
        (st, out, guard), = emu.decompress_units([stream], [len(want)])
        assert st[0] == "Ok" and out == want and guard == b"\xee" * 16


def test_k2_fuzz_against_oracle(oracle):
    from oracle.oracle import OracleError
    rng = random.Random(11)
    base = oracle.compress(corpus("alice29.txt")[:6000])
    streams, caps = [], []
    for _ in range(150):
        s = bytearray(base)
        for _ in range(rng.randrange(1, 4)):
            s[rng.randrange(len(s))] ^= 1 << rng.randrange(8)
        if rng.random() < 0.3:
            s = s[:rng.randrange(1, len(s))]
        s = bytes(s)
        try:
            cap = min(oracle.decompress_len(s), 1 << 18)
        except OracleError:
            cap = 512
        streams.append(s); caps.append(cap)
    for s, cap, (st, out, guard) in zip(streams, caps, emu.decompress_units(streams, caps, block=128)):
        try:
            want = (("Ok", 0, 0, 0), oracle.decompress(s, cap=cap))
        except OracleError as e:
            want = (e.err, b"")
        assert (st, out if st[0] == "Ok" else b"") == want
        assert guard == b"\xee" * 16


@pytest.mark.parametrize("mode", ["multi", "hybrid", "gw", "sm"])
def test_k1_adversarial_blocks_all_layouts(oracle, mode):
    """Every K1 layout (7 pairs per CTA / 7 + 4 pairs with L2-resident tables / global window /
    shared-memory window) on the rare-path blocks."""
    units = adversarial_blocks()
    got = emu.compress_units(units, multi=(mode == "multi"), hybrid=(mode == "hybrid"), global_window=(mode == "gw"),
                             parsers=1, grid=2)
    assert [i for i, (c, u) in enumerate(zip(got, units)) if c != oracle.compress(u)] == []


@pytest.mark.parametrize("chains", [1, 2, 8])
def test_k1_fewer_chains_per_cta(oracle, chains):
    """Small batches are launched with fewer parser/emitter pairs per CTA (one per SM first)."""
    units = adversarial_blocks()[:10] + [corpus("alice29.txt")[:65536], b"", b"xy"]
    assert emu.compress_units(units, hybrid=True, chains=chains, grid=2) == [oracle.compress(u) for u in units]


def test_k1_mbarrier_wakeup(oracle):
    """Experimental emitter wake-up through an mbarrier (-DK1_MBAR build) instead of sleep-polling."""
    units = adversarial_blocks()[:16] + [corpus("alice29.txt")[:65536], corpus("kppkn.gtb")[:65536], b"", b"ab"]
    assert emu.compress_units(units, hybrid=True, mbar=True) == [oracle.compress(u) for u in units]


def test_k1_speculative_slot_reads(oracle):
    """Experimental -DK1_GT_SPEC path: L2-table chains read the next window's slots before the commit and revalidate."""
    units = adversarial_blocks() + [corpus("alice29.txt")[:65536], corpus("html")[:65536], corpus("urls.10K")[:65536]]
    assert emu.compress_units(units, hybrid=True, gt_spec=True) == [oracle.compress(u) for u in units]


@pytest.mark.parametrize("hybrid,aligned,gt", [(False, False, False), (True, False, False), (False, True, False), (True, False, True)])
def test_k1_wide_step(oracle, hybrid, aligned, gt):
    """Experimental -DK1_W64 path: 64 positions per parser step on the shared-memory-table chains
    (gt: also on the L2-table chains, with the match.any commit of -DK1_W64_GT)."""
    units = adversarial_blocks() + [b"", b"a", RANDOM[0]] + small_copy_inputs()[::7]
    for name in ("alice29.txt", "html", "urls.10K", "kppkn.gtb", "geo.protodata", "fireworks.jpeg"):
        units += blocks_of(corpus(name))[:2]
    got = emu.compress_units(units, multi=not hybrid, hybrid=hybrid, w64=True, w64_aligned=aligned, w64_gt=gt, grid=2)
    assert [i for i, (g, u) in enumerate(zip(got, units)) if g != oracle.compress(u)] == []


def test_k1_unaligned_windows(oracle):
    """Experimental -DK1_UNALIGNED path: 32-position windows that start where the parse stands."""
    units = adversarial_blocks() + [b"", b"a"] + small_copy_inputs()[::7]
    for name in ("alice29.txt", "html", "urls.10K", "kppkn.gtb", "fireworks.jpeg"):
        units += blocks_of(corpus(name))[:2]
    got = emu.compress_units(units, hybrid=True, unaligned=True, grid=2)
    assert [i for i, (g, u) in enumerate(zip(got, units)) if g != oracle.compress(u)] == []
    got = emu.compress_units(units[:40], hybrid=True, unaligned=True, w64=True, gt_spec=True)   # all experiments together
    assert [i for i, (g, u) in enumerate(zip(got, units[:40])) if g != oracle.compress(u)] == []


def test_k1_pipelined_parsers(oracle):
    """The NP=2 token-passing variant (kept behind SNAPB200_K1_NP) stays bit-exact."""
    units = adversarial_blocks()[:12] + [corpus("alice29.txt")[:65536], corpus("html")[:65536]]
    assert emu.compress_units(units, parsers=2) == [oracle.compress(u) for u in units]


def test_k2_adversarial_blocks(oracle):
    units = adversarial_blocks()
    res = emu.decompress_units([oracle.compress(u) for u in units], [len(u) for u in units], grid=2, block=128)
    assert [i for i, (r, u) in enumerate(zip(res, units)) if r[0][0] != "Ok" or r[1] != u or r[2] != b"\xee" * 16] == []
