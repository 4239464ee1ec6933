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
    assert emu.compress_units(blocks, grid=2) == want                      # 7 shared-memory-table chains per CTA
    assert emu.compress_units(blocks[:2], hybrid=True) == want[:2]         # ... plus chains with tables in global memory
    for (st, out, guard), b in zip(emu.decompress_units(want, [len(b) for b in blocks], grid=2, block=64), blocks):
        assert st[0] == "Ok" and out == b and guard == b"\xee" * 16


def test_k1_small_inputs(oracle):
    units = [b"", b"\x00"] + RANDOM + small_copy_inputs() + small_regular_inputs()[::9]
    assert emu.compress_units(units) == [oracle.compress(u) for u in units]
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


@pytest.mark.parametrize("hybrid", [False, True])
def test_k1_adversarial_blocks_both_layouts(oracle, hybrid):
    """Shared-memory-table chains and chains with the table in global memory on the rare-path blocks."""
    units = adversarial_blocks()
    got = emu.compress_units(units, hybrid=hybrid, grid=2)
    assert [i for i, (c, u) in enumerate(zip(got, units)) if c != oracle.compress(u)] == []


@pytest.mark.parametrize("chains", [1, 2, 8])
def test_k1_fewer_chains_per_cta(oracle, chains):
    """Small batches are launched with fewer parser/emitter pairs per CTA (one per SM first)."""
    units = adversarial_blocks()[:10] + [corpus("alice29.txt")[:65536], b"", b"xy"]
    assert emu.compress_units(units, hybrid=True, chains=chains, grid=2) == [oracle.compress(u) for u in units]


def test_k1_fused_chunk_checksum(oracle):
    """Frame encode: the emitter warp computes the chunk's masked CRC-32C beside the compress call."""
    units = [corpus("alice29.txt")[:65536], corpus("html")[100:40000], b"", b"a", b"abc" * 11, RANDOM[0], corpus("fireworks.jpeg")[3:65539]]
    got, crc = emu.compress_units(units, hybrid=True, crcs=True)
    assert got == [oracle.compress(u) for u in units]
    assert crc[:2] + crc[3:] == [oracle.crc32c_masked(u) for u in units[:2] + units[3:]]


def test_k1_unit_limits_reported():
    """A unit above 64KB or a slot below max_compress_len is skipped with out_len 0 and the reference's error."""
    units = [b"a" * 100, b"b" * 70000, b"hello hello hello hello"]
    got, st = emu.compress_units(units, statuses=True)
    assert got[1] == b"" and st[1] == ("TooBig", 70000, 65536)
    assert st[0][0] == "Ok" and st[2][0] == "Ok" and got[0] and got[2]
    got, st = emu.compress_units(units[:1], out_cap=100, statuses=True)
    assert got[0] == b"" and st[0] == ("BufferTooSmall", 100, 32 + 100 + 16)


def test_k2_adversarial_blocks(oracle):
    units = adversarial_blocks()
    res = emu.decompress_units([oracle.compress(u) for u in units], [len(u) for u in units], grid=2, block=128)
    assert [i for i, (r, u) in enumerate(zip(res, units)) if r[0][0] != "Ok" or r[1] != u or r[2] != b"\xee" * 16] == []


def _frame_err(e):
    """Oracle error tuple -> the (name, a, b, c) shape of the emulated result record."""
    name = e[0]
    if name == "StreamHeaderMismatch":
        return (name, int.from_bytes(e[1], "little") if isinstance(e[1], (bytes, bytearray)) else e[1], 0, 0)
    return tuple(e)


def test_k4_frame_encode_matches_write_frame_encoder(oracle):
    """fill lens -> K1 with the chunk CRC in the emitter -> two-level scan -> gather == write::FrameEncoder bytes."""
    for name, cut in (("alice29.txt", 150000), ("fireworks.jpeg", 70000), ("html", 65536), ("geo.protodata", 1), ("paper-100k.pdf", 66000)):
        data = corpus(name)[:cut]
        stream, offs, res = emu.frame_encode(data)
        assert res.status.code == 0 and res.nchunks == (len(data) + 65535) // 65536
        assert stream == oracle.frame_encode(data), name
        assert offs[0] == 10 and offs[-1] == len(stream)
    stream, offs, res = emu.frame_encode(corpus("lcet10.txt")[:140000], ident=False)
    assert stream == oracle.frame_encode(corpus("lcet10.txt")[:140000])[10:] and offs[0] == 0


def test_k5_frame_decode_index_walk_and_errors(oracle):
    """K5 under the emulator: parallel parse over the encoder's index, the serial walk, identifier-less fragments, and
    the reference's first error in stream order with the bytes before it."""
    from oracle.oracle import OracleError
    data = corpus("alice29.txt")[:150000]
    good = oracle.frame_encode(data)
    stream, offs, _ = emu.frame_encode(data)
    assert stream == good
    for kw in (dict(index=offs), dict(index=None), dict(index=[10, 50, len(good)])):          # last: wrong index -> serial walk
        st, out, res = emu.frame_decode(good, len(data), **kw)
        assert st[0] == "Ok" and out == data and res.nchunks == 3, kw
    frag = good[10:]
    st, out, _ = emu.frame_decode(frag, len(data), index=[o - 10 for o in offs], fragment=True)
    assert st[0] == "Ok" and out == data
    assert emu.frame_decode(frag, len(data))[0][0] == "StreamHeader"
    st, out, _ = emu.frame_decode(good, 1000)
    assert st[:3] == ("BufferTooSmall", 1000, len(data)) and out == b""
    st, out, _ = emu.frame_decode(good, len(data), max_chunks=2)                                # chunk table too small
    assert st[0] == "Invalid" and st[2] == 1
    ident = b"\xff\x06\x00\x00sNaPpY"
    flip = bytearray(good); flip[len(good) // 2] ^= 0x10
    crc = bytearray(good); crc[14] ^= 1
    streams = [bytes(flip), bytes(crc), good[:-7], good + b"\x00\x07", ident + b"\x02\x00\x00\x00", b"123",
               ident + b"\x80\x03\x00\x00xyz" + b"\xfe\x02\x00\x00\x00\x00" + ident + good[10:],
               ident + b"\x00\x05\x00\x00\x00\x00\x00\x00\x80", ident + b"\x01\x03\x00\x00abc", b"\xff\x05\x00\x00sNaPp",
               b"\xff\x06\x00\x00sNaPpZ", ident + b"\x00\xff\xff\xff", ident + b"\x00\x04\x00\x00\x00\x00\x00\x00"]
    for s in streams:
        try:
            want = (("Ok", 0, 0, 0), oracle.frame_decode(s))
        except OracleError as e:
            want = (_frame_err(e.err), None)
        st, out, _ = emu.frame_decode(s, 200000)
        assert st == want[0], (s[:20], st, want[0])
        if want[1] is not None:
            assert out == want[1]
        else:
            assert data.startswith(out)
