"""Pins the CPU oracle against the reference's own golden vector, decoder KATs
and round-trip set (reference test/tests.rs). CPU only."""
import hashlib
import random

import pytest

from conftest import CORPUS, corpus
from kats import (COPY_CLOSE_TO_END, CORPUS_PINS, DECODE_ERRORS, RANDOM,
                  small_copy_inputs, small_regular_inputs)


def test_golden_encoder_bytes(oracle):
    # test/tests.rs:199-205 -- the one test pinning encoder output bytes
    gold = corpus("Mark.Twain-Tom.Sawyer.txt.rawsnappy")
    assert len(gold) == 9871
    assert oracle.compress(oracle.decompress(gold)) == gold
    assert oracle.compress(corpus("Mark.Twain-Tom.Sawyer.txt")) == gold


@pytest.mark.parametrize("name,data,want,bad_header", DECODE_ERRORS, ids=[k[0] for k in DECODE_ERRORS])
def test_decode_error_kats(oracle, name, data, want, bad_header):
    from oracle.oracle import OracleError
    if bad_header:
        with pytest.raises(OracleError) as ei:
            oracle.decompress_len(data)
        assert ei.value.err == want
        cap = 1024
    else:
        cap = oracle.decompress_len(data)
    with pytest.raises(OracleError) as ei:
        oracle.decompress(data, cap=cap)
    assert ei.value.err == want


@pytest.mark.parametrize("stream,want", COPY_CLOSE_TO_END)
def test_copy_close_to_end(oracle, stream, want):
    assert oracle.decompress(stream) == want


@pytest.mark.parametrize("name", CORPUS)
def test_corpus_roundtrip_and_pins(oracle, name):
    import pyarrow as pa
    data = corpus(name)
    raw = oracle.compress(data)
    frame = oracle.frame_encode(data)
    n, h, fn, fh = CORPUS_PINS[name]
    assert (len(raw), hashlib.sha256(raw).hexdigest()) == (n, h)
    assert (len(frame), hashlib.sha256(frame).hexdigest()) == (fn, fh)
    assert oracle.decompress(raw) == data
    assert oracle.frame_decode(frame) == data
    # independent decoder (Google C++ snappy inside pyarrow), like cpp_decompresses_rust
    assert pa.Codec("snappy").decompress(raw, decompressed_size=len(data)).to_pybytes() == data


def test_simple_and_random_roundtrips(oracle):
    for d in [b"", b"\x00"] + RANDOM + small_copy_inputs() + small_regular_inputs():
        assert oracle.decompress(oracle.compress(d)) == d
        if d:
            assert oracle.frame_decode(oracle.frame_encode(d)) == d
    assert oracle.compress(b"") == b"\x00"
    assert oracle.frame_encode(b"") == b""          # src/write.rs:155-157


def test_property_roundtrip_and_cross_decode(oracle):
    import pyarrow as pa
    rng = random.Random(1234)
    codec = pa.Codec("snappy")
    for _ in range(300):
        n = rng.randrange(0, 10000)
        alpha = rng.choice([2, 4, 16, 256])
        d = bytes(rng.randrange(alpha) for _ in range(n))
        c = oracle.compress(d)
        assert oracle.decompress(c) == d
        if n:
            assert codec.decompress(c, decompressed_size=n).to_pybytes() == d
            # rust_decompresses_cpp: streams from another encoder must decode
            assert oracle.decompress(codec.compress(d).to_pybytes()) == d


def test_crc32c(oracle):
    assert oracle.crc32c(b"123456789") == 0xE3069283
    assert oracle.crc32c_masked(b"") == 0xA282EAD8
    rng = random.Random(7)
    lib = oracle.lib()
    for n in [0, 1, 7, 8, 9, 15, 16, 17, 63, 64, 65, 1000, 65536]:
        d = bytes(rng.randrange(256) for _ in range(n))
        assert lib.orc_crc32c(d, n) == lib.orc_crc32c_bitwise(d, n)


def test_frame_decoder_errors(oracle):
    from oracle.oracle import OracleError

    def err(stream):
        with pytest.raises(OracleError) as ei:
            oracle.frame_decode(stream)
        return ei.value.err

    ident = b"\xff\x06\x00\x00sNaPpY"
    assert err(b"123")[0] == "UnexpectedEof"                      # test/tests.rs:536-545
    assert err(b"\x00\x04\x00\x00abcd") == ("StreamHeader", 0, 0, 0)       # src/read.rs:122-127
    assert err(ident + b"\x02\x00\x00\x00") == ("UnsupportedChunkType", 2, 0, 0)   # :138-142
    assert err(b"\xff\x05\x00\x00sNaPp") == ("UnsupportedChunkLength", 5, 1, 0)    # :160-165
    assert err(b"\xff\x06\x00\x00sNaPpZ")[0] == "StreamHeaderMismatch"              # :167-171
    assert err(ident + b"\x00\xff\xff\xff") == ("UnsupportedChunkLength", 0xFFFFFF, 0, 0)  # :129-135
    assert err(ident + b"\x01\x03\x00\x00abc") == ("UnsupportedChunkLength", 3, 0, 0)      # :174-179
    good = oracle.frame_encode(b"hello world, hello world, hello world")
    bad = bytearray(good); bad[14] ^= 1
    assert err(bytes(bad))[0] == "Checksum"                                              # :189-196
    # skippable + padding chunks are skipped, repeated stream identifiers accepted (:143-172)
    s = ident + b"\x80\x03\x00\x00xyz" + b"\xfe\x02\x00\x00\x00\x00" + ident + good[10:]
    assert oracle.frame_decode(s) == b"hello world, hello world, hello world"


def test_baddata_rejected(oracle):
    # data/baddata*.snappy are referenced by no reference test: "must reject" only
    from oracle.oracle import OracleError
    for i in (1, 2, 3):
        with pytest.raises(OracleError) as ei:
            oracle.decompress(corpus("baddata%d.snappy" % i))
        assert ei.value.err[0] == "Offset"
