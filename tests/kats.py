"""Known-answer vectors restated from the reference's test/tests.rs.

Each entry cites the reference line range it was taken from. Shared by the
oracle tests (CPU) and the CUDA parity tests (GPU).
"""

# (name, input bytes, (variant, a, b, c), bad_header) -- test/tests.rs:345-466
DECODE_ERRORS = [
    ("err_empty", b"", ("Empty", 0, 0, 0), False),                                   # :345
    ("err_header_mismatch", b"\x05\x00a", ("HeaderMismatch", 5, 1, 0), False),        # :348-352
    ("err_varint1", b"\xFF", ("Header", 0, 0, 0), True),                              # :355
    ("err_varint2", b"\xff" * 10 + b"\x00", ("Header", 0, 0, 0), True),               # :358-363
    ("err_varint3", b"\x80\x80\x80\x80\x10", ("TooBig", 4294967296, 4294967295, 0), True),  # :366-371
    ("err_lit", b"\x02\x00hi", ("CopyRead", 1, 0, 0), False),                         # :376-380
    ("err_lit_big1", b"\x02\xechi", ("Literal", 60, 2, 2), False),                    # :382-386
    ("err_lit_big2a", b"\x02\xf0hi", ("Literal", 4, 2, 2), False),                    # :389-393
    ("err_lit_big2b", b"\x02\xf0hi\x00\x00\x00", ("Literal", 105, 4, 2), False),      # :396-404
    ("err_copy1", b"\x02\x00a\x01", ("CopyRead", 1, 0, 0), False),                    # :408-412
    ("err_copy2a", b"\x11\x00a\x3e", ("CopyRead", 2, 0, 0), False),                   # :415-419
    ("err_copy2b", b"\x11\x00a\x3e\x01", ("CopyRead", 2, 1, 0), False),               # :420-424
    ("err_copy3a", b"\x11\x00a\x3f", ("CopyRead", 4, 0, 0), False),                   # :426-430
    ("err_copy3b", b"\x11\x00a\x3f\x00", ("CopyRead", 4, 1, 0), False),               # :431-435
    ("err_copy3c", b"\x11\x00a\x3f\x00\x00", ("CopyRead", 4, 2, 0), False),           # :436-440
    ("err_copy3d", b"\x11\x00a\x3f\x00\x00\x00", ("CopyRead", 4, 3, 0), False),       # :441-445
    ("err_copy_offset_zero", b"\x11\x00a\x01\x00", ("Offset", 0, 1, 0), False),       # :448-452
    ("err_copy_offset_big", b"\x11\x00a\x01\xFF", ("Offset", 255, 1, 0), False),      # :455-459
    ("err_copy_len_big", b"\x05\x00a\x1d\x01", ("CopyWrite", 11, 4, 0), False),       # :462-466
]

# test/tests.rs:232-317 -- exact decoder outputs near the end of the buffer
COPY_CLOSE_TO_END = [
    (bytes([27, 0b000010_00, 1, 2, 3, 0b000_000_10, 3, 0, 0b010110_00] + list(range(4, 27))),
     bytes([1, 2, 3, 1] + list(range(4, 27)))),
    (bytes([28, 0b000010_00, 1, 2, 3, 0b000_000_10, 3, 0, 0b010111_00] + list(range(4, 28))),
     bytes([1, 2, 3, 1] + list(range(4, 28)))),
]

# test/tests.rs:469-504 -- past quickcheck witnesses
RANDOM = [
    bytes([
        0, 0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 4, 0, 0, 0, 5, 0, 0,
        1, 1, 0, 0, 1, 2, 0, 0, 2, 1, 0, 0, 2, 2, 0, 0, 0, 6, 0, 0, 3, 1, 0,
        0, 0, 7, 0, 0, 1, 3, 0, 0, 0, 8, 0, 0, 2, 3, 0, 0, 0, 9, 0, 0, 1, 4,
        0, 0, 1, 0, 0, 3, 0, 0, 1, 0, 1, 0, 0, 0, 10, 0, 0, 0, 0, 2, 4, 0, 0,
        2, 0, 0, 3, 0, 1, 0, 0, 1, 5, 0, 0, 6, 0, 0, 0, 0, 11, 0, 0, 1, 6, 0,
        0, 1, 7, 0, 0, 0, 12, 0, 0, 3, 2, 0, 0, 0, 13, 0, 0, 2, 5, 0, 0, 0, 3,
        3, 0, 0, 0, 1, 8, 0, 0, 1, 0, 1, 0, 0, 0, 4, 1, 0, 0, 0, 0, 14, 0, 0,
        0, 1, 9, 0, 0, 0, 1, 10, 0, 0, 0, 0, 1, 11, 0, 0, 0, 1, 0, 2, 0, 0, 0,
        1, 1, 1, 0, 0, 0, 0, 5, 1, 0, 0, 0, 1, 2, 1, 0, 0, 0, 0, 0, 2, 6, 0,
        0, 0, 0, 0, 1, 12, 0, 0, 0, 0, 0, 3, 4, 0, 0, 0, 0, 0, 7, 0, 0, 0, 0,
        0, 1, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
        0, 0, 0, 0,
    ]),
    bytes([10, 2, 14, 13, 0, 8, 2, 10, 2, 14, 13, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]),
    bytes([0, 0, 0, 4, 1, 4, 0, 0, 0, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]),
    bytes([
        0, 0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 4, 0, 0, 0, 5, 0, 0,
        1, 1, 0, 0, 1, 2, 0, 0, 1, 3, 0, 0, 1, 4, 0, 0, 2, 1, 0, 0, 0, 4, 0,
        1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    ]),
]

# sha256 of raw/frame outputs of the restated encoder.
# Only the Tom.Sawyer raw row is reference-pinned (golden file); the others are
# derived pins that detect any future divergence of the oracle.
CORPUS_PINS = {
    "html": (22843, "c7c94425c2b3516cf3d1c9824391b8453beb544f38dfdfa90eb8126103234b5a", 22872, "565d390c9eaccb758d5bf67314c9cd87cd580338ce6d6e41bf2895794f3848f0"),
    "urls.10K": (335492, "8d578b8cbf000930c09c7d02a6b68aa76731e436a20a036e66c40f63b10ba5ad", 335620, "41d51a0298ee8f9ea1b5060dfa891437003949773fe08d7b37292ee3b389d5c2"),
    "fireworks.jpeg": (123034, "4da5e82d77ebe3d77e4f827a294562df17b5dcf37dcdb30d516ee8544d3164a6", 123119, "c69c85e227e5bb547b773270cc587054a0431cb5fb448abe28234676033eeb07"),
    "paper-100k.pdf": (85304, "ad668e5050689de4486cca4851a67b81731ff77ae920dc78da2e5fc9ca36d7e5", 85327, "c7d486abab1af6d346dfb3bf976cc877908cc0aa8279ddd001350c4a8e67296f"),
    "html_x_4": (92234, "11e53110e963fa6dd4ef3d726cf2d897a7ab689c8edb90ccab3f462ef21872f3", 92318, "4de1903b0360e2b26443635e576b7c92a8fb600438e59c1ba996c5f149c0f514"),
    "alice29.txt": (88034, "d9b27949428e5678cd7a4f00baaba000612d180d9028d28a6ab3a5e308272869", 88074, "e96bd7aceb34fe1db3a696d6b3f92d507872abee05093f12ac063f1a2d00ed0f"),
    "asyoulik.txt": (77503, "4bf8701f8c369f13e679f52e938c8630d2a2920eba4003bfeeced8522d984aa9", 77532, "0a228e9e0a179ad0c7d3e66c6103e9256f721df710938c8990599d004087a379"),
    "lcet10.txt": (234661, "5db82d2428a5b5c747dae15c9b219fffc8093c82a9cc8263bec750d261569c09", 234745, "c72672e5e47458124b6a45b9e132d022ffc32943be0ba4a3b2249607a940e1be"),
    "plrabn12.txt": (319267, "30915f0a26ae2b882e7d8a6951dc3e844c8dd615b0a1a21c6dd69e8c8f958337", 319362, "77820b554a998f8813ad046ec5d966d904595a43647a3903b06e6fb39ef3ba00"),
    "geo.protodata": (23335, "84356d0f45f9cf8547834eabaa8d4ec569c3e71c505828ab3321ffbd35370d11", 23364, "875e5b92256702fa6cc873a0021aeb7bf8245784139045dccec7a45967a96ea9"),
    "kppkn.gtb": (69526, "b6513d28c84b3715f02a2697ddb3f6b56aab8f09f0b5950075762912ae5ae8d9", 69566, "5a9d2497f12d443a1f8a6935f779e4f6ef656a47e54f62b2dd2b8a64d8bd632f"),
    "Mark.Twain-Tom.Sawyer.txt": (9871, "7f1f5878f128aec6140fb135f7ed7e2c29575c7eb5c44431ffc65b22f7a19738", 9889, "10044aa964222a4c0e5c573b332056a6ca88120d85b33842461304103c83f883"),
}


def small_copy_inputs():
    """test/tests.rs:208-216"""
    return [b"aaaa" + b"b" * i + b"aaaabbbb" for i in range(32)]


def small_regular_inputs():
    """test/tests.rs:218-229"""
    out, i = [], 1
    while i < 20000:
        out.append(bytes((j % 10) + ord("a") for j in range(i)))
        i += 23
    return out


def adversarial_blocks(seed=2024):
    """Synthetic blocks aimed at the rare paths of the parsers: zero runs (one giant overlapping copy, copy
    splitting at 64/60), incompressible data (scan stride growth), tiny alphabets (dense slot clashes inside a
    32-position window), periodic data around the copy-1/copy-2 offset and length limits, sizes around the
    table-size and window boundaries."""
    import random
    rng = random.Random(seed)

    def rnd(n, a=256):
        return bytes(rng.randrange(a) for _ in range(n))

    units = [bytes(65536), rnd(65536), rnd(65536, 4)]
    for per in (1, 2, 3, 5, 7, 12, 31, 32, 33, 63, 64, 65, 67, 68, 69, 127, 128, 129, 2047, 2048, 2049, 4095, 4099):
        pat = rnd(per)
        units.append((pat * (65536 // per + 1))[:65536])
    units.append(rnd(30000) + bytes(5536) + rnd(30000))
    units.append((rnd(200) + bytes(100)) * 218 + rnd(136))
    blk = rnd(1000)
    units.append(b"".join(blk[:rng.randrange(4, 80)] + rnd(rng.randrange(0, 20)) for _ in range(1500))[:65536])
    for n in (17, 18, 31, 32, 33, 47, 48, 63, 64, 65, 100, 255, 256, 257, 1023, 1025, 8191, 8192, 8193, 16384, 32768, 65535):
        units.append(rnd(n, 8))
    return units
