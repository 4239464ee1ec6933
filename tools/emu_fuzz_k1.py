#!/usr/bin/env python
"""TEST TOOLING: fuzz the K1 kernel body under the CPU warp emulator against the oracle.
usage: emu_fuzz_k1.py <seed> <nblocks> [multi|hybrid]   (SBEMU_ORDER=reverse|shuffle perturbs the lane schedule)"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import emu_helpers as emu          # noqa: E402
from oracle import oracle          # noqa: E402


def gen(rng, corp):
    kind = rng.randrange(7)
    n = rng.choice([rng.randrange(17, 400), rng.randrange(400, 5000), rng.randrange(5000, 65537), 65536])
    if kind == 0:
        return bytes(rng.randrange(256) for _ in range(min(n, 3000)))
    if kind == 1:
        a = rng.randrange(2, 6)
        return bytes(rng.randrange(a) for _ in range(n))
    if kind == 2:
        pat = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 70)))
        b = bytearray((pat * (n // len(pat) + 1))[:n])
        for _ in range(rng.randrange(0, 40)):
            b[rng.randrange(n)] = rng.randrange(256)
        return bytes(b)
    if kind == 3:
        c = rng.choice(corp)
        o = rng.randrange(0, max(1, len(c) - n))
        return c[o:o + n]
    if kind == 4:
        c = rng.choice(corp)
        o = rng.randrange(0, max(1, len(c) - n))
        b = bytearray(c[o:o + n])
        for _ in range(rng.randrange(1, 200)):
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        return bytes(b)
    if kind == 5:
        words = [bytes(rng.randrange(97, 123) for _ in range(rng.randrange(2, 9))) for _ in range(rng.randrange(3, 60))]
        out = bytearray()
        while len(out) < n:
            out += rng.choice(words) + b" "
        return bytes(out[:n])
    runs = bytearray()
    while len(runs) < n:
        runs += bytes([rng.randrange(256)]) * rng.randrange(1, 300)
    return bytes(runs[:n])


def main():
    seed, nblocks = int(sys.argv[1]), int(sys.argv[2])
    mode = sys.argv[3] if len(sys.argv) > 3 else "multi"
    rng = random.Random(seed)
    gold = os.path.join(ROOT, "tests", "golden", "data")
    corp = [open(os.path.join(gold, f), "rb").read() for f in sorted(os.listdir(gold))]
    corp = [c for c in corp if len(c) > 70000]
    units = [gen(rng, corp) for _ in range(nblocks)]
    got = emu.compress_units(units, hybrid=(mode == "hybrid"), grid=2)
    bad = [i for i, (g, u) in enumerate(zip(got, units)) if g != oracle.compress(u)]
    print("seed", seed, "mode", mode, "blocks", nblocks, "mismatches", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
