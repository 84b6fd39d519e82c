"""Deterministic byte-level mutations of .jxl files shared by the CPU and GPU robustness tests."""
import os
import random


def mutants(paths, seed, count):
    """Yields (name, bytes): truncations, bit flips and random splices of the given files."""
    rng = random.Random(seed)
    blobs = [(os.path.basename(p), open(p, "rb").read()) for p in paths]
    for _ in range(count):
        name, data = rng.choice(blobs)
        d = bytearray(data)
        mode = rng.randrange(3)
        if mode == 0:
            d = d[:rng.randrange(1, len(d))]
        elif mode == 1:
            for _ in range(rng.randrange(1, 6)):
                d[rng.randrange(len(d))] ^= 1 << rng.randrange(8)
        else:
            p = rng.randrange(len(d))
            n = rng.randrange(1, 64)
            d[p:p + n] = bytes(rng.randrange(256) for _ in range(n))
        yield name, bytes(d)
