"""Randomised differential check of the host front-end: for random synthetic frames (sizes around the LF-group and
HF-group boundaries, all transform profiles, both LF codings) the parse digest must not depend on the decode strategy —
LF groups in lockstep pairs, one at a time, through the generic all-properties loop, or on several threads.
Usage: python tools/frontend_sweep.py [cases=60] [seed=7]"""
import ctypes as C
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth  # noqa: E402
from tests import oracle_binding as ob  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
lib = ob.load()
lib.jxo_t_parse_digest.restype = C.c_uint64
lib.jxo_t_parse_digest.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
for it in range(cases):
    w = rng.choice([8, 64, 250, 1000, 2040, 2056, 2600, 4100, 4200, 5000])
    h = rng.choice([8, 100, 520, 1100, 2049, 2300])
    prof, lf, dist, epf = rng.randrange(3), rng.randrange(2), rng.choice([0.3, 0.7, 1.5, 3.0]), rng.randrange(4)
    data = synth.encode_synthetic(w, h, 1000 + it, dist, epf, rng.randrange(2), prof, lf)
    digests = []
    try:
        for pair, generic, threads in ((1, 0, 1), (0, 0, 1), (0, 1, 1), (1, 0, 3)):
            lib.jxo_t_pair_lf_groups(pair)
            lib.jxo_t_force_generic_walk(generic)
            digests.append(lib.jxo_t_parse_digest(data, len(data), threads))
    finally:
        lib.jxo_t_pair_lf_groups(1)
        lib.jxo_t_force_generic_walk(0)
    assert digests[0] != 0 and len(set(digests)) == 1, (w, h, prof, lf, dist, digests)
print(cases, "random frames: paired == single == generic == 3 threads")
