"""Ad-hoc randomised parity sweep on the GPU: 400 random (class, size, seed, acceleration) inputs through the single-block
calls of K1 / K2 (and K3 on every eighth) against the oracle.  Not part of the test suite; a last look at a shipped build."""
import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import zig_lz4_b200 as z
from zig_lz4_b200 import datagen
import b2oracle as o
rnd = random.Random(20261019)
bad = 0; n_cases = 0
for it in range(400):
    mode = rnd.randrange(5)
    n = rnd.choice([rnd.randrange(0, 400), rnd.randrange(400, 70000), rnd.randrange(65000, 66000), rnd.randrange(70000, 400000)])
    seed = rnd.randrange(1 << 30)
    data = datagen.generate(max(n, 1), mode=mode, seed=seed, span=65536)[:n].tobytes()
    accel = rnd.choice([1, 1, 1, 2, 9, 100])
    got = z.lz4.compressFast(data, accel) if hasattr(z.lz4, 'compressFast') else z.lz4.compressDefault(data)
    want = o.compress_fast(data, accel)
    n_cases += 1
    if bytes(got) != bytes(want):
        bad += 1; print("K1 MISMATCH", mode, n, seed, accel)
    back = z.lz4.decompressSafe(bytes(want), n)
    if bytes(back) != data:
        bad += 1; print("K2 MISMATCH", mode, n, seed)
    if it % 8 == 0 and n >= 13:
        lvl = rnd.choice([3, 5, 7, 9])
        g = z.lz4hc.compressHC(data, lvl); w = o.compress_hc(data, lvl)
        if bytes(g) != bytes(w):
            bad += 1; print("K3 MISMATCH", mode, n, seed, lvl)
print("sweep:", n_cases, "cases,", bad, "mismatches")
