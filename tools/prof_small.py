"""Short driver for ncu captures of the non-headline kernels: one Commitment::evaluate sweep, one combine_signatures batch
(t = 10) and one threshold decrypt (t = 64).  Inputs are produced on the GPU.   python tools/prof_small.py [which] ; env N3, N4, N5, DEG"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import fr_bytes, rand_fr
from threshold_crypto_b200._lib import Engine

E = Engine()
rng = np.random.default_rng(1)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "eval"):
    deg = int(os.environ.get("DEG", "127"))
    n = int(os.environ.get("N5", str(1 << 16)))
    comm = E.g1_mul_gen_batch(rand_fr(rng, deg + 1))
    xs = fr_bytes([i + 1 for i in range(n)])
    out = E.commitment_eval_batch(comm, xs)
    print("eval", out.shape)
if which in ("all", "combine"):
    n, t = int(os.environ.get("N3", str(1 << 14))), 10
    m = t + 1
    pts = E.sign_g2_batch(rand_fr(rng, n * m), np.tile(E.hash_g2_batch([b"x"])[0], (n * m, 1)))
    idx = np.stack([np.sort(rng.choice(32, size=m, replace=False)) for _ in range(n)])
    xs = fr_bytes([int(j) + 1 for j in idx.reshape(-1)])
    out, st = E.combine_g2_batch(n, t, xs, pts)
    print("combine", out.shape, int(st.sum()))
if which in ("all", "decrypt"):
    n, t = int(os.environ.get("N4", str(1 << 12))), 64
    m = t + 1
    pts = E.g1_mul_gen_batch(rand_fr(rng, n * m))
    xs = fr_bytes([(i % m) + 1 for i in range(n * m)])
    out, st = E.combine_g1_batch(n, t, xs, pts)
    print("combine_g1", out.shape, int(st.sum()))
