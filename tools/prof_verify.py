"""One pairing-check pass over 2^16 items per engine (shared-memory engine, then the round-1 register engine), for the ncu launch
list / --set full captures: inputs are produced on the GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import rand_fr
from threshold_crypto_b200._lib import Engine, ENGINE_QUAD_REG, ENGINE_QUAD_SMEM, ENGINE_QUAD_SMEM_REGFE

n = int(os.environ.get("N", str(1 << 16)))
E = Engine(devices=[0])
rng = np.random.default_rng(3)
sk = rand_fr(rng, n)
msgs = [i.to_bytes(8, "little") * 4 for i in range(n)]
pk = E.g1_mul_gen_batch(sk)
sig = E.sign_batch(sk, msgs)
h = E.hash_g2_batch(msgs)
for eng in ([ENGINE_QUAD_SMEM, ENGINE_QUAD_SMEM_REGFE, ENGINE_QUAD_REG] if len(sys.argv) < 2 else [int(sys.argv[1])]):
    E.set_engine(eng)
    ok = E.verify_g2_batch(pk, h, None, sig)
    print("engine", eng, "ok", int(ok.sum()), "of", n)
