"""Times BASELINE.json configs #3, #4, #5 end to end through the host-buffer C ABI (one GPU) and checks
a sample of every output against the oracle.  Not the headline (bench.py is); recorded in profiles/."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from conftest import fr_bytes, rand_fr, R
from threshold_crypto_b200._lib import Engine

E = Engine()
O.set_threads(16)
rng = np.random.default_rng(4)
out = {}


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


# ---- config #4: threshold decrypt t = 64, 2^12 ciphertexts
t, n = 64, 1 << 12
m = t + 1
poly = rand_fr(rng, m)
pk = E.g1_mul_gen_batch(poly[:32])
rs = rand_fr(rng, n)
plains = [bytes([i & 0xff]) * 64 for i in range(n)]
dt_enc, (u, v, w) = timed(lambda: E.encrypt_batch(np.tile(pk[0], (n, 1)), rs, plains), 1)
xs_one = fr_bytes([i + 1 for i in range(m)])
sk_shares = O.poly_eval(poly, xs_one)                       # 65 secret key shares
sk_rep = np.tile(sk_shares, (n, 1)).reshape(-1)             # item-major: (item, share)
u_rep = np.repeat(u, m, axis=0)
dt_a, dshares = timed(lambda: E.decrypt_share_batch(sk_rep, u_rep), 2)
xs = np.tile(xs_one, n)
dt_b, (dec, st) = timed(lambda: E.decrypt_batch(n, t, xs, dshares, v), 2)
assert dec == plains and not st.any()
oc, _ = O.decrypt_batch(4, t, xs[:4 * m * 32], dshares[:4 * m], v[:4])
assert oc == plains[:4]
out["config4_decrypt_t64_2^12"] = {"encrypt_per_s": n / dt_enc, "decrypt_shares_per_s": n * m / dt_a, "decrypts_per_s": n / dt_b,
                                  "ms": {"encrypt": 1e3 * dt_enc, "shares": 1e3 * dt_a, "decrypt": 1e3 * dt_b}}
# ---- config #5: Commitment::evaluate, degree 1023 at 2^16 indices
coeff = rand_fr(rng, 1024)
comm = E.g1_mul_gen_batch(coeff)
n5 = 1 << 16
xs5 = fr_bytes([i + 1 for i in range(n5)])
dt5, ev = timed(lambda: E.commitment_eval_batch(comm, xs5), 1)
sel = [0, 1, 4095, 65535]
assert np.array_equal(ev[sel], O.g1_mul_gen_batch(O.poly_eval(coeff, fr_bytes([i + 1 for i in sel]))))
out["config5_commit_eval_deg1023_2^16"] = {"evals_per_s": n5 / dt5, "ms": 1e3 * dt5}
# ---- config #3 e2e: combine_signatures t = 10, 2^14
t3, n3 = 10, 1 << 14
m3 = t3 + 1
poly3 = rand_fr(rng, m3)
hm = E.hash_g2_batch([b"m%d" % i for i in range(n3)])
idx = np.stack([np.sort(rng.choice(32, size=m3, replace=False)) for _ in range(n3)])
sk32 = O.poly_eval(poly3, fr_bytes([j + 1 for j in range(32)]))
sk_rep3 = sk32[idx.reshape(-1)].reshape(-1)
shares3 = E.sign_g2_batch(sk_rep3, np.repeat(hm, m3, axis=0))
xs3 = fr_bytes([int(j) + 1 for j in idx.reshape(-1)])
dt3, (comb, st3) = timed(lambda: E.combine_g2_batch(n3, t3, xs3, shares3), 3)
assert np.array_equal(comb, E.sign_g2_batch(np.tile(poly3[:32], n3), hm)) and not st3.any()
out["config3_combine_t10_2^14_e2e"] = {"combines_per_s": n3 / dt3, "ms": 1e3 * dt3}
print(json.dumps(out, indent=1))
