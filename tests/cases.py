"""Seeded synthetic cases shared by the CPU (hostemu) and GPU parity tests.  Expected values
always come from the oracle; inputs are built with the oracle too (it is the checker)."""
import hashlib

import numpy as np

from conftest import fr_bytes, rand_fr, R

INF1 = np.zeros(96, np.uint8); INF1[0] = 0x40
INF2 = np.zeros(192, np.uint8); INF2[0] = 0x40


def make_sig_batch(O, n, seed, corrupt_every=4):
    """n (pk, sig, msg) triples; item i with i % corrupt_every == 1 is corrupted round-robin
    (wrong key / wrong message / wrong signature) so the expected output is not constant."""
    rng = np.random.default_rng(seed)
    sk = rand_fr(rng, n)
    msgs = [hashlib.sha3_256(b"tcb200/test/msg" + seed.to_bytes(4, "little") + i.to_bytes(8, "little")).digest()[: 1 + (i * 7) % 32]
            for i in range(n)]
    pk = O.g1_mul_gen_batch(sk)
    sig = O.sign_batch(sk, msgs)
    kind = 0
    for i in range(n):
        if i % corrupt_every == 1:
            j = (i + 1) % n
            if kind % 3 == 0:
                pk[i] = pk[j]
            elif kind % 3 == 1:
                msgs[i] = msgs[i] + b"!"
            else:
                sig[i] = sig[j]
            kind += 1
    return sk, pk, sig, msgs


def make_combine_batch(O, n, t, seed, group=2, extra=8):
    """n items of t+1 (x, share) pairs in G1 (group=1) or G2 (group=2): one random polynomial,
    per item a different base point and a random subset of indices from 0..t+extra."""
    rng = np.random.default_rng(seed)
    coeff = rand_fr(rng, t + 1)
    m = t + 1
    xs, shares, expect_master = [], [], []
    base_scalars = rand_fr(rng, n)
    if group == 2:
        bases = O.sign_g2_batch(base_scalars, np.tile(O.g2_generator(), (n, 1)))
    else:
        bases = O.g1_mul_gen_batch(base_scalars)
    for i in range(n):
        idx = rng.choice(t + 1 + extra, size=m, replace=False)
        x = fr_bytes([int(j) + 1 for j in idx])
        ski = O.poly_eval(coeff, x)
        if group == 2:
            sh = O.sign_g2_batch(ski, np.tile(bases[i], (m, 1)))
        else:
            sh = O.decrypt_share_batch(ski, np.tile(bases[i], (m, 1)))
        xs.append(x); shares.append(sh)
    master = coeff[:32]
    if group == 2:
        expect = O.sign_g2_batch(np.tile(master, n), bases)
    else:
        expect = O.decrypt_share_batch(np.tile(master, n), bases)
    return np.concatenate(xs), np.concatenate(shares), expect


def check_all(E, O, n_sig=6, n_comb=3, t=3, deg=6, n_eval=5, seed=7):
    """Every C-ABI entry point of engine E against the oracle on the same inputs (bit-exact)."""
    sk, pk, sig, msgs = make_sig_batch(O, n_sig, seed)
    h = O.hash_g2_batch(msgs)
    for algo in (1, 0):                                                        # one-kernel version, then the default two-kernel one
        E.set_hash_algo(algo)
        assert np.array_equal(E.hash_g2_batch(msgs), h), algo
    assert np.array_equal(E.g1_mul_gen_batch(sk), O.g1_mul_gen_batch(sk))
    long_msgs = [m * (1 + 9 * (k % 2)) for k, m in enumerate(msgs)]          # some > 64 bytes (hashed first), some short
    exp_hg = np.stack([O.hash_g1_g2(pk[k], long_msgs[k]) for k in range(n_sig)])
    exp_sig = O.sign_batch(sk, msgs)
    for algo in (1, 0):
        E.set_hash_algo(algo)
        assert np.array_equal(E.hash_g1_g2_batch(pk, long_msgs), exp_hg), algo
        assert np.array_equal(E.sign_batch(sk, msgs), exp_sig), algo
    assert np.array_equal(E.sign_g2_batch(sk, h), O.sign_g2_batch(sk, h))
    exp = O.verify_batch(pk, sig, msgs)
    assert 0 < exp.sum() < n_sig or n_sig < 3
    assert np.array_equal(E.verify_batch(pk, sig, msgs), exp)
    check_verify_hash_modes(E, O, pk, sig, msgs, exp)
    assert np.array_equal(E.verify_g2_batch(pk, h, None, sig), O.verify_g2_batch(pk, h, None, sig))
    g1 = np.tile(O.g1_generator(), (n_sig, 1))
    assert np.array_equal(E.verify_g2_batch(pk, h, g1, sig), exp)
    # combine in G2 and G1, decrypt
    x2, s2, master2 = make_combine_batch(O, n_comb, t, seed + 1, group=2)
    out, st = E.combine_g2_batch(n_comb, t, x2, s2)
    oout, ost = O.combine_g2_batch(n_comb, t, x2, s2)
    assert np.array_equal(out, oout) and np.array_equal(st, ost) and np.array_equal(out, master2)
    x1, s1, master1 = make_combine_batch(O, n_comb, t, seed + 2, group=1)
    out, st = E.combine_g1_batch(n_comb, t, x1, s1)
    oout, ost = O.combine_g1_batch(n_comb, t, x1, s1)
    assert np.array_equal(out, oout) and np.array_equal(st, ost) and np.array_equal(out, master1)
    vs = [bytes((i * 31 + k) & 0xff for k in range(5 + 20 * i)) for i in range(n_comb)]
    dec, st = E.decrypt_batch(n_comb, t, x1, s1, vs)
    odec, ost = O.decrypt_batch(n_comb, t, x1, s1, vs)
    assert dec == odec and np.array_equal(st, ost)
    # encrypt: one oracle call per item
    r = rand_fr(np.random.default_rng(seed + 9), n_comb)
    plains = [bytes((7 * i + k) & 0xff for k in range(3 + 30 * i)) for i in range(n_comb)]
    pks = O.g1_mul_gen_batch(rand_fr(np.random.default_rng(seed + 10), n_comb))
    u, v, w = E.encrypt_batch(pks, r, plains)
    for i in range(n_comb):
        ou, ov, ow = O.encrypt(pks[i], r[32 * i:32 * i + 32], plains[i])
        assert np.array_equal(u[i], ou) and v[i] == ov and np.array_equal(w[i], ow)
    # decrypt shares
    rng = np.random.default_rng(seed + 3)
    ski = rand_fr(rng, n_comb)
    assert np.array_equal(E.decrypt_share_batch(ski, master1), O.decrypt_share_batch(ski, master1))
    # Commitment::evaluate
    coeff = rand_fr(rng, deg + 1)
    comm = O.g1_mul_gen_batch(coeff)
    xs = fr_bytes([i + 1 for i in range(n_eval - 1)] + [int.from_bytes(rng.bytes(40), "little")])
    out = E.commitment_eval_batch(comm, xs)
    assert np.array_equal(out, O.commitment_eval_batch(comm, xs))
    assert np.array_equal(out, O.g1_mul_gen_batch(O.poly_eval(coeff, xs)))


def check_verify_hash_modes(E, O, pk, sig, msgs, exp):
    """PublicKey::verify with the message point taken up to the unit 3(x^2-1) and the generator scaled by the same unit
    (the default, tcb200.h: tcb_set_verify_hash) must give the booleans of the exact-hash check — also for a wrong message,
    a wrong key and a signature at infinity."""
    n = len(msgs)
    X2 = 0xd201000000010000 ** 2
    for mode, k in ((1, 1), (0, 3 * (X2 - 1))):              # the generator verify pairs with the signature: g1, or [3(x^2-1)] g1
        E.set_verify_hash(mode)
        assert np.array_equal(E.verifier_generator(), O.g1_mul_gen_batch(fr_bytes([k]))[0]), mode
    inf2 = np.zeros(192, np.uint8); inf2[0] = 0x40
    sig_i = sig.copy(); sig_i[0] = inf2
    trials = [(pk, sig, msgs), (pk, sig, msgs[1:] + msgs[:1]), (np.roll(pk, 1, axis=0), sig, msgs), (pk, sig_i, msgs)]
    for k, (p_, s_, m_) in enumerate(trials):
        want = exp if k == 0 else O.verify_batch(p_, s_, m_)
        for mode, algo in ((1, 1), (0, 1), (1, 0), (0, 0)):
            E.set_verify_hash(mode)
            E.set_hash_algo(algo)
            assert np.array_equal(E.verify_batch(p_, s_, m_), want), (k, mode, algo)


def check_poly(E, O, seed=13):
    """Poly::evaluate and Poly * Poly in Fr against the oracle's Horner / Python integers, incl. the
    reference's own known-answer polynomial (src/poly.rs:783-797: 5 x^3 + x - 2 and x^3 + 1... evaluated at small points)."""
    rng = np.random.default_rng(seed)
    coeff = rand_fr(rng, 9)
    xs = fr_bytes([0, 1, 2, R - 1, 65536] + [int.from_bytes(rng.bytes(40), "little") for _ in range(8)])
    assert np.array_equal(E.poly_eval_batch(coeff, xs), O.poly_eval(coeff, xs))
    kat = fr_bytes([R - 2, 1, 0, 5])                      # 5 x^3 + x - 2  (poly.rs:786)
    got = E.poly_eval_batch(kat, fr_bytes([1, 2, 3]))
    assert [int.from_bytes(bytes(g), "little") for g in got] == [4, 40, 136]
    n, da, db = 3, 4, 2
    a = [[int.from_bytes(rng.bytes(40), "little") % R for _ in range(da + 1)] for _ in range(n)]
    b = [[int.from_bytes(rng.bytes(40), "little") % R for _ in range(db + 1)] for _ in range(n)]
    a[1][da] = 0; b[2] = [0] * (db + 1)                   # a leading zero and a zero polynomial
    out = E.poly_mul_batch(n, fr_bytes([c for r_ in a for c in r_]), fr_bytes([c for r_ in b for c in r_]))
    for i in range(n):
        exp = [0] * (da + db + 1)
        for p_, ca in enumerate(a[i]):
            for q_, cb in enumerate(b[i]):
                exp[p_ + q_] = (exp[p_ + q_] + ca * cb) % R
        assert [int.from_bytes(bytes(c), "little") for c in out[i]] == exp


def check_edges(E, O):
    """Edge semantics of SURVEY §7: infinity operands, t == 0, duplicate indices, zero scalars,
    empty batches, ragged messages."""
    g1, g2 = O.g1_generator(), O.g2_generator()
    # infinity operands contribute 1
    a = np.stack([INF1, g1, INF1]); b = np.stack([g2, g2, INF2]); c = np.stack([g1, g1, INF1]); d = np.stack([INF2, INF2, g2])
    assert np.array_equal(E.verify_g2_batch(a, b, c, d), O.verify_g2_batch(a, b, c, d))
    assert list(O.verify_g2_batch(a, b, c, d)) == [1, 0, 1]
    # zero and r-1 scalars, infinity base
    sk = fr_bytes([0, R - 1, 5])
    pts = np.stack([g1, g1, INF1])
    assert np.array_equal(E.decrypt_share_batch(sk, pts), O.decrypt_share_batch(sk, pts))
    h = np.stack([g2, g2, INF2])
    assert np.array_equal(E.sign_g2_batch(sk, h), O.sign_g2_batch(sk, h))
    # t == 0 returns the first sample unchanged
    rng = np.random.default_rng(11)
    s = O.sign_g2_batch(rand_fr(rng, 2), np.tile(g2, (2, 1)))
    out, st = E.combine_g2_batch(2, 0, fr_bytes([1, 2]), s)
    assert np.array_equal(out, s) and not st.any()
    # duplicate indices: the reference's by-value filter (lib.rs:757) never errors; same garbage point expected
    x = fr_bytes([2, 2, 4])
    sh = O.sign_g2_batch(rand_fr(rng, 3), np.tile(g2, (3, 1)))
    out, st = E.combine_g2_batch(1, 2, x, sh)
    oout, ost = O.combine_g2_batch(1, 2, x, sh)
    assert np.array_equal(out, oout) and np.array_equal(st, ost)
    # shares that cancel to infinity: lambda-weighted sum of identical points with x = (1,2): 2P - P... use P, -P check via G1
    x = fr_bytes([1, 2])
    p = O.g1_mul_gen_batch(fr_bytes([7]))[0]
    p2 = O.g1_mul_gen_batch(fr_bytes([14]))[0]     # f(1) = 7, f(2) = 14  =>  f(0) = 0  => infinity
    out, st = E.combine_g1_batch(1, 1, x, np.stack([p, p2]))
    oout, _ = O.combine_g1_batch(1, 1, x, np.stack([p, p2]))
    assert np.array_equal(out, oout) and out[0][0] == 0x40
    # x not canonical (>= r) -> status 3
    bad = np.frombuffer(b"\xff" * 32 + (2).to_bytes(32, "little"), np.uint8).copy()
    _, st = E.combine_g1_batch(1, 1, bad, np.stack([p, p2]))
    assert st[0] == 3
    # empty batch
    z = np.zeros(0, np.uint8)
    assert E.verify_g2_batch(z, z, None, z).size == 0
    assert E.hash_g2_batch([]).shape[0] == 0
    # ragged messages incl. empty, > one Keccak block, exactly 136 bytes
    msgs = [b"", b"a" * 135, b"a" * 136, b"a" * 137, b"b" * 300, bytes(range(64)), bytes(range(65))]
    assert np.array_equal(E.hash_g2_batch(msgs), O.hash_g2_batch(msgs))
    # decrypt with empty and long ciphertext bodies
    xs, sh, _ = make_combine_batch(O, 2, 1, 5, group=1)
    vs = [b"", bytes(range(200))]
    dec, st = E.decrypt_batch(2, 1, xs, sh, vs)
    odec, _ = O.decrypt_batch(2, 1, xs, sh, vs)
    assert dec == odec


def check_msm(E, O, n=3, m=5, seed=21):
    """The shared-doubling multi-scalar multiplication behind combine / decrypt / lincomb against
    sum_i k_i (a_i G) = (sum_i k_i a_i) G computed with the oracle's single scalar multiplications.
    Shares include: zero / one / r-1 / even / odd scalars, the point at infinity, a repeated
    (P, k) pair (accumulator == table entry: doubling branch of the mixed addition) and a
    (P, k), (-P, k) pair (accumulator returns to infinity)."""
    rng = np.random.default_rng(seed)
    for group in (1, 2):
        width = 96 if group == 1 else 192
        inf = INF1 if group == 1 else INF2
        gen = (lambda a: O.g1_mul_gen_batch(fr_bytes(a))) if group == 1 else \
              (lambda a: O.sign_g2_batch(fr_bytes(a), np.tile(O.g2_generator(), (len(a), 1))))
        a_all, k_all, inf_mask = [], [], []
        for i in range(n):
            a = [int.from_bytes(rng.bytes(40), "little") % R for _ in range(m)]
            k = [int.from_bytes(rng.bytes(40), "little") % R for _ in range(m)]
            mask = [False] * m
            if i == 0:
                k[0], k[1], k[2] = 0, 1, R - 1
                k[3] &= ~1; k[4] |= 1
            elif i == 1:
                a[1], k[1] = a[0], k[0]               # repeated pair
                a[3], k[3] = (R - a[2]) % R, k[2]     # P and -P with the same scalar
            else:
                mask[1] = True                        # infinity share
                if m > 3:
                    mask[3] = True
            a_all += a; k_all += k; inf_mask += mask
        pts = gen(a_all)
        for j, isinf in enumerate(inf_mask):
            if isinf:
                pts[j] = inf
        total = [sum(k_all[i * m + s] * a_all[i * m + s] for s in range(m) if not inf_mask[i * m + s]) % R for i in range(n)]
        expect = gen(total)
        for i in range(n):
            if total[i] == 0:
                expect[i] = inf
        fn = E.g1_lincomb_batch if group == 1 else E.g2_lincomb_batch
        out = fn(n, m, fr_bytes(k_all), pts)
        assert out.shape == (n, width) and np.array_equal(out, expect)


def check_codecs(E, O, golden=None, seed=9, n=6):
    """compress / checked decompress against the oracle, incl. infinity and the
    invalid encodings (x >= p, off-curve, wrong flags, on-curve-but-not-in-subgroup)."""
    rng = np.random.default_rng(seed)
    sk = rand_fr(rng, n)
    p1 = np.concatenate([O.g1_mul_gen_batch(sk), INF1[None, :]])
    p2 = np.concatenate([O.sign_g2_batch(sk, np.tile(O.g2_generator(), (n, 1))), INF2[None, :]])
    c1, c2 = O.g1_compress(p1), O.g2_compress(p2)
    assert np.array_equal(E.g1_compress_batch(p1), c1)
    assert np.array_equal(E.g2_compress_batch(p2), c2)
    u1, s1 = E.g1_decompress_batch(c1)
    u2, s2 = E.g2_decompress_batch(c2)
    assert not s1.any() and not s2.any() and np.array_equal(u1, p1) and np.array_equal(u2, p2)
    # tampered encodings: every variant must get the oracle's verdict and output
    bad1 = c1.copy(); bad1[0, 47] ^= 1; bad1[1, 0] &= 0x7f; bad1[2, 0] |= 0x1f; bad1[2, 1:] = 0xff; bad1[3, 0] ^= 0x20
    bad2 = c2.copy(); bad2[0, 95] ^= 1; bad2[1, 0] &= 0x7f; bad2[2, 60] ^= 0x80; bad2[3, 0] ^= 0x20; bad2[4, 48] = 0xff; bad2[4, 49:96] = 0xff
    for bad, dec_e, dec_o in ((bad1, E.g1_decompress_batch, O.g1_decompress), (bad2, E.g2_decompress_batch, O.g2_decompress)):
        ue, se = dec_e(bad)
        uo, so = dec_o(bad)
        assert np.array_equal(se, so), (se, so)
        assert np.array_equal(ue[so == 0], uo[so == 0])
    if golden is not None:
        from conftest import hxs
        u, s = E.g1_decompress_batch(hxs(golden["g1_bad_compressed"]))
        assert list(s) == [3] * len(s)
        u, s = E.g1_decompress_batch(hxs(golden["pk_compressed"]))
        assert not s.any() and [bytes(x).hex() for x in u] == golden["pk"]
        u, s = E.g2_decompress_batch(hxs(golden["hash_g2_compressed"]))
        assert not s.any() and [bytes(x).hex() for x in u] == golden["hash_g2"]
