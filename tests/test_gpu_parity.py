"""GPU (B200): the CUDA kernels through the C ABI (libtcb200.so) against the oracle — bit-exact
on the same seeded inputs (pairing checks on lane quads, G2 work on lane pairs, G1 per thread), the golden
vectors, the edge cases, and size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

import cases
from conftest import fr_bytes, hx, hxs, rand_fr

pytestmark = pytest.mark.gpu


def test_fp_selftest_on_device(gpu_engine):
    """PTX carry-chain Montgomery multiply / dot2 / add / sub and the sliced Fp2 ops against the
    portable CIOS on 2^18 random + edge operands, on the device."""
    assert gpu_engine.selftest_fp(1 << 18, seed=123) == 0


@pytest.mark.parametrize("seed", [21, 22])
def test_all_entry_points_vs_oracle(gpu_engine, O, seed):
    cases.check_all(gpu_engine, O, n_sig=37, n_comb=9, t=4, deg=9, n_eval=33, seed=seed)


def test_edges(gpu_engine, O):
    cases.check_edges(gpu_engine, O)


def test_golden_vectors(gpu_engine, golden):
    E = gpu_engine
    msgs = [bytes.fromhex(m) for m in golden["msgs"]]
    assert [bytes(p).hex() for p in E.hash_g2_batch(msgs)] == golden["hash_g2"]
    sk = np.concatenate([hx(s) for s in golden["sk"]])
    assert [bytes(p).hex() for p in E.g1_mul_gen_batch(sk)] == golden["pk"]
    nsk = len(golden["sk"])
    sks = np.concatenate([hx(golden["sk"][i % nsk]) for i in range(len(msgs))])
    assert [bytes(s).hex() for s in E.sign_batch(sks, msgs)] == golden["sig"]
    for pi, si, mi, exp in golden["verify_cases"]:
        assert bool(E.verify_batch(hx(golden["pk"][pi]), hx(golden["sig"][si]), [msgs[mi]])[0]) == exp
    ts = golden["threshold_sig"]
    for s in ts["sets"]:
        out, st = E.combine_g2_batch(1, ts["t"], fr_bytes([i + 1 for i in s["idx"]]), hxs(s["shares"]))
        assert bytes(out[0]).hex() == s["combined"] and st[0] == 0
    e = golden["enc"]
    out, st = E.decrypt_batch(1, 2, fr_bytes([i + 1 for i in e["idx"]]), hxs(e["dshares"]), [bytes.fromhex(e["v"])])
    assert out[0].hex() == e["plain"]
    ce = golden["commit_eval"]
    out = E.commitment_eval_batch(hxs(ce["coeff"]), np.concatenate([hx(x) for x in ce["x"]]))
    assert [bytes(p).hex() for p in out] == ce["out"]
    shares = E.commitment_eval_batch(hxs(golden["commitment"]), fr_bytes([i + 1 for i in range(5)]))
    assert [bytes(p).hex() for p in shares] == golden["pk_shares"]


def test_config1_threshold_sig_example(gpu_engine, O):
    """BASELINE config #1 (examples/threshold_sig.rs path): t=2, n=5, combine the first 3 shares,
    verify under the master key; a second disjoint-ish subset gives the SAME signature
    (src/lib.rs:823-873)."""
    E = gpu_engine
    rng = np.random.default_rng(99)
    t, nodes = 2, 5
    poly = rand_fr(rng, t + 1)
    msg = b"hey, this is alice"
    sk_shares = O.poly_eval(poly, fr_bytes([i + 1 for i in range(nodes)]))
    sig_shares = E.sign_batch(sk_shares, [msg] * nodes)
    pk_set = E.g1_mul_gen_batch(poly)                     # Poly::commitment
    pk_shares = E.commitment_eval_batch(pk_set, fr_bytes([i + 1 for i in range(nodes)]))
    assert E.verify_batch(pk_shares, sig_shares, [msg] * nodes).all()
    a, _ = E.combine_g2_batch(1, t, fr_bytes([1, 2, 3]), sig_shares[:3])
    b, _ = E.combine_g2_batch(1, t, fr_bytes([5, 3, 4]), sig_shares[[4, 2, 3]])
    assert np.array_equal(a, b)
    assert E.verify_batch(pk_set[0], a[0], [msg])[0] == 1
    assert E.verify_batch(pk_set[0], a[0], [b"another message"])[0] == 0
    assert np.array_equal(a, O.combine_g2_batch(1, t, fr_bytes([1, 2, 3]), sig_shares[:3])[0])


def test_medium_batch_verify_vs_oracle(gpu_engine, O):
    """2^10 verifies, every 4th corrupted; oracle runs on all host cores."""
    O.set_threads(16)
    n = 1 << 10
    sk, pk, sig, msgs = cases.make_sig_batch(O, n, 5)
    exp = O.verify_batch(pk, sig, msgs)
    got = gpu_engine.verify_batch(pk, sig, msgs)
    O.set_threads(1)
    assert np.array_equal(got, exp)
    assert 0 < exp.sum() < n


def test_full_size_properties_config2(gpu_engine, O):
    """BASELINE config #2 size (2^16 verifies) through size-independent properties: signatures
    made on the GPU verify on the GPU, corrupted ones do not, and a 2^9 sample is checked
    bit-exactly against the oracle."""
    E = gpu_engine
    n = 1 << 16
    rng = np.random.default_rng(2)
    sk = rand_fr(rng, n)
    msgs = [i.to_bytes(8, "little") * 4 for i in range(n)]
    pk = E.g1_mul_gen_batch(sk)
    sig = E.sign_batch(sk, msgs)
    bad = np.arange(n) % 16 == 5
    sig_c = sig.copy()
    sig_c[bad] = np.roll(sig, 1, axis=0)[bad]
    ok = E.verify_batch(pk, sig_c, msgs)
    assert np.array_equal(ok.astype(bool), ~bad)
    sel = rng.choice(n, 512, replace=False)
    O.set_threads(16)
    assert np.array_equal(O.sign_batch(sk.reshape(n, 32)[sel].reshape(-1), [msgs[i] for i in sel]), sig[sel])
    assert np.array_equal(O.verify_batch(pk[sel], sig_c[sel], [msgs[i] for i in sel]), ok[sel])
    O.set_threads(1)


def test_full_size_properties_config3_4_5(gpu_engine, O):
    """Config #3 (t=10 combine, 2^14 msgs), #4 (t=64 decrypt, 2^12), #5 (deg-1023 evaluate) at a
    reduced count for the oracle-checked part and the interpolation identity for the rest:
    combining shares of sk_i * B must give master * B."""
    E = gpu_engine
    O.set_threads(16)
    for (n, t, group) in ((256, 10, 2), (64, 64, 1)):
        xs, sh, master = cases.make_combine_batch(O, n, t, 40 + t, group=group, extra=21)
        if group == 2:
            out, st = E.combine_g2_batch(n, t, xs, sh)
        else:
            out, st = E.combine_g1_batch(n, t, xs, sh)
        assert not st.any() and np.array_equal(out, master)
    # config #5: degree 1023 at indices 1..64 vs oracle, and f(x)*g1 identity
    rng = np.random.default_rng(5)
    coeff = rand_fr(rng, 1024)
    comm = E.g1_mul_gen_batch(coeff)
    assert np.array_equal(comm[:8], O.g1_mul_gen_batch(coeff[:8 * 32]))
    xs = fr_bytes([i + 1 for i in range(64)])
    out = E.commitment_eval_batch(comm, xs)
    assert np.array_equal(out, E.g1_mul_gen_batch(O.poly_eval(coeff, xs)))
    assert np.array_equal(out[:4], O.commitment_eval_batch(comm, xs[:4 * 32]))
    O.set_threads(1)


def test_msm_groupings_and_edges(gpu_engine, O):
    """The shared-doubling multi-scalar multiplication with forced 1 / 2 / 3 / m partial sums per item and
    with the automatic choice: zero / one / r-1 scalars, infinity shares, repeated and cancelling pairs."""
    E = gpu_engine
    try:
      for algo in (0, 1, 2, 3, 4, 5, 6):
        E.set_msm_algo(algo)
        for g in (1, 2, 3, 1000, 0):
            E.set_msm_groups(g)
            cases.check_msm(E, O, n=5, m=7, seed=30 + g % 7)
            x, s, master = cases.make_combine_batch(O, 6, 4, 50 + g % 7, group=2)
            out, st = E.combine_g2_batch(6, 4, x, s)
            assert np.array_equal(out, master) and not st.any()
            x, s, master = cases.make_combine_batch(O, 6, 4, 60 + g % 7, group=1)
            out, st = E.combine_g1_batch(6, 4, x, s)
            assert np.array_equal(out, master) and not st.any()
    finally:
        E.set_msm_groups(0)
        E.set_msm_algo(0)


def test_scalar_mul_special_scalars_gpu(gpu_engine, O):
    """The recodings' split points (multiples and neighbours of X^2 and X), 0, 1, r-1 and random scalars through the kernels."""
    from conftest import R
    E = gpu_engine
    rng = np.random.default_rng(321)
    X = 0xd201000000010000
    X2 = X * X
    ks = [0, 1, 2, 3, R - 1, R - 2, X2, X2 - 1, X2 + 1, 2 * X2, (1 << 128) - 1, 1 << 128, (1 << 129) + 1, X, X - 1, X + 1,
          X ** 3 % R, (X ** 3 + 1) % R, (R - 1) // 2, (R + 1) // 2]
    ks += [(a + b * X2) % R for a in (0, 1, (1 << 127) - 1, (1 << 128) - 1) for b in (1, (1 << 126) + 5, (1 << 127) - 1)]
    ks += [int.from_bytes(rng.bytes(40), "little") % R for _ in range(32)]
    n = len(ks)
    sk = fr_bytes(ks)
    O.set_threads(16)
    base1 = O.g1_mul_gen_batch(fr_bytes([int.from_bytes(rng.bytes(40), "little") % R for _ in range(n)]))
    assert np.array_equal(E.decrypt_share_batch(sk, base1), O.decrypt_share_batch(sk, base1))
    base2 = O.sign_g2_batch(fr_bytes([int.from_bytes(rng.bytes(40), "little") % R for _ in range(n)]), np.tile(O.g2_generator(), (n, 1)))
    assert np.array_equal(E.sign_g2_batch(sk, base2), O.sign_g2_batch(sk, base2))
    O.set_threads(1)


def test_two_pass_lagrange_large_batch(gpu_engine, O):
    """Batches of >= 16 x SMs items take the two-pass Lagrange kernels (one shared inversion per item): interpolation
    identity on every item and the oracle on a sample, incl. an item with a repeated index (by-value filter)."""
    E = gpu_engine
    O.set_threads(16)
    n, t = 2500, 2
    xs, sh, master = cases.make_combine_batch(O, n, t, 91, group=1, extra=5)
    xs = xs.copy()
    xs[32 * 3 * 7 + 32: 32 * 3 * 7 + 64] = xs[32 * 3 * 7: 32 * 3 * 7 + 32]        # item 7: x_1 := x_0
    out, st = E.combine_g1_batch(n, t, xs, sh)
    keep = np.arange(n) != 7
    assert not st.any() and np.array_equal(out[keep], master[keep])
    oo, _ = O.combine_g1_batch(16, t, xs[:16 * 3 * 32], sh[:16 * 3])
    assert np.array_equal(out[:16], oo)
    O.set_threads(1)


def test_multi_device_ctx_matches_single(O):
    """A ctx over all visible devices shards contiguous slices and returns the same bytes."""
    import torch
    from threshold_crypto_b200._lib import Engine
    nd = torch.cuda.device_count()
    E = Engine(devices=list(range(nd)))
    sk, pk, sig, msgs = cases.make_sig_batch(O, 23, 77)
    assert np.array_equal(E.verify_batch(pk, sig, msgs), O.verify_batch(pk, sig, msgs))
    xs, sh, master = cases.make_combine_batch(O, 7, 3, 78, group=2)
    out, st = E.combine_g2_batch(7, 3, xs, sh)
    assert np.array_equal(out, master)
    E.close()


def test_codecs(gpu_engine, O, golden):
    cases.check_codecs(gpu_engine, O, golden, n=40)


def test_miller_engines_bit_identical(gpu_engine, O):
    """The shared-memory Miller loop (quadsm.cuh: operands in shared memory, dot-product form, TMA input staging) against
    the register engine's on the same items: f must be bit-identical.  70 items = two full blocks and a ragged tail; the
    second call passes explicit C points, the third infinity operands."""
    E = gpu_engine
    n = 70
    sk, pk, sig, msgs = cases.make_sig_batch(O, n, 91)
    h = O.hash_g2_batch(msgs)
    assert E.selftest_miller(pk, h, None, sig) == 0
    g1 = np.tile(O.g1_generator(), (n, 1))
    assert E.selftest_miller(pk, h, g1, sig) == 0
    a = np.stack([cases.INF1, O.g1_generator(), cases.INF1, pk[0], pk[1]])
    b = np.stack([O.g2_generator(), O.g2_generator(), cases.INF2, cases.INF2, h[1]])
    c = np.stack([O.g1_generator(), cases.INF1, cases.INF1, pk[3], cases.INF1])
    d = np.stack([cases.INF2, sig[0], O.g2_generator(), sig[3], sig[1]])
    assert E.selftest_miller(a, b, c, d) == 0
    assert np.array_equal(E.verify_g2_batch(a, b, c, d), O.verify_g2_batch(a, b, c, d))


def test_both_pairing_engines_vs_oracle(gpu_engine, O):
    """tcb_set_engine: the round-1 register kernel and the shared-memory engine return the oracle's booleans, also for
    encodings >= p (ok = 0) and through device pointers that are NOT 16-byte aligned (plain-copy staging instead of TMA)."""
    import torch
    from threshold_crypto_b200._lib import ENGINE_QUAD_REG, ENGINE_QUAD_SMEM, ENGINE_QUAD_SMEM_REGFE
    E = gpu_engine
    n = 45
    sk, pk, sig, msgs = cases.make_sig_batch(O, n, 92, corrupt_every=3)
    h = O.hash_g2_batch(msgs)
    pk = pk.copy()
    pk[7, :48] = 0xff                      # x >= p: invalid encoding -> false
    exp = O.verify_g2_batch(pk, h, None, sig)
    assert exp[7] == 0 and 0 < exp.sum() < n
    try:
        for eng in (ENGINE_QUAD_REG, ENGINE_QUAD_SMEM_REGFE, ENGINE_QUAD_SMEM):
            E.set_engine(eng)
            assert np.array_equal(E.verify_g2_batch(pk, h, None, sig), exp)
            assert np.array_equal(E.verify_batch(pk, sig, msgs), exp)
        dev = torch.device("cuda", 0)
        st = torch.cuda.current_stream().cuda_stream

        def off1(arr):
            t = torch.zeros(arr.size + 17, dtype=torch.uint8, device=dev)
            t[1:1 + arr.size] = torch.from_numpy(arr.reshape(-1)).to(dev)
            return t[1:1 + arr.size]
        d_pk, d_h, d_sig = off1(pk), off1(h), off1(sig)
        assert d_pk.data_ptr() % 16 == 1
        d_ok = torch.zeros(n, dtype=torch.uint8, device=dev)
        E.dev_call("tcb_verify_g2_batch_dev", st, ("size", n), d_pk.data_ptr(), d_h.data_ptr(), 0, d_sig.data_ptr(), d_ok.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_ok.cpu().numpy(), exp)
    finally:
        E.set_engine(ENGINE_QUAD_SMEM)


def test_hash_g2_algorithms_and_verify_modes(gpu_engine, O):
    """tcb_set_hash_algo / tcb_set_verify_hash: the two-kernel hash_g2 (point per thread + cofactor clearing on lane pairs) and the
    one-kernel version give the oracle's points on ragged messages (0..300 bytes, a count that leaves a partial warp), and
    PublicKey::verify with the message point taken up to the unit 3(x^2-1) against the scaled generator gives the oracle's booleans
    (valid, wrong message, wrong key, signature at infinity)."""
    E = gpu_engine
    O.set_threads(16)
    n = 333
    rng = np.random.default_rng(77)
    msgs = [bytes(rng.integers(0, 256, int(l), dtype=np.uint8)) for l in rng.integers(0, 301, n)]
    msgs[0] = b""; msgs[1] = bytes(135); msgs[2] = bytes(136); msgs[3] = bytes(137)
    exp = O.hash_g2_batch(msgs)
    try:
        for algo in (1, 0):
            E.set_hash_algo(algo)
            assert np.array_equal(E.hash_g2_batch(msgs), exp), algo
        sk, pk, sig, vm = cases.make_sig_batch(O, 200, 93, corrupt_every=5)
        cases.check_verify_hash_modes(E, O, pk, sig, vm, O.verify_batch(pk, sig, vm))
    finally:
        E.set_hash_algo(0); E.set_verify_hash(0)
        O.set_threads(1)


def test_verifier_pieces_compose_to_verify(gpu_engine, O):
    """include/tcb200.h: tcb_verify_batch == tcb_verify_g2_batch(pk, tcb_verifier_hash_g2 points, tcb_verifier_generator, sig) in both
    tcb_set_verify_hash modes; in the default mode the points are [3(x^2-1)] hash_g2(msg) (checked through the oracle's G2
    multiplication), in mode 1 the oracle's hash_g2 itself."""
    import torch
    from threshold_crypto_b200._lib import pack_msgs
    E = gpu_engine
    n = 37
    sk, pk, sig, msgs = cases.make_sig_batch(O, n, 94, corrupt_every=4)
    exp = O.verify_batch(pk, sig, msgs)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    buf, off = pack_msgs(msgs)
    d_msg, d_off = torch.from_numpy(buf).to(dev), torch.from_numpy(off.view(np.int64)).to(dev)
    d_h = torch.zeros(n * 192, dtype=torch.uint8, device=dev)
    c = 3 * (0xd201000000010000 ** 2 - 1)
    try:
        for mode in (0, 1):
            E.set_verify_hash(mode)
            E.dev_call("tcb_verifier_hash_g2_batch_dev", st, ("size", n), d_msg.data_ptr(), d_off.data_ptr(), d_h.data_ptr())
            torch.cuda.synchronize()
            h = d_h.cpu().numpy().reshape(n, 192)
            want = O.hash_g2_batch(msgs)
            if mode == 0:
                want = O.sign_g2_batch(np.tile(fr_bytes([c]), n), want)          # [c] H(m)
            assert np.array_equal(h, want), mode
            gen = np.tile(E.verifier_generator(), (n, 1))
            assert np.array_equal(E.verify_g2_batch(pk, h, gen, sig), exp), mode
            assert np.array_equal(E.verify_batch(pk, sig, msgs), exp), mode
    finally:
        E.set_verify_hash(0)


def _poly_shares(coeff_bytes, xs_ints):
    """host big-int Horner: f(x) for every x (canonical 32-byte little-endian scalars)"""
    from conftest import R
    cs = [int.from_bytes(bytes(coeff_bytes[32 * k:32 * k + 32]), "little") for k in range(coeff_bytes.size // 32)]
    out = []
    for x in xs_ints:
        acc = 0
        for c in reversed(cs):
            acc = (acc * x + c) % R
        out.append(acc)
    return out


def test_config3_combine_full_size(gpu_engine, O):
    """BASELINE config #3 at full size: combine_signatures t = 10 over 2^14 distinct messages (per-item signer subsets out of 32):
    every item against the interpolation identity (the master signature), 512 items spread over the batch bit-exact against the
    oracle's interpolate (src/lib.rs:719-767)."""
    E = gpu_engine
    n, t = 1 << 14, 10
    m = t + 1
    rng = np.random.default_rng(33)
    poly = rand_fr(rng, m)
    idx = np.stack([np.sort(rng.choice(32, size=m, replace=False)) for _ in range(n)])
    sk32 = fr_bytes(_poly_shares(poly, [j + 1 for j in range(32)])).reshape(32, 32)
    hm = E.hash_g2_batch([b"cfg3/%d" % i for i in range(n)])
    shares = E.sign_g2_batch(sk32[idx.reshape(-1)].reshape(-1), np.repeat(hm, m, axis=0))
    xs = fr_bytes([int(j) + 1 for j in idx.reshape(-1)])
    out, st = E.combine_g2_batch(n, t, xs, shares)
    assert not st.any() and np.array_equal(out, E.sign_g2_batch(np.tile(poly[:32], n), hm))
    sel = np.sort(rng.choice(n, 512, replace=False))
    O.set_threads(16)
    xs_s = xs.reshape(n, m * 32)[sel].reshape(-1)
    sh_s = shares.reshape(n, m, 192)[sel].reshape(-1, 192)
    oo, ost = O.combine_g2_batch(len(sel), t, xs_s, sh_s)
    O.set_threads(1)
    assert np.array_equal(out[sel], oo) and not ost.any()


def test_config4_decrypt_full_size(gpu_engine, O):
    """BASELINE config #4: threshold decrypt t = 64 over 2^12 ciphertexts through tcb_decrypt_batch (src/lib.rs:618-626, test
    :908-939): every plaintext must come back (encrypt -> 65 decryption shares -> decrypt identity) and 256 items are bit-exact
    against the oracle's decrypt (interpolate::<G1> + xor_with_hash)."""
    E = gpu_engine
    n, t = 1 << 12, 64
    m = t + 1
    rng = np.random.default_rng(44)
    poly = rand_fr(rng, m)
    pkm = E.g1_mul_gen_batch(poly[:32])
    rs = rand_fr(rng, n)
    plains = [bytes(rng.integers(0, 256, size=64, dtype=np.uint8)) for _ in range(n)]
    u, v, w = E.encrypt_batch(np.tile(pkm[0], (n, 1)), rs, plains)
    assert E.verify_g2_batch(u, E.hash_g1_g2_batch(u, v), None, w).all()          # Ciphertext::verify on every item
    sk_shares = fr_bytes(_poly_shares(poly, [j + 1 for j in range(m)]))
    dshares = E.decrypt_share_batch(np.tile(sk_shares, n), np.repeat(u, m, axis=0))
    xs = np.tile(fr_bytes([j + 1 for j in range(m)]), n)
    dec, st = E.decrypt_batch(n, t, xs, dshares, v)
    assert dec == plains and not st.any()
    sel = np.sort(rng.choice(n, 256, replace=False))
    O.set_threads(16)
    odec, ost = O.decrypt_batch(len(sel), t, xs.reshape(n, m * 32)[sel].reshape(-1), dshares.reshape(n, m, 96)[sel].reshape(-1, 96), [v[i] for i in sel])
    assert np.array_equal(O.decrypt_share_batch(np.tile(sk_shares, 4), np.repeat(u[:4], m, axis=0)), dshares[:4 * m])
    O.set_threads(1)
    assert odec == [plains[i] for i in sel] and not ost.any()


def test_config5_commit_eval_full_size(gpu_engine, O):
    """BASELINE config #5: Commitment::evaluate of a degree-1023 commitment at the 2^16 indices 1..65536 (src/poly.rs:497-508,
    src/lib.rs:570-573): every output against f(x) * g1 (host Fr Horner + fixed-base multiplication), 512 points spread over the
    range bit-exact against the oracle's Horner; the 8192-point call takes the split path (coefficient blocks) and must give the
    same bytes; secondary run with 2^12 random 255-bit x."""
    from conftest import R
    E = gpu_engine
    rng = np.random.default_rng(55)
    coeff = rand_fr(rng, 1024)
    comm = E.g1_mul_gen_batch(coeff)
    n = 1 << 16
    xs = fr_bytes([i + 1 for i in range(n)])
    out = E.commitment_eval_batch(comm, xs)
    assert np.array_equal(out, E.g1_mul_gen_batch(fr_bytes(_poly_shares(coeff, [i + 1 for i in range(n)]))))
    sel = np.sort(np.concatenate([[0, 1, n - 2, n - 1], rng.choice(n, 508, replace=False)]))
    O.set_threads(16)
    assert np.array_equal(out[sel], O.commitment_eval_batch(comm, xs.reshape(n, 32)[sel].reshape(-1)))
    # small batch -> split path; forced single-unit walk gives the same bytes
    small = E.commitment_eval_batch(comm, xs[:8192 * 32])
    assert np.array_equal(small, out[:8192])
    try:
        E.set_eval_split(1)
        assert np.array_equal(E.commitment_eval_batch(comm, xs[:256 * 32]), out[:256])
        E.set_eval_split(5)         # a block count that does not divide 1024
        assert np.array_equal(E.commitment_eval_batch(comm, xs[:256 * 32]), out[:256])
    finally:
        E.set_eval_split(0)
    # random full-size x
    n2 = 1 << 12
    xr_int = [int.from_bytes(rng.bytes(40), "little") % R for _ in range(n2)]
    xr = fr_bytes(xr_int)
    out2 = E.commitment_eval_batch(comm, xr)
    assert np.array_equal(out2, E.g1_mul_gen_batch(fr_bytes(_poly_shares(coeff, xr_int))))
    assert np.array_equal(out2[:16], O.commitment_eval_batch(comm, xr[:16 * 32]))
    O.set_threads(1)


def test_fr_poly_algebra(gpu_engine, O):
    """SURVEY §8(f) row 4 on the device: Poly::evaluate / Poly * Poly kernels against the oracle and Python integers, and a
    degree-1023 polynomial at 2^14 points against the oracle's Horner."""
    cases.check_poly(gpu_engine, O)
    rng = np.random.default_rng(66)
    coeff = rand_fr(rng, 1024)
    xs = rand_fr(rng, 1 << 14)
    O.set_threads(16)
    assert np.array_equal(gpu_engine.poly_eval_batch(coeff, xs), O.poly_eval(coeff, xs))
    O.set_threads(1)
