"""CPU: the device task code (threshold_crypto_b200/csrc/*.cuh) compiled for the host
(tests/hostemu, scalar engine, emulated carry flag) against the oracle and the golden vectors.
This checks the LOGIC of what the kernels run; the kernels themselves are checked on the GPU
in test_gpu_parity.py."""
import numpy as np

import cases
from conftest import fr_bytes, hx, hxs


def test_hostemu_all_entry_points(emu, O):
    cases.check_all(emu, O, n_sig=5, n_comb=2, t=2, deg=4, n_eval=4, seed=3)


def test_hostemu_edges(emu, O):
    cases.check_edges(emu, O)


def test_hostemu_golden(emu, golden):
    msgs = [bytes.fromhex(m) for m in golden["msgs"]]
    assert [bytes(p).hex() for p in emu.hash_g2_batch(msgs)] == golden["hash_g2"]
    sk = np.concatenate([hx(s) for s in golden["sk"]])
    assert [bytes(p).hex() for p in emu.g1_mul_gen_batch(sk)] == golden["pk"]
    for pi, si, mi, exp in golden["verify_cases"]:
        assert bool(emu.verify_batch(hx(golden["pk"][pi]), hx(golden["sig"][si]), [msgs[mi]])[0]) == exp
    ts = golden["threshold_sig"]
    for s in ts["sets"]:
        out, st = emu.combine_g2_batch(1, ts["t"], fr_bytes([i + 1 for i in s["idx"]]), hxs(s["shares"]))
        assert bytes(out[0]).hex() == s["combined"]
    e = golden["enc"]
    out, st = emu.decrypt_batch(1, 2, fr_bytes([i + 1 for i in e["idx"]]), hxs(e["dshares"]), [bytes.fromhex(e["v"])])
    assert out[0].hex() == e["plain"]
    ce = golden["commit_eval"]
    out = emu.commitment_eval_batch(hxs(ce["coeff"]), np.concatenate([hx(x) for x in ce["x"]]))
    assert [bytes(p).hex() for p in out] == ce["out"]


def test_hostemu_msm_groupings(emu, O):
    """Shared-doubling MSM with 1, 2 and m groups per item (m groups = one share per unit)."""
    try:
        for algo in (0, 1, 6):       # Straus with mixed additions, batch-affine tree, Straus in the spill layout
            emu.set_msm_algo(algo)
            for g in (1, 2, 64):
                emu.set_msm_groups(g)
                cases.check_msm(emu, O, n=3, m=5, seed=21 + g)
                x, s, master = cases.make_combine_batch(O, 2, 3, 40 + g, group=2)
                out, st = emu.combine_g2_batch(2, 3, x, s)
                assert np.array_equal(out, master) and not st.any()
    finally:
        emu.set_msm_groups(0)
        emu.set_msm_algo(0)


def test_msm_plan_for_the_baseline_shapes(emu):
    """Host-side planning of the multi-scalar multiplication (csrc/msm_plan.h, the code tcb200.cu runs): groups per item and the
    spill factor for the BASELINE shapes on a B200 (148 SMs; G2 accumulation 2 blocks x 64 lane pairs per SM, G1 3 blocks x 128
    threads): config #3 (2^14 items, 11 shares) runs one unit per item with the last share spilled to the spare units, 7 items
    each; one eighth of it (what a GPU gets in the 8-way sharded record) is split into 6 partial sums; config #4 (2^12 items,
    65 shares) into 13."""
    import ctypes as C
    lib = emu.lib

    def plan(n, m, wave, g2):
        G, q = C.c_size_t(0), C.c_size_t(0)
        lib.tcb_emu_msm_plan(C.c_size_t(n), C.c_size_t(m), C.c_size_t(wave), int(g2), C.byref(G), C.byref(q))
        return G.value, q.value
    w2, w1 = 2 * 64 * 148, 3 * 128 * 148
    assert plan(1 << 14, 11, w2, True) == (1, 7)
    assert plan(2048, 11, w2, True) == (6, 0)
    assert plan(w2, 11, w2, True) == (1, 0)                  # a full wave: nothing to spill to
    assert plan(9000, 11, w2, True) == (2, 0)                # two units per item still fit one wave
    assert plan(1 << 12, 65, w1, False) == (13, 0)
    assert plan(1 << 14, 2, w2, True)[1] == 0                # too few shares for the layout to pay
    for n in (1, 100, 5000, 1 << 14, 1 << 18):
        for m in (1, 2, 11, 65):
            G, q = plan(n, m, w2, True)
            assert 1 <= G <= m and (q == 0 or (G == 1 and n + (n + q - 1) // q <= w2))


def test_binary_gcd_inverse_and_legendre_symbol(emu):
    """fp_inv (branch-free binary GCD) == Fermat inverse, fp_is_square (binary Jacobi) == Euler
    criterion, on random elements, every bit length, +-2^k and the edge values 0, 1, p-1 (host instantiation of tower.cuh);
    the production inverse (batched divsteps) is also compared with the limb-by-limb binary GCD."""
    import ctypes as C
    lib = emu.lib
    assert lib.tcb_emu_inv_check(500, C.c_uint64(99)) == 0
    r = lib.tcb_emu_issquare_check(1000, C.c_uint64(98))
    assert r // 100000 == 0 and 400 < r % 100000 < 620


def test_scalar_mul_special_scalars(emu, O):
    """GLV2 (width-2 window) / GLS4 regular recodings on the scalars that stress them: 0, 1, r-1, multiples and neighbours of
    X^2 and X (the split points), all-ones halves, plus random ones — against the oracle's double-and-add."""
    from conftest import R
    rng = np.random.default_rng(123)
    X = 0xd201000000010000
    X2 = X * X
    ks = [0, 1, 2, 3, R - 1, R - 2, X2, X2 - 1, X2 + 1, 2 * X2, (1 << 128) - 1, 1 << 128, (1 << 129) + 1, X, X - 1, X + 1,
          X ** 3 % R, (X ** 3 + 1) % R, (R - 1) // 2, (R + 1) // 2]
    ks += [(a + b * X2) % R for a in (0, 1, (1 << 127) - 1, (1 << 128) - 1) for b in (1, (1 << 126) + 5, (1 << 127) - 1)]
    ks += [int.from_bytes(rng.bytes(40), "little") % R for _ in range(24)]
    n = len(ks)
    sk = fr_bytes(ks)
    base1 = O.g1_mul_gen_batch(fr_bytes([int.from_bytes(rng.bytes(40), "little") % R for _ in range(n)]))
    assert np.array_equal(emu.decrypt_share_batch(sk, base1), O.decrypt_share_batch(sk, base1))
    base2 = O.sign_g2_batch(fr_bytes([int.from_bytes(rng.bytes(40), "little") % R for _ in range(n)]), np.tile(O.g2_generator(), (n, 1)))
    assert np.array_equal(emu.sign_g2_batch(sk, base2), O.sign_g2_batch(sk, base2))


def test_karabina_compressed_squarings(emu):
    """The compressed-squaring x-power (what quad.cuh runs on the device) equals the Granger-Scott loop on random elements of
    the cyclotomic subgroup and on 1 (fallback when a saved z2 is zero)."""
    import ctypes as C
    assert emu.lib.tcb_emu_karabina_check(6, C.c_uint64(5)) == 0


def test_hostemu_codecs(emu, O, golden):
    cases.check_codecs(emu, O, golden)


def test_dot_product_core(emu):
    """mont_dotk (fp.cuh): K-term dot product with one interleaved Montgomery reduction, the multiplier core of the
    shared-memory pairing engine, for K = 1, 2, 3, 4, 6, 8 against sums of portable products (incl. 0, 1, p - 1 and the
    non-canonical operand p)."""
    import ctypes as C
    assert emu.lib.tcb_emu_dotk_check(300, C.c_uint64(17)) == 0


def test_commit_eval_split_blocks(emu, O):
    """Commitment::evaluate with B units per point (coefficient blocks, recombined with x^(b L)): same bytes as the oracle's Horner
    for B = 2, 3, 4 incl. blocks past the end (deg + 1 = 7 not divisible), x = 0 and a full-size x."""
    import conftest
    rng = np.random.default_rng(12)
    coeff = conftest.rand_fr(rng, 7)
    comm = O.g1_mul_gen_batch(coeff)
    xs = fr_bytes([0, 1, 2, 65536, int.from_bytes(rng.bytes(40), "little")])
    exp = O.commitment_eval_batch(comm, xs)
    try:
        for B in (2, 3, 4, 8):
            emu.set_eval_split(B)
            assert np.array_equal(emu.commitment_eval_batch(comm, xs), exp), B
    finally:
        emu.set_eval_split(0)


def test_hostemu_poly_algebra(emu, O):
    cases.check_poly(emu, O)
