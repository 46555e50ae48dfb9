// hostemu.cpp — TEST INFRASTRUCTURE ONLY.
// Compiles the __host__ __device__ task code of threshold_crypto_b200/csrc with g++ (scalar
// Fp2 engine, PTX carry primitives emulated bit-exactly) and exposes the same tcb_* entry
// points, so the field/curve/pairing/hash LOGIC can be checked against the oracle on a box
// without a GPU.  It is never loaded by the threshold_crypto_b200 package; the shipped
// library (libtcb200.so) has no CPU path.
#include <cstring>
#include <cstdlib>
#include <vector>
#include "../../include/tcb200.h"
#include "../../threshold_crypto_b200/csrc/scheme.cuh"
#include "../../threshold_crypto_b200/csrc/msm_plan.h"
using namespace tcb;

struct tcb_ctx { int dummy; };
static bool g_ready = false;
static void ensure() { if (!g_ready) { Consts C; build_consts(C); g_ready = true; } }

extern "C" int tcb_init(tcb_ctx **ctx, const int *, int) { ensure(); *ctx = new tcb_ctx(); return 0; }
extern "C" void tcb_free(tcb_ctx *ctx) { delete ctx; }
extern "C" const char *tcb_last_error(const tcb_ctx *) { return ""; }
extern "C" int tcb_set_engine(tcb_ctx *, int) { return 0; }
extern "C" uint64_t tcb_launch_count(const tcb_ctx *) { return 0; }

extern "C" int tcb_verify_g2_batch(tcb_ctx *, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, u8 *ok) {
    for (size_t i = 0; i < n; i++) task_verify_g2<Fp2>(i, a, b, c, d, ok);
    return 0;
}
static int g_hash_algo = 0;
extern "C" int tcb_set_hash_algo(tcb_ctx *, int a) { g_hash_algo = a; return 0; }
// the two-kernel composition of tcb200.cu: point per item, cofactor clearing, one-kernel pass over the flagged items
static void emu_hash_g2(size_t n, const u8 *msgs, const u64 *off, u8 *out, bool exact) {
    if (g_hash_algo == 1) { for (size_t i = 0; i < n; i++) task_hash_g2<Fp2>(i, msgs, off, out, exact); return; }
    std::vector<G2PointStore> pts(n + 1);
    std::vector<u8> redo(n);
    for (size_t i = 0; i < n; i++) task_hash_g2_point(i, msgs, off, &pts[i]);
    for (size_t i = 0; i < n; i++) task_g2_clear<Fp2>(i, pts.data(), out, exact, redo.data());
    for (size_t i = 0; i < n; i++) task_hash_g2<Fp2>(i, msgs, off, out, exact, redo.data());
}
extern "C" int tcb_hash_g2_batch(tcb_ctx *, size_t n, const u8 *msgs, const u64 *off, u8 *out) {
    emu_hash_g2(n, msgs, off, out, true);
    return 0;
}
extern "C" int tcb_hash_g1_g2_batch(tcb_ctx *, size_t n, const u8 *g1, const u8 *msgs, const u64 *off, u8 *out) {
    if (g_hash_algo == 1) { for (size_t i = 0; i < n; i++) task_hash_g1_g2<Fp2>(i, g1, msgs, off, out); return 0; }
    std::vector<G2PointStore> pts(n + 1);
    std::vector<u8> redo(n);
    for (size_t i = 0; i < n; i++) task_hash_g1_g2_point(i, g1, msgs, off, &pts[i]);
    for (size_t i = 0; i < n; i++) task_g2_clear<Fp2>(i, pts.data(), out, true, redo.data());
    for (size_t i = 0; i < n; i++) task_hash_g1_g2<Fp2>(i, g1, msgs, off, out, redo.data());
    return 0;
}
static bool g_verify_exact = false;
extern "C" int tcb_set_verify_hash(tcb_ctx *, int exact) { g_verify_exact = exact != 0; return 0; }
extern "C" int tcb_verifier_generator(const tcb_ctx *, u8 *out_g1) {
    ensure();
    Aff<Fp> g;
    g.x = g_verify_exact ? CONSTS().g1x : CONSTS().g1cx; g.y = g_verify_exact ? CONSTS().g1y : CONSTS().g1cy; g.inf = false;
    store_g1(out_g1, g);
    return 0;
}
extern "C" int tcb_verify_batch(tcb_ctx *, size_t n, const u8 *pk, const u8 *sig, const u8 *msgs, const u64 *off, u8 *ok) {
    // as tcb200.cu: message points (exact, or up to the unit 3 (x^2 - 1)) into scratch, then the pairing check against the
    // generator (or its multiple by the same unit)
    std::vector<u8> h(192 * n);
    emu_hash_g2(n, msgs, off, h.data(), g_verify_exact);
    for (size_t i = 0; i < n; i++) task_verify_g2<Fp2>(i, pk, h.data(), nullptr, sig, ok, !g_verify_exact);
    return 0;
}
extern "C" int tcb_sign_batch(tcb_ctx *, size_t n, const u8 *sk, const u8 *msgs, const u64 *off, u8 *out) {
    if (g_hash_algo == 1) { for (size_t i = 0; i < n; i++) task_sign<Fp2>(i, sk, msgs, off, nullptr, out); return 0; }
    std::vector<u8> h(192 * n);                       // as tcb200.cu: two-kernel hash into scratch, then the multiplication
    emu_hash_g2(n, msgs, off, h.data(), true);
    for (size_t i = 0; i < n; i++) task_sign<Fp2>(i, sk, nullptr, nullptr, h.data(), out);
    return 0;
}
extern "C" int tcb_sign_g2_batch(tcb_ctx *, size_t n, const u8 *sk, const u8 *h, u8 *out) {
    for (size_t i = 0; i < n; i++) task_sign<Fp2>(i, sk, nullptr, nullptr, h, out);
    return 0;
}
// groups per item used by the emulated multi-scalar multiplication (set by tests to cover G = 1, 2, m)
static size_t g_groups = 2;
extern "C" void tcb_emu_set_groups(size_t g) { g_groups = g ? g : 1; }
extern "C" int tcb_set_msm_groups(tcb_ctx *, size_t g) { g_groups = g ? g : 2; return 0; }
static int g_algo = 0;
extern "C" int tcb_set_msm_algo(tcb_ctx *, int a) { g_algo = a; return 0; }
// the planning functions tcb200.cu uses (msm_plan.h): groups per item and the spill factor for a batch shape
extern "C" void tcb_emu_msm_plan(size_t n, size_t m, size_t units_per_wave, int g2, size_t *G, size_t *q) {
    double fixed, share;
    tcbk::straus_costs(g2 != 0, fixed, share);
    *G = tcbk::pick_groups(n, m, units_per_wave, fixed, share);
    *q = (g2 && *G == 1) ? tcbk::pick_spill(n, m, units_per_wave, fixed, share) : 0;
}
static size_t g_eval_split = 0;
extern "C" int tcb_set_eval_split(tcb_ctx *, size_t u) { g_eval_split = u > 1 ? u : 0; return 0; }
template <class M, class JS>
static void msm_acc_ba(size_t n, size_t m, size_t G, const typename M::PS *tab, const typename M::DG *dg, JS *part) {
    size_t cnt_max = (m + G - 1) / G;
    std::vector<typename M::PS> a(n * G * ba_points_per_unit<M>(cnt_max)), b(a.size());
    std::vector<typename M::FS> pre(n * G * ba_prefix_per_unit<M>(cnt_max));
    for (size_t w = 0; w < n * G; w++) task_msm_acc_ba<M>(w, m, G, tab, dg, a.data(), b.data(), pre.data(), cnt_max, part);
}
static void g2_msm(size_t n, size_t m, const u32 *k, const u8 *pts, u8 *out, u8 *status) {
    size_t G = g_groups < m ? g_groups : m;
    std::vector<AffStore<Fp2>> tab(n * m * 8);
    std::vector<Gls4Digits> dg(n * m);
    std::vector<JacStore<Fp2>> part(n * G);
    for (size_t u = 0; u < n * m; u++) task_g2_msm_prep<Fp2>(u, k, pts, tab.data(), dg.data(), status, m);
    if (g_algo == 6 && m >= 2) {            // spill layout (forced): main units + units that take the last share of q = 2 items
        const size_t q = 2, units = n + (n + q - 1) / q;
        part.assign(n * 2, JacStore<Fp2>());
        for (size_t w = 0; w < units; w++) task_g2_msm_acc_spill<Fp2>(w, n, m, q, tab.data(), dg.data(), part.data());
        for (size_t i = 0; i < n; i++) task_g2_sum<Fp2>(i, 2, part.data(), out);
        return;
    }
    if (g_algo == 1) msm_acc_ba<MsmG2<Fp2>>(n, m, G, tab.data(), dg.data(), part.data());
    else for (size_t w = 0; w < n * G; w++) task_g2_msm_acc<Fp2>(w, m, G, tab.data(), dg.data(), part.data());
    for (size_t i = 0; i < n; i++) task_g2_sum<Fp2>(i, G, part.data(), out);
}
static void g1_msm(size_t n, size_t m, const u32 *k, const u8 *pts, Jac1Store *part, size_t G, u8 *status) {
    std::vector<Aff1Store> tab(n * m * 8);
    std::vector<Glv2Digits> dg(n * m);
    for (size_t u = 0; u < n * m; u++) task_g1_msm_prep(u, k, pts, tab.data(), dg.data(), status, m);
    if (g_algo == 1) msm_acc_ba<MsmG1>(n, m, G, tab.data(), dg.data(), part);
    else for (size_t w = 0; w < n * G; w++) task_g1_msm_acc(w, m, G, tab.data(), dg.data(), part);
}
extern "C" int tcb_combine_g2_batch(tcb_ctx *, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    memset(status, 0, n);
    if (t == 0) { memcpy(out, shares, n * 192); return 0; }
    size_t m = t + 1;
    std::vector<u32> lam(n * m * 8);
    for (size_t i = 0; i < n; i++)
        { std::vector<LagrangeND> nd(m);
          for (size_t k = 0; k < m; k++) lagrange_num_den(x + i * m * 32, m, k, nd[k], status[i]);
          lagrange_finish_item(nd.data(), m, &lam[8 * i * m]); }
    g2_msm(n, m, lam.data(), shares, out, status);
    return 0;
}
static int combine_g1(size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status, int mode, const u8 *v, const u64 *voff) {
    memset(status, 0, n);
    size_t m = t + 1;
    std::vector<u32> lam(n * m * 8);
    size_t G = g_groups < m ? g_groups : m;
    std::vector<Jac1Store> terms(n * G);
    if (t > 0) {
        for (size_t i = 0; i < n; i++)
            for (size_t k = 0; k < m; k++) lagrange_coeff(x + i * m * 32, m, k, &lam[8 * (i * m + k)], status[i]);
        g1_msm(n, m, lam.data(), shares, terms.data(), G, status);
    }
    for (size_t i = 0; i < n; i++) {
        Aff<Fp> g;
        bool ok = true;
        if (t > 0) g = g1_sum(i, G, terms.data()); else g = load_g1(shares + 96 * i, ok);
        if (mode == 0) store_g1(out + 96 * i, g);
        else xor_with_hash(out + voff[i], g, v + voff[i], (size_t)(voff[i + 1] - voff[i]));
    }
    return 0;
}
extern "C" int tcb_combine_g1_batch(tcb_ctx *, size_t n, size_t t, const u8 *x, const u8 *shares, u8 *out, u8 *status) {
    return combine_g1(n, t, x, shares, out, status, 0, nullptr, nullptr);
}
extern "C" int tcb_decrypt_batch(tcb_ctx *, size_t n, size_t t, const u8 *x, const u8 *shares, const u8 *v, const u64 *voff, u8 *out, u8 *status) {
    return combine_g1(n, t, x, shares, out, status, 1, v, voff);
}
extern "C" int tcb_decrypt_share_batch(tcb_ctx *, size_t n, const u8 *sk, const u8 *u, u8 *out) {
    for (size_t i = 0; i < n; i++) task_g1_mul(i, sk, u, out);
    return 0;
}
extern "C" int tcb_g1_mul_gen_batch(tcb_ctx *, size_t n, const u8 *sk, u8 *out) {
    for (size_t i = 0; i < n; i++) task_g1_mul(i, sk, nullptr, out);
    return 0;
}
extern "C" int tcb_commitment_eval_batch(tcb_ctx *, size_t deg, const u8 *coeff, size_t n, const u8 *x, u8 *out) {
    std::vector<Aff1Store> tab(deg + 1);
    for (size_t c = 0; c <= deg; c++) task_g1_decode(c, coeff, tab.data());
    if (g_eval_split > 1) {       // the split evaluation of small batches (coefficient blocks + x^(b L) recombination + sum)
        size_t B = g_eval_split, L = (deg + B) / B;
        std::vector<Jac1Store> terms(n * B);
        for (size_t u = 0; u < n * B; u++) task_commit_eval_part(u, B, L, deg, tab.data(), x, terms.data());
        for (size_t i = 0; i < n; i++) store_g1(out + 96 * i, g1_sum(i, B, terms.data()));
        return 0;
    }
    for (size_t i = 0; i < n; i++) task_commit_eval(i, deg, tab.data(), x, out);
    return 0;
}

extern "C" int tcb_poly_eval_batch(tcb_ctx *, size_t deg, const u8 *coeff, size_t n, const u8 *x, u8 *out) {
    std::vector<Fr> cm(deg + 1);
    u8 bad = 0;
    for (size_t c = 0; c <= deg; c++) task_fr_to_mont(c, coeff, cm.data(), &bad);
    for (size_t i = 0; i < n; i++) task_poly_eval(i, deg, cm.data(), x, out, &bad);
    return bad ? -10 : 0;
}
extern "C" int tcb_poly_mul_batch(tcb_ctx *, size_t n, size_t da, const u8 *a, size_t db, const u8 *b, u8 *out) {
    std::vector<Fr> am(n * (da + 1)), bm(n * (db + 1));
    u8 bad = 0;
    for (size_t i = 0; i < am.size(); i++) task_fr_to_mont(i, a, am.data(), &bad);
    for (size_t i = 0; i < bm.size(); i++) task_fr_to_mont(i, b, bm.data(), &bad);
    for (size_t u = 0; u < n * (da + db + 1); u++) task_poly_mul(u, da, db, am.data(), bm.data(), out);
    return bad ? -10 : 0;
}
extern "C" int tcb_g1_lincomb_batch(tcb_ctx *, size_t n, size_t m, const u8 *sc, const u8 *pts, u8 *out) {
    size_t G = g_groups < m ? g_groups : m;
    std::vector<Jac1Store> terms(n * G);
    std::vector<u8> st(n);
    g1_msm(n, m, (const u32 *)sc, pts, terms.data(), G, st.data());
    for (size_t i = 0; i < n; i++) store_g1(out + 96 * i, g1_sum(i, G, terms.data()));
    return 0;
}
extern "C" int tcb_g2_lincomb_batch(tcb_ctx *, size_t n, size_t m, const u8 *sc, const u8 *pts, u8 *out) {
    std::vector<u8> st(n);
    g2_msm(n, m, (const u32 *)sc, pts, out, st.data());
    return 0;
}
extern "C" int tcb_encrypt_batch(tcb_ctx *, size_t n, const u8 *pk, const u8 *r, const u8 *msgs, const u64 *off, u8 *u_out, u8 *v_out, u8 *w_out) {
    std::vector<u8> h(192 * n);
    for (size_t i = 0; i < n; i++) task_encrypt_uv(i, pk, r, msgs, off, u_out, v_out);
    for (size_t i = 0; i < n; i++) task_hash_g1_g2<Fp2>(i, u_out, v_out, off, h.data());
    for (size_t i = 0; i < n; i++) task_sign<Fp2>(i, r, nullptr, nullptr, h.data(), w_out);
    return 0;
}
extern "C" int tcb_g1_compress_batch(tcb_ctx *, size_t n, const u8 *unc, u8 *out) { for (size_t i = 0; i < n; i++) task_g1_compress(i, unc, out); return 0; }
extern "C" int tcb_g2_compress_batch(tcb_ctx *, size_t n, const u8 *unc, u8 *out) { for (size_t i = 0; i < n; i++) task_g2_compress<Fp2>(i, unc, out); return 0; }
extern "C" int tcb_g1_decompress_batch(tcb_ctx *, size_t n, const u8 *in, u8 *out, u8 *st) { for (size_t i = 0; i < n; i++) task_g1_decompress(i, in, out, st); return 0; }
extern "C" int tcb_g2_decompress_batch(tcb_ctx *, size_t n, const u8 *in, u8 *out, u8 *st) { for (size_t i = 0; i < n; i++) task_g2_decompress<Fp2>(i, in, out, st); return 0; }
extern "C" uint64_t tcb_emu_mac_count(int reset) { uint64_t v = g_mac_count; if (reset) g_mac_count = 0; return v; }

// test hook: number of mismatches between the binary-GCD inverse and the Fermat inverse on n
// pseudo-random elements plus the edge values 0, 1, p-1
extern "C" int tcb_emu_inv_check(int n, uint64_t seed) {
    ensure();
    int bad = 0;
    auto check = [&](const Fp &a, bool fermat) {
        Fp x = fp_inv(a);                          // batched-divstep binary GCD (production)
        if (x != fp_inv_basic(a)) bad++;           // limb-by-limb binary GCD
        if (fermat && x != fp_inv_fermat(a)) bad++;
        if (!a.is_zero() && (x * a) != fp_one()) bad++;
    };
    for (int i = 0; i < n + 3; i++) {
        Fp a;
        for (;;) {
            for (int k = 0; k < 12; k++) { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; a.l[k] = (u32)(seed >> 32); }
            a.l[11] &= 0x1fffffffu;
            if (limbs_lt_mod<FpParams>(a.l)) break;
        }
        if (i == n) a = Fp::zero();
        if (i == n + 1) a = fp_one();
        if (i == n + 2) a = -fp_one();
        check(a, i < 64 || i >= n);
    }
    // every bit length (the 64-bit approximations of the batched version slide with the top bit) and +-2^k
    for (int bits = 1; bits <= 381; bits++) {
        Fp a = Fp::zero();
        for (int k = 0; k < 12; k++) { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; a.l[k] = (u32)(seed >> 32); }
        int top = (bits - 1) / 32;
        for (int k = top + 1; k < 12; k++) a.l[k] = 0;
        a.l[top] &= 0xffffffffu >> (31 - ((bits - 1) % 32));
        a.l[top] |= 1u << ((bits - 1) % 32);
        if (limbs_lt_mod<FpParams>(a.l)) check(a, false);
        if (bits <= 380) { Fp p2 = Fp::zero(); p2.l[(bits - 1) / 32] = 1u << ((bits - 1) % 32); check(p2, false); check(-p2, false); }
    }
    return bad;
}

// test hook: Karabina compressed-squaring exponentiation (tower.cuh, the scalar statement of quad.cuh's device code) against the
// Granger-Scott loop on n random elements of the cyclotomic subgroup, for |x| and |x| >> 1, plus f = 1 (fallback path)
extern "C" int tcb_emu_karabina_check(int n, uint64_t seed) {
    ensure();
    int bad = 0;
    for (int i = 0; i < n + 1; i++) {
        Fp12T<Fp2> a;
        Fp *c = (Fp *)&a;
        for (int k = 0; k < 12; k++) {
            for (;;) {
                for (int w = 0; w < 12; w++) { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; c[k].l[w] = (u32)(seed >> 32); }
                c[k].l[11] &= 0x1fffffffu;
                if (limbs_lt_mod<FpParams>(c[k].l)) break;
            }
        }
        Fp12T<Fp2> r = fp12_mul(fp12_conj(a), fp12_inv(a));     // easy part: a^((p^6 - 1)(p^2 + 1))
        r = fp12_mul(fp12_frob(r, 2), r);
        if (i == n) r = fp12_one<Fp2>();
        const u64 x = TCB_BLS_X;
        for (u64 e : {x, x >> 1}) {
            Fp12T<Fp2> g = fp12_exp_by_x(r, e), k = fp12_exp_by_x_karabina(r, e);
            const Fp *pg = (const Fp *)&g, *pk = (const Fp *)&k;
            for (int w = 0; w < 12; w++) if (pg[w] != pk[w]) { bad++; break; }
        }
    }
    return bad;
}
// K-term dot product with one interleaved reduction (fp.cuh, mont_dotk: the multiplier core of the shared-memory pairing
// engine) against the sum of portable single products; operands include 0, 1, p - 1 and the non-canonical value p (what a
// conditional negation of 0 produces).
template <int K>
static int dotk_check_one(uint64_t &seed, int mode) {
    Fp x[K], y[K];
    Fp ref = Fp::zero();
    for (int k = 0; k < K; k++) {
        for (int side = 0; side < 2; side++) {
            Fp &v = side ? y[k] : x[k];
            for (;;) {
                for (int w = 0; w < 12; w++) { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; v.l[w] = (u32)(seed >> 32); }
                v.l[11] &= 0x1fffffffu;
                if (limbs_lt_mod<FpParams>(v.l)) break;
            }
            int sel = (mode + 3 * k + side) % 7;
            if (mode && sel == 0) v = Fp::zero();
            if (mode && sel == 1) { v = Fp::zero(); v.l[0] = 1; }
            if (mode && sel == 2) { for (int w = 0; w < 12; w++) v.l[w] = FpParams::mod(w); v.l[0] -= 1; }
            if (mode == 2) { for (int w = 0; w < 12; w++) v.l[w] = FpParams::mod(w); if (side) v.l[0] -= 1; }   // x = p (non-canonical), y = p - 1: the largest admissible sum
        }
        Fp xc = x[k];
        if (!limbs_lt_mod<FpParams>(xc.l)) xc = Fp::zero();       // p == 0
        ref = ref + mont_mul_portable<FpParams>(xc, y[k]);
    }
    Fp got = mont_dotk<FpParams, K>(x, y);
    return (got != ref || !limbs_lt_mod<FpParams>(got.l)) ? 1 : 0;
}
extern "C" int tcb_emu_dotk_check(int n, uint64_t seed) {
    ensure();
    int bad = 0;
    for (int i = 0; i < n; i++) {
        int mode = i < 16 ? (i % 3) : 0;
        bad += dotk_check_one<1>(seed, mode) + dotk_check_one<2>(seed, mode) + dotk_check_one<3>(seed, mode) + dotk_check_one<4>(seed, mode) +
               dotk_check_one<6>(seed, mode) + dotk_check_one<8>(seed, mode);
    }
    return bad;
}
extern "C" int tcb_emu_issquare_check(int n, uint64_t seed) {
    ensure();
    int bad = 0, squares = 0;
    for (int i = 0; i < n + 3; i++) {
        Fp a;
        for (;;) {
            for (int k = 0; k < 12; k++) { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; a.l[k] = (u32)(seed >> 32); }
            a.l[11] &= 0x1fffffffu;
            if (limbs_lt_mod<FpParams>(a.l)) break;
        }
        if (i == n) a = Fp::zero();
        if (i == n + 1) a = fp_one();
        if (i == n + 2) a = -fp_one();          // -1 is a non-residue (p = 3 mod 4)
        Fp s = fp_pow<ExpPp1d4>(a);
        bool want = sqr(s) == a;
        if (fp_is_square(a) != want || fp_is_square_basic(a) != want) bad++;
        squares += want;
    }
    // values with long runs of trailing zeros (the production version removes up to 31 per iteration): +-2^k, 2^k * odd, low limbs zero
    for (int k = 0; k < 381; k++) {
        Fp pw = Fp::zero();
        pw.l[k >> 5] = 1u << (k & 31);
        if (!limbs_lt_mod<FpParams>(pw.l)) continue;
        Fp cand[3] = {pw, -pw, pw};
        for (int j = 0; j < 12; j++) if (j > (k >> 5)) { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; cand[2].l[j] = (u32)(seed >> 32); }
        cand[2].l[11] &= 0x0fffffffu;
        for (int c = 0; c < 3; c++) {
            Fp s = fp_pow<ExpPp1d4>(cand[c]);
            bool want = sqr(s) == cand[c];
            if (fp_is_square(cand[c]) != want || fp_is_square_basic(cand[c]) != want) bad++;
        }
    }
    return bad * 100000 + squares;
}
