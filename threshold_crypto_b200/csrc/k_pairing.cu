// k_pairing.cu — the pairing-equality kernel (quad engine, quad.cuh) and the on-device self-tests.
#if defined(TCB_PAIRING_FP_NOINLINE)   // experiment build: Fp multiply / dot2 as functions in the pairing kernel
#define TCB_FP_NOINLINE 1
#endif
#include "kern.h"
#include "quad.cuh"
#include "quadsm.cuh"
using namespace tcb;

// one item per QUAD of lanes (quad.cuh); h_g2 replaces b_g2 when the hash was computed on device
#ifndef TCB_QUAD_MINB
#define TCB_QUAD_MINB 2
#endif
__global__ void __launch_bounds__(128, TCB_QUAD_MINB) k_verify_g2_quad(size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, u8 *ok) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    bool live = i < n;
    if (!live) i = n - 1;          // tail lanes recompute the last item: no early exit (block-wide phase barriers)
    bool enc_ok;
    bool res = pairing_eq_quad(a + 96 * i, b + 192 * i, c ? c + 96 * i : nullptr, d + 192 * i, enc_ok);
    if (live && (threadIdx.x & 3) == 0) ok[i] = (res && enc_ok) ? 1 : 0;
}
// (Measured and dropped, profiles/kbench_r2i_fecells.json: keeping this kernel's saved compressed powers and prefix products in
// shared-memory cells instead of its local frame.  The 110 KB carve-out leaves ~28 KB of L1 for the remaining 4 KB frames per
// thread: 37.8 instead of 32.2 ms, DRAM write-back unchanged at 4.1 GB.  The register engine lives off the L1.)
// Final exponentiation + "== 1" of the Miller-loop values produced by k_miller_quad (k_miller.cu) or k_miller_quad_reg:
// f of item i, lane l, coefficient k at fin[(4 i + l) * 3 + k]  (576 B per item through HBM/L2).
#ifndef TCB_FE_MINB
#define TCB_FE_MINB TCB_QUAD_MINB
#endif
__global__ void __launch_bounds__(128, TCB_FE_MINB) k_final_exp_quad(size_t n, const Fp *fin, const u8 *enc_ok, u8 *ok, Fp *fe_out) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    bool live = i < n;
    if (!live) i = n - 1;
    const Fp *p = fin + (i * 4 + (threadIdx.x & 3u)) * 3;
    Fp12Q f;
    f.h.c0.h = ldg_fp(p); f.h.c1.h = ldg_fp(p + 1); f.h.c2.h = ldg_fp(p + 2);
    Fp12Q g = final_exponentiation(fp12_conj(f));
    bool res = fp12_is_one(g);
    if (live && fe_out) { Fp *o = fe_out + (i * 4 + (threadIdx.x & 3u)) * 3; stg_fp(o, g.h.c0.h); stg_fp(o + 1, g.h.c1.h); stg_fp(o + 2, g.h.c2.h); }
    if (live && (threadIdx.x & 3) == 0) ok[i] = (res && enc_ok[i]) ? 1 : 0;
}
// ---- final exponentiation ENTIRELY on shared-memory cells (quadsm.cuh: qf_mul12, qf_comp_sqr + the pieces below): no Fp12 /
// Fp6 value crosses a function boundary through the local frame.  Slots: V0 (running value), V1 (second operand), six scratch.
// Same operations as quad.cuh's chain (same values, checked bit for bit by tcb_selftest_miller); the only local arrays are the
// six saved compressed powers and the prefix products of an x-run (own halves: 864 B per thread).
TCB_D Fp12Q qf_ld12(u32 v) {
    Fp12Q r;
    u32 t = q_tid();
    r.h.c0.h = q_ld(v, t); r.h.c1.h = q_ld(v + 1, t); r.h.c2.h = q_ld(v + 2, t);
    return r;
}
TCB_D void qf_st12(u32 v, const Fp12Q &a) {      // the caller hands off with __syncwarp()
    q_st(v, a.h.c0.h); q_st(v + 1, a.h.c1.h); q_st(v + 2, a.h.c2.h);
}
TCB_D void qf_copy12(u32 dst, u32 src) {         // own cells
    u32 t = q_tid();
    Fp a = q_ld(src, t), b = q_ld(src + 1, t), c = q_ld(src + 2, t);
    __syncwarp();
    q_st(dst, a); q_st(dst + 1, b); q_st(dst + 2, c);
    __syncwarp();
}
TCB_D void qf_conj_inplace(u32 v) {              // own cells only
    if (q_pair()) { u32 t = q_tid(); for (u32 k = 0; k < 3; k++) q_st(v + k, -q_ld(v + k, t)); }
    __syncwarp();
}
static __device__ __noinline__ Fp q_fp_inv(Fp a) { return fp_inv(a); }
// my half of the inverse of the Fp2 value whose my-half is `h` (exchange through cell `tmp`): conj(a) / norm(a)
TCB_D Fp qf_inv2(const Fp &h, u32 tmp) {
    Fp sq = q_fmul(h, h);
    __syncwarp();
    q_st(tmp, sq);
    __syncwarp();
    Fp n = q_fp_inv(sq + q_ld(tmp, q_tid() ^ 1u));
    Fp r = q_fmul(h, n);
    return q_role() ? -r : r;
}
// xi-multiples of the coefficients 1, 2 of V[a] (my pair) into S0, S1: what q_mul3x3 streams beside a Fp6 operand
TCB_D void qf_xi12(u32 a) {
    const u32 t = q_tid();
    const bool e = q_role();
#pragma unroll
    for (int k = 0; k < 2; k++) {
        Fp own = q_ld(a + 1 + k, t), part = q_ld(a + 1 + k, t ^ 1u);
        q_st(QF_S0 + k, q_addsub(e, own, part));
    }
    __syncwarp();
}
// dst (my column) <- A * B in Fp6, A = slots a.. at re-column ac (its xi-multiples in S0, S1 at the same column), B = slots b.. at bc
TCB_D void qf_mul6(u32 dst, u32 a, u32 ac, u32 b, u32 bc) {
    q_mul3x3(q_cell(b, bc), q_cell(b + 1, bc), q_cell(b + 2, bc),
             q_cell(a, ac), q_cell(QF_S1, ac), q_cell(QF_S0, ac),
             q_cell(a + 1, ac), q_cell(a, ac), q_cell(QF_S1, ac),
             q_cell(a + 2, ac), q_cell(a + 1, ac), q_cell(a, ac), dst, dst + 1, dst + 2);
}
// V0 <- V0^-1 (uses V1 and the scratch): 1 / (c0 + c1 w) = (c0 - c1 w) / (c0^2 - v c1^2)
static __device__ __noinline__ void qf_inv12() {
    const u32 t = q_tid(), o = t ^ 2u, me = q_col_re(q_pair());
    const bool e = q_role(), p0 = q_pair() == 0;
    qf_xi12(QF_V0);
    qf_mul6(QF_S2, QF_V0, me, QF_V0, me);                         // s = (my half)^2 in S2..S4
    {
        Fp s0 = q_ld(QF_S2, t), s1 = q_ld(QF_S3, t), s2 = q_ld(QF_S4, t);
        Fp o0 = q_ld(QF_S2, o), o1 = q_ld(QF_S3, o), o2 = q_ld(QF_S4, o);
        Fp X = fp_select(p0, o2, s2), Xp = q_ld(QF_S4, (p0 ? o : t) ^ 1u);
        Fp xX = q_addsub(e, X, Xp);                                // xi * (other's s2 | my s2)
        Fp t0 = fp_select(p0, s0, o0) - xX;                        // c0^2 - v c1^2 on both pairs
        Fp t1 = fp_select(p0, s1 - o0, o1 - s0);
        Fp t2 = fp_select(p0, s2 - o1, o2 - s1);
        __syncwarp();
        q_st(QF_V1, t0); q_st(QF_V1 + 1, t1); q_st(QF_V1 + 2, t2);
        __syncwarp();
    }
    // Fp6 inverse of V1 (the same value on both pairs)
    const u32 c0 = q_cell(QF_V1, me), c1 = q_cell(QF_V1 + 1, me), c2 = q_cell(QF_V1 + 2, me);
    Fp sq0 = q_sqr(QF_V1), sq1 = q_sqr(QF_V1 + 1), sq2 = q_sqr(QF_V1 + 2);
    Fp m12 = q_mul2(c1, c2), m01 = q_mul2(c0, c1), m02 = q_mul2(c0, c2);
    q_st(QF_S0, m12); q_st(QF_S1, sq2);
    __syncwarp();
    Fp A0 = sq0 - q_addsub(e, m12, q_ld(QF_S0, t ^ 1u));
    Fp A1 = q_addsub(e, sq2, q_ld(QF_S1, t ^ 1u)) - m01;
    Fp A2 = sq1 - m02;
    __syncwarp();
    q_st(QF_S2, A0); q_st(QF_S3, A1); q_st(QF_S4, A2);
    __syncwarp();
    Fp u = q_mul2(c2, q_cell(QF_S3, me)) + q_mul2(c1, q_cell(QF_S4, me));
    Fp w = q_mul2(c0, q_cell(QF_S2, me));
    q_st(QF_S0, u);
    __syncwarp();
    Fp D = q_addsub(e, u, q_ld(QF_S0, t ^ 1u)) + w;
    Fp dinv = qf_inv2(D, QF_S1);
    __syncwarp();
    q_st(QF_S0, dinv);
    __syncwarp();
    const u32 dc = q_cell(QF_S0, me);
    Fp i0 = q_mul2(q_cell(QF_S2, me), dc), i1 = q_mul2(q_cell(QF_S3, me), dc), i2 = q_mul2(q_cell(QF_S4, me), dc);
    __syncwarp();
    q_st(QF_V1, i0); q_st(QF_V1 + 1, i1); q_st(QF_V1 + 2, i2);
    __syncwarp();
    qf_xi12(QF_V1);
    qf_mul6(QF_V0, QF_V1, me, QF_V0, me);                         // my half * (c0^2 - v c1^2)^-1
    qf_conj_inplace(QF_V0);
}
// V[v] <- V[v]^(p^k) in place: coefficient v^i w^j is conjugated k times and multiplied by frob[k][2 i + j]
static __device__ __noinline__ void qf_frob(u32 v, int k) {
    const Consts &C = CONSTS();
    const u32 t = q_tid(), me = q_col_re(q_pair()), j = q_pair();
    const bool e = q_role(), odd = k & 1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Fp x = q_ld(v + i, t);
        if (odd && e) x = -x;
        const Fp2c &c = C.frob[k][2 * i + j];
        q_st(v + i, x);
        q_st(QF_S0 + i, e ? c.c1 : c.c0);
    }
    __syncwarp();
    Fp r0 = q_mul2(q_cell(v, me), q_cell(QF_S0, me)), r1 = q_mul2(q_cell(v + 1, me), q_cell(QF_S1, me)), r2 = q_mul2(q_cell(v + 2, me), q_cell(QF_S2, me));
    __syncwarp();
    q_st(v, r0); q_st(v + 1, r1); q_st(v + 2, r2);
    __syncwarp();
}
// V0 <- product of the decompressed saved powers; false (warp-uniform result is formed by the caller) if a z2 is zero.
// z1 = (xi z5^2 + 3 z4^2 - 2 z3) / (4 z2),  z0 = (2 z1^2 + z2 z5 - 3 z3 z4) xi + 1; one shared inversion (Montgomery's trick).
static __device__ __noinline__ bool qf_decompress_product(const Fp *sa, const Fp *sb, int n) {
    const u32 t = q_tid(), o = t ^ 2u, me = q_col_re(q_pair());
    const bool e = q_role(), p0 = q_pair() == 0;
    Fp pre[COMP_MAX];
    Fp one = e ? Fp::zero() : fp_one();
    bool bad = false;
    q_st(QF_S2, one);                                              // running product of the denominators
    __syncwarp();
    for (int k = 0; k < n; k++) {
        Fp den = dbl(dbl(sa[k]));                                  // pair 1: 4 z2
        bool z = pair_and(den.is_zero());
        bad = bad || (!p0 && z);
        pre[k] = q_ld(QF_S2, t);
        q_st(QF_S3, den);
        __syncwarp();
        Fp run = q_mul2(q_cell(QF_S2, me), q_cell(QF_S3, me));
        __syncwarp();
        q_st(QF_S2, run);
        __syncwarp();
    }
    if (!q_quad_and(!bad)) bad = true;                             // quad-uniform
    if (__any_sync(0xffffffffu, bad)) return false;                // warp-uniform: the caller redoes the whole warp the plain way
    Fp rinv = qf_inv2(q_ld(QF_S2, t), QF_S3);
    for (int k = n - 1; k >= 0; k--) {
        __syncwarp();
        q_st(QF_S2, rinv); q_st(QF_S3, pre[k]); q_st(QF_S0, dbl(dbl(sa[k]))); q_st(QF_CA, sa[k]); q_st(QF_CB, sb[k]);
        __syncwarp();
        Fp dinv = q_mul2(q_cell(QF_S2, me), q_cell(QF_S3, me));
        rinv = q_mul2(q_cell(QF_S2, me), q_cell(QF_S0, me));
        Fp s = q_sqr(p0 ? QF_CA : QF_CB);                          // pair 0: z4^2 ; pair 1: z5^2
        Fp m = q_mul2(q_cell(QF_CA, me), q_cell(QF_CB, me));       // pair 0: z4 z3 ; pair 1: z2 z5
        __syncwarp();
        q_st(QF_S3, dinv); q_st(QF_S1, s);
        __syncwarp();
        Fp os = q_ld(QF_S1, o);
        Fp num = q_addsub(e, s, q_ld(QF_S1, t ^ 1u)) + (dbl(os) + os) - dbl(q_ld(QF_CB, o));     // pair 1: xi z5^2 + 3 z4^2 - 2 z3
        q_st(QF_S0, num);
        __syncwarp();
        Fp z1 = q_mul2(q_cell(QF_S0, me), q_cell(QF_S3, me));
        __syncwarp();
        q_st(QF_S1, z1);
        __syncwarp();
        Fp w = dbl(q_sqr(QF_S1)) + m;                              // pair 1: 2 z1^2 + z2 z5
        q_st(QF_S0, w);
        __syncwarp();
        Fp u = q_ld(QF_S0, o) - (dbl(m) + m);                      // pair 0: 2 z1^2 + z2 z5 - 3 z3 z4
        q_st(QF_S3, u);
        __syncwarp();
        Fp z0 = q_addsub(e, u, q_ld(QF_S3, t ^ 1u)) + one;
        const u32 dst = (k == n - 1) ? QF_V0 : QF_V1;
        q_st(dst, fp_select(p0, z0, sa[k])); q_st(dst + 1, fp_select(p0, sa[k], z1)); q_st(dst + 2, sb[k]);     // (z0, z4, z3) | (z2, z1, z5)
        __syncwarp();
        if (k != n - 1) qf_mul12(QF_V0, QF_V0, QF_V1);
    }
    return true;
}
// V0 <- conj(V0^x) in place
static __device__ __noinline__ void qf_exp_by_x(u64 x) {
    const u32 t = q_tid();
    const bool p0 = q_pair() == 0;
    int top = 63;
    while (!((x >> top) & 1)) top--;
    Fp12Q f = qf_ld12(QF_V0);                                      // kept for the (rare) plain path
    q_st(QF_CA, p0 ? f.h.c1.h : f.h.c0.h); q_st(QF_CB, f.h.c2.h);
    __syncwarp();
    Fp sa[COMP_MAX], sb[COMP_MAX];
    int ns = 0;
#pragma unroll 1
    for (int i = 1; i <= top; i++) {
        qf_comp_sqr();
        if (((x >> i) & 1) && ns < COMP_MAX) { sa[ns] = q_ld(QF_CA, t); sb[ns] = q_ld(QF_CB, t); ns++; }
    }
    int want = 0;
    for (int i = 1; i <= top; i++) want += (int)((x >> i) & 1);
    bool ok = ns > 0 && want == ns && qf_decompress_product(sa, sb, ns);        // warp-uniform
    if (ok) {
        if (x & 1) { __syncwarp(); qf_st12(QF_V1, f); __syncwarp(); qf_mul12(QF_V0, QF_V0, QF_V1); }
    } else {
        // some quad of the warp holds a value with z2 = 0 (f = 1 when both pairings are skipped): plain square-and-multiply, whole warp
        __syncwarp();
        qf_st12(QF_V0, f); qf_st12(QF_V1, f);
        __syncwarp();
        for (int i = top - 1; i >= 0; i--) {
            qf_mul12(QF_V0, QF_V0, QF_V0);
            if ((x >> i) & 1) qf_mul12(QF_V0, QF_V0, QF_V1);
        }
    }
    qf_conj_inplace(QF_V0);
}
TCB_D void qf_put(u32 v, const Fp12Q &a) { __syncwarp(); qf_st12(v, a); __syncwarp(); }
// same chain as quad.cuh final_exponentiation; the running value is V0, the second operand of every product goes through V1;
// the other live values (r, y0..y3) wait in the local frame as own halves
// want_value == false: only the boolean is needed — the last product frob(y3) * y1 == 1 is tested as frob(y3) == conj(y1) (both factors are
// in the cyclotomic subgroup, where the inverse is the conjugate) and `out` is left unset
static __device__ __noinline__ bool qf_final_exp_is_one(const Fp12Q &in, Fp12Q &out, bool want_value) {
    const u64 x = TCB_BLS_X;
    // easy part: r = (conj(in) * in^-1)^(p^2 + 1)
    qf_put(QF_V0, in);
    qf_inv12();
    qf_put(QF_V1, fp12_conj(in));
    qf_mul12(QF_V0, QF_V0, QF_V1);
    qf_copy12(QF_V1, QF_V0);
    qf_frob(QF_V0, 2);
    qf_mul12(QF_V0, QF_V0, QF_V1);
    Fp12Q r = qf_ld12(QF_V0);
    qf_mul12(QF_V0, QF_V0, QF_V0);                 // y0 = r^2 (r is in the cyclotomic subgroup: the plain square is the cyclotomic one)
    Fp12Q y0 = qf_ld12(QF_V0);
    qf_exp_by_x(x);
    Fp12Q y1 = qf_ld12(QF_V0);                     // y1 = y0^x
    qf_exp_by_x(x >> 1);
    Fp12Q y2 = qf_ld12(QF_V0);                     // y2 = y1^(x/2)
    qf_put(QF_V0, y1);
    qf_put(QF_V1, fp12_conj(r));
    qf_mul12(QF_V0, QF_V0, QF_V1);                 // y1 * conj(r)
    qf_conj_inplace(QF_V0);
    qf_put(QF_V1, y2);
    qf_mul12(QF_V0, QF_V0, QF_V1);                 // y1 = conj(y1 conj(r)) * y2
    y1 = qf_ld12(QF_V0);
    qf_exp_by_x(x);
    y2 = qf_ld12(QF_V0);                           // y2 = y1^x
    qf_exp_by_x(x);                                // y3 = y2^x
    qf_put(QF_V1, fp12_conj(y1));
    qf_mul12(QF_V0, QF_V0, QF_V1);                 // y3 = y3 * conj(y1)
    Fp12Q y3 = qf_ld12(QF_V0);
    qf_put(QF_V0, y2);
    qf_frob(QF_V0, 2);
    qf_copy12(QF_V1, QF_V0);
    qf_put(QF_V0, y1);
    qf_frob(QF_V0, 3);
    qf_mul12(QF_V0, QF_V0, QF_V1);
    y1 = qf_ld12(QF_V0);                           // y1 = frob3(y1) * frob2(y2)
    qf_put(QF_V0, y3);
    qf_exp_by_x(x);                                // y3^x
    qf_put(QF_V1, y0);
    qf_mul12(QF_V0, QF_V0, QF_V1);
    qf_put(QF_V1, r);
    qf_mul12(QF_V0, QF_V0, QF_V1);                 // y2 = y3^x * y0 * r
    qf_put(QF_V1, y1);
    qf_mul12(QF_V0, QF_V0, QF_V1);                 // y1 * y2
    y1 = qf_ld12(QF_V0);
    qf_put(QF_V0, y3);
    qf_frob(QF_V0, 1);
    bool p0 = q_pair() == 0, e = q_role();
    if (!want_value) {
        Fp12Q f3 = qf_ld12(QF_V0);
        Fp12Q c1 = fp12_conj(y1);
        return q_quad_and(f3.h.c0.h == c1.h.c0.h && f3.h.c1.h == c1.h.c1.h && f3.h.c2.h == c1.h.c2.h);
    }
    qf_put(QF_V1, y1);
    qf_mul12(QF_V0, QF_V0, QF_V1);
    out = qf_ld12(QF_V0);
    Fp want = (p0 && !e) ? fp_one() : Fp::zero();
    return q_quad_and(out.h.c0.h == want && out.h.c1.h.is_zero() && out.h.c2.h.is_zero());
}
#ifndef TCB_FESM_MINB
#define TCB_FESM_MINB 2
#endif
// fe_out (optional, self-test): the value of the final exponentiation in the layout of fin
__global__ void __launch_bounds__(QNT, TCB_FESM_MINB) k_final_exp_sm(size_t n, const Fp *fin, const u8 *enc_ok, u8 *ok, Fp *fe_out) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    bool live = i < n;
    if (!live) i = n - 1;
    const Fp *p = fin + (i * 4 + (threadIdx.x & 3u)) * 3;
    Fp12Q f;
    f.h.c0.h = ldg_fp(p); f.h.c1.h = ldg_fp(p + 1); f.h.c2.h = ldg_fp(p + 2);
    Fp12Q g;
    bool res = qf_final_exp_is_one(fp12_conj(f), g, fe_out != nullptr);
    if (live && fe_out) { Fp *o = fe_out + (i * 4 + (threadIdx.x & 3u)) * 3; stg_fp(o, g.h.c0.h); stg_fp(o + 1, g.h.c1.h); stg_fp(o + 2, g.h.c2.h); }
    if (live && (threadIdx.x & 3) == 0) ok[i] = (res && enc_ok[i]) ? 1 : 0;
}
// the register engine's Miller loop alone, same output format (self-test reference for the shared-memory engine)
__global__ void __launch_bounds__(128, TCB_QUAD_MINB) k_miller_quad_reg(size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, Fp *fout, u8 *enc_ok) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    bool live = i < n;
    if (!live) i = n - 1;
    bool enc;
    Fp12Q f = miller_quad(a + 96 * i, b + 192 * i, c ? c + 96 * i : nullptr, d + 192 * i, enc);
    if (live) {
        Fp *o = fout + (i * 4 + (threadIdx.x & 3u)) * 3;
        stg_fp(o, f.h.c0.h); stg_fp(o + 1, f.h.c1.h); stg_fp(o + 2, f.h.c2.h);
        if ((threadIdx.x & 3) == 0) enc_ok[i] = enc ? 1 : 0;
    }
}
__global__ void k_count_diff(size_t words, const u32 *x, const u32 *y, unsigned long long *bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words && x[i] != y[i]) atomicAdd(bad, 1ULL);
}
// ---- self-test and roofline probes
static __device__ __forceinline__ u64 splitmix(u64 &s) {
    u64 z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static __device__ Fp rand_fp(u64 &s, int mode) {
    Fp r;
    for (;;) {
        for (int i = 0; i < 12; i += 2) { u64 v = splitmix(s); r.l[i] = (u32)v; r.l[i + 1] = (u32)(v >> 32); }
        if (mode == 1) { for (int i = 0; i < 12; i++) r.l[i] = FpParams::mod(i); r.l[0] -= 1; return r; }
        if (mode == 2) { for (int i = 0; i < 12; i++) r.l[i] = 0; return r; }
        if (mode == 3) { for (int i = 0; i < 12; i++) r.l[i] = 0; r.l[0] = 1; return r; }
        r.l[11] &= 0x1fffffffu;
        if (limbs_lt_mod<FpParams>(r.l)) return r;
    }
}
__global__ void k_selftest_fp(size_t n, u64 seed, unsigned long long *bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 s = seed + i * 0x632be59bd9b4e019ULL;
    int mode_a = i < 64 ? (int)(i & 3) : 0, mode_b = i < 64 ? (int)((i >> 2) & 3) : 0;
    Fp a = rand_fp(s, mode_a), b = rand_fp(s, mode_b), c = rand_fp(s, 0), d = rand_fp(s, mode_a);
    int errs = 0;
    Fp ref = mont_mul_portable<FpParams>(a, b);
    if ((a * b) != ref) errs++;
    if (sqr(a) != mont_mul_portable<FpParams>(a, a)) errs++;
    Fp d2 = dot2(a, b, c, d);
    if (d2 != (ref + mont_mul_portable<FpParams>(c, d))) errs++;
    if (((a + b) - b) != a) errs++;
    if (!(a + (-a)).is_zero()) errs++;
    if (!limbs_lt_mod<FpParams>((a + b).l) || !limbs_lt_mod<FpParams>((a - b).l) || !limbs_lt_mod<FpParams>(d2.l)) errs++;
    if (((i >> 5) & 15) == 0) {   // warp-uniform subset: binary-GCD inverse and Legendre symbol against the Fermat / Euler powers
        Fp am = fp_to_mont(a);
        if (fp_inv(am) != fp_inv_fermat(am)) errs++;
        Fp rt = fp_pow<ExpPp1d4>(am);
        if (fp_is_square(am) != (sqr(rt) == am)) errs++;
        if (fp_half(am) + fp_half(am) != am) errs++;
    }
    // sliced Fp2 against scalar Fp2: lanes of a pair share (a,b,c,d) of the even lane
    {
        u32 m = 3u << (threadIdx.x & 30u);
        Fp2 x, y;
        x.c0 = a; x.c1 = b; y.c0 = c; y.c1 = d;
        for (int k = 0; k < 12; k++) {
            x.c0.l[k] = __shfl_sync(m, x.c0.l[k], threadIdx.x & 30u); x.c1.l[k] = __shfl_sync(m, x.c1.l[k], threadIdx.x & 30u);
            y.c0.l[k] = __shfl_sync(m, y.c0.l[k], threadIdx.x & 30u); y.c1.l[k] = __shfl_sync(m, y.c1.l[k], threadIdx.x & 30u);
        }
        Fp2S xs = Fp2S::from_halves(x.c0, x.c1), ys = Fp2S::from_halves(y.c0, y.c1);
        Fp2 pm = x * y, ps = sqr(x), px = mul_xi(x);
        Fp2S qm = xs * ys, qs = sqr(xs), qx = mul_xi(xs);
        bool role = threadIdx.x & 1;
        if (qm.h != (role ? pm.c1 : pm.c0)) errs++;
        if (qs.h != (role ? ps.c1 : ps.c0)) errs++;
        if (qx.h != (role ? px.c1 : px.c0)) errs++;
        // Fp2 mul against the schoolbook with portable multiplies
        Fp t0 = mont_mul_portable<FpParams>(x.c0, y.c0) - mont_mul_portable<FpParams>(x.c1, y.c1);
        Fp t1 = mont_mul_portable<FpParams>(x.c0, y.c1) + mont_mul_portable<FpParams>(x.c1, y.c0);
        if (pm.c0 != t0 || pm.c1 != t1) errs++;
        // predicates and rarely-used ops of the sliced engine against the scalar one
        Fp2 z0 = x; z0.c1 = Fp::zero();          // only one half zero: exercises pair_and
        Fp2S z0s = Fp2S::from_halves(z0.c0, z0.c1);
        if (is_zero(z0s) != is_zero(z0)) errs++;
        if (is_zero(Fp2S::zero()) != true) errs++;
        if (eq(xs, ys) != eq(x, y) || !eq(xs, xs)) errs++;
        if (eq(z0s, xs) != eq(z0, x)) errs++;
        Fp2 pc = conj(x), pi = inv(x), pn = -x, pf = mul_fp(x, c);
        Fp2S qc = conj(xs), qi = inv(xs), qn = -xs, qf = mul_fp(xs, c);
        if (qc.h != (role ? pc.c1 : pc.c0)) errs++;
        if (qi.h != (role ? pi.c1 : pi.c0)) errs++;
        if (qn.h != (role ? pn.c1 : pn.c0)) errs++;
        if (qf.h != (role ? pf.c1 : pf.c0)) errs++;
        if (fp2_cmp(xs, ys) != fp2_cmp(x, y)) errs++;
        if (norm(xs) != norm(x)) errs++;
        Fp2 one_s; Fp2S::one().gather(one_s.c0, one_s.c1);
        if (!eq(one_s, Fp2::one())) errs++;
        if (((i >> 1) & 31) == 0) {   // a few square roots (expensive); the condition is pair-uniform
            Fp2 sq = sqr(x), r1;
            Fp2S r2;
            bool ok1 = fp2_sqrt(r1, sq), ok2 = fp2_sqrt(r2, Fp2S::from_halves(sq.c0, sq.c1));
            if (!ok1 || !ok2) errs++;
            if (r2.h != (role ? r1.c1 : r1.c0)) errs++;
            Fp2 nr;
            bool ok3 = fp2_sqrt(nr, x), ok4 = fp2_sqrt(r2, xs);
            if (ok3 != ok4) errs++;
        }
    }
    if (errs) atomicAdd(bad, (unsigned long long)errs);
}
// quad-engine Fp12 ops (quad.cuh) against the scalar tower on random operands
__device__ Fp12T<Fp2> rand_fp12(u64 &s) {
    Fp12T<Fp2> r;
    Fp *c = (Fp *)&r;
    for (int k = 0; k < 12; k++) c[k] = rand_fp(s, 0);
    return r;
}
__device__ Fp12Q to_quad(const Fp12T<Fp2> &a) {
    const Fp6T<Fp2> &h = quad_pair() ? a.c1 : a.c0;
    Fp12Q r;
    r.h.c0 = Fp2S::from_halves(h.c0.c0, h.c0.c1);
    r.h.c1 = Fp2S::from_halves(h.c1.c0, h.c1.c1);
    r.h.c2 = Fp2S::from_halves(h.c2.c0, h.c2.c1);
    return r;
}
__device__ int cmp_quad(const Fp12Q &q, const Fp12T<Fp2> &a) {
    const Fp6T<Fp2> &h = quad_pair() ? a.c1 : a.c0;
    bool role = threadIdx.x & 1;
    int e = 0;
    if (q.h.c0.h != (role ? h.c0.c1 : h.c0.c0)) e++;
    if (q.h.c1.h != (role ? h.c1.c1 : h.c1.c0)) e++;
    if (q.h.c2.h != (role ? h.c2.c1 : h.c2.c0)) e++;
    return e;
}
__global__ void k_selftest_quad(size_t n, u64 seed, unsigned long long *bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 s = seed + (i >> 2) * 0x9e3779b97f4a7c15ULL;      // same stream on the 4 lanes of a quad
    Fp12T<Fp2> a = rand_fp12(s), b = rand_fp12(s);
    Fp2 l0, l1, l4;
    l0.c0 = rand_fp(s, 0); l0.c1 = rand_fp(s, 0); l1.c0 = rand_fp(s, 0); l1.c1 = rand_fp(s, 0); l4.c0 = rand_fp(s, 0); l4.c1 = rand_fp(s, 0);
    Fp12Q qa = to_quad(a), qb = to_quad(b);
    int errs = 0;
    errs += cmp_quad(fp12_sqr(qa), fp12_sqr(a));
    errs += cmp_quad(fp12_mul(qa, qb), fp12_mul(a, b));
    errs += cmp_quad(fp12_conj(qa), fp12_conj(a));
    errs += cmp_quad(fp12_cyclo_sqr(qa), fp12_cyclo_sqr(a));
    for (int k = 1; k <= 3; k++) errs += cmp_quad(fp12_frob(qa, k), fp12_frob(a, k));
    {
        Fp12T<Fp2> t = a; Fp12Q qt = qa;
        fp12_mul_by_014(t, l0, l1, l4);
        fp12_mul_by_014(qt, Fp2S::from_halves(l0.c0, l0.c1), Fp2S::from_halves(l1.c0, l1.c1), Fp2S::from_halves(l4.c0, l4.c1));
        errs += cmp_quad(qt, t);
        // two lines at once (line product first) against two successive sparse multiplications of the scalar engine
        Fp12T<Fp2> t2 = a; Fp12Q qt2 = qa;
        fp12_mul_by_014(t2, l0, l1, l4);
        fp12_mul_by_014(t2, l4, l0, l1);
        LineS LA, LB;
        LA.c0 = Fp2S::from_halves(l0.c0, l0.c1); LA.c1 = Fp2S::from_halves(l1.c0, l1.c1); LA.c4 = Fp2S::from_halves(l4.c0, l4.c1);
        LB.c0 = LA.c4; LB.c1 = LA.c0; LB.c4 = LA.c1;
        fp12_mul_by_two_lines(qt2, LA, LB);
        errs += cmp_quad(qt2, t2);
    }
    if (((i >> 2) & 15) == 0) {
        errs += cmp_quad(fp12_inv(qa), fp12_inv(a));
        if (fp12_is_one(fp12_mul(qa, fp12_inv(qa))) != true) errs++;
        if (fp12_is_one(qa) != false) errs++;
        if (fp12_is_one(q12_one()) != true) errs++;
    }
    if (((i >> 2) & 255) == 0) errs += cmp_quad(final_exponentiation(qa), final_exponentiation(a));
    if (errs) atomicAdd(bad, (unsigned long long)errs);
}

namespace tcbk {
cudaError_t upload_consts_pairing(const Consts &c) {
    // the pairing kernel keeps its small call frames in L1: no shared-memory carve-out
    cudaFuncSetAttribute(k_verify_g2_quad, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncSetAttribute(k_final_exp_quad, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    if (cudaFuncSetAttribute(k_final_exp_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QF_SMEM_BYTES) != cudaSuccess) return cudaErrorInvalidValue;
    cudaFuncSetAttribute(k_miller_quad_reg, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    return cudaMemcpyToSymbol(d_consts, &c, sizeof c);
}
void run_verify_g2_quad(cudaStream_t st, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, u8 *ok) {
    if (!n) return;
    size_t threads = n * 4;
    k_verify_g2_quad<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(n, a, b, c, d, ok);
}
void run_final_exp_quad(cudaStream_t st, size_t n, const void *fbuf, const u8 *enc_ok, u8 *ok, void *fe_out) {
    if (!n) return;
    k_final_exp_quad<<<(unsigned)((n * 4 + 127) / 128), 128, 0, st>>>(n, (const Fp *)fbuf, enc_ok, ok, (Fp *)fe_out);
}
void run_final_exp_sm(cudaStream_t st, size_t n, const void *fbuf, const u8 *enc_ok, u8 *ok, void *fe_out) {
    if (!n) return;
    k_final_exp_sm<<<(unsigned)((n * 4 + QNT - 1) / QNT), QNT, QF_SMEM_BYTES, st>>>(n, (const Fp *)fbuf, enc_ok, ok, (Fp *)fe_out);
}
void run_miller_quad_reg(cudaStream_t st, size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, void *fbuf, u8 *enc_ok) {
    if (!n) return;
    k_miller_quad_reg<<<(unsigned)((n * 4 + 127) / 128), 128, 0, st>>>(n, a, b, c, d, (Fp *)fbuf, enc_ok);
}
void run_count_diff(cudaStream_t st, size_t bytes, const void *x, const void *y, unsigned long long *bad) {
    size_t words = bytes / 4;
    if (words) k_count_diff<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(words, (const u32 *)x, (const u32 *)y, bad);
}
void run_selftest(cudaStream_t st, size_t n, u64 seed, unsigned long long *bad) {
    k_selftest_fp<<<(unsigned)(n / 128), 128, 0, st>>>(n, seed, bad);
    k_selftest_quad<<<(unsigned)(n / 128 / 8 + 1), 128, 0, st>>>(n / 8, seed ^ 0x5555, bad);
}
}  // namespace tcbk
