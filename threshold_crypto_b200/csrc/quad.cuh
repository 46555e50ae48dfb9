// quad.cuh — register-resident pairing-equality engine: FOUR lanes per item (device only).
//
// Why: with one item per thread (or per lane pair) an Fp12 accumulator is 144 (72) words per
// thread, the Miller-loop state spills to local memory and the kernel becomes bound by L1/L2
// latency at 2 warps/SMSP (the first version of round 1).  Here an Fp12 value f = c0 + c1 w is
// spread over a quad of lanes:
//      lane = 2*j + e   (j = "pair", e = "role")   holds the Fp2-half e of every Fp2
//      coefficient of c_j  ->  3 Fp = 36 registers per Fp12 value per lane.
// Pair j additionally owns the G2 running point / line computation of pairing j of the
// two-pairing product e(A,B) * e(-C,D), so both Miller loops advance in the same instruction
// stream.  All exchanges are warp shuffles (xor 1 inside a pair, xor 2 across pairs); nothing
// but the inputs and the final boolean touches memory.
//
// Fp6 arithmetic inside a pair is the generic Fp6T<Fp2S> code of tower.cuh.
#pragma once
#include "scheme.cuh"

#if defined(__CUDACC__)
namespace tcb {

typedef Fp6T<Fp2S> Fp6S;

// Phase lock: the hot loop is a ~450 KB straight-line instruction stream per Miller iteration, far
// beyond the instruction caches, and every warp fetches it from L2 on its own.  The four warps of
// a block sit on the four SMSPs of the SM (no pipe contention between them), so re-aligning them at
// block-uniform points lets them share instruction-cache fills.  All threads of the block must
// reach every TCB_PHASE() (no early exit, no call under a non-uniform condition).
#if defined(TCB_PHASE_SYNC)
#define TCB_PHASE() __syncthreads()
#else
#define TCB_PHASE() ((void)0)
#endif

TCB_D u32 quad_pair() { return (threadIdx.x >> 1) & 1u; }
TCB_D u32 quad_mask() { return 0xFu << (threadIdx.x & 28u); }
TCB_D Fp xq(const Fp &a) {   // exchange with the same role in the other pair of the quad
    Fp r;
    u32 m = quad_mask();
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_xor_sync(m, a.l[i], 2);
    return r;
}
TCB_D Fp2S xq(const Fp2S &a) { Fp2S r; r.h = xq(a.h); return r; }
TCB_D Fp6S xq(const Fp6S &a) { Fp6S r; r.c0 = xq(a.c0); r.c1 = xq(a.c1); r.c2 = xq(a.c2); return r; }
TCB_D bool quad_and(bool v) {
    u32 m = quad_mask();
    int a = __shfl_xor_sync(m, (int)v, 1);
    bool t = v & (a != 0);
    int b = __shfl_xor_sync(m, (int)t, 2);
    return t & (b != 0);
}
TCB_D bool xq_flag(bool v) { return __shfl_xor_sync(quad_mask(), (int)v, 2) != 0; }
TCB_D Fp6S sel6(bool c, const Fp6S &a, const Fp6S &b) {
    Fp6S r; r.c0 = select(c, a.c0, b.c0); r.c1 = select(c, a.c1, b.c1); r.c2 = select(c, a.c2, b.c2); return r;
}

struct Fp12Q { Fp6S h; };   // pair 0: c0, pair 1: c1

TCB_D Fp12Q q12_one() {
    Fp12Q r;
    r.h.c0 = quad_pair() ? Fp2S::zero() : Fp2S::one();
    r.h.c1 = Fp2S::zero(); r.h.c2 = Fp2S::zero();
    return r;
}
// complex squaring: pair 0 computes (a0+a1)(a0+v a1), pair 1 computes a0 a1
__device__ __noinline__ Fp12Q fp12_sqr(const Fp12Q &a) {
    bool p0 = quad_pair() == 0;
    Fp6S o = xq(a.h);
    Fp6S a0 = sel6(p0, a.h, o), a1 = sel6(p0, o, a.h);
    Fp6S lhs = sel6(p0, a0 + a1, a0);
    Fp6S rhs = sel6(p0, a0 + mul_v(a1), a1);
    Fp6S x = fp6_mul(lhs, rhs);
    Fp6S y = xq(x);                     // pair 0 receives ab
    Fp12Q r;
    r.h = sel6(p0, x - y - mul_v(y), x + x);
    return r;
}
// pair 0: a0 b0 + v a1 b1;  pair 1: a0 b1 + a1 b0
__device__ __noinline__ Fp12Q fp12_mul(const Fp12Q &a, const Fp12Q &b) {
    bool p0 = quad_pair() == 0;
    Fp6S oa = xq(a.h), ob = xq(b.h);
    Fp6S m1 = fp6_mul(sel6(p0, a.h, oa), b.h);
    Fp6S m2 = fp6_mul(sel6(p0, oa, a.h), ob);
    Fp12Q r;
    r.h = m1 + sel6(p0, mul_v(m2), m2);
    return r;
}
// f * (l0 + l1 v + l4 v w); the line coefficients are pair-sliced Fp2S values present on BOTH pairs
__device__ __noinline__ void fp12_mul_by_014(Fp12Q &f, const Fp2S &l0, const Fp2S &l1, const Fp2S &l4) {
    bool p0 = quad_pair() == 0;
    Fp6S o = xq(f.h);
    Fp6S a = fp6_mul_by_01(f.h, l0, l1);
    Fp6S b = fp6_mul_by_1(o, l4);
    f.h = a + sel6(p0, mul_v(b), b);
}
TCB_D Fp12Q fp12_conj(const Fp12Q &a) {
    Fp12Q r;
    if (quad_pair()) r.h = -a.h; else r.h = a.h;
    return r;
}
__device__ __noinline__ Fp12Q fp12_frob(const Fp12Q &a, int k) {
    const Consts &C = CONSTS();
    u32 j = quad_pair();
    bool odd = k & 1;
    Fp12Q r;
    Fp2S t0 = odd ? conj(a.h.c0) : a.h.c0, t1 = odd ? conj(a.h.c1) : a.h.c1, t2 = odd ? conj(a.h.c2) : a.h.c2;
    // coefficient v^i w^j has w-degree m = 2 i + j
    r.h.c0 = j ? t0 * Fp2S::load(C.frob[k][1]) : t0;
    r.h.c1 = t1 * Fp2S::load(C.frob[k][2 + j]);
    r.h.c2 = t2 * Fp2S::load(C.frob[k][4 + j]);
    return r;
}
__device__ __noinline__ Fp12Q fp12_inv(const Fp12Q &a) {
    bool p0 = quad_pair() == 0;
    Fp6S s = fp6_sqr(a.h);
    Fp6S o = xq(s);
    Fp6S t = sel6(p0, s - mul_v(o), o - mul_v(s));    // c0^2 - v c1^2 on both pairs
    Fp6S ti = fp6_inv(t);
    Fp12Q r;
    Fp6S m = fp6_mul(a.h, ti);
    if (p0) r.h = m; else r.h = -m;
    return r;
}
TCB_D bool fp12_is_one(const Fp12Q &a) {
    bool p0 = quad_pair() == 0;
    Fp2S want = p0 ? Fp2S::one() : Fp2S::zero();
    bool z = eq(a.h.c0, want);
    z = is_zero(a.h.c1) & z;
    z = is_zero(a.h.c2) & z;
    return quad_and(z);
}
// Granger-Scott cyclotomic squaring.  With z0=c0.c0 z4=c0.c1 z3=c0.c2 (pair 0) and
// z2=c1.c0 z1=c1.c1 z5=c1.c2 (pair 1):
//   z0' = 3(z0^2 + xi z1^2) - 2 z0      z1' = 3(2 z0 z1) + 2 z1
//   z4' = 3(z2^2 + xi z3^2) - 2 z4      z5' = 3(2 z2 z3) + 2 z5
//   z3' = 3(z4^2 + xi z5^2) - 2 z3      z2' = 3 xi (2 z4 z5) + 2 z2
// Each pair squares its own three coefficients; the three cross terms are (a+b)^2 - a^2 - b^2,
// two of them computed on pair 0 and one on pair 1 (5 squaring slots per lane).
__device__ __noinline__ Fp12Q fp12_cyclo_sqr(const Fp12Q &f) {
    bool p0 = quad_pair() == 0;
    Fp6S o = xq(f.h);                       // the other pair's coefficients
    // own squares: pair 0: q0,q4,q3 ; pair 1: q2,q1,q5  (named by z index)
    Fp6S q; q.c0 = sqr(f.h.c0); q.c1 = sqr(f.h.c1); q.c2 = sqr(f.h.c2);
    Fp6S oq = xq(q);
    // cross squares: slot A: pair 0 -> (z0+z1)^2 = (own.c0 + o.c1)^2 ; pair 1 -> (z2+z3)^2 = (own.c0 + o.c2)^2
    Fp2S sa = sqr(f.h.c0 + select(p0, o.c1, o.c2));
    // slot B: pair 0 -> (z4+z5)^2 = (own.c1 + o.c2)^2 ; pair 1 idles on the same operands (result unused)
    Fp2S sb = sqr(f.h.c1 + o.c2);
    // 2ab terms, all needed on pair 1:  2 z0 z1 = sa(p0) - q0 - q1 ; 2 z2 z3 = sa(p1) - q2 - q3 ; 2 z4 z5 = sb(p0) - q4 - q5
    Fp2S sa_o = xq(sa), sb_o = xq(sb);
    Fp12Q r;
    if (p0) {
        // own q = (q0, q4, q3), other oq = (q2, q1, q5)
        Fp2S t0 = q.c0 + mul_xi(oq.c1);      // z0^2 + xi z1^2
        Fp2S t1 = oq.c0 + mul_xi(q.c2);      // z2^2 + xi z3^2
        Fp2S t2 = q.c1 + mul_xi(oq.c2);      // z4^2 + xi z5^2
        Fp2S d;
        d = t0 - f.h.c0; r.h.c0 = d + d + t0;        // z0'
        d = t1 - f.h.c1; r.h.c1 = d + d + t1;        // z4'
        d = t2 - f.h.c2; r.h.c2 = d + d + t2;        // z3'
    } else {
        // own q = (q2, q1, q5), other oq = (q0, q4, q3); own z = (z2, z1, z5)
        Fp2S c01 = sa_o - oq.c0 - q.c1;      // 2 z0 z1
        Fp2S c23 = sa - q.c0 - oq.c2;        // 2 z2 z3
        Fp2S c45 = mul_xi(sb_o - oq.c1 - q.c2);   // xi * 2 z4 z5
        Fp2S d;
        d = c45 + f.h.c0; r.h.c0 = d + d + c45;      // z2'
        d = c01 + f.h.c1; r.h.c1 = d + d + c01;      // z1'
        d = c23 + f.h.c2; r.h.c2 = d + d + c23;      // z5'
    }
    return r;
}
__device__ __noinline__ Fp12Q fp12_exp_by_x_plain(const Fp12Q &f, u64 x) {
    Fp12Q acc = f;
    int top = 63;
    while (!((x >> top) & 1)) top--;
    for (int i = top - 1; i >= 0; i--) {
        TCB_PHASE();
        acc = fp12_cyclo_sqr(acc);
        if ((x >> i) & 1) { TCB_PHASE(); acc = fp12_mul(acc, f); }
    }
    return fp12_conj(acc);
}
// ---- Karabina's compressed squarings (eprint 2010/542) for the x-power runs.
// In the cyclotomic subgroup (z0..z5 as in fp12_cyclo_sqr) the four coordinates (z2, z3, z4, z5) square among
// themselves:  z4' = 3(z2^2 + xi z3^2) - 2 z4,  z3' = 3(z4^2 + xi z5^2) - 2 z3,  z2' = 3 xi (2 z4 z5) + 2 z2,
// z5' = 3 (2 z2 z3) + 2 z5  — 3 product slots per lane instead of 5, a third of the exchanges and two thirds of
// the linear work of the full Granger-Scott squaring (which executes 39 % of this kernel's instructions,
// profiles/r1s2c_*).  f^|x| = prod over the set bits i of x of f^(2^i): the compressed chain is walked once, the
// (six) needed powers are saved and decompressed together with ONE inversion:
//     z1 = (xi z5^2 + 3 z4^2 - 2 z3) / (4 z2),    z0 = (2 z1^2 + z2 z5 - 3 z3 z4) xi + 1        (z2 != 0).
// If some saved z2 is zero (f = 1 when both pairings are skipped; otherwise probability ~1/p^2) the whole power
// falls back to the uncompressed loop, so every input is handled.
struct CompQ { Fp2S a, b; };     // pair 0: (z4, z3); pair 1: (z2, z5)
TCB_D CompQ q12_compress(const Fp12Q &f) {
    bool p0 = quad_pair() == 0;
    CompQ c;
    c.a = select(p0, f.h.c1, f.h.c0);
    c.b = f.h.c2;
    return c;
}
// Both pairs run ONE instruction stream (no pair-divergent branch): each pair multiplies by xi what only the other pair
// needs in that form before sending it, the sums W and Q are common, and pair 1's extra steps are selected in.
__device__ __noinline__ CompQ comp_sqr(const CompQ &c) {
    bool p0 = quad_pair() == 0;
    Fp2S ob = xq(c.b);                                   // pair 0 gets z5, pair 1 gets z3
    Fp2S sA = sqr(c.a), sB = sqr(c.b), sC = sqr(c.a + ob);   // pair 0: z4^2, z3^2, (z4+z5)^2 ; pair 1: z2^2, z5^2, (z2+z3)^2
    Fp2S xsB = mul_xi(sB);
    Fp2S own = select(p0, xsB, sB);                      // kept:  pair 0: xi z3^2 ; pair 1: z5^2
    Fp2S snd = select(p0, sB, xsB);                      // sent:  pair 0: z3^2    ; pair 1: xi z5^2
    Fp2S o_sA = xq(sA), rcv = xq(snd), o_sC = xq(sC);
    Fp2S W = o_sA + own;                                 // pair 0: z2^2 + xi z3^2 ; pair 1: z4^2 + z5^2
    Fp2S Q = sA + rcv;                                   // pair 0: z4^2 + xi z5^2 ; pair 1: z2^2 + z3^2
    Fp2S Ta = mul_xi(o_sC - W);                          // pair 1: xi * 2 z4 z5
    Fp2S Tb = sC - Q;                                    // pair 1: 2 z2 z3
    Fp2S Pa = select(p0, W, Ta), Pb = select(p0, Q, Tb);
    CompQ r;
    Fp2S da = select(p0, Pa - c.a, Pa + c.a);            // z4' = 3 Pa - 2 z4 | z2' = 3 Pa + 2 z2
    Fp2S db = select(p0, Pb - c.b, Pb + c.b);            // z3' = 3 Pb - 2 z3 | z5' = 3 Pb + 2 z5
    r.a = da + da + Pa;
    r.b = db + db + Pb;
    return r;
}
constexpr int COMP_MAX = 6;      // set bits of the exponent above bit 0 (6 for |x| and |x| >> 1)
// product of the decompressed values c[0..n); false (quad-uniform) if a z2 is zero
__device__ __noinline__ bool comp_decompress_product(const CompQ *c, int n, Fp12Q &prod) {
    bool p0 = quad_pair() == 0;
    Fp2S num[COMP_MAX], den[COMP_MAX], m[COMP_MAX], pre[COMP_MAX];
    Fp2S run = Fp2S::one();
    bool bad1 = false;
    for (int k = 0; k < n; k++) {
        Fp2S s = sqr(select(p0, c[k].a, c[k].b));        // pair 0: z4^2 ; pair 1: z5^2
        m[k] = c[k].a * c[k].b;                          // pair 0: z4 z3 ; pair 1: z2 z5
        Fp2S os = xq(s), ob = xq(c[k].b);                // pair 1 receives z4^2 and z3
        num[k] = mul_xi(s) + (dbl(os) + os) - dbl(ob);   // pair 1: xi z5^2 + 3 z4^2 - 2 z3
        den[k] = dbl(dbl(c[k].a));                       // pair 1: 4 z2
        bool z = is_zero(den[k]);
        bad1 = bad1 || (!p0 && z);
        pre[k] = run;
        run = run * den[k];
    }
    bool bad_o = xq_flag(bad1);                          // executed by all four lanes (no short-circuit around the shuffle)
    if (bad1 | bad_o) return false;
    Fp2S rinv = inv(run);
    for (int k = n - 1; k >= 0; k--) {
        Fp2S dinv = rinv * pre[k];
        rinv = rinv * den[k];
        Fp2S z1 = num[k] * dinv;                         // pair 1
        Fp2S w = dbl(sqr(z1)) + m[k];                    // pair 1: 2 z1^2 + z2 z5
        Fp2S ow = xq(w);
        Fp2S z0 = mul_xi(ow - (dbl(m[k]) + m[k])) + Fp2S::one();   // pair 0: (2 z1^2 + z2 z5 - 3 z3 z4) xi + 1
        Fp12Q dk;
        dk.h.c0 = select(p0, z0, c[k].a);                // pair 0: (z0, z4, z3) ; pair 1: (z2, z1, z5)
        dk.h.c1 = select(p0, c[k].a, z1);
        dk.h.c2 = c[k].b;
        prod = (k == n - 1) ? dk : fp12_mul(prod, dk);
    }
    return true;
}
__device__ __noinline__ Fp12Q fp12_exp_by_x(const Fp12Q &f, u64 x) {
    int top = 63;
    while (!((x >> top) & 1)) top--;
    CompQ saved[COMP_MAX];
    int ns = 0;
    CompQ c = q12_compress(f);
    for (int i = 1; i <= top; i++) {
        TCB_PHASE();
        c = comp_sqr(c);
        if (((x >> i) & 1) && ns < COMP_MAX) saved[ns++] = c;
    }
    int want = 0;
    for (int i = 1; i <= top; i++) want += (int)((x >> i) & 1);
    Fp12Q acc;
    if (want != ns || ns == 0 || !comp_decompress_product(saved, ns, acc)) return fp12_exp_by_x_plain(f, x);
    if (x & 1) acc = fp12_mul(acc, f);
    return fp12_conj(acc);
}
// same chain as final_exponentiation<F2> in tower.cuh; X::exp(f, x) = conj(f^|x|) (the x-power runs are where the storage policy
// differs: local arrays here, shared-memory cells in k_final_exp_quad)
struct ExpLocal { TCB_D static Fp12Q exp(const Fp12Q &f, u64 x) { return fp12_exp_by_x(f, x); } };
template <class X>
__device__ __noinline__ Fp12Q final_exponentiation_with(const Fp12Q &in) {
    Fp12Q f2 = fp12_inv(in);
    TCB_PHASE();
    Fp12Q r = fp12_mul(fp12_conj(in), f2);
    f2 = r;
    r = fp12_mul(fp12_frob(r, 2), f2);
    TCB_PHASE();
    const u64 x = TCB_BLS_X;
    Fp12Q y0 = fp12_cyclo_sqr(r);
    Fp12Q y1 = X::exp(y0, x);
    Fp12Q y2 = X::exp(y1, x >> 1);
    Fp12Q y3 = fp12_conj(r);
    y1 = fp12_mul(y1, y3);
    y1 = fp12_conj(y1);
    y1 = fp12_mul(y1, y2);
    y2 = X::exp(y1, x);
    y3 = X::exp(y2, x);
    y1 = fp12_conj(y1);
    y3 = fp12_mul(y3, y1);
    y1 = fp12_conj(y1);
    y1 = fp12_frob(y1, 3);
    y2 = fp12_frob(y2, 2);
    y1 = fp12_mul(y1, y2);
    y2 = X::exp(y3, x);
    y2 = fp12_mul(y2, y0);
    y2 = fp12_mul(y2, r);
    y1 = fp12_mul(y1, y2);
    y2 = fp12_frob(y3, 1);
    return fp12_mul(y1, y2);
}
TCB_D Fp12Q final_exponentiation(const Fp12Q &in) { return final_exponentiation_with<ExpLocal>(in); }

// ----------------------------------------------------------------------------- two-pairing Miller loop on a quad
// Pair j holds (P_j, Q_j) and the running point T_j.  Per step each pair evaluates its own
// line, the lines are swapped across pairs (36 words) and f is multiplied by both.
struct LineS { Fp2S c0, c1, c4; };   // l0 + l1 v + l4 v w, already scaled by P
TCB_D LineS xq(const LineS &l) { LineS r; r.c0 = xq(l.c0); r.c1 = xq(l.c1); r.c4 = xq(l.c4); return r; }
TCB_D LineS scale_line(const Line<Fp2S> &l, const Aff<Fp> &p) {
    LineS r;
    r.c0 = l.c; r.c1 = mul_fp(l.b, p.x); r.c4 = mul_fp(l.a, p.y);
    return r;
}
// a * (y1 v + y2 v^2): 5 products
TCB_D Fp6S fp6_mul_by_12(const Fp6S &a, const Fp2S &y1, const Fp2S &y2) {
    Fp2S t1 = a.c1 * y1, t2 = a.c2 * y2;
    Fp6S r;
    r.c0 = mul_xi((a.c1 + a.c2) * (y1 + y2) - t1 - t2);
    r.c1 = a.c0 * y1 + mul_xi(t2);
    r.c2 = a.c0 * y2 + t1;
    return r;
}
// f * A * B for two line elements (l0 + l1 v + l4 v w), both present on both pairs: the two lines are multiplied first
// (6 products, 3 per pair, exchanged) into  L0 + L1 w,  L0 = (a0 b0 + xi a4 b4, a0 b1 + a1 b0, a1 b1),
// L1 = (0, a0 b4 + a4 b0, a1 b4 + a4 b1), then each pair needs one full Fp6 product and one product by (0, y1, y2):
// 14 product slots per lane instead of the 16 of two successive sparse multiplications.
__device__ __noinline__ void fp12_mul_by_two_lines(Fp12Q &f, const LineS &A, const LineS &B) {
    bool p0 = quad_pair() == 0;
    // slot operands: pair 0: a0 b0, a1 b1, (a0+a1)(b0+b1) ; pair 1: a4 b4, (a0+a4)(b0+b4), (a1+a4)(b1+b4)
    Fp2S x1 = select(p0, A.c0, A.c4), y1 = select(p0, B.c0, B.c4);
    Fp2S x2 = select(p0, A.c1, A.c0 + A.c4), y2 = select(p0, B.c1, B.c0 + B.c4);
    Fp2S x3 = select(p0, A.c0, A.c4) + A.c1, y3 = select(p0, B.c0, B.c4) + B.c1;
    Fp6S q;
    q.c0 = x1 * y1; q.c1 = x2 * y2; q.c2 = x3 * y3;
    Fp6S o = xq(q);
    Fp2S P00 = select(p0, q.c0, o.c0), P11 = select(p0, q.c1, o.c1), K01 = select(p0, q.c2, o.c2);
    Fp2S P44 = select(p0, o.c0, q.c0), K04 = select(p0, o.c1, q.c1), K14 = select(p0, o.c2, q.c2);
    Fp6S L0;
    L0.c0 = P00 + mul_xi(P44);
    L0.c1 = K01 - P00 - P11;
    L0.c2 = P11;
    Fp2S m1 = K04 - P00 - P44, m2 = K14 - P11 - P44;
    Fp6S oth = xq(f.h);
    Fp6S r1 = fp6_mul(f.h, L0);
    Fp6S r2 = fp6_mul_by_12(oth, m1, m2);
    f.h = r1 + sel6(p0, mul_v(r2), r2);
}
TCB_D void apply_lines(Fp12Q &f, const LineS &mine, bool act_mine, bool act_other) {
    bool p0 = quad_pair() == 0;
    LineS other = xq(mine);
    // line of pairing 0 first, then pairing 1 (same order on all four lanes)
    bool act0 = p0 ? act_mine : act_other, act1 = p0 ? act_other : act_mine;
    // Measured (profiles/r1s2g_*): the line-product form needs 14 instead of 16 product slots per lane but the pairing kernel is
    // SLOWER with it (79.5 vs 75.7 ms per 2^16): more live values around the Fp6 product at 255 registers.  Kept, self-tested,
    // behind TCB_TWO_LINE_PRODUCT.
#if defined(TCB_TWO_LINE_PRODUCT)
    if (act0 && act1) {
        LineS A, B;
        A.c0 = select(p0, mine.c0, other.c0); A.c1 = select(p0, mine.c1, other.c1); A.c4 = select(p0, mine.c4, other.c4);
        B.c0 = select(p0, other.c0, mine.c0); B.c1 = select(p0, other.c1, mine.c1); B.c4 = select(p0, other.c4, mine.c4);
        fp12_mul_by_two_lines(f, A, B);
        return;
    }
#endif
    if (act0) fp12_mul_by_014(f, select(p0, mine.c0, other.c0), select(p0, mine.c1, other.c1), select(p0, mine.c4, other.c4));
    if (act1) fp12_mul_by_014(f, select(p0, other.c0, mine.c0), select(p0, other.c1, mine.c1), select(p0, other.c4, mine.c4));
}
// The two-pairing Miller loop of e(a,b) * e(-c,d) on one quad (before the final conjugation).  Every lane passes the
// pointers of the item; pair 0 loads (a, b), pair 1 loads (-c, d).
__device__ __noinline__ Fp12Q miller_quad(const u8 *a_g1, const u8 *b_g2, const u8 *c_g1, const u8 *d_g2, bool &ok_enc) {
    bool p0 = quad_pair() == 0;
    bool ok = true;
    Aff<Fp> p;
    if (p0) p = load_g1(a_g1, ok);
    else if (c_g1) { p = load_g1(c_g1, ok); p.y = -p.y; }
    else { p.x = CONSTS().g1x; p.y = -CONSTS().g1y; p.inf = false; }
    Aff<Fp2S> q = load_g2<Fp2S>(p0 ? b_g2 : d_g2, ok);
    bool act = !(p.inf || q.inf);
    bool act_o = xq_flag(act);
    ok_enc = quad_and(ok);
    Jac<Fp2S> t = jac_from_aff(q);
    Fp12Q f = q12_one();
    const u64 xs = TCB_BLS_X >> 1;
    for (int i = 61; i >= 0; i--) {
        TCB_PHASE();
        LineS l = scale_line(doubling_step(t), p);
        TCB_PHASE();
        apply_lines(f, l, act, act_o);
        if ((xs >> i) & 1) {
            TCB_PHASE();
            l = scale_line(addition_step(t, q), p);
            TCB_PHASE();
            apply_lines(f, l, act, act_o);
        }
        TCB_PHASE();
        f = fp12_sqr(f);
    }
    TCB_PHASE();
    LineS l = scale_line(doubling_step(t), p);
    apply_lines(f, l, act, act_o);
    return f;
}
// e(a,b) == e(c,d), evaluated by one quad (the round-1 register engine, kept for the self-test and for A/B measurements)
__device__ __noinline__ bool pairing_eq_quad(const u8 *a_g1, const u8 *b_g2, const u8 *c_g1, const u8 *d_g2, bool &ok_enc) {
    Fp12Q f = fp12_conj(miller_quad(a_g1, b_g2, c_g1, d_g2, ok_enc));
    TCB_PHASE();
    return fp12_is_one(final_exponentiation(f));
}
// an Fp in global memory as three 128-bit vectors (scratch buffers are 16-byte aligned)
TCB_D Fp ldg_fp(const Fp *p) {
    const uint4 *q = (const uint4 *)p;
    uint4 a = q[0], b = q[1], c = q[2];
    Fp r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    return r;
}
TCB_D void stg_fp(Fp *p, const Fp &v) {
    uint4 *q = (uint4 *)p;
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]); q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]); q[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}

}  // namespace tcb
#endif
