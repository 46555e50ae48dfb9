// quadsm.cuh — the Miller loop of the pairing-equality check with its operands STAGED IN SHARED MEMORY
// ("dot-product form"; device only).  Same lane-quad layout as quad.cuh (lane = 2*pair + role holds the
// Fp2-half `role` of the three Fp2 coefficients of the Fp6-half `pair` of an Fp12 value; pair j also owns
// the running G2 point and the line of pairing j of e(A,B) * e(-C,D)), but every value that another lane
// needs lives in a CELL of shared memory instead of being exchanged by warp shuffles:
//
//   * a slot is one Fp (48 B = three 128-bit vectors) per thread of the block, lane-interleaved
//     (vector v of thread t of slot s at  smem[(3 s + v) * 128 + t]): every access is a conflict-free
//     LDS.128 / STS.128, and "the partner's half" / "the other pair's coefficient" is just another column
//     (t ^ 1, t ^ 2), i.e. an ADDRESS computation instead of 12 shuffles + 12 selects per operand;
//   * products are K-term dot products with ONE interleaved Montgomery reduction (fp.cuh, dot_rows4):
//     an Fp2 product is a 2-term dot per lane, the sparse line multiplication f * (l0 + l1 v + l4 v w) and the Fp6
//     product of the complex squaring are three 6-term dots per lane (schoolbook over Fp2 with the
//     xi-multiples xi*f1, xi*f2 kept beside f) — no Karatsuba temporaries, no additions between products.  The x operands of a dot
//     are register-resident (with the real-part lane holding -im of the left operand), the y operands
//     stream from shared memory four limbs at a time;
//   * the whole multiplier core is ONE loop body of four CIOS rows (~6 KB of SASS) that the instruction
//     cache keeps; the previous engine's Miller iteration was ~290 KB of straight-line code.
//
// Values are bit-identical to the register engine's (same formulas: Costello-Lange-Naehrig projective steps,
// M-type lines, complex squaring), which the device self-test checks on random inputs (k_selftest_miller).
// The accumulator f leaves through global memory (576 B per item) to the final-exponentiation kernel.
#pragma once
#include "scheme.cuh"

#if defined(__CUDACC__)
// The single Fp products of the cell engine (q_fmul, the product inside q_sqr) and the 2-term dot q_mul2 are INLINED at their call
// sites: the by-value call ABI moved 24 + 12 registers per product and kept ptxas from scheduling the additions / selects of a step
// between the multiplies.  Measured on one box (profiles/kbench_r2v*.json, r2w*): pairing check 61.3 ms with all three as calls,
// 58.8 (q_sqr's product inline), 58.65 (+ q_fmul), 58.2 (+ q_mul2); unrolling the chunk loop of q_dot on top: 59.8 (255 registers
// and spills in the Miller kernel); two q_mul2 of the doubling step jammed into one chunk loop: 58.1 vs 58.1 (no change, removed).
// -DTCB_Q_CALLS restores the calls.
#if !defined(TCB_Q_CALLS)
#define TCB_QSQR_INLINE 1
#define TCB_QFMUL_INLINE 1
#define TCB_QMUL2_INLINE 1
#endif
namespace tcb {

constexpr int QNT = 128;                 // threads per block (32 quads)
enum { Q_F0 = 0, Q_F1, Q_F2, Q_XF1, Q_XF2, Q_TX, Q_TY, Q_TZ, Q_L0, Q_L1, Q_L4, Q_S0, Q_S1, Q_S2, Q_S3, Q_S4, Q_S5, Q_P, Q_NSLOT };
constexpr size_t Q_SMEM_BYTES = (size_t)Q_NSLOT * 3 * QNT * 16 + 16;      // + the mbarrier of the input staging

TCB_D bool q_quad_and(bool v) {
    u32 m = 0xFu << (threadIdx.x & 28u);
    int a = __shfl_xor_sync(m, (int)v, 1);
    bool t = v & (a != 0);
    int b = __shfl_xor_sync(m, (int)t, 2);
    return t & (b != 0);
}
TCB_D uint4 *q_sm() { extern __shared__ uint4 q_smem[]; return q_smem; }
TCB_D u32 q_tid() { return threadIdx.x; }
TCB_D u32 q_role() { return threadIdx.x & 1u; }
TCB_D u32 q_pair() { return (threadIdx.x >> 1) & 1u; }
// cell index (in 128-bit units) of vector 0 of slot s, column c
TCB_D u32 q_cell(u32 s, u32 c) { return s * (3 * QNT) + c; }
// the "re" column of pair pv of my quad (the "im" column is + 1)
TCB_D u32 q_col_re(u32 pv) { return (threadIdx.x & ~3u) | (pv << 1); }
TCB_D Fp q_ldc(u32 cell) {
    const uint4 *p = q_sm() + cell;
    uint4 a = p[0], b = p[QNT], c = p[2 * QNT];
    Fp r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    return r;
}
TCB_D void q_stc(u32 cell, const Fp &v) {
    uint4 *p = q_sm() + cell;
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[QNT] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    p[2 * QNT] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}
TCB_D Fp q_ld(u32 s, u32 col) { return q_ldc(q_cell(s, col)); }
TCB_D void q_st(u32 s, const Fp &v) { q_stc(q_cell(s, threadIdx.x), v); }      // own column
// p - a (a in [0, p]; the result of 0 is the non-canonical p, which the dot products accept)
TCB_D void q_neg_raw(u32 *r, const u32 *a) {
    sub_cc(r[0], FpParams::mod(0), a[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) subc_cc(r[i], FpParams::mod(i), a[i]);
    subc(r[11], FpParams::mod(11), a[11]);
}
// Resident (left) operands of T Fp2 values given by the cell index of their "re" column: x[2t] = my half,
// x[2t+1] = the partner's half, negated on the real-part lane (re = a0 b0 - a1 b1, im = a1 b0 + a0 b1).
template <int T>
TCB_D void q_load_x(u32 (*x)[12], const u32 *ure) {
    const u32 e = q_role();
#pragma unroll
    for (int t = 0; t < T; t++) {
        Fp own = q_ldc(ure[t] + e), part = q_ldc(ure[t] + (e ^ 1u));
        u32 n[12];
        q_neg_raw(n, part.l);
#pragma unroll
        for (int i = 0; i < 12; i++) { x[2 * t][i] = own.l[i]; x[2 * t + 1][i] = e ? part.l[i] : n[i]; }
    }
}
// sum_t U_t * V_t (my half of the Fp2 result); V_t given by the cell index of its "re" column.
template <int T>
TCB_D Fp q_dot(const u32 (*x)[12], const u32 *vre) {
    u32 even[12], odd[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { even[i] = 0; odd[i] = 0; }
#if defined(TCB_QDOT_UNROLL)
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int c = 0; c < 3; c++) {
        u32 y4[2 * T][4];
#pragma unroll
        for (int t = 0; t < T; t++) {
            uint4 a = q_sm()[vre[t] + c * QNT], b = q_sm()[vre[t] + 1 + c * QNT];
            y4[2 * t][0] = a.x; y4[2 * t][1] = a.y; y4[2 * t][2] = a.z; y4[2 * t][3] = a.w;
            y4[2 * t + 1][0] = b.x; y4[2 * t + 1][1] = b.y; y4[2 * t + 1][2] = b.z; y4[2 * t + 1][3] = b.w;
        }
        dot_rows4<FpParams, 2 * T>(even, odd, x, y4);
    }
    return dot_finish<FpParams, 2 * T>(even, odd);
}
// an Fp in global memory as three 128-bit vectors (scratch buffers are 16-byte aligned)
TCB_D Fp ldg_fp2(const Fp *p) {
    const uint4 *q = (const uint4 *)p;
    uint4 a = q[0], b = q[1], c = q[2];
    Fp r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    return r;
}
TCB_D void stg_fp2(Fp *p, const Fp &v) {
    uint4 *q = (uint4 *)p;
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]); q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]); q[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}
// single Fp product as a call (by-value ABI: operands and result in registers) whatever the translation unit inlines elsewhere
#if defined(TCB_QFMUL_INLINE)
TCB_D Fp q_fmul(Fp a, Fp b) { return mmul<FpParams>(a, b); }
#else
static __device__ __noinline__ Fp q_fmul(Fp a, Fp b) { return mmul<FpParams>(a, b); }
#endif
// Fp2 product of two cell-resident values (my half)
#if defined(TCB_QMUL2_INLINE)
TCB_D Fp q_mul2(u32 ure, u32 vre) {
#else
static __device__ __noinline__ Fp q_mul2(u32 ure, u32 vre) {
#endif
    u32 x[2][12];
    q_load_x<1>(x, &ure);
    return q_dot<1>(x, &vre);
}
// a + b and a + (p - b) WITHOUT the conditional subtraction (a, b canonical: results < 2p < 2^382)
TCB_D Fp q_add_raw(const Fp &a, const Fp &b) {
    Fp r;
    add_cc(r.l[0], a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) addc_cc(r.l[i], a.l[i], b.l[i]);
    addc(r.l[11], a.l[11], b.l[11]);
    return r;
}
TCB_D Fp q_sub_raw(const Fp &a, const Fp &b) {
    Fp n;
    q_neg_raw(n.l, b.l);
    return q_add_raw(a, n);
}
// Fp2 square of a cell-resident value: (a0 + a1)(a0 - a1) | (2 a0) a1 — one Fp product per lane.  The two factors stay unreduced
// (< 2p each): the Montgomery product of x, y < 2p is < 4 p^2 / R + p < 1.41 p (R / p = 9.8), which its final subtraction brings
// to the canonical range — same value as with reduced factors, 26 instructions less per factor.
TCB_D Fp q_sqr(u32 s) {
    Fp own = q_ld(s, q_tid()), part = q_ld(s, q_tid() ^ 1u);
    bool e = q_role();
    Fp x = q_add_raw(own, fp_select(e, own, part));
    Fp y = fp_select(e, part, q_sub_raw(own, part));
#if defined(TCB_QSQR_INLINE)
    return mmul<FpParams>(x, y);      // experiment: let ptxas interleave the independent squares of a step (profiles/README.md)
#else
    return q_fmul(x, y);
#endif
}
// Three 3-term Fp2 dot products with the SAME left operands (u0, u1, u2):  r_d = sum_t U_t * V_{d,t}.  The results are written
// to the slots d0..d2 (my column) after every lane of the warp has finished reading, so they may alias the operands.
static __device__ __noinline__ void q_mul3x3(u32 u0, u32 u1, u32 u2, u32 v00, u32 v01, u32 v02, u32 v10, u32 v11, u32 v12,
                                              u32 v20, u32 v21, u32 v22, u32 d0, u32 d1, u32 d2) {
    u32 x[6][12];
    {
        u32 ure[3] = {u0, u1, u2};
        q_load_x<3>(x, ure);
    }
    Fp r0 = Fp::zero(), r1 = Fp::zero(), r2 = Fp::zero();
    u32 va = v00, vb = v01, vc = v02;
#pragma unroll 1
    for (int d = 0; d < 3; d++) {
        u32 vre[3] = {va, vb, vc};
        Fp r = q_dot<3>(x, vre);
        r0 = r1; r1 = r2; r2 = r;
        va = d ? v20 : v10; vb = d ? v21 : v11; vc = d ? v22 : v12;
    }
    __syncwarp();
    q_st(d0, r0); q_st(d1, r1); q_st(d2, r2);
    __syncwarp();
}
// xi * F1, xi * F2 of my pair, kept beside F (xi (a0 + a1 u) = (a0 - a1) + (a0 + a1) u)
TCB_D void q_update_xf() {
    bool e = q_role();
    u32 t = q_tid();
#pragma unroll
    for (int k = 0; k < 2; k++) {
        Fp own = q_ld(Q_F1 + k, t), part = q_ld(Q_F1 + k, t ^ 1u);
        q_st(Q_XF1 + k, e ? own + part : own - part);
    }
    __syncwarp();
}
// f <- f * (l0 + l1 v + l4 v w) with the line of pairing lp (cells L0, L1, L4 of pair lp).  Per Fp6 half:
//   pair 0: r0 = l0 f0 + l1 xi f2 + l4 xi o1   r1 = l0 f1 + l1 f0 + l4 xi o2   r2 = l0 f2 + l1 f1 + l4 o0
//   pair 1: r0 = l0 f0 + l1 xi f2 + l4 xi o2   r1 = l0 f1 + l1 f0 + l4 o0      r2 = l0 f2 + l1 f1 + l4 o1      (o = the other half)
static __device__ __noinline__ void q_mul_by_line(u32 lp) {
    const u32 me = q_col_re(q_pair()), ot = q_col_re(q_pair() ^ 1u), lc = q_col_re(lp);
    const bool p0 = q_pair() == 0;
    u32 t0 = q_cell(p0 ? Q_XF1 : Q_XF2, ot), t1 = q_cell(p0 ? Q_XF2 : Q_F0, ot), t2 = q_cell(p0 ? Q_F0 : Q_F1, ot);
    q_mul3x3(q_cell(Q_L0, lc), q_cell(Q_L1, lc), q_cell(Q_L4, lc),
             q_cell(Q_F0, me), q_cell(Q_XF2, me), t0,
             q_cell(Q_F1, me), q_cell(Q_F0, me), t1,
             q_cell(Q_F2, me), q_cell(Q_F1, me), t2, Q_F0, Q_F1, Q_F2);
    q_update_xf();
}
// f <- f^2 (complex squaring): pair 0 computes x = (a0 + a1)(a0 + v a1), pair 1 computes y = a0 a1;
// f0' = x - y - v y, f1' = 2 y.  The Fp6 product is A * B with B resident: r0 = B0 A0 + B1 xi A2 + B2 xi A1, ...
static __device__ __noinline__ void q_sqr12() {
    const u32 t = q_tid(), o = t ^ 2u;
    const bool p0 = q_pair() == 0;
    {
        Fp a0 = q_ld(Q_F0, o), a1 = q_ld(Q_F1, o), a2 = q_ld(Q_F2, o), xa1 = q_ld(Q_XF1, o), xa2 = q_ld(Q_XF2, o);
        Fp b0 = q_ld(Q_F0, t), b1 = q_ld(Q_F1, t), b2 = q_ld(Q_F2, t);
        if (p0) {
            Fp m0 = b0, m1 = b1, m2 = b2;                 // my a0
            b0 = m0 + xa2; b1 = m1 + a0; b2 = m2 + a1;    // B = a0 + v a1   (v a1 = (xi a1_2, a1_0, a1_1))
            xa1 = xa1 + q_ld(Q_XF1, t); xa2 = xa2 + q_ld(Q_XF2, t);
            a0 = a0 + m0; a1 = a1 + m1; a2 = a2 + m2;     // A = a0 + a1
        }
        q_st(Q_L0, a0); q_st(Q_L1, a1); q_st(Q_L4, a2); q_st(Q_S3, xa1); q_st(Q_S4, xa2);
        q_st(Q_S0, b0); q_st(Q_S1, b1); q_st(Q_S2, b2);
    }
    __syncwarp();
    const u32 me = q_col_re(q_pair());
    const u32 A0 = q_cell(Q_L0, me), A1 = q_cell(Q_L1, me), A2 = q_cell(Q_L4, me), XA1 = q_cell(Q_S3, me), XA2 = q_cell(Q_S4, me);
    q_mul3x3(q_cell(Q_S0, me), q_cell(Q_S1, me), q_cell(Q_S2, me), A0, XA2, XA1, A1, A0, XA2, A2, A1, A0, Q_S0, Q_S1, Q_S2);
    Fp x0 = q_ld(Q_S0, t), x1 = q_ld(Q_S1, t), x2 = q_ld(Q_S2, t);
    Fp y0 = q_ld(Q_S0, o), y1 = q_ld(Q_S1, o), y2 = q_ld(Q_S2, o), y2p = q_ld(Q_S2, o ^ 1u);
    Fp r0, r1, r2;
    if (p0) {
        Fp xy2 = q_role() ? y2 + y2p : y2 - y2p;          // xi * y2 (my half)
        r0 = x0 - y0 - xy2; r1 = x1 - y1 - y0; r2 = x2 - y2 - y1;
    } else {
        r0 = dbl(x0); r1 = dbl(x1); r2 = dbl(x2);
    }
    __syncwarp();
    q_st(Q_F0, r0); q_st(Q_F1, r1); q_st(Q_F2, r2);
    __syncwarp();
    q_update_xf();
}
// line cells of my pairing from the unscaled line (c, b, a) and P = (px, py): l0 = c, l1 = b px, l4 = a py; an inactive pairing
// (an operand at infinity contributes 1) stores the constant 1 instead.
TCB_D void q_store_line(const Fp &lc, const Fp &lb, const Fp &la, bool act) {
    const u32 pc = q_col_re(q_pair());
    Fp px = q_ld(Q_P, pc), py = q_ld(Q_P, pc + 1);
    Fp l1 = q_fmul(lb, px), l4 = q_fmul(la, py);
    Fp one = q_role() ? Fp::zero() : fp_one();
    q_st(Q_L0, act ? lc : one);
    q_st(Q_L1, act ? l1 : Fp::zero());
    q_st(Q_L4, act ? l4 : Fp::zero());
}
// Costello-Lange-Naehrig doubling of my pairing's running point T (homogeneous projective) + its line; same values as
// tower.cuh doubling_step.
static __device__ __noinline__ void q_dbl_step(bool act) {
    const u32 t = q_tid(), me = q_col_re(q_pair());
    const bool e = q_role();
    Fp mxy = q_mul2(q_cell(Q_TX, me), q_cell(Q_TY, me));
    Fp myz = q_mul2(q_cell(Q_TY, me), q_cell(Q_TZ, me));
    Fp sx = q_sqr(Q_TX), sy = q_sqr(Q_TY), sz = q_sqr(Q_TZ);
    Fp c3 = dbl(sz) + sz;
    Fp c12 = dbl(dbl(c3));
    q_st(Q_S0, c12);
    __syncwarp();
    Fp c12p = q_ld(Q_S0, t ^ 1u);
    Fp ee = e ? c12 + c12p : c12 - c12p;          // e = xi * 12 c = 3 b' c
    Fp f = dbl(ee) + ee;
    Fp g = fp_half(sy + f);
    Fp h = dbl(myz);                              // (y + z)^2 - y^2 - z^2
    q_store_line(ee - sy, dbl(sx) + sx, -h, act);
    __syncwarp();                                 // the partner has read c12 from S0
    q_st(Q_S0, ee); q_st(Q_S1, g); q_st(Q_S2, fp_half(mxy)); q_st(Q_S3, sy - f); q_st(Q_S4, sy); q_st(Q_S5, h);
    __syncwarp();
    Fp e2 = q_sqr(Q_S0), g2 = q_sqr(Q_S1);
    Fp nx = q_mul2(q_cell(Q_S2, me), q_cell(Q_S3, me));
    Fp nz = q_mul2(q_cell(Q_S4, me), q_cell(Q_S5, me));
    Fp ny = g2 - (dbl(e2) + e2);
    q_st(Q_TX, nx); q_st(Q_TY, ny); q_st(Q_TZ, nz);
    __syncwarp();
}
// mixed addition T += Q + its line; same values as tower.cuh addition_step.  (qx, qy): my halves of the affine Q.
static __device__ __noinline__ void q_add_step(Fp qx, Fp qy, bool act) {
    const u32 t = q_tid(), me = q_col_re(q_pair());
    q_st(Q_S4, qx); q_st(Q_S5, qy);
    __syncwarp();
    Fp theta = q_ld(Q_TY, t) - q_mul2(q_cell(Q_S5, me), q_cell(Q_TZ, me));
    Fp lambda = q_ld(Q_TX, t) - q_mul2(q_cell(Q_S4, me), q_cell(Q_TZ, me));
    q_st(Q_S0, theta); q_st(Q_S1, lambda);
    __syncwarp();
    Fp c = q_sqr(Q_S0), d = q_sqr(Q_S1);
    Fp lc = q_mul2(q_cell(Q_S0, me), q_cell(Q_S4, me)) - q_mul2(q_cell(Q_S1, me), q_cell(Q_S5, me));
    q_st(Q_S2, c); q_st(Q_S3, d);
    __syncwarp();
    Fp ee = q_mul2(q_cell(Q_S1, me), q_cell(Q_S3, me));
    Fp ff = q_mul2(q_cell(Q_TZ, me), q_cell(Q_S2, me));
    Fp gg = q_mul2(q_cell(Q_TX, me), q_cell(Q_S3, me));
    Fp hh = ee + ff - dbl(gg);
    __syncwarp();                                 // S2, S3 have been read by both lanes
    q_st(Q_S2, hh); q_st(Q_S3, gg - hh); q_st(Q_L0, ee);
    __syncwarp();
    Fp nx = q_mul2(q_cell(Q_S1, me), q_cell(Q_S2, me));
    Fp ny = q_mul2(q_cell(Q_S0, me), q_cell(Q_S3, me)) - q_mul2(q_cell(Q_L0, me), q_cell(Q_TY, me));
    Fp nz = q_mul2(q_cell(Q_TZ, me), q_cell(Q_L0, me));
    __syncwarp();
    q_st(Q_TX, nx); q_st(Q_TY, ny); q_st(Q_TZ, nz);
    q_store_line(lc, -theta, lambda, act);
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------------------
// Final exponentiation on cells: the two hot pieces (Fp12 products, Karabina compressed squarings) of quad.cuh's chain with their
// operands in shared memory.  Slot map of that kernel: two Fp12 values V0, V1 (three slots each: my pair's coefficient halves)
// and six scratch slots = 12 slots = 72 KB per 128-thread block (three blocks per SM); the other live values of the chain wait in
// the thread's registers / local frame.
enum { QF_V0 = 0, QF_V1 = 3, QF_S0 = 6, QF_S1, QF_S2, QF_S3, QF_S4, QF_S5, QF_NSLOT };
constexpr u32 QF_CA = QF_S4, QF_CB = QF_S5;         // the compressed value of the squaring chain: pair 0 (z4, z3), pair 1 (z2, z5)
constexpr size_t QF_SMEM_BYTES = (size_t)QF_NSLOT * 3 * QNT * 16;
// r = x + y on the imaginary lane, x - y on the real lane (xi * (a + b u) = (a - b) + (a + b) u and friends)
TCB_D Fp q_addsub(bool e, const Fp &x, const Fp &y) { return x + fp_select(e, y, -y); }
// V[dst] <- V[a] * V[b]  (dst may be a or b).  pair 0: a0 b0 + v a1 b1;  pair 1: a0 b1 + a1 b0 — two Fp6 products per pair,
// each three 6-term dots with the b-half resident (q_mul3x3), the xi-multiples of a's coefficients in scratch.
static __device__ __noinline__ void qf_mul12(u32 dst, u32 a, u32 b) {
    const u32 t = q_tid();
    const bool e = q_role(), p0 = q_pair() == 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        Fp own = q_ld(a + 1 + k, t), part = q_ld(a + 1 + k, t ^ 1u);
        q_st(QF_S0 + k, q_addsub(e, own, part));            // xi a_{1+k}
    }
    __syncwarp();
    const u32 me = q_col_re(q_pair()), ot = q_col_re(q_pair() ^ 1u);
    u32 ya = p0 ? me : ot;                                  // P: pair 0 a0 * b0 (own, own);  pair 1 a0 * b1 (a other, b own)
    q_mul3x3(q_cell(b, me), q_cell(b + 1, me), q_cell(b + 2, me),
             q_cell(a, ya), q_cell(QF_S1, ya), q_cell(QF_S0, ya),
             q_cell(a + 1, ya), q_cell(a, ya), q_cell(QF_S1, ya),
             q_cell(a + 2, ya), q_cell(a + 1, ya), q_cell(a, ya), QF_S2, QF_S3, QF_S4);
    u32 yq = p0 ? ot : me;                                  // Q: pair 0 a1 * b1 (other, other);  pair 1 a1 * b0 (a own, b other)
    q_mul3x3(q_cell(b, ot), q_cell(b + 1, ot), q_cell(b + 2, ot),
             q_cell(a, yq), q_cell(QF_S1, yq), q_cell(QF_S0, yq),
             q_cell(a + 1, yq), q_cell(a, yq), q_cell(QF_S1, yq),
             q_cell(a + 2, yq), q_cell(a + 1, yq), q_cell(a, yq), dst, dst + 1, dst + 2);
    Fp P0 = q_ld(QF_S2, t), P1 = q_ld(QF_S3, t), P2 = q_ld(QF_S4, t);
    Fp Q0 = q_ld(dst, t), Q1 = q_ld(dst + 1, t), Q2 = q_ld(dst + 2, t), Q2p = q_ld(dst + 2, t ^ 1u);
    Fp r0, r1, r2;
    if (p0) { r0 = P0 + q_addsub(e, Q2, Q2p); r1 = P1 + Q0; r2 = P2 + Q1; }       // P + v Q,  v Q = (xi Q2, Q0, Q1)
    else { r0 = P0 + Q0; r1 = P1 + Q1; r2 = P2 + Q2; }
    __syncwarp();
    q_st(dst, r0); q_st(dst + 1, r1); q_st(dst + 2, r2);
    __syncwarp();
}
// One Karabina compressed squaring of (CA, CB) in place; same values as quad.cuh comp_sqr.
static __device__ __noinline__ void qf_comp_sqr() {
    const u32 t = q_tid(), o = t ^ 2u;
    const bool e = q_role(), p0 = q_pair() == 0;
    Fp a = q_ld(QF_CA, t);
    q_st(QF_S0, a + q_ld(QF_CB, o));                        // pair 0: z4 + z5 ; pair 1: z2 + z3
    __syncwarp();
    Fp sA = q_sqr(QF_CA), sB = q_sqr(QF_CB), sC = q_sqr(QF_S0);
    q_st(QF_S1, sA); q_st(QF_S2, sB); q_st(QF_S3, sC);
    __syncwarp();
    Fp xsB = q_addsub(e, sB, q_ld(QF_S2, t ^ 1u));          // xi * sB (mine)
    Fp own = fp_select(p0, xsB, sB);                        // pair 0: xi z3^2 ; pair 1: z5^2
    Fp osB = q_ld(QF_S2, o);
    Fp rcv = fp_select(p0, q_addsub(e, osB, q_ld(QF_S2, o ^ 1u)), osB);     // pair 0: xi z5^2 ; pair 1: z3^2
    Fp W = q_ld(QF_S1, o) + own;                            // pair 0: z2^2 + xi z3^2 ; pair 1: z4^2 + z5^2
    Fp Q = sA + rcv;                                        // pair 0: z4^2 + xi z5^2 ; pair 1: z2^2 + z3^2
    Fp d = q_ld(QF_S3, o) - W;
    q_st(QF_S0, d);                                         // S0 was last read before the previous hand-off
    __syncwarp();
    Fp Ta = q_addsub(e, d, q_ld(QF_S0, t ^ 1u));            // pair 1: xi * 2 z4 z5
    Fp Tb = sC - Q;                                         // pair 1: 2 z2 z3
    Fp Pa = fp_select(p0, W, Ta), Pb = fp_select(p0, Q, Tb);
    Fp b = q_ld(QF_CB, t);
    Fp da = Pa + fp_select(p0, -a, a);                      // z4' = 3 Pa - 2 z4 | z2' = 3 Pa + 2 z2
    Fp db = Pb + fp_select(p0, -b, b);                      // z3' = 3 Pb - 2 z3 | z5' = 3 Pb + 2 z5
    Fp na = dbl(da) + Pa, nb = dbl(db) + Pb;
    __syncwarp();
    q_st(QF_CA, na); q_st(QF_CB, nb);
    __syncwarp();
}

// ---- input staging: the block's slice of the four input arrays comes in as bulk asynchronous copies (TMA unit, 1-D) into the
// (still unused) first slots, completion on an mbarrier; the lanes then decode their coordinates from shared memory with 128-bit
// loads.  Unaligned caller pointers take a plain cooperative copy instead.
TCB_D u32 q_smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
TCB_D void q_bulk_g2s(void *dst_smem, const void *src, u32 bytes, void *mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(q_smem_addr(dst_smem)), "l"(src), "r"(bytes), "r"(q_smem_addr(mbar)) : "memory");
}
// Fp from 48 big-endian bytes at a 16-byte aligned shared-memory address -> canonical limbs; ok = ok && value < p
TCB_D Fp q_load_be_sm(const u8 *b, bool &ok) {
    const uint4 *p = (const uint4 *)b;
    uint4 v0 = p[0], v1 = p[1], v2 = p[2];
    Fp t;
    t.l[11] = __byte_perm(v0.x, 0, 0x0123); t.l[10] = __byte_perm(v0.y, 0, 0x0123); t.l[9] = __byte_perm(v0.z, 0, 0x0123); t.l[8] = __byte_perm(v0.w, 0, 0x0123);
    t.l[7] = __byte_perm(v1.x, 0, 0x0123); t.l[6] = __byte_perm(v1.y, 0, 0x0123); t.l[5] = __byte_perm(v1.z, 0, 0x0123); t.l[4] = __byte_perm(v1.w, 0, 0x0123);
    t.l[3] = __byte_perm(v2.x, 0, 0x0123); t.l[2] = __byte_perm(v2.y, 0, 0x0123); t.l[1] = __byte_perm(v2.z, 0, 0x0123); t.l[0] = __byte_perm(v2.w, 0, 0x0123);
    ok = ok && limbs_lt_mod<FpParams>(t.l);
    return fp_to_mont(t);
}
struct QStage {          // byte offsets inside the staging area (slots 0..2 = 18 432 B)
    static constexpr u32 A = 0, B = 3072, C = 9216, D = 12288, END = 18432;
};
static_assert(QStage::END <= 3 * 3 * QNT * 16, "staging area fits the first three slots");

// The Miller loop of e(a,b) * e(-c,d) for the block's 32 items; f goes to fout[(item * 4 + lane) * 3 + k], the encoding flag
// (all field elements < p) to enc_ok[item].
TCB_D void q_miller_block(size_t n, const u8 *a, const u8 *b, const u8 *c, const u8 *d, Fp *fout, u8 *enc_ok, bool gen_scaled = false) {
    const u32 t = q_tid();
    const size_t first = (size_t)blockIdx.x * (QNT / 4);
    const u32 cnt = (u32)((n - first) < (size_t)(QNT / 4) ? (n - first) : (size_t)(QNT / 4));
    u8 *stage = (u8 *)q_sm();
    unsigned long long *mbar = (unsigned long long *)(q_sm() + (size_t)Q_NSLOT * 3 * QNT);
    const bool aligned = ((((size_t)a | (size_t)b | (size_t)d | (c ? (size_t)c : 0)) & 15) == 0);
    if (aligned) {
        if (t == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(q_smem_addr(mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (t == 0) {
            u32 bytes = cnt * (96 + 192 + 192 + (c ? 96 : 0));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(q_smem_addr(mbar)), "r"(bytes) : "memory");
            q_bulk_g2s(stage + QStage::A, a + 96 * first, cnt * 96, mbar);
            q_bulk_g2s(stage + QStage::B, b + 192 * first, cnt * 192, mbar);
            if (c) q_bulk_g2s(stage + QStage::C, c + 96 * first, cnt * 96, mbar);
            q_bulk_g2s(stage + QStage::D, d + 192 * first, cnt * 192, mbar);
        }
        u32 done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(q_smem_addr(mbar)) : "memory");
        }
    } else {
        for (u32 k = t; k < cnt * 96; k += QNT) { stage[QStage::A + k] = a[96 * first + k]; if (c) stage[QStage::C + k] = c[96 * first + k]; }
        for (u32 k = t; k < cnt * 192; k += QNT) { stage[QStage::B + k] = b[192 * first + k]; stage[QStage::D + k] = d[192 * first + k]; }
        __syncthreads();
    }
    // decode: pair 0 takes (a, b), pair 1 takes (-c, d); tail quads recompute the block's last item
    u32 it = t >> 2;
    const bool live = it < cnt;
    if (!live) it = cnt - 1;
    const bool p0 = q_pair() == 0, e = q_role();
    bool ok = true;
    const u8 *pg = stage + (p0 ? QStage::A : QStage::C) + 96 * it;
    const u8 *qg = stage + (p0 ? QStage::B : QStage::D) + 192 * it;
    bool pinf, qinf = (qg[0] & 0x40) != 0;
    Fp pcoord, qx = Fp::zero(), qy = Fp::zero();       // pcoord: role 0 keeps px, role 1 keeps py
    if (p0 || c) {
        pinf = (pg[0] & 0x40) != 0;
        pcoord = pinf ? Fp::zero() : q_load_be_sm(pg + (e ? 48 : 0), ok);
    } else {
        pinf = false;
        pcoord = gen_scaled ? (e ? CONSTS().g1cy : CONSTS().g1cx) : (e ? CONSTS().g1y : CONSTS().g1x);
    }
    if (!p0 && e) pcoord = -pcoord;                    // -C
    if (!qinf) { qx = q_load_be_sm(qg + (e ? 0 : 48), ok); qy = q_load_be_sm(qg + 96 + (e ? 0 : 48), ok); }
    const bool act = !(pinf || qinf);
    ok = q_quad_and(ok);
    __syncthreads();                                   // the staging area becomes slots 0..2
    Fp one = e ? Fp::zero() : fp_one();
    q_st(Q_F0, p0 ? one : Fp::zero()); q_st(Q_F1, Fp::zero()); q_st(Q_F2, Fp::zero());
    q_st(Q_XF1, Fp::zero()); q_st(Q_XF2, Fp::zero());
    q_st(Q_TX, qinf ? Fp::zero() : qx); q_st(Q_TY, qinf ? one : qy); q_st(Q_TZ, qinf ? Fp::zero() : one);
    q_st(Q_P, pcoord);
    __syncwarp();
    const u64 xs = TCB_BLS_X >> 1;
#pragma unroll 1
    for (int i = 61; i >= 0; i--) {
        q_dbl_step(act);
        q_mul_by_line(0);
        q_mul_by_line(1);
        if ((xs >> i) & 1) {
            q_add_step(qx, qy, act);
            q_mul_by_line(0);
            q_mul_by_line(1);
        }
        q_sqr12();
    }
    q_dbl_step(act);
    q_mul_by_line(0);
    q_mul_by_line(1);
    if (live) {
        Fp *o = fout + ((first + it) * 4 + (t & 3u)) * 3;
        o[0] = q_ld(Q_F0, t); o[1] = q_ld(Q_F1, t); o[2] = q_ld(Q_F2, t);
        if ((t & 3u) == 0) enc_ok[first + it] = ok ? 1 : 0;
    }
}

}  // namespace tcb
#endif
