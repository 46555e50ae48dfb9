// mulcs.cuh — prototype: "carry-save" CIOS Montgomery multiply on the split even/odd 64-bit lanes of fp.cuh.
//
// fp.cuh's multiply ripples the carry of every row through one predicate: 36 serially dependent
// IMAD.WIDE.U32.X per dot2 row at 6.35 cycles each, so one warp reaches 62 % of the pipe and the kernels
// (2 warps/SMSP, 45 % non-multiply instructions) sit at 56 %.  Here every multiply-accumulate is
//     mad.lo.cc  lo, a, b, lo ;  madc.hi.cc  hi, a, b, hi ;  addc  cnt, cnt, 0
// i.e. one IMAD.WIDE.U32 with a carry-OUT only, and the carry is counted in a per-lane counter (weight =
// bit 0 of the next lane of the same parity).  No MAC depends on another MAC's predicate, so the twelve
// products of a row are independent and pipeline back to back.  Counters ride along with the implicit
// >> 32 of each row (pure renaming) and are folded in when their lane reaches the bottom.
#pragma once
#include "fp.cuh"
namespace tcb {

// (lo,hi) += a*b ; cnt += carry-out
TCB_HD void mac_cnt(u32 &lo, u32 &hi, u32 &cnt, u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                 : "+r"(lo), "+r"(hi), "+r"(cnt) : "r"(a), "r"(b));
#else
    mad_pair_cc(lo, hi, a, b); addc(cnt, cnt, 0);
#endif
}
// (lo,hi) += a*b, the lane cannot overflow (top lane)
TCB_HD void mac_top(u32 &lo, u32 &hi, u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
#else
    mad_pair_cc(lo, hi, a, b);
#endif
}
// (lo,hi) += x (32 bit) ; cnt += carry-out
TCB_HD void add32_cnt(u32 &lo, u32 &hi, u32 &cnt, u32 x) {
#if defined(__CUDA_ARCH__)
    asm volatile("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, 0; addc.u32 %2, %2, 0;" : "+r"(lo), "+r"(hi), "+r"(cnt) : "r"(x));
#else
    add_cc(lo, lo, x); addc_cc(hi, hi, 0); addc(cnt, cnt, 0);
#endif
}

// Lane state: E[k] = (e[2k], e[2k+1]) at relative limbs (2k, 2k+1), O[k] = (o[2k], o[2k+1]) at (2k+1, 2k+2);
// ce[k] / co[k] count carries out of E[k] / O[k].
template <class P, bool DOT2>
TCB_HD Mont<P> mont_mul_cs(const Mont<P> &a, const Mont<P> &b, const Mont<P> &c, const Mont<P> &d) {
    constexpr int N = P::N, H = N / 2;
    u32 e[N], o[N], ce[H], co[H];
#pragma unroll
    for (int k = 0; k < H; k++) { ce[k] = 0; co[k] = 0; }
#pragma unroll
    for (int i = 0; i < N; i++) {
        if (i == 0) {
#pragma unroll
            for (int k = 0; k < H; k++) {
                mul_wide_pair(e[2 * k], e[2 * k + 1], a.l[2 * k], b.l[0]);
                mul_wide_pair(o[2 * k], o[2 * k + 1], a.l[2 * k + 1], b.l[0]);
            }
        } else {
            // shift by one limb: E' = O, O'[k] = E[k+1], O'[H-1] = 0; ce' = co, co'[k] = ce[k+1];
            // E[0].hi is added into E'[0], ce[0] into O'[0].
            u32 e1 = e[1], c0 = ce[0];
            u32 ne[N], no[N], nce[H], nco[H];
#pragma unroll
            for (int k = 0; k < H; k++) { ne[2 * k] = o[2 * k]; ne[2 * k + 1] = o[2 * k + 1]; nce[k] = co[k]; }
#pragma unroll
            for (int k = 0; k < H - 1; k++) { no[2 * k] = e[2 * k + 2]; no[2 * k + 1] = e[2 * k + 3]; nco[k] = ce[k + 1]; }
            no[N - 2] = 0; no[N - 1] = 0; nco[H - 1] = 0;
#pragma unroll
            for (int k = 0; k < N; k++) { e[k] = ne[k]; o[k] = no[k]; }
#pragma unroll
            for (int k = 0; k < H; k++) { ce[k] = nce[k]; co[k] = nco[k]; }
            add32_cnt(e[0], e[1], ce[0], e1);
            add32_cnt(o[0], o[1], co[0], c0);
#pragma unroll
            for (int k = 0; k < H; k++) {
                mac_cnt(e[2 * k], e[2 * k + 1], ce[k], a.l[2 * k], b.l[i]);
                if (k < H - 1) mac_cnt(o[2 * k], o[2 * k + 1], co[k], a.l[2 * k + 1], b.l[i]);
                else mac_top(o[2 * k], o[2 * k + 1], a.l[2 * k + 1], b.l[i]);
            }
        }
        if (DOT2) {
#pragma unroll
            for (int k = 0; k < H; k++) {
                mac_cnt(e[2 * k], e[2 * k + 1], ce[k], c.l[2 * k], d.l[i]);
                if (k < H - 1) mac_cnt(o[2 * k], o[2 * k + 1], co[k], c.l[2 * k + 1], d.l[i]);
                else mac_top(o[2 * k], o[2 * k + 1], c.l[2 * k + 1], d.l[i]);
            }
        }
        u32 m = e[0] * P::INV;
#pragma unroll
        for (int k = 0; k < H; k++) {
            mac_cnt(e[2 * k], e[2 * k + 1], ce[k], P::mod(2 * k), m);
            if (k < H - 1) mac_cnt(o[2 * k], o[2 * k + 1], co[k], P::mod(2 * k + 1), m);
            else mac_top(o[2 * k], o[2 * k + 1], P::mod(2 * k + 1), m);
        }
    }
    // final shift + merge:  r = (E >> 32) + O + counters, with E[0].lo == 0
    //   word w of (E >> 32) = e[w + 1]; word w of O = o[w]; ce[k] has weight word 2k+1, co[k] word 2k+2 (after the shift)
    u32 t[N], r[N];
    add_cc(t[0], e[1], o[0]);
#pragma unroll
    for (int w = 1; w < N - 1; w++) addc_cc(t[w], e[w + 1], o[w]);
    addc(t[N - 1], 0, o[N - 1]);
    r[0] = t[0];
    add_cc(r[1], t[1], ce[0]);
#pragma unroll
    for (int w = 2; w < N - 1; w++) addc_cc(r[w], t[w], (w & 1) ? ce[w >> 1] : co[(w >> 1) - 1]);
    addc(r[N - 1], t[N - 1], ce[H - 1]);
    Mont<P> out;
    if (DOT2) {
        Mont<P> u;
        final_sub<P>(u, r, 0);
        final_sub<P>(out, u.l, 0);
    } else {
        final_sub<P>(out, r, 0);
    }
    return out;
}

}  // namespace tcb
