// mul28.cuh — prototype: Montgomery multiply for the 381-bit field with 14 x 28-bit limbs and carry-free
// 64-bit column accumulators (plain IMAD.WIDE.U32, no predicate carries).  Interface: 12 x u32 limbs in /
// out, Montgomery factor 2^384 exactly as fp.cuh (operand a is pre-shifted by 2^8, 14 reduction steps
// divide by 2^392).
#pragma once
#include "fp.cuh"
namespace tcb {
// split a 384-bit value (12 x u32) shifted left by SH bits into 14 x 28-bit limbs
template <int SH>
TCB_HD void to28(u32 *o, const u32 *x) {
#pragma unroll
    for (int k = 0; k < 14; k++) {
        int bit = 28 * k - SH;           // position of limb k's bit 0 in x
        u32 v;
        if (bit + 28 <= 0) v = 0;
        else if (bit < 0) v = x[0] << (-bit);
        else {
            int w = bit >> 5, s = bit & 31;
            u32 lo = w < 12 ? x[w] : 0u, hi = (w + 1) < 12 ? x[w + 1] : 0u;
            v = s ? ((lo >> s) | (hi << (32 - s))) : lo;
        }
        o[k] = v & 0x0fffffffu;
    }
}
struct P28 {   // p in 28-bit limbs and -p^-1 mod 2^28
    TCB_HD static constexpr u32 p(int i) {
        return i == 0 ? 0xfffaaabu : i == 1 ? 0xfefffffu : i == 2 ? 0x3ffffb9u : i == 3 ? 0xfffeb15u : i == 4 ? 0x6241eabu
             : i == 5 ? 0xa0f6b0fu : i == 6 ? 0xf6730d2u : i == 7 ? 0xf38512bu : i == 8 ? 0x4774b84u : i == 9 ? 0x4bacd76u
             : i == 10 ? 0xba7b643u : i == 11 ? 0xe69a4b1u : i == 12 ? 0x1ea397fu : 0x001a011u;
    }
    static constexpr u32 INV = 0xffcfffdu;   // -p^-1 mod 2^28
};
// acc += a * b as ONE IMAD.WIDE.U32 (the C form leaves a dead high-word add behind when b is an immediate)
TCB_HD void mac64(u64 &acc, u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#else
    acc += (u64)a * b;
#endif
}
template <bool DOT2>
TCB_HD Fp mul28(const Fp &a, const Fp &b, const Fp &c, const Fp &d) {
    u32 al[14], bl[14], cl[14], dl[14];
    to28<8>(al, a.l); to28<0>(bl, b.l);
    if (DOT2) { to28<8>(cl, c.l); to28<0>(dl, d.l); }
    u64 t[28];
#pragma unroll
    for (int k = 0; k < 28; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < 14; i++)
#pragma unroll
        for (int j = 0; j < 14; j++) {
            mac64(t[i + j], al[i], bl[j]);
            if (DOT2) mac64(t[i + j], cl[i], dl[j]);
        }
#pragma unroll
    for (int i = 0; i < 14; i++) {
        if (i) t[i] += t[i - 1] >> 28;
        u32 m = ((u32)t[i] * P28::INV) & 0x0fffffffu;
#pragma unroll
        for (int j = 0; j < 14; j++) mac64(t[i + j], m, P28::p(j));
    }
    // carry-propagate the upper half and repack to 12 x u32
    u32 r28[14];
    u64 carry = t[13] >> 28;
#pragma unroll
    for (int k = 0; k < 14; k++) { u64 v = t[14 + k] + carry; r28[k] = (u32)v & 0x0fffffffu; carry = v >> 28; }
    u32 r[12];
#pragma unroll
    for (int w = 0; w < 12; w++) {
        int bit = 32 * w, k = bit / 28, s = bit % 28;
        u64 v = ((u64)r28[k] >> s);
        if (k + 1 < 14) v |= (u64)r28[k + 1] << (28 - s);
        if (k + 2 < 14 && 56 - s < 32) v |= (u64)r28[k + 2] << (56 - s);
        r[w] = (u32)v;
    }
    Fp out;
    final_sub<FpParams>(out, r, 0);
    return out;
}
}  // namespace tcb
