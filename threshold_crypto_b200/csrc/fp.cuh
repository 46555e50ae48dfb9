// fp.cuh — BLS12-381 base field Fp (381 bit) and scalar field Fr (255 bit) for sm_100a.
//
// Storage layout is the one EXTERNAL pairing 0.16 uses (FqRepr = 6 x u64 little-endian,
// Montgomery form, R = 2^384), viewed as 12 x u32 limbs because the B200 integer
// multiplier is 32x32->64 (IMAD.WIDE).  One element per thread, limbs in registers.
// Montgomery multiplication is a CIOS with split even/odd 64-bit accumulator lanes so
// every multiply-accumulate is a carry-chained IMAD.WIDE.U32 (mad.lo.cc/madc.hi.cc pairs
// that ptxas fuses); `fp_dot2` computes a*b + c*d with ONE reduction (lazy reduction for
// the Fp2 products).
//
// Everything is __host__ __device__: the host instantiation exists only so the logic can
// be exercised without a GPU by tests/hostemu (test infrastructure); the carry-flag PTX
// primitives have a bit-exact C emulation for that purpose.  The shipped library never
// runs the host instantiation (see capi.cu: no CPU fallback).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define TCB_HD __host__ __device__ __forceinline__
#define TCB_HDN static __host__ __device__ __noinline__
#define TCB_D __device__ __forceinline__
#else
#define TCB_HD inline __attribute__((always_inline))
#define TCB_HDN static __attribute__((noinline))
#define TCB_D inline
#endif

namespace tcb {

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;

// ----------------------------------------------------------------------------- carry primitives
#if defined(__CUDA_ARCH__)
#define TCB_ASM asm volatile
// d(lo,hi) = a*b (no carry)
TCB_D void mul_wide_pair(u32 &lo, u32 &hi, u32 a, u32 b) {
    TCB_ASM("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b, carry-out only (starts a chain)
TCB_D void mad_pair_cc(u32 &lo, u32 &hi, u32 a, u32 b) {
    TCB_ASM("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b + CC, carry-out
TCB_D void madc_pair_cc(u32 &lo, u32 &hi, u32 a, u32 b) {
    TCB_ASM("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) = a*b + (c_lo,c_hi) + CC, carry-out   (accumulate-and-shift form)
TCB_D void madc_pair_cc_from(u32 &lo, u32 &hi, u32 a, u32 b, u32 c_lo, u32 c_hi) {
    TCB_ASM("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c_lo), "r"(c_hi));
}
// (lo,hi) = a*b + CC, no carry-out (top lane)
TCB_D void madc_pair_top(u32 &lo, u32 &hi, u32 a, u32 b) {
    TCB_ASM("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
TCB_D void add_cc(u32 &d, u32 a, u32 b) { TCB_ASM("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
TCB_D void addc_cc(u32 &d, u32 a, u32 b) { TCB_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
TCB_D void addc(u32 &d, u32 a, u32 b) { TCB_ASM("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
TCB_D void sub_cc(u32 &d, u32 a, u32 b) { TCB_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
TCB_D void subc_cc(u32 &d, u32 a, u32 b) { TCB_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
TCB_D void subc(u32 &d, u32 a, u32 b) { TCB_ASM("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
#else
// Bit-exact host emulation of the PTX carry flag (test infrastructure for tests/hostemu).
static thread_local u32 g_cc;
static thread_local u64 g_mac_count;   // algorithmic 32x32->64 MACs issued by mont_mul_impl (host only)
inline void mul_wide_pair(u32 &lo, u32 &hi, u32 a, u32 b) { u64 p = (u64)a * b; lo = (u32)p; hi = (u32)(p >> 32); }
inline void mad_pair_cc(u32 &lo, u32 &hi, u32 a, u32 b) {
    u64 p = (u64)a * b;
    u64 s = (u64)(u32)p + lo; lo = (u32)s;
    u64 h = (p >> 32) + hi + (s >> 32); hi = (u32)h; g_cc = (u32)(h >> 32);
}
inline void madc_pair_cc(u32 &lo, u32 &hi, u32 a, u32 b) {
    u64 p = (u64)a * b;
    u64 s = (u64)(u32)p + lo + g_cc; lo = (u32)s;
    u64 h = (p >> 32) + hi + (s >> 32); hi = (u32)h; g_cc = (u32)(h >> 32);
}
inline void madc_pair_cc_from(u32 &lo, u32 &hi, u32 a, u32 b, u32 c_lo, u32 c_hi) {
    u64 p = (u64)a * b;
    u64 s = (u64)(u32)p + c_lo + g_cc; lo = (u32)s;
    u64 h = (p >> 32) + c_hi + (s >> 32); hi = (u32)h; g_cc = (u32)(h >> 32);
}
inline void madc_pair_top(u32 &lo, u32 &hi, u32 a, u32 b) {
    u64 p = (u64)a * b;
    u64 s = (u64)(u32)p + g_cc; lo = (u32)s;
    hi = (u32)((p >> 32) + (s >> 32)); g_cc = 0;
}
inline void add_cc(u32 &d, u32 a, u32 b) { u64 s = (u64)a + b; d = (u32)s; g_cc = (u32)(s >> 32); }
inline void addc_cc(u32 &d, u32 a, u32 b) { u64 s = (u64)a + b + g_cc; d = (u32)s; g_cc = (u32)(s >> 32); }
inline void addc(u32 &d, u32 a, u32 b) { d = a + b + g_cc; }
inline void sub_cc(u32 &d, u32 a, u32 b) { u64 s = (u64)a - b; d = (u32)s; g_cc = (u32)(s >> 63); }
inline void subc_cc(u32 &d, u32 a, u32 b) { u64 s = (u64)a - b - g_cc; d = (u32)s; g_cc = (u32)(s >> 63); }
inline void subc(u32 &d, u32 a, u32 b) { d = a - b - g_cc; }
#endif

// ----------------------------------------------------------------------------- generic N-limb Montgomery field
// P::N limbs, P::mod(i) modulus limb, P::INV = -mod^-1 mod 2^32.
template <class P>
struct Mont {
    static constexpr int N = P::N;
    u32 l[N];

    TCB_HD static Mont zero() { Mont r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
    TCB_HD bool is_zero() const { u32 t = 0; for (int i = 0; i < N; i++) t |= l[i]; return t == 0; }
    TCB_HD bool operator==(const Mont &o) const { u32 t = 0; for (int i = 0; i < N; i++) t |= l[i] ^ o.l[i]; return t == 0; }
    TCB_HD bool operator!=(const Mont &o) const { return !(*this == o); }
};

// r = a - mod if a >= mod else a   (a < 2*mod, possibly with an extra carry word `hi`)
template <class P>
TCB_HD void final_sub(Mont<P> &r, const u32 *a, u32 hi) {
    constexpr int N = P::N;
    u32 s[N], bw;
    sub_cc(s[0], a[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(s[i], a[i], P::mod(i));
    subc(bw, hi, 0);   // bw = hi - borrow: 0xffffffff iff (hi == 0 and borrow)
    bool keep = (bw >> 31) != 0;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = keep ? a[i] : s[i];
}

template <class P>
TCB_HD Mont<P> madd(const Mont<P> &a, const Mont<P> &b) {
    constexpr int N = P::N;
    u32 t[N], hi;
    add_cc(t[0], a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) addc_cc(t[i], a.l[i], b.l[i]);
    addc(hi, 0, 0);
    Mont<P> r;
    final_sub<P>(r, t, hi);
    return r;
}
template <class P>
TCB_HD Mont<P> msub(const Mont<P> &a, const Mont<P> &b) {
    constexpr int N = P::N;
    u32 t[N], bw;
    sub_cc(t[0], a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(t[i], a.l[i], b.l[i]);
    subc(bw, 0, 0);    // 0xffffffff on borrow
    Mont<P> r;
    add_cc(r.l[0], t[0], P::mod(0) & bw);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r.l[i], t[i], P::mod(i) & bw);
    addc(r.l[N - 1], t[N - 1], P::mod(N - 1) & bw);
    return r;
}
template <class P>
TCB_HD Mont<P> mneg(const Mont<P> &a) {
    constexpr int N = P::N;
    Mont<P> r;
    u32 nz = 0;
#pragma unroll
    for (int i = 0; i < N; i++) nz |= a.l[i];
    u32 mask = nz ? 0xffffffffu : 0u;
    sub_cc(r.l[0], P::mod(0) & mask, a.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) subc_cc(r.l[i], P::mod(i) & mask, a.l[i]);
    subc(r.l[N - 1], P::mod(N - 1) & mask, a.l[N - 1]);
    return r;
}

// --- CIOS building blocks on split even/odd accumulators (see DESIGN.md "Fp multiply").
// Credit: this formulation of the row (mul_n / cmad_n / madc_n_rshift / mad_row_redc, the a+1 views, the even/odd role swap
// that makes the per-row shift free) follows Supranational's sppark, ff/mont_t.cuh (Apache-2.0), as recalled from its public
// source; dot2 and the K-term dot_rows4 / dot_finish below are additions of this repository.
template <int N>
TCB_HD void mul_n(u32 *acc, const u32 *a, u32 bi) {
#pragma unroll
    for (int j = 0; j < N; j += 2) mul_wide_pair(acc[j], acc[j + 1], a[j], bi);
}
template <int N>
TCB_HD void cmad_n(u32 *acc, const u32 *a, u32 bi) {
    mad_pair_cc(acc[0], acc[1], a[0], bi);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_pair_cc(acc[j], acc[j + 1], a[j], bi);
}
template <int N>
TCB_HD void madc_n_rshift(u32 *odd, const u32 *a, u32 bi) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) madc_pair_cc_from(odd[j], odd[j + 1], a[j], bi, odd[j + 2], odd[j + 3]);
    madc_pair_top(odd[N - 2], odd[N - 1], a[N - 2], bi);
}
// modulus limbs as a compile-time indexed pseudo array (so they become immediates)
template <class P, int OFF>
struct ModView {
    TCB_HD u32 operator[](int j) const { return P::mod(j + OFF); }
};
template <class P, int OFF>
TCB_HD void cmad_mod(u32 *acc, u32 mi) {
    constexpr int N = P::N;
    mad_pair_cc(acc[0], acc[1], P::mod(OFF), mi);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_pair_cc(acc[j], acc[j + 1], P::mod(j + OFF), mi);
}

// One CIOS row: T += a*bi (+ c*di), then T += m*mod and the implicit >> 32 (role swap).
template <class P, bool FIRST, bool DOT2>
TCB_HD void mad_row_redc(u32 *even, u32 *odd, const u32 *a, u32 bi, const u32 *c, u32 di) {
    constexpr int N = P::N;
    if (FIRST) {
        mul_n<N>(odd, a + 1, bi);
        mul_n<N>(even, a, bi);
    } else {
        add_cc(even[0], even[0], odd[1]);
        madc_n_rshift<N>(odd, a + 1, bi);
        cmad_n<N>(even, a, bi);
        addc(odd[N - 1], odd[N - 1], 0);
    }
    if (DOT2) {
        cmad_n<N>(odd, c + 1, di);
        cmad_n<N>(even, c, di);
        addc(odd[N - 1], odd[N - 1], 0);
    }
    u32 mi = even[0] * P::INV;
    cmad_mod<P, 1>(odd, mi);
    cmad_mod<P, 0>(even, mi);
    addc(odd[N - 1], odd[N - 1], 0);
}

// r = a*b (+ c*d) * R^-1 mod p.  Inputs < p.  (DOT2 result < 3p before the final subtractions.)
template <class P, bool DOT2>
TCB_HD Mont<P> mont_mul_impl(const Mont<P> &a, const Mont<P> &b, const Mont<P> &c, const Mont<P> &d) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
#if !defined(__CUDA_ARCH__)
    g_mac_count += (DOT2 ? 3 : 2) * N * N + N;   // product rows + reduction rows + the N m_i multiplies
#endif
    u32 even[N], odd[N];
    // the high limb above a[N-1] used by a+1 views is never read: views use indices j+1 <= N-1 for j even <= N-2
    mad_row_redc<P, true, DOT2>(even, odd, a.l, b.l[0], c.l, d.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i += 2) {
        mad_row_redc<P, false, DOT2>(odd, even, a.l, b.l[i], c.l, d.l[i]);
        mad_row_redc<P, false, DOT2>(even, odd, a.l, b.l[i + 1], c.l, d.l[i + 1]);
    }
    mad_row_redc<P, false, DOT2>(odd, even, a.l, b.l[N - 1], c.l, d.l[N - 1]);
    // merge: result = even + (odd >> 32)
    add_cc(even[0], even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(even[i], even[i], odd[i + 1]);
    addc(even[N - 1], even[N - 1], 0);
    Mont<P> r;
    if (DOT2) {
        Mont<P> t;
        final_sub<P>(t, even, 0);
        final_sub<P>(r, t.l, 0);
    } else {
        final_sub<P>(r, even, 0);
    }
    return r;
}

// ---- K-term dot product  sum_k x_k * y_k * R^-1 mod p  with ONE interleaved Montgomery reduction (K <= 8 for BLS12-381:
// the running value stays below (K + 1) p < 2^384).  The x operands are register-resident (K x N limbs); the y operands are
// consumed four limbs at a time (one 128-bit vector per operand and chunk), so that a caller can stream them from shared
// memory with 128-bit loads while the rows of the previous chunk are still being issued.  `dot_rows4` is one chunk (four
// CIOS rows: the even/odd accumulator roles are back where they started), `dot_finish` merges the two accumulator arrays and
// brings the result from [0, (K + 1) p) to the canonical range.
template <class P, int K>
TCB_HD void dot_row(u32 *even, u32 *odd, const u32 (*x)[P::N], const u32 *b) {
    constexpr int N = P::N;
    add_cc(even[0], even[0], odd[1]);
    madc_n_rshift<N>(odd, x[0] + 1, b[0]);
    cmad_n<N>(even, x[0], b[0]);
    addc(odd[N - 1], odd[N - 1], 0);
#pragma unroll
    for (int k = 1; k < K; k++) {
        cmad_n<N>(odd, x[k] + 1, b[k]);
        cmad_n<N>(even, x[k], b[k]);
        addc(odd[N - 1], odd[N - 1], 0);
    }
    u32 mi = even[0] * P::INV;
    cmad_mod<P, 1>(odd, mi);
    cmad_mod<P, 0>(even, mi);
    addc(odd[N - 1], odd[N - 1], 0);
}
// y4[k][0..3]: limbs 4c .. 4c+3 of y_k
template <class P, int K>
TCB_HD void dot_rows4(u32 *even, u32 *odd, const u32 (*x)[P::N], const u32 (*y4)[4]) {
    u32 b[K];
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int k = 0; k < K; k++) b[k] = y4[k][r];
        if (r & 1) dot_row<P, K>(odd, even, x, b);
        else dot_row<P, K>(even, odd, x, b);
    }
}
// k * mod as compile-time limbs (k a power of two <= 4: a shift)
template <class P, int SH>
struct ModShl {
    TCB_HD static constexpr u32 limb(int i) {
        return SH == 0 ? P::mod(i) : ((P::mod(i) << SH) | (i ? (P::mod(i - 1) >> (32 - SH)) : 0u));
    }
};
template <class P, int SH>
TCB_HD void cond_sub_shl(u32 *a) {          // a -= (mod << SH) if a >= (mod << SH)
    constexpr int N = P::N;
    u32 s[N], bw;
    sub_cc(s[0], a[0], ModShl<P, SH>::limb(0));
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(s[i], a[i], ModShl<P, SH>::limb(i));
    subc(bw, 0, 0);
    bool keep = bw != 0;
#pragma unroll
    for (int i = 0; i < N; i++) a[i] = keep ? a[i] : s[i];
}
template <class P, int K>
TCB_HD Mont<P> dot_finish(u32 *even, const u32 *odd) {
    constexpr int N = P::N;
    static_assert(K >= 1 && K <= 8, "dot product length");
    add_cc(even[0], even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(even[i], even[i], odd[i + 1]);
    addc(even[N - 1], even[N - 1], 0);
    if (K + 1 > 4) cond_sub_shl<P, 2>(even);
    if (K + 1 > 2) cond_sub_shl<P, 1>(even);
    cond_sub_shl<P, 0>(even);
    Mont<P> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = even[i];
    return r;
}
// reference use of the pieces (host emulation self-check and small device callers): all operands in registers
template <class P, int K>
TCB_HD Mont<P> mont_dotk(const Mont<P> *x, const Mont<P> *y) {
    constexpr int N = P::N;
#if !defined(__CUDA_ARCH__)
    g_mac_count += (u64)(K + 1) * N * N + N;
#endif
    u32 xs[K][N], even[N], odd[N];
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int i = 0; i < N; i++) xs[k][i] = x[k].l[i];
#pragma unroll
    for (int i = 0; i < N; i++) { even[i] = 0; odd[i] = 0; }
#pragma unroll
    for (int c = 0; c < N / 4; c++) {
        u32 y4[K][4];
#pragma unroll
        for (int k = 0; k < K; k++)
#pragma unroll
            for (int r = 0; r < 4; r++) y4[k][r] = y[k].l[4 * c + r];
        dot_rows4<P, K>(even, odd, xs, y4);
    }
    return dot_finish<P, K>(even, odd);
}

// Portable (no carry-flag tricks) CIOS used (a) by the host emulation as an independent
// check of the PTX path and (b) on the device in the self-test kernel only.
template <class P>
TCB_HD Mont<P> mont_mul_portable(const Mont<P> &a, const Mont<P> &b) {
    constexpr int N = P::N;
    u32 t[N + 2];
    for (int i = 0; i < N + 2; i++) t[i] = 0;
    for (int i = 0; i < N; i++) {
        u64 c = 0;
        for (int j = 0; j < N; j++) {
            u64 s = (u64)a.l[j] * b.l[i] + t[j] + c;
            t[j] = (u32)s; c = s >> 32;
        }
        u64 s = (u64)t[N] + c;
        t[N] = (u32)s; t[N + 1] = (u32)(s >> 32);
        u32 m = t[0] * P::INV;
        c = ((u64)m * P::mod(0) + t[0]) >> 32;
        for (int j = 1; j < N; j++) {
            u64 s2 = (u64)m * P::mod(j) + t[j] + c;
            t[j - 1] = (u32)s2; c = s2 >> 32;
        }
        s = (u64)t[N] + c;
        t[N - 1] = (u32)s;
        t[N] = t[N + 1] + (u32)(s >> 32);
    }
    // conditional subtract
    u32 sres[N];
    u64 bw = 0;
    for (int i = 0; i < N; i++) {
        u64 d = (u64)t[i] - P::mod(i) - bw;
        sres[i] = (u32)d; bw = (d >> 63) & 1;
    }
    bool keep = (t[N] == 0) && bw;
    Mont<P> r;
    for (int i = 0; i < N; i++) r.l[i] = keep ? t[i] : sres[i];
    return r;
}

template <class P>
TCB_HD Mont<P> mmul(const Mont<P> &a, const Mont<P> &b) { return mont_mul_impl<P, false>(a, b, a, b); }
template <class P>
TCB_HD Mont<P> msqr(const Mont<P> &a) { return mont_mul_impl<P, false>(a, a, a, a); }
// a*b + c*d with a single Montgomery reduction
template <class P>
TCB_HD Mont<P> mdot2(const Mont<P> &a, const Mont<P> &b, const Mont<P> &c, const Mont<P> &d) {
    return mont_mul_impl<P, true>(a, b, c, d);
}

// ----------------------------------------------------------------------------- parameters
struct FpParams {
    static constexpr int N = 12;
    static constexpr u32 INV = 0xfffcfffdu;
    TCB_HD static constexpr u32 mod(int i) {
        return i == 0 ? 0xffffaaabu : i == 1 ? 0xb9feffffu : i == 2 ? 0xb153ffffu : i == 3 ? 0x1eabfffeu
             : i == 4 ? 0xf6b0f624u : i == 5 ? 0x6730d2a0u : i == 6 ? 0xf38512bfu : i == 7 ? 0x64774b84u
             : i == 8 ? 0x434bacd7u : i == 9 ? 0x4b1ba7b6u : i == 10 ? 0x397fe69au : 0x1a0111eau;
    }
};
struct FrParams {
    static constexpr int N = 8;
    static constexpr u32 INV = 0xffffffffu;
    TCB_HD static constexpr u32 mod(int i) {
        return i == 0 ? 0x00000001u : i == 1 ? 0xffffffffu : i == 2 ? 0xfffe5bfeu : i == 3 ? 0x53bda402u
             : i == 4 ? 0x09a1d805u : i == 5 ? 0x3339d808u : i == 6 ? 0x299d7d48u : 0x73eda753u;
    }
};
typedef Mont<FpParams> Fp;
typedef Mont<FrParams> Fr;

TCB_HD Fp operator+(const Fp &a, const Fp &b) { return madd<FpParams>(a, b); }
TCB_HD Fp operator-(const Fp &a, const Fp &b) { return msub<FpParams>(a, b); }
TCB_HD Fp operator-(const Fp &a) { return mneg<FpParams>(a); }
// TCB_FP_NOINLINE (set per translation unit): the Fp multiply (and dot2) are real functions instead of ~390
// (~590) inlined instructions per use.  The instruction caches are small (L1.5: 32 KB); a curve operation with
// every multiply inlined is 50-100 KB of straight-line code and the G1 kernels spent 71 % of their stall
// samples waiting for instructions (profiles/r1s2_*).  The pairing kernel keeps them inlined (measured).
#if defined(TCB_FP_NOINLINE) && defined(__CUDACC__)
static __device__ __noinline__ Fp fp_mul_call(Fp a, Fp b) { return mmul<FpParams>(a, b); }
static __device__ __noinline__ Fp fp_dot2_call(Fp a, Fp b, Fp c, Fp d) { return mdot2<FpParams>(a, b, c, d); }
#endif
#if defined(TCB_FP_NOINLINE) && defined(__CUDA_ARCH__)
TCB_HD Fp operator*(const Fp &a, const Fp &b) { return fp_mul_call(a, b); }
TCB_HD Fp sqr(const Fp &a) { return fp_mul_call(a, a); }
TCB_HD Fp dot2(const Fp &a, const Fp &b, const Fp &c, const Fp &d) { return fp_dot2_call(a, b, c, d); }
#else
TCB_HD Fp operator*(const Fp &a, const Fp &b) { return mmul<FpParams>(a, b); }
TCB_HD Fp sqr(const Fp &a) { return msqr<FpParams>(a); }
TCB_HD Fp dot2(const Fp &a, const Fp &b, const Fp &c, const Fp &d) { return mdot2<FpParams>(a, b, c, d); }
#endif
TCB_HD Fp dbl(const Fp &a) { return madd<FpParams>(a, a); }

TCB_HD Fr operator+(const Fr &a, const Fr &b) { return madd<FrParams>(a, b); }
TCB_HD Fr operator-(const Fr &a, const Fr &b) { return msub<FrParams>(a, b); }
TCB_HD Fr operator-(const Fr &a) { return mneg<FrParams>(a); }
TCB_HD Fr operator*(const Fr &a, const Fr &b) { return mmul<FrParams>(a, b); }

// canonical-integer comparison helpers on raw limb arrays
template <int N>
TCB_HD int limbs_cmp(const u32 *a, const u32 *b) {
    for (int i = N - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return -1;
    }
    return 0;
}
template <class P>
TCB_HD bool limbs_lt_mod(const u32 *a) {
    for (int i = P::N - 1; i >= 0; i--) {
        if (a[i] < P::mod(i)) return true;
        if (a[i] > P::mod(i)) return false;
    }
    return false;
}

// from Montgomery form to the canonical integer (multiply by 1)
template <class P>
TCB_HD Mont<P> from_mont(const Mont<P> &a) {
    Mont<P> one = Mont<P>::zero();
    one.l[0] = 1;
    return mmul<P>(a, one);
}

}  // namespace tcb
