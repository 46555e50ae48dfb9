// tower.cuh — Fp2 / Fp6 / Fp12 tower, Jacobian curve arithmetic and the optimal-ate pairing
// check, templated over the Fp2 implementation:
//   * Fp2   — one element per thread (c0, c1 both in the thread's registers);
//   * Fp2S  — "lane-pair sliced": lane 2k holds c0, lane 2k+1 holds c1 of the SAME element,
//             products are one lazy-reduced dot product per lane (a0*b0 + a1*(-b1) |
//             a0*b1 + a1*b0) with the partner half fetched by warp shuffles.  Halves the
//             per-thread state of every G2 / Fp12 value and doubles the threads per item.
// All tower code above Fp2 is written once against that interface.
//
// What the reference computes with these (EXTERNAL pairing 0.16, call sites in
// /root/reference/src/lib.rs:108-110,185,511): e(A,B) == e(C,D).  Here that is evaluated as
// FE( ML(A,B) * ML(-C,D) ) == 1 with one shared Miller-loop accumulator and one final
// exponentiation (cyclotomic squarings in the hard part) — same boolean, ~half the work.
#pragma once
#include "fp.cuh"

namespace tcb {

// ----------------------------------------------------------------------------- constants (host-built, device __constant__)
struct Fp2c { Fp c0, c1; };   // plain storage form of an Fp2 element
struct Consts {
    Fp r1;            // R mod p  (Montgomery one)
    Fp r2;            // R^2 mod p
    Fp r3;            // R^3 mod p (Montgomery fix-up of the binary-GCD inverse)
    Fp b1;            // 4 (G1 curve b), Montgomery
    Fp2c frob[4][6];  // frob[k][m] = xi^(m (p^k - 1)/6)
    Fp g1x, g1y;      // G1 generator
    Fp2c g2x, g2y;    // G2 generator
    Fr fr_r1, fr_r2;  // Fr Montgomery constants
    Fp2c psi_x, psi_y;  // untwist-Frobenius-twist endomorphism: psi(x,y) = (conj(x) psi_x, conj(y) psi_y)
    Fp2c psi_cx[4], psi_cy[4];   // psi^i(x,y) = (conj^i(x) psi_cx[i], conj^i(y) psi_cy[i])
    Fp beta;            // cube root of unity in Fp with (beta x, y) = [-x^2] (x, y) on G1
    Fp g1cx, g1cy;      // [3 (x^2 - 1)] G1 generator: partner of the h_eff-cleared message point inside PublicKey::verify (g2_clear_cofactor)
};
#if defined(__CUDACC__)
static __device__ __constant__ Consts d_consts;   // one copy per translation unit (uploaded by each)
#endif
static Consts h_consts;
TCB_HD const Consts &CONSTS() {
#if defined(__CUDA_ARCH__)
    return d_consts;
#else
    return h_consts;
#endif
}

// big-exponent tables (little-endian u32 limbs), same on host and device
#define TCB_EXP_TABLE(NAME, NLIMBS, ...)                              \
    struct NAME { static constexpr int N = NLIMBS;                    \
                  TCB_HD static u32 get(int i) { const u32 t[NLIMBS] = {__VA_ARGS__}; return t[i]; } };
// p - 2
TCB_EXP_TABLE(ExpPm2, 12, 0xffffaaa9u, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau)
// (p - 3) / 4
TCB_EXP_TABLE(ExpPm3d4, 12, 0xffffeaaau, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u, 0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au)
// (p - 1) / 2
TCB_EXP_TABLE(ExpPm1d2, 12, 0xffffd555u, 0xdcff7fffu, 0x58a9ffffu, 0x0f55ffffu, 0x7b587b12u, 0xb3986950u, 0x79c2895fu, 0xb23ba5c2u, 0x21a5d66bu, 0x258dd3dbu, 0x1cbff34du, 0x0d0088f5u)
// (p + 1) / 4
TCB_EXP_TABLE(ExpPp1d4, 12, 0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u, 0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au)
// G2 cofactor h2 (EXTERNAL pairing scale_by_cofactor), 507 bits
TCB_EXP_TABLE(ExpH2, 16, 0x1c7238e5u, 0xcf1c38e3u, 0x786f0c70u, 0x1616ec6eu, 0x3a6691aeu, 0x21537e29u, 0x4d9e82efu, 0xa628f1cbu, 0x2e5a7ddfu, 0xa68a205bu, 0x47085abau, 0xcd91de45u, 0x2876a202u, 0x091d5079u, 0x5414e7f1u, 0x05d543a9u)
// r (group order) for subgroup checks
TCB_EXP_TABLE(ExpR, 8, 0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u)
// r - 2
TCB_EXP_TABLE(ExpRm2, 8, 0xffffffffu, 0xfffffffeu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u)

#define TCB_BLS_X 0xd201000000010000ULL

TCB_HD Fp fp_one() { return CONSTS().r1; }

// a^E for a compile-time exponent: fixed 4-bit windows (the exponents used here, (p+1)/4, (p-3)/4, p-2,
// have ~190 one bits: 380 squarings + 95 window products + 14 for the table instead of 380 + 190).
// TCB_FP_POW_CALL (k_g2.cu): the five multiplies of the loop body are calls to one shared function; inlined they are
// 38 KB of straight-line code against a 32 KB instruction cache (33-38 % instruction-fetch stalls in k_hash_g2).
#if defined(TCB_FP_POW_CALL) && defined(__CUDACC__) && !defined(TCB_FP_NOINLINE)
static __device__ __noinline__ Fp fp_pow_mul_call(const Fp &a, const Fp &b) { return mmul<FpParams>(a, b); }
#endif
TCB_HD Fp pw_mul(const Fp &a, const Fp &b) {
#if defined(TCB_FP_POW_CALL) && defined(__CUDA_ARCH__) && !defined(TCB_FP_NOINLINE)
    return fp_pow_mul_call(a, b);
#else
    return a * b;
#endif
}
template <class E>
TCB_HDN Fp fp_pow(const Fp &a) {
    Fp tab[16];
    tab[0] = fp_one();
    tab[1] = a;
    for (int i = 2; i < 16; i++) tab[i] = pw_mul(tab[i - 1], a);
    int top = E::N * 8 - 1;                     // nibble index
    while (top > 0 && ((E::get(top >> 3) >> ((top & 7) * 4)) & 15u) == 0) top--;
    Fp acc = tab[(E::get(top >> 3) >> ((top & 7) * 4)) & 15u];
    for (int i = top - 1; i >= 0; i--) {
        for (int k = 0; k < 4; k++) acc = pw_mul(acc, acc);
        u32 nib = (E::get(i >> 3) >> ((i & 7) * 4)) & 15u;
        if (nib) acc = pw_mul(acc, tab[nib]);
    }
    return acc;
}
TCB_HD Fp fp_inv_fermat(const Fp &a) { return fp_pow<ExpPm2>(a); }
// Modular inverse by a branch-free binary GCD, limb-by-limb reference version (Pornin, "Optimized binary GCD for modular
// inversion", basic variant): 2*381 iterations of { if a odd: (swap if a < b); a -= b; u -= v }
// a >>= 1; u /= 2 } with invariants a = u*y, b = v*y (mod p).  No multiplications: the work runs on
// the ALU pipe and overlaps with other warps' IMAD.WIDE chains (a Fermat inverse is 570 Montgomery
// multiplies, ~11% of a pairing check when replicated on the 4 lanes of a quad).
// Input and output in Montgomery form: inv(aR) = a^-1 R^-1, then * R^3 * R^-1.  inv(0) = 0.
TCB_HDN Fp fp_inv_basic(const Fp &y) {
    u32 a[12], b[12], u[12], v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { a[i] = y.l[i]; b[i] = FpParams::mod(i); u[i] = 0; v[i] = 0; }
    u[0] = 1;
    for (int it = 0; it < 2 * 381; it++) {
        u32 odd = 0u - (a[0] & 1u);                       // all-ones if a is odd
        // lt = (a < b)
        u32 t[12], bw;
        sub_cc(t[0], a[0], b[0]);
#pragma unroll
        for (int i = 1; i < 12; i++) subc_cc(t[i], a[i], b[i]);
        subc(bw, 0, 0);                                    // all-ones if a < b
        u32 sw = odd & bw;                                 // swap (a,b), (u,v)
#pragma unroll
        for (int i = 0; i < 12; i++) {
            u32 x = (a[i] ^ b[i]) & sw; a[i] ^= x; b[i] ^= x;
            u32 z = (u[i] ^ v[i]) & sw; u[i] ^= z; v[i] ^= z;
        }
        // a -= b (if odd)
        sub_cc(a[0], a[0], b[0] & odd);
#pragma unroll
        for (int i = 1; i < 11; i++) subc_cc(a[i], a[i], b[i] & odd);
        subc(a[11], a[11], b[11] & odd);
        // u = u - v mod p (if odd)
        sub_cc(u[0], u[0], v[0] & odd);
#pragma unroll
        for (int i = 1; i < 12; i++) subc_cc(u[i], u[i], v[i] & odd);
        subc(bw, 0, 0);
        add_cc(u[0], u[0], FpParams::mod(0) & bw);
#pragma unroll
        for (int i = 1; i < 11; i++) addc_cc(u[i], u[i], FpParams::mod(i) & bw);
        addc(u[11], u[11], FpParams::mod(11) & bw);
        // a >>= 1
#pragma unroll
        for (int i = 0; i < 11; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
        a[11] >>= 1;
        // u = u / 2 mod p
        u32 uo = 0u - (u[0] & 1u), hi;
        add_cc(u[0], u[0], FpParams::mod(0) & uo);
#pragma unroll
        for (int i = 1; i < 12; i++) addc_cc(u[i], u[i], FpParams::mod(i) & uo);
        addc(hi, 0, 0);
#pragma unroll
        for (int i = 0; i < 11; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
        u[11] = (u[11] >> 1) | (hi << 31);
    }
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = v[i];
    return r * CONSTS().r3;
}
// Modular inverse, production version: Pornin's OPTIMISED binary GCD (eprint 2020/972, Algorithm 2).  The 2x381
// divsteps are done 30 at a time on 64-bit approximations (top 34 bits + low 30 bits of a and b), which yields
// a 2x2 transition matrix (f0 g0; f1 g1) with |f| + |g| <= 2^30; the matrix is then applied to the full-length
// (a, b) exactly and to (u, v) modulo p (a Montgomery-style division by 2^30).  26 rounds x (30 cheap steps + eight
// 12-limb x 1-limb products) ~ 30 k instructions instead of the ~145 k of the limb-by-limb version below, which
// is kept as an independent check (self-tests).  Same contract: Montgomery form in and out, inv(0) = 0.
typedef int32_t s32;
typedef int64_t s64;
TCB_HD int clz32(u32 v) {
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return __builtin_clz(v);
#endif
}
// t (14 limbs, two's complement) = f*x + g*y + k*m   for signed f, g with |f| + |g| <= 2^30, 0 <= k < 2^30
TCB_HD void inv_lin(u32 *t, s32 f, const u32 *x, s32 g, const u32 *y, u32 k) {
    s64 acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        acc += (s64)f * (s64)(u64)x[i] + (s64)g * (s64)(u64)y[i] + (s64)((u64)k * (u64)FpParams::mod(i));
        t[i] = (u32)acc;
        acc >>= 32;
    }
    t[12] = (u32)acc;
    t[13] = (u32)(acc >> 32);
}
TCB_HDN Fp fp_inv(const Fp &yin) {
    const int K = 30, ROUNDS = 26;                 // 26 * 30 = 780 >= 2 * 381 - 1
    u32 a[12], b[12], u[12], v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { a[i] = yin.l[i]; b[i] = FpParams::mod(i); u[i] = 0; v[i] = 0; }
    u[0] = 1;
    for (int rd = 0; rd < ROUNDS; rd++) {
        // 64-bit window whose top bit is the top bit of (a | b), branch-free scan from the top limb
        u32 a2 = a[11], a1 = a[10], a0 = a[9], b2 = b[11], b1 = b[10], b0 = b[9];
#pragma unroll
        for (int top = 11; top >= 3; top--) {
            bool z = (a2 | b2) == 0;                // limb `top` empty in both: slide the window down one limb
            a2 = z ? a1 : a2; a1 = z ? a0 : a1; a0 = z ? a[top - 3] : a0;
            b2 = z ? b1 : b2; b1 = z ? b0 : b1; b0 = z ? b[top - 3] : b0;
        }
        bool small = (a2 | b2) == 0;                // a and b fit 64 bits: use them exactly
        u64 ah, bh;
        {
            u32 m = a2 | b2;
            int lz = m ? clz32(m) : 0;
            u64 at = ((u64)a2 << 32) | a1, bt = ((u64)b2 << 32) | b1;
            u64 a64 = lz ? ((at << lz) | (u64)(a0 >> (32 - lz))) : at;
            u64 b64 = lz ? ((bt << lz) | (u64)(b0 >> (32 - lz))) : bt;
            ah = (a64 & ~(u64)0x3fffffffu) | (a[0] & 0x3fffffffu);
            bh = (b64 & ~(u64)0x3fffffffu) | (b[0] & 0x3fffffffu);
            if (small) { ah = ((u64)a[1] << 32) | a[0]; bh = ((u64)b[1] << 32) | b[0]; }
        }
        s32 f0 = 1, g0 = 0, f1 = 0, g1 = 1;
        for (int i = 0; i < K; i++) {
            bool odd = (ah & 1) != 0;
            bool sw = odd && ah < bh;
            u64 th = sw ? bh : ah; bh = sw ? ah : bh; ah = th;
            s32 tf = sw ? f1 : f0; f1 = sw ? f0 : f1; f0 = tf;
            s32 tg = sw ? g1 : g0; g1 = sw ? g0 : g1; g0 = tg;
            ah -= odd ? bh : 0;
            f0 -= odd ? f1 : 0;
            g0 -= odd ? g1 : 0;
            ah >>= 1;
            f1 += f1; g1 += g1;
        }
        // (a, b) <- |(f0 a + g0 b, f1 a + g1 b)| / 2^30 (exact), the signs go into the matrix rows
        u32 ta[14], tb[14];
        inv_lin(ta, f0, a, g0, b, 0);
        inv_lin(tb, f1, a, g1, b, 0);
        bool na = (ta[13] >> 31) != 0, nb = (tb[13] >> 31) != 0;
        {
            u32 ca = na ? 1u : 0u, cb = nb ? 1u : 0u, ma = na ? 0xffffffffu : 0u, mb = nb ? 0xffffffffu : 0u;
#pragma unroll
            for (int i = 0; i < 14; i++) {
                u64 sa = (u64)(ta[i] ^ ma) + ca; ta[i] = (u32)sa; ca = (u32)(sa >> 32);
                u64 sb = (u64)(tb[i] ^ mb) + cb; tb[i] = (u32)sb; cb = (u32)(sb >> 32);
            }
#pragma unroll
            for (int i = 0; i < 12; i++) { a[i] = (ta[i] >> 30) | (ta[i + 1] << 2); b[i] = (tb[i] >> 30) | (tb[i + 1] << 2); }
        }
        if (na) { f0 = -f0; g0 = -g0; }
        if (nb) { f1 = -f1; g1 = -g1; }
        // (u, v) <- (f0 u + g0 v, f1 u + g1 v) / 2^30 mod p: add k p with k = -t p^-1 mod 2^30, shift, one correction
        u32 tu[14], tv[14];
        {
            u32 lu = (u32)f0 * u[0] + (u32)g0 * v[0], lv = (u32)f1 * u[0] + (u32)g1 * v[0];
            inv_lin(tu, f0, u, g0, v, (lu * FpParams::INV) & 0x3fffffffu);
            inv_lin(tv, f1, u, g1, v, (lv * FpParams::INV) & 0x3fffffffu);
        }
#pragma unroll
        for (int i = 0; i < 12; i++) { u[i] = (tu[i] >> 30) | (tu[i + 1] << 2); v[i] = (tv[i] >> 30) | (tv[i + 1] << 2); }
        // values are in (-p, 2p): top word (bits 384.. of the shifted value) tells the sign
        s32 su = (s32)((tu[12] >> 30) | (tu[13] << 2)), sv = (s32)((tv[12] >> 30) | (tv[13] << 2));
        {
            // u: if negative add p; else subtract p if >= p
            u32 t[12], bw;
            sub_cc(t[0], u[0], FpParams::mod(0));
#pragma unroll
            for (int i = 1; i < 12; i++) subc_cc(t[i], u[i], FpParams::mod(i));
            subc(bw, 0, 0);                                   // all-ones if u(low 384 bits) < p
            bool neg = su < 0, ge = !neg && (su > 0 || bw == 0);
            u32 am = neg ? 0xffffffffu : 0u;
            u32 r[12];
            add_cc(r[0], u[0], FpParams::mod(0) & am);
#pragma unroll
            for (int i = 1; i < 11; i++) addc_cc(r[i], u[i], FpParams::mod(i) & am);
            addc(r[11], u[11], FpParams::mod(11) & am);
#pragma unroll
            for (int i = 0; i < 12; i++) u[i] = ge ? t[i] : r[i];
        }
        {
            u32 t[12], bw;
            sub_cc(t[0], v[0], FpParams::mod(0));
#pragma unroll
            for (int i = 1; i < 12; i++) subc_cc(t[i], v[i], FpParams::mod(i));
            subc(bw, 0, 0);
            bool neg = sv < 0, ge = !neg && (sv > 0 || bw == 0);
            u32 am = neg ? 0xffffffffu : 0u;
            u32 r[12];
            add_cc(r[0], v[0], FpParams::mod(0) & am);
#pragma unroll
            for (int i = 1; i < 11; i++) addc_cc(r[i], v[i], FpParams::mod(i) & am);
            addc(r[11], v[11], FpParams::mod(11) & am);
#pragma unroll
            for (int i = 0; i < 12; i++) v[i] = ge ? t[i] : r[i];
        }
    }
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = v[i];
    return r * CONSTS().r3;
}
// Legendre symbol of a (any representative < p; Montgomery form is fine because R = (2^192)^2 is
// a square): binary Jacobi algorithm, no multiplications.  Returns true iff a is a nonzero square
// or zero (i.e. a has a square root in Fp).
TCB_HDN bool fp_is_square_basic(const Fp &y) {      // one halving per iteration (reference version for the self-tests)
    u32 a[12], b[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { a[i] = y.l[i]; b[i] = FpParams::mod(i); }
    u32 neg = 0;                                           // parity of the number of sign flips
    for (int it = 0; it < 2 * 381; it++) {
        u32 nz = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) nz |= a[i];
        if (nz == 0) break;
        u32 odd = 0u - (a[0] & 1u);
        u32 t[12], bw;
        sub_cc(t[0], a[0], b[0]);
#pragma unroll
        for (int i = 1; i < 12; i++) subc_cc(t[i], a[i], b[i]);
        subc(bw, 0, 0);
        u32 sw = odd & bw;
        neg ^= ((a[0] & b[0]) >> 1) & 1u & sw;             // reciprocity: both = 3 mod 4
#pragma unroll
        for (int i = 0; i < 12; i++) { u32 x = (a[i] ^ b[i]) & sw; a[i] ^= x; b[i] ^= x; }
        sub_cc(a[0], a[0], b[0] & odd);
#pragma unroll
        for (int i = 1; i < 11; i++) subc_cc(a[i], a[i], b[i] & odd);
        subc(a[11], a[11], b[11] & odd);
        // a is even now (possibly zero): one halving, (2/b) = -1 iff b = 3, 5 mod 8
        u32 nz2 = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) nz2 |= a[i];
        u32 b8 = b[0] & 7u;
        neg ^= (nz2 != 0 && (b8 == 3u || b8 == 5u)) ? 1u : 0u;
#pragma unroll
        for (int i = 0; i < 11; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
        a[11] >>= 1;
    }
    // gcd is in b: b == 1 unless y == 0 (then b = p and the symbol is 0 -> has the root 0)
    return neg == 0;
}
// Production version: the same binary Jacobi algorithm with ALL trailing zero bits of a removed per iteration (one funnel shift
// per limb) and the swap as selects, so an iteration is one subtraction: ~0.7 * 381 iterations of ~80 instructions instead of
// up to 762 of ~110.  Invariant: b odd, (y/p) = (-1)^neg * (a/b).
TCB_HD u32 ctz32(u32 v) {
#if defined(__CUDA_ARCH__)
    return (u32)(__ffs((int)v) - 1);
#else
    return (u32)__builtin_ctz(v);
#endif
}
TCB_HDN bool fp_is_square(const Fp &y) {
    u32 a[12], b[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { a[i] = y.l[i]; b[i] = FpParams::mod(i); }
    u32 neg = 0;
    for (int it = 0; it < 2 * 381 + 12; it++) {
        u32 nz = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) nz |= a[i];
        if (nz == 0) break;
        // a <- a / 2^k, k = number of trailing zeros (at most 31 per iteration); (2/b) = -1 iff b = 3, 5 mod 8
        u32 k = a[0] ? ctz32(a[0]) : 31u;
        u32 b8 = b[0] & 7u;
        neg ^= (k & 1u) & ((b8 == 3u || b8 == 5u) ? 1u : 0u);
        if (k) {
#if defined(__CUDA_ARCH__)
#pragma unroll
            for (int i = 0; i < 11; i++) a[i] = __funnelshift_r(a[i], a[i + 1], k);
#else
#pragma unroll
            for (int i = 0; i < 11; i++) a[i] = (a[i] >> k) | (a[i + 1] << (32 - k));
#endif
            a[11] >>= k;
        }
        if (!(a[0] & 1u)) continue;                        // the low limb was zero: more zeros to remove
        // a odd: order the pair (reciprocity: a sign flip iff both = 3 mod 4), then a <- a - b (even again)
        u32 t[12], bw;
        sub_cc(t[0], a[0], b[0]);
#pragma unroll
        for (int i = 1; i < 12; i++) subc_cc(t[i], a[i], b[i]);
        subc(bw, 0, 0);                                    // all-ones if a < b
        neg ^= ((a[0] & b[0]) >> 1) & 1u & bw;
        // a < b: (a, b) <- (b - a, a) = (-t, a); else a <- t.  -t = (t ^ m) + (m & 1) with m = all-ones
#pragma unroll
        for (int i = 0; i < 12; i++) { b[i] = bw ? a[i] : b[i]; t[i] ^= bw; }
        add_cc(a[0], t[0], bw & 1u);
#pragma unroll
        for (int i = 1; i < 11; i++) addc_cc(a[i], t[i], 0);
        addc(a[11], t[11], 0);
    }
    return neg == 0;
}
TCB_HD Fp fp_to_mont(const Fp &a) { return a * CONSTS().r2; }
TCB_HD Fp fp_from_mont(const Fp &a) { return from_mont<FpParams>(a); }
// canonical compare of two Montgomery-form elements
TCB_HD int fp_cmp(const Fp &a, const Fp &b) {
    Fp ca = fp_from_mont(a), cb = fp_from_mont(b);
    return limbs_cmp<12>(ca.l, cb.l);
}

// a / 2 mod p (valid on Montgomery representatives as well)
TCB_HD Fp fp_half(const Fp &a) {
    u32 mask = (a.l[0] & 1u) ? 0xffffffffu : 0u;
    u32 t[12], hi;
    add_cc(t[0], a.l[0], FpParams::mod(0) & mask);
#pragma unroll
    for (int i = 1; i < 12; i++) addc_cc(t[i], a.l[i], FpParams::mod(i) & mask);
    addc(hi, 0, 0);
    Fp r;
#pragma unroll
    for (int i = 0; i < 11; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
    r.l[11] = (t[11] >> 1) | (hi << 31);
    return r;
}

// ----------------------------------------------------------------------------- lane helpers
#if defined(__CUDACC__)
TCB_D u32 lane_role() { return threadIdx.x & 1u; }
TCB_D u32 pair_mask() { return 3u << (threadIdx.x & 30u); }
TCB_D Fp partner(const Fp &a) {
    Fp r;
    u32 m = pair_mask();
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_xor_sync(m, a.l[i], 1);
    return r;
}
// NB: the shuffle must be executed by BOTH lanes unconditionally (no short-circuit)
TCB_D bool pair_and(bool v) { u32 m = pair_mask(); int o = __shfl_xor_sync(m, (int)v, 1); return v & (o != 0); }
#endif

// ----------------------------------------------------------------------------- Fp2, one element per thread
struct Fp2 {
    Fp c0, c1;
    static constexpr bool SLICED = false;
    TCB_HD static Fp2 zero() { Fp2 r; r.c0 = Fp::zero(); r.c1 = Fp::zero(); return r; }
    TCB_HD static Fp2 one() { Fp2 r; r.c0 = fp_one(); r.c1 = Fp::zero(); return r; }
    TCB_HD static Fp2 load(const Fp2c &c) { Fp2 r; r.c0 = c.c0; r.c1 = c.c1; return r; }
    TCB_HD static Fp2 from_halves(const Fp &c0, const Fp &c1) { Fp2 r; r.c0 = c0; r.c1 = c1; return r; }
    TCB_HD void store(Fp2c &c) const { c.c0 = c0; c.c1 = c1; }
    TCB_HD void gather(Fp &o0, Fp &o1) const { o0 = c0; o1 = c1; }
};
TCB_HD Fp2 operator+(const Fp2 &a, const Fp2 &b) { Fp2 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
TCB_HD Fp2 operator-(const Fp2 &a, const Fp2 &b) { Fp2 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
TCB_HD Fp2 operator-(const Fp2 &a) { Fp2 r; r.c0 = -a.c0; r.c1 = -a.c1; return r; }
TCB_HD Fp2 dbl(const Fp2 &a) { return a + a; }
TCB_HD Fp2 conj(const Fp2 &a) { Fp2 r; r.c0 = a.c0; r.c1 = -a.c1; return r; }
TCB_HD Fp2 operator*(const Fp2 &a, const Fp2 &b) {
    Fp2 r;
    r.c0 = dot2(a.c0, b.c0, a.c1, -b.c1);
    r.c1 = dot2(a.c0, b.c1, a.c1, b.c0);
    return r;
}
TCB_HD Fp2 sqr(const Fp2 &a) {
    Fp2 r;
    Fp m = a.c0 * a.c1;
    r.c0 = (a.c0 + a.c1) * (a.c0 - a.c1);
    r.c1 = dbl(m);
    return r;
}
TCB_HD Fp2 mul_fp(const Fp2 &a, const Fp &k) { Fp2 r; r.c0 = a.c0 * k; r.c1 = a.c1 * k; return r; }
TCB_HD Fp2 half(const Fp2 &a) { Fp2 r; r.c0 = fp_half(a.c0); r.c1 = fp_half(a.c1); return r; }
TCB_HD Fp2 mul_xi(const Fp2 &a) { Fp2 r; r.c0 = a.c0 - a.c1; r.c1 = a.c0 + a.c1; return r; }
TCB_HD bool is_zero(const Fp2 &a) { return a.c0.is_zero() && a.c1.is_zero(); }
TCB_HD bool eq(const Fp2 &a, const Fp2 &b) { return a.c0 == b.c0 && a.c1 == b.c1; }
TCB_HD Fp2 select(bool c, const Fp2 &a, const Fp2 &b) { return c ? a : b; }
TCB_HD Fp norm(const Fp2 &a) { return dot2(a.c0, a.c0, a.c1, a.c1); }
TCB_HD Fp2 inv(const Fp2 &a) {
    Fp n = fp_inv(norm(a));
    Fp2 r; r.c0 = a.c0 * n; r.c1 = -(a.c1 * n); return r;
}

#if defined(__CUDACC__)
// ----------------------------------------------------------------------------- Fp2S, sliced over a lane pair (device only)
struct Fp2S {
    Fp h;   // lane role 0: c0, role 1: c1
    static constexpr bool SLICED = true;
    TCB_D static Fp2S zero() { Fp2S r; r.h = Fp::zero(); return r; }
    TCB_D static Fp2S one() { Fp2S r; r.h = lane_role() ? Fp::zero() : fp_one(); return r; }
    TCB_D static Fp2S load(const Fp2c &c) { Fp2S r; r.h = lane_role() ? c.c1 : c.c0; return r; }
    TCB_D static Fp2S from_halves(const Fp &c0, const Fp &c1) { Fp2S r; r.h = lane_role() ? c1 : c0; return r; }
    TCB_D void store(Fp2c &c) const { if (lane_role()) c.c1 = h; else c.c0 = h; }
    TCB_D void gather(Fp &o0, Fp &o1) const { Fp o = partner(h); bool r = lane_role(); o0 = r ? o : h; o1 = r ? h : o; }
};
TCB_D Fp fp_select(bool c, const Fp &a, const Fp &b) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
TCB_D Fp2S operator+(const Fp2S &a, const Fp2S &b) { Fp2S r; r.h = a.h + b.h; return r; }
TCB_D Fp2S operator-(const Fp2S &a, const Fp2S &b) { Fp2S r; r.h = a.h - b.h; return r; }
TCB_D Fp2S operator-(const Fp2S &a) { Fp2S r; r.h = -a.h; return r; }
TCB_D Fp2S dbl(const Fp2S &a) { Fp2S r; r.h = dbl(a.h); return r; }
TCB_D Fp2S conj(const Fp2S &a) { Fp2S r; r.h = lane_role() ? -a.h : a.h; return r; }
// TCB_FP2S_CALL: build the sliced multiply / square as real functions instead of inlining them at
// every use (hot-loop code size vs. call overhead; see DESIGN.md "instruction footprint").
#if defined(TCB_FP2S_NOINLINE)
#define TCB_FP2S_CALL static __device__ __noinline__
#else
#define TCB_FP2S_CALL TCB_D
#endif
// As real functions the operands are passed BY VALUE: the device ABI then hands the 12 limbs over in registers;
// by reference they would have to be spilled to the local stack around every call.
#if defined(TCB_FP2S_NOINLINE) && !defined(TCB_FP2S_BYREF)
#define TCB_FP2S_ARG Fp2S
#else
#define TCB_FP2S_ARG const Fp2S &
#endif
TCB_FP2S_CALL Fp2S operator*(TCB_FP2S_ARG a, TCB_FP2S_ARG b) {
    bool role = lane_role();
    // role 0: a0*b0 + a1*(-b1) = h_a*h_b + o_a*(-o_b);  role 1: a0*b1 + a1*b0 = o_a*h_b + h_a*o_b.
    // The c1 lane SENDS its half of b already negated: the negation runs on the lane's own data before the
    // exchange instead of sitting between the shuffle and the multiply.
    Fp oa = partner(a.h), ob = partner(fp_select(role, -b.h, b.h));
    Fp y1 = fp_select(role, ob, b.h);
    Fp y2 = fp_select(role, b.h, ob);
    Fp2S r; r.h = dot2(a.h, y1, oa, y2);
    return r;
}
TCB_FP2S_CALL Fp2S sqr(TCB_FP2S_ARG a) {
    Fp o = partner(a.h);
    bool role = lane_role();
    // role 0: (a0 + a1)(a0 - a1);  role 1: (a0 + a0) a1   (one addition, one subtraction, two selects)
    Fp x = fp_select(role, o, a.h) + o;
    Fp y = fp_select(role, a.h, a.h - o);
    Fp2S r; r.h = x * y;
    return r;
}
TCB_D Fp2S half(const Fp2S &a) { Fp2S r; r.h = fp_half(a.h); return r; }
TCB_D Fp2S mul_fp(const Fp2S &a, const Fp &k) { Fp2S r; r.h = a.h * k; return r; }
TCB_D Fp2S mul_xi(const Fp2S &a) {
    Fp o = partner(a.h);
    Fp2S r; r.h = a.h + fp_select(lane_role(), o, -o);   // role 0: c0 - c1; role 1: c1 + c0
    return r;
}
TCB_D bool is_zero(const Fp2S &a) { return pair_and(a.h.is_zero()); }
TCB_D bool eq(const Fp2S &a, const Fp2S &b) { return pair_and(a.h == b.h); }
TCB_D Fp2S select(bool c, const Fp2S &a, const Fp2S &b) { Fp2S r; r.h = fp_select(c, a.h, b.h); return r; }
TCB_D Fp norm(const Fp2S &a) { Fp s = sqr(a.h); return s + partner(s); }
TCB_D Fp2S inv(const Fp2S &a) {
    Fp n = fp_inv(norm(a));
    Fp t = a.h * n;
    Fp2S r; r.h = lane_role() ? -t : t;
    return r;
}
#endif

// canonical ordering of EXTERNAL pairing Fq2: by c1 then c0 (A3). Works for both engines.
template <class F2>
TCB_HD int fp2_cmp(const F2 &a, const F2 &b) {
    Fp a0, a1, b0, b1;
    a.gather(a0, a1);
    b.gather(b0, b1);
    int c = fp_cmp(a1, b1);
    return c ? c : fp_cmp(a0, b0);
}

template <class F2, class E>
TCB_HDN F2 fp2_pow(const F2 &a) {
    F2 acc = F2::one();
    bool started = false;
    for (int i = E::N * 32 - 1; i >= 0; i--) {
        if (started) acc = sqr(acc);
        if ((E::get(i >> 5) >> (i & 31)) & 1) {
            if (started) acc = acc * a; else { acc = a; started = true; }
        }
    }
    return acc;
}
// Algorithm 9 of eprint 2012/685; returns false when `a` is a non-residue.
template <class F2>
TCB_HD bool fp2_sqrt(F2 &out, const F2 &a) {
    if (is_zero(a)) { out = F2::zero(); return true; }
    F2 a1 = fp2_pow<F2, ExpPm3d4>(a);
    F2 alpha = sqr(a1) * a;
    F2 a0 = conj(alpha) * alpha;
    F2 neg1 = -F2::one();
    if (eq(a0, neg1)) return false;
    a1 = a1 * a;
    if (eq(alpha, neg1)) {
        a1 = a1 * F2::from_halves(Fp::zero(), fp_one());
    } else {
        alpha = fp2_pow<F2, ExpPm1d2>(alpha + F2::one());
        a1 = a1 * alpha;
    }
    out = a1;
    return true;
}

// ----------------------------------------------------------------------------- Fp6 = Fp2[v]/(v^3 - xi)
template <class F2>
struct Fp6T { F2 c0, c1, c2; };
template <class F2> TCB_HD Fp6T<F2> operator+(const Fp6T<F2> &a, const Fp6T<F2> &b) { Fp6T<F2> r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; r.c2 = a.c2 + b.c2; return r; }
template <class F2> TCB_HD Fp6T<F2> operator-(const Fp6T<F2> &a, const Fp6T<F2> &b) { Fp6T<F2> r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; r.c2 = a.c2 - b.c2; return r; }
template <class F2> TCB_HD Fp6T<F2> operator-(const Fp6T<F2> &a) { Fp6T<F2> r; r.c0 = -a.c0; r.c1 = -a.c1; r.c2 = -a.c2; return r; }
template <class F2> TCB_HD Fp6T<F2> mul_v(const Fp6T<F2> &a) { Fp6T<F2> r; r.c0 = mul_xi(a.c2); r.c1 = a.c0; r.c2 = a.c1; return r; }
template <class F2>
TCB_HDN Fp6T<F2> fp6_mul(const Fp6T<F2> &a, const Fp6T<F2> &b) {
    F2 v0 = a.c0 * b.c0, v1 = a.c1 * b.c1, v2 = a.c2 * b.c2;
    Fp6T<F2> r;
    r.c0 = mul_xi((a.c1 + a.c2) * (b.c1 + b.c2) - v1 - v2) + v0;
    r.c1 = (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1 + mul_xi(v2);
    r.c2 = (a.c0 + a.c2) * (b.c0 + b.c2) - v0 - v2 + v1;
    return r;
}
template <class F2>
TCB_HDN Fp6T<F2> fp6_sqr(const Fp6T<F2> &a) {   // CH-SQR2
    F2 s0 = sqr(a.c0);
    F2 ab = a.c0 * a.c1;
    F2 s1 = dbl(ab);
    F2 s2 = sqr(a.c0 - a.c1 + a.c2);
    F2 bc = a.c1 * a.c2;
    F2 s3 = dbl(bc);
    F2 s4 = sqr(a.c2);
    Fp6T<F2> r;
    r.c0 = s0 + mul_xi(s3);
    r.c1 = s1 + mul_xi(s4);
    r.c2 = s1 + s2 + s3 - s0 - s4;
    return r;
}
// a * (b1 v)
template <class F2>
TCB_HD Fp6T<F2> fp6_mul_by_1(const Fp6T<F2> &a, const F2 &b1) {
    Fp6T<F2> r;
    r.c0 = mul_xi(a.c2 * b1);
    r.c1 = a.c0 * b1;
    r.c2 = a.c1 * b1;
    return r;
}
// a * (b0 + b1 v)
template <class F2>
TCB_HD Fp6T<F2> fp6_mul_by_01(const Fp6T<F2> &a, const F2 &b0, const F2 &b1) {
    F2 v0 = a.c0 * b0, v1 = a.c1 * b1;
    Fp6T<F2> r;
    r.c0 = mul_xi((a.c1 + a.c2) * b1 - v1) + v0;
    r.c1 = (a.c0 + a.c1) * (b0 + b1) - v0 - v1;
    r.c2 = (a.c0 + a.c2) * b0 - v0 + v1;
    return r;
}
template <class F2>
TCB_HDN Fp6T<F2> fp6_inv(const Fp6T<F2> &a) {
    F2 t0 = sqr(a.c0) - mul_xi(a.c1 * a.c2);
    F2 t1 = mul_xi(sqr(a.c2)) - a.c0 * a.c1;
    F2 t2 = sqr(a.c1) - a.c0 * a.c2;
    F2 d = inv(mul_xi(a.c2 * t1 + a.c1 * t2) + a.c0 * t0);
    Fp6T<F2> r;
    r.c0 = t0 * d; r.c1 = t1 * d; r.c2 = t2 * d;
    return r;
}

// ----------------------------------------------------------------------------- Fp12 = Fp6[w]/(w^2 - v)
template <class F2>
struct Fp12T { Fp6T<F2> c0, c1; };
template <class F2>
TCB_HD Fp12T<F2> fp12_one() {
    Fp12T<F2> r;
    r.c0.c0 = F2::one(); r.c0.c1 = F2::zero(); r.c0.c2 = F2::zero();
    r.c1.c0 = F2::zero(); r.c1.c1 = F2::zero(); r.c1.c2 = F2::zero();
    return r;
}
template <class F2>
TCB_HDN Fp12T<F2> fp12_mul(const Fp12T<F2> &a, const Fp12T<F2> &b) {
    Fp6T<F2> aa = fp6_mul(a.c0, b.c0);
    Fp6T<F2> bb = fp6_mul(a.c1, b.c1);
    Fp12T<F2> r;
    r.c1 = fp6_mul(a.c0 + a.c1, b.c0 + b.c1) - aa - bb;
    r.c0 = aa + mul_v(bb);
    return r;
}
template <class F2>
TCB_HDN Fp12T<F2> fp12_sqr(const Fp12T<F2> &a) {
    Fp6T<F2> ab = fp6_mul(a.c0, a.c1);
    Fp12T<F2> r;
    r.c0 = fp6_mul(a.c0 + a.c1, a.c0 + mul_v(a.c1)) - ab - mul_v(ab);
    r.c1 = ab + ab;
    return r;
}
template <class F2> TCB_HD Fp12T<F2> fp12_conj(const Fp12T<F2> &a) { Fp12T<F2> r; r.c0 = a.c0; r.c1 = -a.c1; return r; }
template <class F2>
TCB_HDN Fp12T<F2> fp12_inv(const Fp12T<F2> &a) {
    Fp6T<F2> t = fp6_inv(fp6_sqr(a.c0) - mul_v(fp6_sqr(a.c1)));
    Fp12T<F2> r;
    r.c0 = fp6_mul(a.c0, t);
    r.c1 = -fp6_mul(a.c1, t);
    return r;
}
// Frobenius p^k: the coefficient of v^i w^j (w-degree m = 2i + j) is conjugated k times and
// multiplied by frob[k][m].
template <class F2>
TCB_HDN Fp12T<F2> fp12_frob(const Fp12T<F2> &a, int k) {
    const Consts &C = CONSTS();
    Fp12T<F2> r;
    bool odd = k & 1;
    r.c0.c0 = odd ? conj(a.c0.c0) : a.c0.c0;
    r.c1.c0 = (odd ? conj(a.c1.c0) : a.c1.c0) * F2::load(C.frob[k][1]);
    r.c0.c1 = (odd ? conj(a.c0.c1) : a.c0.c1) * F2::load(C.frob[k][2]);
    r.c1.c1 = (odd ? conj(a.c1.c1) : a.c1.c1) * F2::load(C.frob[k][3]);
    r.c0.c2 = (odd ? conj(a.c0.c2) : a.c0.c2) * F2::load(C.frob[k][4]);
    r.c1.c2 = (odd ? conj(a.c1.c2) : a.c1.c2) * F2::load(C.frob[k][5]);
    return r;
}
// f * (l0 + l1 v + l4 v w)   — the sparse line element of the M-type twist
template <class F2>
TCB_HDN void fp12_mul_by_014(Fp12T<F2> &f, const F2 &l0, const F2 &l1, const F2 &l4) {
    Fp6T<F2> aa = fp6_mul_by_01(f.c0, l0, l1);
    Fp6T<F2> bb = fp6_mul_by_1(f.c1, l4);
    Fp6T<F2> s = fp6_mul_by_01(f.c0 + f.c1, l0, l1 + l4);
    f.c1 = s - aa - bb;
    f.c0 = mul_v(bb) + aa;
}
template <class F2>
TCB_HD bool fp12_is_one(const Fp12T<F2> &a) {
    bool z = is_zero(a.c0.c1);
    z = is_zero(a.c0.c2) && z;
    z = is_zero(a.c1.c0) && z;
    z = is_zero(a.c1.c1) && z;
    z = is_zero(a.c1.c2) && z;
    z = eq(a.c0.c0, F2::one()) && z;
    return z;
}
// Granger-Scott squaring, valid in the cyclotomic subgroup (after the easy part)
template <class F2>
TCB_HD void fp4_sqr(F2 &c0, F2 &c1, const F2 &a, const F2 &b) {
    F2 t0 = sqr(a), t1 = sqr(b);
    c0 = mul_xi(t1) + t0;
    c1 = sqr(a + b) - t0 - t1;
}
template <class F2>
TCB_HDN Fp12T<F2> fp12_cyclo_sqr(const Fp12T<F2> &f) {
    F2 z0 = f.c0.c0, z4 = f.c0.c1, z3 = f.c0.c2, z2 = f.c1.c0, z1 = f.c1.c1, z5 = f.c1.c2;
    F2 t0, t1, t2, t3;
    fp4_sqr(t0, t1, z0, z1);
    z0 = t0 - z0; z0 = z0 + z0 + t0;
    z1 = t1 + z1; z1 = z1 + z1 + t1;
    fp4_sqr(t0, t1, z2, z3);
    fp4_sqr(t2, t3, z4, z5);
    z4 = t0 - z4; z4 = z4 + z4 + t0;
    z5 = t1 + z5; z5 = z5 + z5 + t1;
    t0 = mul_xi(t3);
    z2 = t0 + z2; z2 = z2 + z2 + t0;
    z3 = t2 - z3; z3 = z3 + z3 + t2;
    Fp12T<F2> r;
    r.c0.c0 = z0; r.c0.c1 = z4; r.c0.c2 = z3;
    r.c1.c0 = z2; r.c1.c1 = z1; r.c1.c2 = z5;
    return r;
}
// f^|x| then conjugate (x < 0), for f in the cyclotomic subgroup
template <class F2>
TCB_HDN Fp12T<F2> fp12_exp_by_x(const Fp12T<F2> &f, u64 x) {
    Fp12T<F2> acc = f;
    int top = 63;
    while (!((x >> top) & 1)) top--;
    for (int i = top - 1; i >= 0; i--) {
        acc = fp12_cyclo_sqr(acc);
        if ((x >> i) & 1) acc = fp12_mul(acc, f);
    }
    return fp12_conj(acc);
}
// Karabina's compressed squarings, scalar-engine statement of what quad.cuh runs on the device (there spread over a lane
// quad): (z2, z3, z4, z5) square among themselves; z1 = (xi z5^2 + 3 z4^2 - 2 z3) / (4 z2), z0 = (2 z1^2 + z2 z5 - 3 z3 z4) xi + 1.
// Used by the host emulation to check the formulas and the fallback (tests/hostemu: tcb_emu_karabina_check).
template <class F2> struct Comp4 { F2 z2, z3, z4, z5; };
template <class F2>
TCB_HD Comp4<F2> fp12_compress(const Fp12T<F2> &f) { Comp4<F2> c; c.z2 = f.c1.c0; c.z3 = f.c0.c2; c.z4 = f.c0.c1; c.z5 = f.c1.c2; return c; }
template <class F2>
TCB_HDN Comp4<F2> fp12_comp_sqr(const Comp4<F2> &c) {
    F2 s2 = sqr(c.z2), s3 = sqr(c.z3), s4 = sqr(c.z4), s5 = sqr(c.z5);
    F2 c23 = sqr(c.z2 + c.z3) - s2 - s3;      // 2 z2 z3
    F2 c45 = sqr(c.z4 + c.z5) - s4 - s5;      // 2 z4 z5
    F2 t1 = s2 + mul_xi(s3), t2 = s4 + mul_xi(s5), u1 = mul_xi(c45);
    Comp4<F2> r;
    F2 d;
    d = t1 - c.z4; r.z4 = d + d + t1;
    d = t2 - c.z3; r.z3 = d + d + t2;
    d = u1 + c.z2; r.z2 = d + d + u1;
    d = c23 + c.z5; r.z5 = d + d + c23;
    return r;
}
// f^|x| then conjugate, through the compressed chain; falls back to the Granger-Scott loop when a saved z2 is zero
template <class F2>
TCB_HDN Fp12T<F2> fp12_exp_by_x_karabina(const Fp12T<F2> &f, u64 x) {
    int top = 63;
    while (!((x >> top) & 1)) top--;
    Comp4<F2> saved[64];
    int ns = 0;
    Comp4<F2> c = fp12_compress(f);
    for (int i = 1; i <= top; i++) {
        c = fp12_comp_sqr(c);
        if ((x >> i) & 1) saved[ns++] = c;
    }
    if (ns == 0) return fp12_exp_by_x(f, x);
    F2 pre[64], den[64];
    F2 run = F2::one();
    for (int k = 0; k < ns; k++) {
        den[k] = dbl(dbl(saved[k].z2));
        if (is_zero(den[k])) return fp12_exp_by_x(f, x);
        pre[k] = run;
        run = run * den[k];
    }
    F2 rinv = inv(run);
    Fp12T<F2> acc = fp12_one<F2>();
    for (int k = ns - 1; k >= 0; k--) {
        const Comp4<F2> &s = saved[k];
        F2 dinv = rinv * pre[k];
        rinv = rinv * den[k];
        F2 s4 = sqr(s.z4);
        F2 z1 = (mul_xi(sqr(s.z5)) + (dbl(s4) + s4) - dbl(s.z3)) * dinv;
        F2 m34 = s.z3 * s.z4;
        F2 z0 = mul_xi(dbl(sqr(z1)) + s.z2 * s.z5 - (dbl(m34) + m34)) + F2::one();
        Fp12T<F2> d;
        d.c0.c0 = z0; d.c0.c1 = s.z4; d.c0.c2 = s.z3;
        d.c1.c0 = s.z2; d.c1.c1 = z1; d.c1.c2 = s.z5;
        acc = (k == ns - 1) ? d : fp12_mul(acc, d);
    }
    if (x & 1) acc = fp12_mul(acc, f);
    return fp12_conj(acc);
}
// f^(3 (p^12 - 1)/r): easy part, then the hard part by the x-chain (same chain as EXTERNAL
// pairing 0.16 final_exponentiation, with cyclotomic squarings).  Only == 1 is ever observed.
template <class F2>
TCB_HDN Fp12T<F2> final_exponentiation(const Fp12T<F2> &in) {
    Fp12T<F2> f2 = fp12_inv(in);
    Fp12T<F2> r = fp12_mul(fp12_conj(in), f2);
    f2 = r;
    r = fp12_mul(fp12_frob(r, 2), f2);
    const u64 x = TCB_BLS_X;
    Fp12T<F2> y0 = fp12_cyclo_sqr(r);
    Fp12T<F2> y1 = fp12_exp_by_x(y0, x);
    Fp12T<F2> y2 = fp12_exp_by_x(y1, x >> 1);
    Fp12T<F2> y3 = fp12_conj(r);
    y1 = fp12_mul(y1, y3);
    y1 = fp12_conj(y1);
    y1 = fp12_mul(y1, y2);
    y2 = fp12_exp_by_x(y1, x);
    y3 = fp12_exp_by_x(y2, x);
    y1 = fp12_conj(y1);
    y3 = fp12_mul(y3, y1);
    y1 = fp12_conj(y1);
    y1 = fp12_frob(y1, 3);
    y2 = fp12_frob(y2, 2);
    y1 = fp12_mul(y1, y2);
    y2 = fp12_exp_by_x(y3, x);
    y2 = fp12_mul(y2, y0);
    y2 = fp12_mul(y2, r);
    y1 = fp12_mul(y1, y2);
    y2 = fp12_frob(y3, 1);
    return fp12_mul(y1, y2);
}

// ----------------------------------------------------------------------------- curves (Jacobian), F = Fp (G1) or F2 (G2)
TCB_HD bool is_zero(const Fp &a) { return a.is_zero(); }
TCB_HD bool eq(const Fp &a, const Fp &b) { return a == b; }
TCB_HD Fp inv(const Fp &a) { return fp_inv(a); }
TCB_HD Fp select(bool c, const Fp &a, const Fp &b) { return c ? a : b; }
template <class F> struct FieldOne;
template <> struct FieldOne<Fp> { TCB_HD static Fp one() { return fp_one(); } TCB_HD static Fp zero() { return Fp::zero(); } };
template <> struct FieldOne<Fp2> { TCB_HD static Fp2 one() { return Fp2::one(); } TCB_HD static Fp2 zero() { return Fp2::zero(); } };
#if defined(__CUDACC__)
template <> struct FieldOne<Fp2S> { TCB_D static Fp2S one() { return Fp2S::one(); } TCB_D static Fp2S zero() { return Fp2S::zero(); } };
#endif

template <class F> struct Aff { F x, y; bool inf; };
template <class F> struct Jac { F x, y, z; };   // z == 0 <=> infinity

template <class F> TCB_HD Jac<F> jac_inf() { Jac<F> r; r.x = FieldOne<F>::zero(); r.y = FieldOne<F>::one(); r.z = FieldOne<F>::zero(); return r; }
template <class F> TCB_HD bool jac_is_inf(const Jac<F> &p) { return is_zero(p.z); }
template <class F> TCB_HD Jac<F> jac_from_aff(const Aff<F> &a) {
    Jac<F> r;
    if (a.inf) return jac_inf<F>();
    r.x = a.x; r.y = a.y; r.z = FieldOne<F>::one();
    return r;
}
template <class F> TCB_HD Jac<F> jac_neg(const Jac<F> &p) { Jac<F> r = p; r.y = -p.y; return r; }
template <class F>
TCB_HDN Aff<F> jac_to_aff(const Jac<F> &p) {
    Aff<F> r;
    if (jac_is_inf(p)) { r.inf = true; r.x = FieldOne<F>::zero(); r.y = FieldOne<F>::zero(); return r; }
    F zi = inv(p.z), zi2 = sqr(zi);
    r.x = p.x * zi2; r.y = p.y * (zi2 * zi); r.inf = false;
    return r;
}
template <class F>
TCB_HDN Jac<F> jac_dbl(const Jac<F> &p) {   // dbl-2009-l (a = 0); also correct for infinity (z stays 0)
    F a = sqr(p.x), b = sqr(p.y), c = sqr(b);
    F d = dbl(sqr(p.x + b) - a - c);
    F e = dbl(a) + a;
    F f = sqr(e);
    Jac<F> r;
    r.z = dbl(p.y * p.z);
    r.x = f - dbl(d);
    r.y = e * (d - r.x) - dbl(dbl(dbl(c)));
    return r;
}
template <class F>
TCB_HDN Jac<F> jac_add_mixed(const Jac<F> &p, const Aff<F> &q) {   // madd-2007-bl + exceptional cases
    if (q.inf) return p;
    if (jac_is_inf(p)) return jac_from_aff(q);
    F z1z1 = sqr(p.z);
    F u2 = q.x * z1z1;
    F s2 = q.y * p.z * z1z1;
    if (eq(p.x, u2)) {
        if (eq(p.y, s2)) return jac_dbl(p);
        return jac_inf<F>();
    }
    F h = u2 - p.x, hh = sqr(h);
    F i = dbl(dbl(hh));
    F j = h * i;
    F rr = dbl(s2 - p.y);
    F v = p.x * i;
    Jac<F> r;
    r.x = sqr(rr) - j - dbl(v);
    r.y = rr * (v - r.x) - dbl(p.y * j);
    r.z = sqr(p.z + h) - z1z1 - hh;
    return r;
}
template <class F>
TCB_HDN Jac<F> jac_add(const Jac<F> &p, const Jac<F> &q) {   // add-2007-bl + exceptional cases
    if (jac_is_inf(p)) return q;
    if (jac_is_inf(q)) return p;
    F z1z1 = sqr(p.z), z2z2 = sqr(q.z);
    F u1 = p.x * z2z2, u2 = q.x * z1z1;
    F s1 = p.y * q.z * z2z2, s2 = q.y * p.z * z1z1;
    if (eq(u1, u2)) {
        if (eq(s1, s2)) return jac_dbl(p);
        return jac_inf<F>();
    }
    F h = u2 - u1;
    F i = sqr(dbl(h));
    F j = h * i;
    F rr = dbl(s2 - s1);
    F v = u1 * i;
    Jac<F> r;
    r.x = sqr(rr) - j - dbl(v);
    r.y = rr * (v - r.x) - dbl(s1 * j);
    r.z = (sqr(p.z + q.z) - z1z1 - z2z2) * h;
    return r;
}
// p + q with q's z^2 and z^3 supplied (a base that is added many times: Horner steps of Commitment::evaluate): 10M + 4S
template <class F>
TCB_HDN Jac<F> jac_add_cached(const Jac<F> &p, const Jac<F> &q, const F &z2z2, const F &z2c) {
    if (jac_is_inf(p)) return q;
    if (jac_is_inf(q)) return p;
    F z1z1 = sqr(p.z);
    F u1 = p.x * z2z2, u2 = q.x * z1z1;
    F s1 = p.y * z2c, s2 = q.y * p.z * z1z1;
    if (eq(u1, u2)) {
        if (eq(s1, s2)) return jac_dbl(p);
        return jac_inf<F>();
    }
    F h = u2 - u1;
    F i = sqr(dbl(h));
    F j = h * i;
    F rr = dbl(s2 - s1);
    F v = u1 * i;
    Jac<F> r;
    r.x = sqr(rr) - j - dbl(v);
    r.y = rr * (v - r.x) - dbl(s1 * j);
    r.z = (sqr(p.z + q.z) - z1z1 - z2z2) * h;
    return r;
}
// k * P for an affine base, k given as NL little-endian u32 limbs (canonical integer).
// MSB-first double-and-add; the group element is unique whatever the algorithm (App. A).
template <class F, int NL>
TCB_HDN Jac<F> jac_mul_aff(const Aff<F> &p, const u32 *k) {
    Jac<F> acc = jac_inf<F>();
    bool started = false;
    for (int i = NL * 32 - 1; i >= 0; i--) {
        if (started) acc = jac_dbl(acc);
        if ((k[i >> 5] >> (i & 31)) & 1) { acc = jac_add_mixed(acc, p); started = true; }
    }
    return acc;
}
template <class F, class E>
TCB_HDN Jac<F> jac_mul_const(const Aff<F> &p) {
    Jac<F> acc = jac_inf<F>();
    bool started = false;
    for (int i = E::N * 32 - 1; i >= 0; i--) {
        if (started) acc = jac_dbl(acc);
        if ((E::get(i >> 5) >> (i & 31)) & 1) { acc = jac_add_mixed(acc, p); started = true; }
    }
    return acc;
}
// k * P for a Jacobian base (Commitment::evaluate's `result.mul_assign(x)`)
template <class F, int NL>
TCB_HDN Jac<F> jac_mul_jac(const Jac<F> &p, const u32 *k) {
    Jac<F> acc = jac_inf<F>();
    bool started = false;
    for (int i = NL * 32 - 1; i >= 0; i--) {
        if (started) acc = jac_dbl(acc);
        if ((k[i >> 5] >> (i & 31)) & 1) { acc = started ? jac_add(acc, p) : p; started = true; }
    }
    return acc;
}

// ----------------------------------------------------------------------------- psi endomorphism and small-scalar multipliers
template <class F2>
TCB_HD Jac<F2> jac_psi(const Jac<F2> &p) {
    const Consts &C = CONSTS();
    Jac<F2> r;
    r.x = conj(p.x) * F2::load(C.psi_x);
    r.y = conj(p.y) * F2::load(C.psi_y);
    r.z = conj(p.z);
    return r;
}
// |k| * P for a 64-bit constant, Jacobian base (MSB-first double-and-add)
template <class F>
TCB_HDN Jac<F> jac_mul_u64(const Jac<F> &p, u64 k) {
    Jac<F> acc = p;
    int top = 63;
    while (top > 0 && !((k >> top) & 1)) top--;
    for (int i = top - 1; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((k >> i) & 1) acc = jac_add(acc, p);
    }
    return acc;
}
// the same for an affine base: every addition is a mixed addition
template <class F>
TCB_HDN Jac<F> jac_mul_u64_aff(const Aff<F> &p, u64 k) {
    Jac<F> acc = jac_from_aff(p);
    int top = 63;
    while (top > 0 && !((k >> top) & 1)) top--;
    for (int i = top - 1; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((k >> i) & 1) acc = jac_add_mixed(acc, p);
    }
    return acc;
}
// sum_j (plus_j - minus_j) 2^j * P, digits given as NAF bit masks
template <class F>
TCB_HDN Jac<F> jac_mul_naf64(const Jac<F> &p, u64 plus, u64 minus) {
    Jac<F> acc = jac_inf<F>();
    Jac<F> np = jac_neg(p);
    for (int i = 63; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((plus >> i) & 1) acc = jac_add(acc, p);
        if ((minus >> i) & 1) acc = jac_add(acc, np);
    }
    return acc;
}
// Exact multiplication by the G2 cofactor h2 (EXTERNAL pairing scale_by_cofactor) without a
// 507-bit scalar: with x the (negative) curve parameter and psi the endomorphism,
//   Q0 = h_eff P = [x^2-x-1]P + [x-1]psi(P) + psi^2(2P)        (Budroni-Pintore, h_eff = 3(x^2-1) h2)
//   [h2]P = [1/(3(x^2-1)) mod r] Q0,  and on G2 (psi = [x]):  1/(x^2-1) = -x^2,  1/3 = 1 + 2 x^2 (x+1) (x-1)/3
//   => R = -psi^2(Q0),  T = psi^2(psi(R) + R),  [h2]P = R - [w] T,  w = 2(|x|+1)/3 = 0x8c00aaaaaaab5556.
// Same group element as the reference's 507-bit double-and-add (checked against it in the tests).
// exact == false stops at Q0 = [3 (x^2 - 1)] [h2]P: a verifier that only needs e(pk, H) == e(g1, sig) can test
// e(pk, Q0) == e([3 (x^2 - 1)] g1, sig) instead (3 (x^2 - 1) is a unit mod r, so the two equalities are equivalent) and
// skip the third 64-bit multiplication; [3 (x^2 - 1)] g1 is the constant Consts::g1cx/g1cy.
#define TCB_W_PLUS 0x9001000000000000ULL
#define TCB_W_MINUS 0x040055555554aaaaULL
template <class F2>
TCB_HDN Jac<F2> g2_clear_cofactor(const Aff<F2> &p, bool exact = true) {
    const u64 X = TCB_BLS_X;
    Jac<F2> pj = jac_from_aff(p);
    Jac<F2> t1 = jac_neg(jac_mul_u64_aff(p, X));          // [x]P
    Jac<F2> t2 = jac_psi(pj);
    Jac<F2> t3 = jac_psi(jac_psi(jac_dbl(pj)));
    t3 = jac_add(t3, jac_neg(t2));
    t2 = jac_add(t1, t2);
    t2 = jac_neg(jac_mul_u64(t2, X));                      // [x](…)
    t3 = jac_add(t3, t2);
    t3 = jac_add(t3, jac_neg(t1));
    Jac<F2> q0 = jac_add(t3, jac_neg(pj));
    if (!exact) return q0;
    Jac<F2> r = jac_neg(jac_psi(jac_psi(q0)));
    Jac<F2> t = jac_psi(jac_psi(jac_add(jac_psi(r), r)));
    return jac_add(r, jac_neg(jac_mul_naf64(t, TCB_W_PLUS, TCB_W_MINUS)));
}

// ----------------------------------------------------------------------------- GLS / GLV scalar multiplication
// SIMT note: with one scalar per lane (pair), a data-dependent "if (bit) add" diverges across
// the warp and every bit position ends up paying for an addition.  The multipliers below use
// *regular* signed recoding (Faz-Hernandez, Longa, Sanchez): every step is one doubling and
// exactly one table addition, so no slot is wasted.
//
// k < 2^256 canonical little-endian limbs.  Returns k mod r in place (at most two subtractions).
TCB_HD void scalar_reduce(u32 *k) {
    for (int rep = 0; rep < 2; rep++) {
        if (limbs_lt_mod<FrParams>(k)) return;
        u64 bw = 0;
        for (int i = 0; i < 8; i++) {
            u64 d = (u64)k[i] - FrParams::mod(i) - bw;
            k[i] = (u32)d; bw = (d >> 63) & 1;
        }
    }
}
// q (8 limbs) := q / d, returns q % d, for a 64-bit divisor with its top bit set
TCB_HD u64 divmod_u64(u32 *q, u64 d) {
    u64 rem = 0;
    for (int i = 255; i >= 0; i--) {
        u32 hi = (u32)(rem >> 63);
        rem = (rem << 1) | ((q[i >> 5] >> (i & 31)) & 1u);
        u32 bit = (hi || rem >= d) ? 1u : 0u;
        if (bit) rem -= d;
        q[i >> 5] = (q[i >> 5] & ~(1u << (i & 31))) | (bit << (i & 31));
    }
    return rem;
}
// psi^i applied to an affine point, i in 0..3
template <class F2>
TCB_HD Aff<F2> aff_psi_i(const Aff<F2> &p, int i) {
    const Consts &C = CONSTS();
    Aff<F2> r;
    r.inf = p.inf;
    bool odd = i & 1;
    r.x = (odd ? conj(p.x) : p.x) * F2::load(C.psi_cx[i]);
    r.y = (odd ? conj(p.y) : p.y) * F2::load(C.psi_cy[i]);
    return r;
}
// k * P on G2 (P of order r) through the 4-dimensional decomposition k = sum_i d_i X^i,
// X = |x| = -lambda_psi, i.e. k P = sum_i (-1)^i d_i psi^i(P); 66 doublings + 66 additions from
// an 8-entry table  T[b3 b2 b1] = P0 + b1 P1 + b2 P2 + b3 P3,  P_i = (-1)^i psi^i(P).
// The three pieces (recoding, table, main loop) are separate so that a multi-scalar
// multiplication can share the 66 doublings between several points (scheme.cuh, *_msm_*).
//
// Recoded scalar: digit j (0..65) is a 4-bit nibble, bit 0 = sign of the a0 digit (1 = -1), bits 1-3 = table index.
struct Gls4Digits {
    u32 nib[9];
    u32 flags;   // bit 0: a0 was even (subtract P at the end); bit 1: the point is the point at infinity (skip)
    TCB_HD u32 digit(int j) const { return (nib[j >> 3] >> ((j & 7) * 4)) & 15u; }
};
constexpr int GLS4_L = 65;   // digits 0..GLS4_L
TCB_HD void gls4_recode(const u32 *k_in, Gls4Digits &dg) {
    u32 k[8];
    for (int i = 0; i < 8; i++) k[i] = k_in[i];
    scalar_reduce(k);
    u64 a[4];
    u32 ahi0 = 0;                   // bit 64 of a0 (only a0 + 1 can carry)
    const u64 X = TCB_BLS_X;
    a[0] = divmod_u64(k, X);
    a[1] = divmod_u64(k, X);
    a[2] = divmod_u64(k, X);
    a[3] = (u64)k[0] | ((u64)k[1] << 32);        // k < r < X^4  =>  the last quotient fits 64 bits
    bool even = !(a[0] & 1);
    if (even) { a[0] += 1; if (a[0] == 0) ahi0 = 1; }     // make a0 odd; fixed up at the end
    for (int i = 0; i < 9; i++) dg.nib[i] = 0;
    dg.flags = even ? 1u : 0u;
    // b0[j] = 2 * bit_{j+1}(a0) - 1 for j < L, b0[L] = +1: sign bit j = !bit_{j+1}(a0)
    u64 a0s = (a[0] >> 1) | ((u64)ahi0 << 63);       // bits 1..64 of a0 at positions 0..63
    u64 av[3] = {a[1], a[2], a[3]};
    u32 avh[3] = {0, 0, 0};
    for (int j = 0; j <= GLS4_L; j++) {
        bool neg = j < 64 ? !((a0s >> j) & 1) : (j == 64);   // j = 64: bit 65 of a0 is 0 -> -1 ; j = 65 (top) is +1
        u32 d = neg ? 1u : 0u;
        for (int t = 0; t < 3; t++) {
            u32 bit = (u32)(av[t] & 1);
            d |= bit << (t + 1);
            // a = floor(a / 2) - floor(b / 2),  b = bit * (+1 | -1):  b = -1 -> floor(-1/2) = -1 -> +1
            av[t] = (av[t] >> 1) | ((u64)avh[t] << 63);
            avh[t] = 0;
            if (bit && neg && j < GLS4_L) { av[t] += 1; if (av[t] == 0) avh[t] = 1; }
        }
        dg.nib[j >> 3] |= d << ((j & 7) * 4);
    }
}
template <class F2>
TCB_HD void gls4_table(const Aff<F2> &p, Jac<F2> *T) {
    Aff<F2> P1 = aff_psi_i(p, 1), P2 = aff_psi_i(p, 2), P3 = aff_psi_i(p, 3);
    P1.y = -P1.y; P3.y = -P3.y;
    T[0] = jac_from_aff(p);
    T[1] = jac_add_mixed(T[0], P1);
    T[2] = jac_add_mixed(T[0], P2);
    T[3] = jac_add_mixed(T[1], P2);
    T[4] = jac_add_mixed(T[0], P3);
    T[5] = jac_add_mixed(T[1], P3);
    T[6] = jac_add_mixed(T[2], P3);
    T[7] = jac_add_mixed(T[3], P3);
}
// Simultaneous conversion of N Jacobian points to affine with ONE field inversion (Montgomery's
// trick).  None of the inputs may be the point at infinity.
template <class F, int N>
TCB_HD void jac_batch_to_aff(const Jac<F> *in, Aff<F> *out) {
    F pre[N];
    pre[0] = in[0].z;
    for (int i = 1; i < N; i++) pre[i] = pre[i - 1] * in[i].z;
    F acc = inv(pre[N - 1]);
    for (int i = N - 1; i >= 0; i--) {
        F zi = i ? acc * pre[i - 1] : acc;
        if (i) acc = acc * in[i].z;
        F zi2 = sqr(zi);
        out[i].x = in[i].x * zi2;
        out[i].y = in[i].y * (zi2 * zi);
        out[i].inf = false;
    }
}
// the same table normalised to affine coordinates with ONE inversion, so that every addition of the main
// loop is a mixed addition (7M + 4S instead of 11M + 5S)
template <class F2>
TCB_HD void gls4_table_affine(const Aff<F2> &p, Aff<F2> *A) {
    Jac<F2> T[8];
    gls4_table(p, T);
    A[0] = p;
    jac_batch_to_aff<F2, 7>(T + 1, A + 1);
}
template <class F2>
TCB_HDN Jac<F2> jac_mul_gls4(const Aff<F2> &p, const u32 *k_in) {
    if (p.inf) return jac_inf<F2>();
    Gls4Digits dg;
    gls4_recode(k_in, dg);
    Aff<F2> A[8];
    gls4_table_affine(p, A);
    Jac<F2> acc = jac_inf<F2>();
    for (int j = GLS4_L; j >= 0; j--) {
        acc = jac_dbl(acc);
        u32 d = dg.digit(j);
        Aff<F2> t = A[d >> 1];
        if (d & 1) t.y = -t.y;
        acc = jac_add_mixed(acc, t);
    }
    if (dg.flags & 1) { Aff<F2> np = p; np.y = -p.y; acc = jac_add_mixed(acc, np); }
    return acc;
}
// q (8 limbs) := q / d for a 128-bit divisor d = (dhi, dlo) with its top bit set; remainder in (rhi, rlo)
TCB_HD void divmod_u128(u32 *q, u64 dhi, u64 dlo, u64 &rhi, u64 &rlo) {
    rhi = 0; rlo = 0;
    for (int i = 255; i >= 0; i--) {
        u32 carry = (u32)(rhi >> 63);
        rhi = (rhi << 1) | (rlo >> 63);
        rlo = (rlo << 1) | ((q[i >> 5] >> (i & 31)) & 1u);
        bool ge = carry || rhi > dhi || (rhi == dhi && rlo >= dlo);
        if (ge) { u64 b = rlo < dlo; rlo -= dlo; rhi = rhi - dhi - b; }
        q[i >> 5] = (q[i >> 5] & ~(1u << (i & 31))) | ((ge ? 1u : 0u) << (i & 31));
    }
}
// k * P on G1 (P of order r): k = a + b X^2 and [X^2]P = -phi(P), phi(x, y) = (beta x, y);
// regular recoding on (a, b): 130 doublings + 130 additions from {P0, P0 + P1}, P1 = -phi(P).
// Recoded scalar: digit j (0..129) is 2 bits, bit 0 = sign (1 = -1), bit 1 = table index.
struct Glv2Digits {
    u32 w[9];
    u32 flags;   // bit 0: a was even (subtract P at the end); bit 1: point at infinity (skip)
    TCB_HD u32 digit(int j) const { return (w[j >> 4] >> ((j & 15) * 2)) & 3u; }
};
constexpr int GLV2_L = 129;   // digits 0..GLV2_L
TCB_HD void glv2_recode(const u32 *k_in, Glv2Digits &dg) {
    u32 k[8];
    for (int i = 0; i < 8; i++) k[i] = k_in[i];
    scalar_reduce(k);
    // X^2 = 0xac45a4010001a402 0000000100000000 (128 bit, top bit set)
    const u64 MU_HI = 0xac45a4010001a402ULL, MU_LO = 0x0000000100000000ULL;
    u64 ahi, alo;
    divmod_u128(k, MU_HI, MU_LO, ahi, alo);
    u64 blo = (u64)k[0] | ((u64)k[1] << 32), bhi = (u64)k[2] | ((u64)k[3] << 32);
    bool even = !(alo & 1);
    u32 atop = 0;
    if (even) { alo += 1; if (alo == 0) { ahi += 1; if (ahi == 0) atop = 1; } }
    for (int i = 0; i < 9; i++) dg.w[i] = 0;
    dg.flags = even ? 1u : 0u;
    // bits 1..129 of a at positions 0..128; sign bit j = !bit_{j+1}(a), top digit +1
    u64 s0 = (alo >> 1) | (ahi << 63), s1 = (ahi >> 1) | ((u64)atop << 63);
    u64 b0 = blo, b1 = bhi; u32 b2 = 0;
    for (int j = 0; j <= GLV2_L; j++) {
        bool neg = j < 64 ? !((s0 >> j) & 1) : (j < 128 ? !((s1 >> (j - 64)) & 1) : (j == 128));
        u32 bit = (u32)(b0 & 1);
        dg.w[j >> 4] |= ((neg ? 1u : 0u) | (bit << 1)) << ((j & 15) * 2);
        b0 = (b0 >> 1) | (b1 << 63); b1 = (b1 >> 1) | ((u64)b2 << 63); b2 = 0;
        if (bit && neg && j < GLV2_L) { b0 += 1; if (b0 == 0) { b1 += 1; if (b1 == 0) b2 = 1; } }
    }
}
// Two digit positions per table look-up (width-2 window on the regular recoding): positions (2k, 2k+1) contribute
//   s0 (P + i0 P1) + 2 s1 (P + i1 P1) = s1 [ (2 + sg) P + (2 i1 + sg i0) P1 ],   sg = s0 s1,
// i.e. one of 8 points  E[4 n + 2 i1 + i0],  n = (sg == -1):  {3P + k P1, k = 0..3}  and  {P, P - P1, P + 2 P1, P + P1},
// times the sign s1.  65 additions instead of 130 per scalar for a table of 8 affine points (1 doubling + 7 mixed additions
// and one shared inversion to build).
constexpr int GLV2_W = 64;   // window positions 0..GLV2_W
TCB_HD u32 glv2_wdigit(const Glv2Digits &dg, int k) { return (dg.w[k >> 3] >> ((k & 7) * 4)) & 15u; }   // s0 | i0 << 1 | s1 << 2 | i1 << 3
TCB_HD u32 glv2_windex(u32 nib) { return ((((nib ^ (nib >> 2)) & 1u)) << 2) | ((nib >> 2) & 2u) | ((nib >> 1) & 1u); }
TCB_HD bool glv2_wneg(u32 nib) { return (nib & 4u) != 0; }
TCB_HD void glv2_table8(const Aff<Fp> &p, Aff<Fp> *E) {
    Aff<Fp> P1, nP1;
    P1.x = p.x * CONSTS().beta; P1.y = -p.y; P1.inf = false;     // P1 = -phi(P) = [X^2] P
    nP1 = P1; nP1.y = p.y;
    Jac<Fp> J[7];
    Jac<Fp> pj = jac_from_aff(p);
    J[0] = jac_add_mixed(jac_dbl(pj), p);       // 3P
    J[1] = jac_add_mixed(J[0], P1);             // 3P + P1
    J[2] = jac_add_mixed(J[1], P1);             // 3P + 2 P1
    J[3] = jac_add_mixed(J[2], P1);             // 3P + 3 P1
    J[4] = jac_add_mixed(pj, nP1);              // P - P1
    J[6] = jac_add_mixed(pj, P1);               // P + P1
    J[5] = jac_add_mixed(J[6], P1);             // P + 2 P1
    Aff<Fp> A[7];
    jac_batch_to_aff<Fp, 7>(J, A);
    E[0] = A[0]; E[1] = A[1]; E[2] = A[2]; E[3] = A[3];
    E[4] = p; E[5] = A[4]; E[6] = A[5]; E[7] = A[6];
}
TCB_HDN Jac<Fp> jac_mul_glv2(const Aff<Fp> &p, const u32 *k_in) {
    if (p.inf) return jac_inf<Fp>();
    Glv2Digits dg;
    glv2_recode(k_in, dg);
    Aff<Fp> E[8];
    glv2_table8(p, E);
    Jac<Fp> acc = jac_inf<Fp>();
    for (int k = GLV2_W; k >= 0; k--) {
        acc = jac_dbl(jac_dbl(acc));
        u32 nib = glv2_wdigit(dg, k);
        Aff<Fp> t = E[glv2_windex(nib)];
        if (glv2_wneg(nib)) t.y = -t.y;
        acc = jac_add_mixed(acc, t);
    }
    if (dg.flags & 1) { Aff<Fp> np = p; np.y = -p.y; acc = jac_add_mixed(acc, np); }
    return acc;
}

// ----------------------------------------------------------------------------- Miller loop (M-type twist, projective lines)
// Running point in HOMOGENEOUS projective coordinates (x = X/Z, y = Y/Z) with the Costello-Lange-Naehrig
// doubling (3M + 6S) and mixed addition (11M + 2S); line = a * yP * (v w) + b * xP * v + c.  The reference's
// Jacobian steps (EXTERNAL pairing 0.16) cost 3M + 8S per doubling; the two differ by factors in Fp2 per
// line, which the final exponentiation removes, and only "== 1" is ever observed.
template <class F2> struct Line { F2 a, b, c; };
template <class F2>
TCB_HDN Line<F2> doubling_step(Jac<F2> &r) {
    F2 a = half(r.x * r.y);
    F2 b = sqr(r.y), c = sqr(r.z);
    F2 c3 = dbl(c) + c;
    F2 e = mul_xi(dbl(dbl(c3)));            // 3 b' c, b' = 4 (1 + u)
    F2 f = dbl(e) + e;
    F2 g = half(b + f);
    F2 h = sqr(r.y + r.z) - (b + c);
    F2 j = sqr(r.x);
    F2 e2 = sqr(e);
    Line<F2> l;
    l.c = e - b;
    l.b = dbl(j) + j;
    l.a = -h;
    r.x = a * (b - f);
    r.y = sqr(g) - (dbl(e2) + e2);
    r.z = b * h;
    return l;
}
template <class F2>
TCB_HDN Line<F2> addition_step(Jac<F2> &r, const Aff<F2> &q) {
    F2 theta = r.y - q.y * r.z;
    F2 lambda = r.x - q.x * r.z;
    F2 c = sqr(theta), d = sqr(lambda);
    F2 e = lambda * d;
    F2 f = r.z * c;
    F2 g = r.x * d;
    F2 h = e + f - dbl(g);
    Line<F2> l;
    l.c = theta * q.x - lambda * q.y;
    l.b = -theta;
    l.a = lambda;
    r.x = lambda * h;
    r.y = theta * (g - h) - e * r.y;
    r.z = r.z * e;
    return l;
}
template <class F2>
TCB_HD void ell(Fp12T<F2> &f, const Line<F2> &l, const Aff<Fp> &p) {
    fp12_mul_by_014(f, l.c, mul_fp(l.b, p.x), mul_fp(l.a, p.y));
}
// Product of up to two Miller loops with one shared accumulator.  A pair with an infinity
// operand contributes 1 (A8).
template <class F2>
TCB_HDN Fp12T<F2> miller_loop2(const Aff<Fp> &p0, const Aff<F2> &q0, const Aff<Fp> &p1, const Aff<F2> &q1) {
    Fp12T<F2> f = fp12_one<F2>();
    bool act0 = !(p0.inf || q0.inf), act1 = !(p1.inf || q1.inf);
    Jac<F2> r0 = jac_from_aff(q0), r1 = jac_from_aff(q1);
    const u64 xs = TCB_BLS_X >> 1;
    for (int i = 61; i >= 0; i--) {   // bit 62 is the leading one of x >> 1
        bool bit = (xs >> i) & 1;
        if (act0) { Line<F2> l = doubling_step(r0); ell(f, l, p0); }
        if (act1) { Line<F2> l = doubling_step(r1); ell(f, l, p1); }
        if (bit) {
            if (act0) { Line<F2> l = addition_step(r0, q0); ell(f, l, p0); }
            if (act1) { Line<F2> l = addition_step(r1, q1); ell(f, l, p1); }
        }
        f = fp12_sqr(f);
    }
    if (act0) { Line<F2> l = doubling_step(r0); ell(f, l, p0); }
    if (act1) { Line<F2> l = doubling_step(r1); ell(f, l, p1); }
    return fp12_conj(f);
}
// e(a,b) == e(c,d)
template <class F2>
TCB_HD bool pairing_eq(const Aff<Fp> &a, const Aff<F2> &b, const Aff<Fp> &c, const Aff<F2> &d) {
    Aff<Fp> nc = c;
    nc.y = -c.y;
    Fp12T<F2> f = miller_loop2(a, b, nc, d);
    return fp12_is_one(final_exponentiation(f));
}

}  // namespace tcb
