// scheme.cuh — per-item "tasks" of the threshold_crypto hot path,
// written once as __host__ __device__ templates over the Fp2 engine.  kernels.cu launches
// them one item per thread (Fp2) or one item per lane pair (Fp2S); tests/hostemu runs the
// scalar instantiation on the CPU for logic checks only.
//
// Byte formats are the C-ABI's (include/tcb200.h): EXTERNAL pairing's uncompressed affine
// big-endian encodings and canonical little-endian Fr (src/serde_impl.rs:109,296).
#pragma once
#include "tower.cuh"

namespace tcb {

template <class F2> TCB_HD bool is_writer() {
#if defined(__CUDA_ARCH__)
    return !F2::SLICED || (threadIdx.x & 1u) == 0;
#else
    return true;
#endif
}
template <class F2> TCB_HD u32 my_role() {
#if defined(__CUDA_ARCH__)
    return F2::SLICED ? (threadIdx.x & 1u) : 0u;
#else
    return 0u;
#endif
}

// ----------------------------------------------------------------------------- codecs
// 48-byte big-endian canonical -> Montgomery.  ok=false if the integer is >= p.
// On the device a 16-byte aligned source is read as three 128-bit vectors and byte-swapped with PRMT (the ABI's arrays
// are 16-byte aligned whenever the caller's base pointer is: every record size is a multiple of 16); anything else (and the
// host emulation) takes the byte loop.
TCB_HD u32 bswap32(u32 v) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0, 0x0123);
#else
    return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
#endif
}
TCB_HD void load_limbs_be(u32 *l, const u8 *b) {
#if defined(__CUDA_ARCH__)
    if ((((size_t)b) & 15) == 0) {
        const uint4 *p = (const uint4 *)b;
        uint4 v0 = p[0], v1 = p[1], v2 = p[2];
        l[11] = bswap32(v0.x); l[10] = bswap32(v0.y); l[9] = bswap32(v0.z); l[8] = bswap32(v0.w);
        l[7] = bswap32(v1.x); l[6] = bswap32(v1.y); l[5] = bswap32(v1.z); l[4] = bswap32(v1.w);
        l[3] = bswap32(v2.x); l[2] = bswap32(v2.y); l[1] = bswap32(v2.z); l[0] = bswap32(v2.w);
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const u8 *q = b + (11 - i) * 4;
        l[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
    }
}
TCB_HD void store_limbs_be(u8 *b, const u32 *l) {
#if defined(__CUDA_ARCH__)
    if ((((size_t)b) & 15) == 0) {
        uint4 *p = (uint4 *)b;
        p[0] = make_uint4(bswap32(l[11]), bswap32(l[10]), bswap32(l[9]), bswap32(l[8]));
        p[1] = make_uint4(bswap32(l[7]), bswap32(l[6]), bswap32(l[5]), bswap32(l[4]));
        p[2] = make_uint4(bswap32(l[3]), bswap32(l[2]), bswap32(l[1]), bswap32(l[0]));
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u8 *q = b + (11 - i) * 4;
        q[0] = (u8)(l[i] >> 24); q[1] = (u8)(l[i] >> 16); q[2] = (u8)(l[i] >> 8); q[3] = (u8)l[i];
    }
}
TCB_HD Fp load_fp_be(const u8 *b, bool &ok) {
    Fp t;
    load_limbs_be(t.l, b);
    ok = ok && limbs_lt_mod<FpParams>(t.l);
    return fp_to_mont(t);
}
TCB_HD void store_fp_be(u8 *b, const Fp &a) {
    Fp t = fp_from_mont(a);
    store_limbs_be(b, t.l);
}
// G1 uncompressed 96 B: x || y, byte0 bit 0x40 = infinity
TCB_HD Aff<Fp> load_g1(const u8 *b, bool &ok) {
    Aff<Fp> a;
    if (b[0] & 0x40) { a.inf = true; a.x = Fp::zero(); a.y = Fp::zero(); return a; }
    a.inf = false;
    a.x = load_fp_be(b, ok);
    a.y = load_fp_be(b + 48, ok);
    return a;
}
TCB_HD void store_g1(u8 *b, const Aff<Fp> &a) {
    if (a.inf) { for (int i = 0; i < 96; i++) b[i] = 0; b[0] = 0x40; return; }
    store_fp_be(b, a.x);
    store_fp_be(b + 48, a.y);
}
// Fp2 as c1 || c0 (96 B); a sliced engine only touches its own half
template <class F2>
TCB_HD F2 load_f2_be(const u8 *b, bool &ok) {
    if (F2::SLICED) {
        Fp h = load_fp_be(b + (my_role<F2>() ? 0 : 48), ok);
        return F2::from_halves(h, h);
    }
    Fp c1 = load_fp_be(b, ok), c0 = load_fp_be(b + 48, ok);
    return F2::from_halves(c0, c1);
}
template <class F2>
TCB_HD void store_f2_be(u8 *b, const F2 &a) {
    if (F2::SLICED) {
        // each lane writes its own half
        Fp2c t; a.store(t);
        if (my_role<F2>()) store_fp_be(b, t.c1); else store_fp_be(b + 48, t.c0);
        return;
    }
    Fp2c t; a.store(t);
    store_fp_be(b, t.c1);
    store_fp_be(b + 48, t.c0);
}
// G2 uncompressed 192 B: x.c1 || x.c0 || y.c1 || y.c0
template <class F2>
TCB_HD Aff<F2> load_g2(const u8 *b, bool &ok) {
    Aff<F2> a;
    if (b[0] & 0x40) { a.inf = true; a.x = F2::zero(); a.y = F2::zero(); return a; }
    a.inf = false;
    a.x = load_f2_be<F2>(b, ok);
    a.y = load_f2_be<F2>(b + 96, ok);
    return a;
}
template <class F2>
TCB_HD void store_g2(u8 *b, const Aff<F2> &a) {
    if (a.inf) {
        if (is_writer<F2>()) { for (int i = 0; i < 192; i++) b[i] = 0; b[0] = 0x40; }
        return;
    }
    store_f2_be<F2>(b, a.x);
    store_f2_be<F2>(b + 96, a.y);
}
// canonical LE 32-byte scalar -> 8 u32 limbs (no reduction; caller guarantees < r or accepts k mod group order)
TCB_HD void load_scalar_le(u32 *k, const u8 *b) {
#if defined(__CUDA_ARCH__)
    if ((((size_t)b) & 15) == 0) {       // two 128-bit loads
        const uint4 *p = (const uint4 *)b;
        uint4 v0 = p[0], v1 = p[1];
        k[0] = v0.x; k[1] = v0.y; k[2] = v0.z; k[3] = v0.w; k[4] = v1.x; k[5] = v1.y; k[6] = v1.z; k[7] = v1.w;
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = (u32)b[4 * i] | ((u32)b[4 * i + 1] << 8) | ((u32)b[4 * i + 2] << 16) | ((u32)b[4 * i + 3] << 24);
}
// G1 compressed 48 B: flags 0x80 compressed, 0x40 infinity, 0x20 y > -y
TCB_HD void g1_compress(u8 *out, const Aff<Fp> &a) {
    if (a.inf) { for (int i = 0; i < 48; i++) out[i] = 0; out[0] = 0xc0; return; }
    store_fp_be(out, a.x);
    out[0] |= 0x80;
    if (fp_cmp(a.y, -a.y) > 0) out[0] |= 0x20;
}

// ----------------------------------------------------------------------------- SHA3-256 (FIPS 202) and ChaCha20 word stream
TCB_HD u64 rotl64(u64 v, int n) { return n ? ((v << n) | (v >> (64 - n))) : v; }
TCB_HDN void keccak_f(u64 *s) {
    const u64 RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int rnd = 0; rnd < 24; rnd++) {
        u64 c[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) {
            u64 d = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
            for (int y = 0; y < 25; y += 5) s[y + x] ^= d;
        }
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(s[x + 5 * y], ROT[x + 5 * y]);
        for (int y = 0; y < 25; y += 5)
            for (int x = 0; x < 5; x++) s[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
        s[0] ^= RC[rnd];
    }
}
// SHA3-256 of the concatenation m1 || m2 (either may be empty); src/util.rs:3-9
TCB_HDN void sha3_256(const u8 *m1, size_t l1, const u8 *m2, size_t l2, u8 *out) {
    u64 s[25];
    for (int i = 0; i < 25; i++) s[i] = 0;
    size_t pos = 0;   // byte position inside the 136-byte rate block
    for (size_t i = 0; i < l1 + l2; i++) {
        u8 v = i < l1 ? m1[i] : m2[i - l1];
        s[pos >> 3] ^= (u64)v << (8 * (pos & 7));
        if (++pos == 136) { keccak_f(s); pos = 0; }
    }
    s[pos >> 3] ^= (u64)0x06 << (8 * (pos & 7));
    s[16] ^= 0x8000000000000000ULL;
    keccak_f(s);
    for (int i = 0; i < 32; i++) out[i] = (u8)(s[i >> 3] >> (8 * (i & 7)));
}
// rand_chacha 0.2 ChaChaRng::from_seed + BlockRng word stream
struct ChaChaRng {
    u32 key[8];
    u32 buf[16];
    u64 ctr;
    int idx;
};
TCB_HD u32 rotl32(u32 v, int n) { return (v << n) | (v >> (32 - n)); }
#define TCB_QR(a, b, c, d) a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); a += b; d ^= a; d = rotl32(d, 8); c += d; b ^= c; b = rotl32(b, 7);
TCB_HDN void chacha_refill(ChaChaRng &g) {
    u32 s[16], w[16];
    s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
    for (int i = 0; i < 8; i++) s[4 + i] = g.key[i];
    s[12] = (u32)g.ctr; s[13] = (u32)(g.ctr >> 32); s[14] = 0; s[15] = 0;
    for (int i = 0; i < 16; i++) w[i] = s[i];
    for (int i = 0; i < 10; i++) {
        TCB_QR(w[0], w[4], w[8], w[12]) TCB_QR(w[1], w[5], w[9], w[13]) TCB_QR(w[2], w[6], w[10], w[14]) TCB_QR(w[3], w[7], w[11], w[15])
        TCB_QR(w[0], w[5], w[10], w[15]) TCB_QR(w[1], w[6], w[11], w[12]) TCB_QR(w[2], w[7], w[8], w[13]) TCB_QR(w[3], w[4], w[9], w[14])
    }
    for (int i = 0; i < 16; i++) g.buf[i] = w[i] + s[i];
    g.ctr++;
    g.idx = 0;
}
TCB_HD void rng_seed(ChaChaRng &g, const u8 *seed) {
    for (int i = 0; i < 8; i++) g.key[i] = (u32)seed[4 * i] | ((u32)seed[4 * i + 1] << 8) | ((u32)seed[4 * i + 2] << 16) | ((u32)seed[4 * i + 3] << 24);
    g.ctr = 0;
    g.idx = 16;
}
TCB_HD u32 rng_u32(ChaChaRng &g) {
    if (g.idx >= 16) chacha_refill(g);
    return g.buf[g.idx++];
}
// ff_derive 0.6 Field::random for Fq (A1): 6 next_u64 = 12 consecutive words, top limb >> 3,
// reject >= p; the raw limbs ARE the Montgomery representation.
TCB_HD Fp fp_random(ChaChaRng &g) {
    Fp r;
    for (;;) {
        for (int i = 0; i < 12; i++) r.l[i] = rng_u32(g);
        r.l[11] &= 0x1fffffffu;
        if (limbs_lt_mod<FpParams>(r.l)) return r;
    }
}
// ---- helpers that make the sampling loop work for both engines (1 or 2 lanes per item)
template <class F2>
TCB_HD bool unit_and(bool v) {   // AND over the lanes of the unit
#if defined(__CUDA_ARCH__)
    if (F2::SLICED) return pair_and(v);
#endif
    return v;
}
template <class F2>
TCB_HD int vote_first(bool ok_me) {   // lowest lane of the unit whose flag is set, or -1 (uniform over the unit)
#if defined(__CUDA_ARCH__)
    if (F2::SLICED) {
        bool o = __shfl_xor_sync(pair_mask(), (int)ok_me, 1) != 0;
        bool role = lane_role();
        bool ok0 = role ? o : ok_me, ok1 = role ? ok_me : o;
        return ok0 ? 0 : (ok1 ? 1 : -1);
    }
#endif
    return ok_me ? 0 : -1;
}
template <class F2>
TCB_HD Fp bcast_fp(const Fp &v, int src) {   // value held by lane `src` of the unit
#if defined(__CUDA_ARCH__)
    if (F2::SLICED) {
        Fp r;
        u32 m = pair_mask();
        int lane = (int)((threadIdx.x & 30u) | (u32)src);
#pragma unroll
        for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(m, v.l[i], lane);
        return r;
    }
#endif
    return v;
}
template <class F2>
TCB_HD bool bcast_flag(bool v, int src) {
#if defined(__CUDA_ARCH__)
    if (F2::SLICED) return __shfl_sync(pair_mask(), (int)v, (int)((threadIdx.x & 30u) | (u32)src)) != 0;
#endif
    return v;
}
// EXTERNAL pairing 0.16 G2::random (A2, A3): x = Fq2::random, greatest = next_u32() % 2,
// y = sqrt(x^3 + b) picked by (y < -y) ^ greatest, multiplied by the exact cofactor h2, retry on
// no root / zero.  The candidates are consumed from the ChaCha stream in the reference's order;
// HOW a candidate is tested and the root extracted is free (any root gives the same point after
// the ordering rule):
//   * residuosity by the Legendre symbol of norm(a) (binary Jacobi algorithm on the ALU pipe, no
//     multiplications); an engine with two lanes per item tests two consecutive candidates per
//     round; only the winning candidate pays for s = sqrt(norm(a)) (one Fp power);
//   * y0 = sqrt((a0 +- s)/2), y1 = a1 / (2 y0) with e = delta^((p-3)/4): y0 = delta e, 1/y0 = e;
//     the two lanes try the two signs at once;
//   * the cofactor multiplication runs AFTER the rejection loop (all lanes converged) and uses
//     the psi-endomorphism identity of g2_clear_cofactor.
template <class F2>
TCB_HDN Jac<F2> g2_random(ChaChaRng &g, bool exact = true) {
    const Consts &C = CONSTS();
    constexpr int L = F2::SLICED ? 2 : 1;
    const int me = (int)my_role<F2>();
    Fp2 b2;
    b2.c0 = C.b1; b2.c1 = C.b1;   // 4 (1 + u)
    for (;;) {
        Fp2 x, a;
        Fp s;
        bool greatest;
        for (;;) {
            Fp2 xc;
            bool gr = false;
            xc.c0 = Fp::zero(); xc.c1 = Fp::zero();
            for (int k = 0; k < L; k++) {
                Fp c0 = fp_random(g);
                Fp c1 = fp_random(g);
                bool gk = (rng_u32(g) & 1u) != 0;
                if (k == me) { xc.c0 = c0; xc.c1 = c1; gr = gk; }
            }
            Fp2 ac = sqr(xc) * xc + b2;
            Fp nrm = norm(ac);
            int w = vote_first<F2>(fp_is_square(nrm));       // a is a square in Fp2 iff its norm is one in Fp
            if (w < 0) continue;
            x.c0 = bcast_fp<F2>(xc.c0, w); x.c1 = bcast_fp<F2>(xc.c1, w);
            a.c0 = bcast_fp<F2>(ac.c0, w); a.c1 = bcast_fp<F2>(ac.c1, w);
            s = fp_pow<ExpPp1d4>(bcast_fp<F2>(nrm, w));       // sqrt(norm), after the loop has converged
            greatest = bcast_flag<F2>(gr, w);
            break;
        }
        F2 y;
        if (a.c1.is_zero()) {
            fp2_sqrt(y, F2::from_halves(a.c0, a.c1));   // a in Fp: generic Algorithm 9 (never hit by random x)
        } else {
            Fp y0 = Fp::zero(), y1 = Fp::zero();
            for (int r = 0; r < 2 / L; r++) {   // exactly one of (a0 + s)/2, (a0 - s)/2 is a square
                int which = r * L + me;
                Fp d = fp_half(which == 0 ? a.c0 + s : a.c0 - s);
                Fp e = fp_pow<ExpPm3d4>(d);
                Fp c = d * e;
                int w = vote_first<F2>(sqr(c) == d);
                if (w >= 0) {
                    y0 = bcast_fp<F2>(c, w);
                    y1 = bcast_fp<F2>(fp_half(a.c1 * e), w);
                    break;
                }
            }
            y = F2::from_halves(y0, y1);
        }
        F2 ny = -y;
        bool pick_y = (fp2_cmp(y, ny) < 0) != greatest;
        Aff<F2> pt;
        pt.x = F2::from_halves(x.c0, x.c1); pt.y = select(pick_y, y, ny); pt.inf = false;
        Jac<F2> p = g2_clear_cofactor(pt, exact);
        if (!jac_is_inf(p)) return p;
    }
}
template <class F2>
TCB_HD Jac<F2> hash_g2(const u8 *m1, size_t l1, const u8 *m2, size_t l2, bool exact = true) {   // src/lib.rs:691-694
    u8 digest[32];
    sha3_256(m1, l1, m2, l2, digest);
    ChaChaRng g;
    rng_seed(g, digest);
    return g2_random<F2>(g, exact);
}
// ---- hash_g2 in two kernels.  Everything up to the curve point (SHA3, ChaCha, the candidate search, the two Fp powers of the
// square root) is Fp-level work with nothing for the second lane of a pair to do — inside the lane-pair kernel both powers ran
// redundantly on the two lanes.  g2_random_point is that first half for ONE thread per item (the residuosity of (a0 + s)/2 is
// decided by a Legendre symbol, so exactly two powers per hash); task_g2_clear is the second half on the item's Fp2 engine.
// The reference retries with further candidates when the cleared point is the identity (probability ~2^-255): such an item is
// flagged in `redo` and recomputed by the one-kernel path (k_hash_g2 with its `only` filter), which continues the ChaCha stream.
TCB_HDN void g2_random_point(ChaChaRng &g, Fp2 &px, Fp2 &py) {
    const Consts &C = CONSTS();
    Fp2 b2;
    b2.c0 = C.b1; b2.c1 = C.b1;   // 4 (1 + u)
    for (;;) {
        Fp2 x, a;
        Fp nrm;
        bool greatest;
        for (;;) {                                   // the rejection loop holds only the cheap part; lanes leave it at different trips
            x.c0 = fp_random(g);
            x.c1 = fp_random(g);
            greatest = (rng_u32(g) & 1u) != 0;
            a = sqr(x) * x + b2;
            nrm = norm(a);
            if (fp_is_square(nrm)) break;            // a is a square in Fp2 iff its norm is one in Fp
        }
#if defined(__CUDA_ARCH__)
        __syncwarp();                                // the kernel keeps all 32 lanes alive: the powers below run converged
#endif
        Fp2 y;
        if (a.c1.is_zero()) {
            fp2_sqrt(y, a);                          // a in Fp: generic Algorithm 9 (never hit by random x)
        } else {
            Fp s = fp_pow<ExpPp1d4>(nrm);            // sqrt(norm)
            Fp d = fp_half(a.c0 + s);                // exactly one of (a0 + s)/2, (a0 - s)/2 is a square (their product is -a1^2/4)
            if (!fp_is_square(d)) d = fp_half(a.c0 - s);
            Fp e = fp_pow<ExpPm3d4>(d);              // y0 = d e = sqrt(d), 1/y0 = e
            y.c0 = d * e;
            y.c1 = fp_half(a.c1 * e);
        }
        Fp2 ny = -y;
        bool pick_y = (fp2_cmp(y, ny) < 0) != greatest;
        px = x; py = pick_y ? y : ny;
        return;
    }
}
struct G2PointStore { Fp2c x, y; };      // affine twist point, Montgomery limbs (scratch between the two kernels)
TCB_HD void task_hash_g2_point(size_t i, const u8 *msgs, const u64 *off, G2PointStore *dst) {   // dst: where item i's point goes
    u8 digest[32];
    sha3_256(msgs + off[i], (size_t)(off[i + 1] - off[i]), msgs, 0, digest);
    ChaChaRng g;
    rng_seed(g, digest);
    Fp2 x, y;
    g2_random_point(g, x, y);
    x.store(dst->x); y.store(dst->y);
}
// the same first half for hash_g1_g2 (src/lib.rs:697-707): SHA3 over msg (or its digest when longer than 64 bytes) || compressed g1
TCB_HD void task_hash_g1_g2_point(size_t i, const u8 *g1_pts, const u8 *msgs, const u64 *off, G2PointStore *dst) {
    bool ok = true;
    Aff<Fp> g1 = load_g1(g1_pts + 96 * i, ok);
    const u8 *msg = msgs + off[i];
    size_t len = (size_t)(off[i + 1] - off[i]);
    u8 comp[48], d[32], digest[32];
    g1_compress(comp, g1);
    if (len > 64) {
        sha3_256(msg, len, msg, 0, d);
        sha3_256(d, 32, comp, 48, digest);
    } else {
        sha3_256(msg, len, comp, 48, digest);
    }
    ChaChaRng g;
    rng_seed(g, digest);
    Fp2 x, y;
    g2_random_point(g, x, y);
    x.store(dst->x); y.store(dst->y);
}
template <class F2>
TCB_HD void task_g2_clear(size_t i, const G2PointStore *pts, u8 *out_g2, bool exact, u8 *redo) {
    Aff<F2> pt;
    pt.x = F2::load(pts[i].x); pt.y = F2::load(pts[i].y); pt.inf = false;
    Jac<F2> p = g2_clear_cofactor(pt, exact);
    bool inf = jac_is_inf(p);
    if (is_writer<F2>()) redo[i] = inf ? 1 : 0;
    if (!inf) store_g2<F2>(out_g2 + 192 * i, jac_to_aff(p));
}
template <class F2>
TCB_HD Jac<F2> hash_g1_g2(const Aff<Fp> &g1, const u8 *msg, size_t len) {   // src/lib.rs:697-707
    u8 comp[48], d[32];
    g1_compress(comp, g1);
    if (len > 64) {
        sha3_256(msg, len, msg, 0, d);
        return hash_g2<F2>(d, 32, comp, 48);
    }
    return hash_g2<F2>(msg, len, comp, 48);
}
// src/lib.rs:710-715: one keystream u32 per output byte (`next_u32() as u8`, A5)
TCB_HD void xor_with_hash(u8 *out, const Aff<Fp> &g1, const u8 *in, size_t len) {
    u8 comp[48], d[32];
    g1_compress(comp, g1);
    sha3_256(comp, 48, comp, 0, d);
    ChaChaRng g;
    rng_seed(g, d);
    for (size_t i = 0; i < len; i++) out[i] = (u8)rng_u32(g) ^ in[i];
}

// ----------------------------------------------------------------------------- Fr helpers
TCB_HD Fr fr_one() { return CONSTS().fr_r1; }
TCB_HD Fr fr_load_le(const u8 *b, bool &ok) {
    Fr t;
    load_scalar_le(t.l, b);
    ok = ok && limbs_lt_mod<FrParams>(t.l);
    return t * CONSTS().fr_r2;
}
TCB_HD Fr fr_inv(const Fr &a) {
    Fr acc = fr_one();
    for (int i = 255; i >= 0; i--) {
        acc = acc * acc;
        if ((ExpRm2::get(i >> 5) >> (i & 31)) & 1) acc = acc * a;
    }
    return acc;
}
// lambda_i(0) for sample i of item `xs` (m = t+1 canonical LE scalars), exactly the value
// src/lib.rs:739-765 produces: numerator = prod_{j != i} x_j (by position), denominator =
// prod_{x_j != x_i} (x_j - x_i) (by value — duplicates are skipped, never a zero divisor).
// Output: canonical little-endian limbs.  status 3 if an x is not a canonical Fr.
TCB_HDN void lagrange_coeff(const u8 *xs, size_t m, size_t i, u32 *out, u8 &status) {
    bool ok = true;
    Fr xi = fr_load_le(xs + 32 * i, ok);
    Fr num = fr_one(), den = fr_one();
    for (size_t j = 0; j < m; j++) {
        Fr xj = fr_load_le(xs + 32 * j, ok);
        if (j != i) num = num * xj;
        if (xj != xi) den = den * (xj - xi);
    }
    Fr l = from_mont<FrParams>(num * fr_inv(den));
    for (int k = 0; k < 8; k++) out[k] = l.l[k];
    if (!ok) status = 3;
}

// The same coefficients in two passes with ONE inversion per item instead of one per (item, share):
//   pass A (one unit per (item, share)): numerator and denominator products, kept in Montgomery form in scratch;
//   pass B (one unit per item): Montgomery's simultaneous inversion of the t+1 denominators (never zero: equal
//   x values are skipped), lambda_i = num_i / den_i written as canonical little-endian limbs.
// ~2 m + 3 + 380 / m products per share instead of 2 m + 380.
struct LagrangeND { Fr num, den, pre; };
TCB_HD void lagrange_num_den(const u8 *xs, size_t m, size_t i, LagrangeND &out, u8 &status) {
    bool ok = true;
    Fr xi = fr_load_le(xs + 32 * i, ok);
    Fr num = fr_one(), den = fr_one();
    for (size_t j = 0; j < m; j++) {
        Fr xj = fr_load_le(xs + 32 * j, ok);
        if (j != i) num = num * xj;
        if (xj != xi) den = den * (xj - xi);
    }
    out.num = num; out.den = den;
    if (!ok) status = 3;
}
TCB_HD void lagrange_finish_item(LagrangeND *nd, size_t m, u32 *out) {
    Fr run = fr_one();
    for (size_t k = 0; k < m; k++) { nd[k].pre = run; run = run * nd[k].den; }
    Fr rinv = fr_inv(run);
    for (size_t k = m; k-- > 0;) {
        Fr dinv = rinv * nd[k].pre;
        rinv = rinv * nd[k].den;
        Fr l = from_mont<FrParams>(nd[k].num * dinv);
        for (int w = 0; w < 8; w++) out[8 * k + w] = l.l[w];
    }
}

// ----------------------------------------------------------------------------- §8(f) row 4: Fr-side Poly algebra
// Poly::evaluate (src/poly.rs:358-369: Horner from the leading coefficient) and Poly * Poly (src/poly.rs:173-194: schoolbook
// convolution) on canonical little-endian Fr coefficients; `cm` / `am` / `bm` are the coefficients already in Montgomery form
// (k_fr_to_mont), outputs are canonical again.
TCB_HD void task_fr_to_mont(size_t i, const u8 *in, Fr *out, u8 *bad) {
    bool ok = true;
    out[i] = fr_load_le(in + 32 * i, ok);
    if (!ok) *bad = 3;
}
TCB_HD void fr_store_le(u8 *b, const Fr &a) {
    Fr c = from_mont<FrParams>(a);
    for (int i = 0; i < 8; i++) { b[4 * i] = (u8)c.l[i]; b[4 * i + 1] = (u8)(c.l[i] >> 8); b[4 * i + 2] = (u8)(c.l[i] >> 16); b[4 * i + 3] = (u8)(c.l[i] >> 24); }
}
TCB_HD void task_poly_eval(size_t i, size_t deg, const Fr *cm, const u8 *x_fr, u8 *out_fr, u8 *bad) {
    bool ok = true;
    Fr x = fr_load_le(x_fr + 32 * i, ok);
    Fr acc = cm[deg];
    for (size_t c = deg; c-- > 0;) acc = acc * x + cm[c];
    fr_store_le(out_fr + 32 * i, acc);
    if (!ok) *bad = 3;
}
// out_{item,k} = sum_i a_{item,i} * b_{item,k-i}
TCB_HD void task_poly_mul(size_t u, size_t da, size_t db, const Fr *am, const Fr *bm, u8 *out_fr) {
    size_t w = da + db + 1, item = u / w, k = u % w;
    const Fr *a = am + item * (da + 1), *b = bm + item * (db + 1);
    size_t lo = k > db ? k - db : 0, hi = k < da ? k : da;
    Fr acc = Fr::zero();
    for (size_t i = lo; i <= hi; i++) acc = acc + a[i] * b[k - i];
    fr_store_le(out_fr + 32 * u, acc);
}

// ----------------------------------------------------------------------------- per-item tasks
// a1: e(a,b) == e(c,d); c == nullptr means the G1 generator (src/lib.rs:108-110,182-186,508-512)
template <class F2>
TCB_HD void task_verify_g2(size_t i, const u8 *a_g1, const u8 *b_g2, const u8 *c_g1, const u8 *d_g2, u8 *ok_out, bool gen_scaled = false) {
    bool ok = true;
    Aff<Fp> a = load_g1(a_g1 + 96 * i, ok);
    Aff<F2> b = load_g2<F2>(b_g2 + 192 * i, ok);
    Aff<Fp> c;
    if (c_g1) c = load_g1(c_g1 + 96 * i, ok);
    else { c.x = gen_scaled ? CONSTS().g1cx : CONSTS().g1x; c.y = gen_scaled ? CONSTS().g1cy : CONSTS().g1y; c.inf = false; }
    Aff<F2> d = load_g2<F2>(d_g2 + 192 * i, ok);
    bool res = pairing_eq<F2>(a, b, c, d);
    if (is_writer<F2>()) ok_out[i] = (res && ok) ? 1 : 0;
}
// a2: hash_g2 -> uncompressed affine G2
template <class F2>
TCB_HD void task_hash_g2(size_t i, const u8 *msgs, const u64 *off, u8 *out_g2, bool exact = true, const u8 *only = nullptr) {   // exact == false: [3 (x^2 - 1)] H(m), see g2_clear_cofactor
    if (only && !only[i]) return;            // second pass of the two-kernel path: only the flagged items
    Jac<F2> h = hash_g2<F2>(msgs + off[i], (size_t)(off[i + 1] - off[i]), msgs, 0, exact);
    store_g2<F2>(out_g2 + 192 * i, jac_to_aff(h));
}
// a2: hash_g1_g2 (src/lib.rs:697-707) -> uncompressed affine G2
template <class F2>
TCB_HD void task_hash_g1_g2(size_t i, const u8 *g1_pts, const u8 *msgs, const u64 *off, u8 *out_g2, const u8 *only = nullptr) {
    if (only && !only[i]) return;            // second pass of the two-kernel path: only the flagged items
    bool ok = true;
    Aff<Fp> g = load_g1(g1_pts + 96 * i, ok);
    Jac<F2> h = hash_g1_g2<F2>(g, msgs + off[i], (size_t)(off[i + 1] - off[i]));
    store_g2<F2>(out_g2 + 192 * i, jac_to_aff(h));
}
// a3: PublicKey::verify = hash_g2 then a1 (src/lib.rs:115-117)
template <class F2>
TCB_HD void task_verify(size_t i, const u8 *pk_g1, const u8 *sig_g2, const u8 *msgs, const u64 *off, u8 *ok_out) {
    bool ok = true;
    Aff<Fp> pk = load_g1(pk_g1 + 96 * i, ok);
    Aff<F2> sig = load_g2<F2>(sig_g2 + 192 * i, ok);
    Aff<F2> h = jac_to_aff(hash_g2<F2>(msgs + off[i], (size_t)(off[i + 1] - off[i]), msgs, 0));
    Aff<Fp> g;
    g.x = CONSTS().g1x; g.y = CONSTS().g1y; g.inf = false;
    bool res = pairing_eq<F2>(pk, h, g, sig);
    if (is_writer<F2>()) ok_out[i] = (res && ok) ? 1 : 0;
}
// a4: sign = sk * hash_g2(msg) (src/lib.rs:372-381); h_g2 != nullptr selects sign_g2
template <class F2>
TCB_HD void task_sign(size_t i, const u8 *sk, const u8 *msgs, const u64 *off, const u8 *h_g2, u8 *out_g2) {
    u32 k[8];
    load_scalar_le(k, sk + 32 * i);
    bool ok = true;
    Aff<F2> h = h_g2 ? load_g2<F2>(h_g2 + 192 * i, ok)
                     : jac_to_aff(hash_g2<F2>(msgs + off[i], (size_t)(off[i + 1] - off[i]), msgs, 0));
    store_g2<F2>(out_g2 + 192 * i, jac_to_aff(jac_mul_gls4<F2>(h, k)));
}
// generic k * P over G2 with the result kept Jacobian in a Montgomery scratch array
// (used for the per-share terms of interpolate, src/lib.rs:753-765)
template <class F2> struct JacStore { Fp2c x, y, z; };
template <class F2>
TCB_HD void task_g2_mul_store(size_t i, const u32 *k_limbs, const u8 *pts_g2, JacStore<F2> *out, u8 *status, size_t per_item) {
    bool ok = true;
    Aff<F2> p = load_g2<F2>(pts_g2 + 192 * i, ok);
    Jac<F2> r = jac_mul_gls4<F2>(p, k_limbs + 8 * i);
    r.x.store(out[i].x); r.y.store(out[i].y); r.z.store(out[i].z);
    if (!ok && is_writer<F2>()) status[i / per_item] = 3;
}
template <class F2>
TCB_HD void task_g2_sum(size_t i, size_t m, const JacStore<F2> *terms, u8 *out_g2) {
    Jac<F2> acc = jac_inf<F2>();
    for (size_t k = 0; k < m; k++) {
        const JacStore<F2> &t = terms[i * m + k];
        Jac<F2> p;
        p.x = F2::load(t.x); p.y = F2::load(t.y); p.z = F2::load(t.z);
        acc = jac_add(acc, p);
    }
    store_g2<F2>(out_g2 + 192 * i, jac_to_aff(acc));
}
// ---- interpolate / linear combinations as ONE multi-scalar multiplication per item (Straus with the
// GLS4 recoding): sum_s k_s P_s with the 66 doublings SHARED by the shares of a group and every addition a
// MIXED addition from a per-share table normalised to affine with one inversion (Montgomery's trick).
// Unit u = (item, share) prepares digits + table; unit w = (item, group g) accumulates the shares
// s = g, g + G, ... of its item; the G partial sums go through task_g2_sum.  Same group element as
// the reference's per-share double-and-add (src/lib.rs:753-765), ~1.6x fewer multiplications.
template <class F2> struct AffStore { Fp2c x, y; };
template <class F2>
TCB_HD void task_g2_msm_prep(size_t u, const u32 *k_limbs, const u8 *pts_g2, AffStore<F2> *tab, Gls4Digits *dgs, u8 *status, size_t per_item) {
    bool ok = true;
    Aff<F2> p = load_g2<F2>(pts_g2 + 192 * u, ok);
    Gls4Digits dg;
    gls4_recode(k_limbs + 8 * u, dg);
    if (p.inf) dg.flags |= 2u;
    else {
        Aff<F2> A[8];
        gls4_table_affine(p, A);
        for (int e = 0; e < 8; e++) { A[e].x.store(tab[8 * u + e].x); A[e].y.store(tab[8 * u + e].y); }
    }
    if (is_writer<F2>()) {
        dgs[u] = dg;
        if (!ok) status[u / per_item] = 3;
    }
}
template <class F2>
TCB_HD Aff<F2> g2_msm_fetch(const AffStore<F2> *tab, const Gls4Digits *dgs, size_t u, int j) {
    u32 d = dgs[u].digit(j);
    const AffStore<F2> &e = tab[8 * u + (d >> 1)];
    Aff<F2> t;
    t.x = F2::load(e.x); t.y = F2::load(e.y);
    t.inf = (dgs[u].flags & 2u) != 0;      // infinity share: its (unwritten) table entry is ignored by the addition
    if (d & 1) t.y = -t.y;
    return t;
}
// sum of the shares base, base + G, ..., base + (cnt - 1) G (units of the prep arrays) with shared doublings
template <class F2>
TCB_HD Jac<F2> g2_msm_walk(size_t base, size_t G, size_t cnt, const AffStore<F2> *tab, const Gls4Digits *dgs) {
    Jac<F2> acc = jac_inf<F2>();
    // flattened (digit, share) sequence with the next table entry fetched before the current addition
    Aff<F2> nxt = g2_msm_fetch<F2>(tab, dgs, base, GLS4_L);
    int j = GLS4_L;
    size_t si = 0;
    for (;;) {
        Aff<F2> cur = nxt;
        bool row_start = si == 0;
        size_t si2 = si + 1;
        int j2 = j;
        if (si2 == cnt) { si2 = 0; j2 = j - 1; }
        bool more = j2 >= 0;
        if (more) nxt = g2_msm_fetch<F2>(tab, dgs, base + si2 * G, j2);
        if (row_start) acc = jac_dbl(acc);
        acc = jac_add_mixed(acc, cur);
        if (!more) break;
        si = si2; j = j2;
    }
    for (size_t s = 0; s < cnt; s++) {
        size_t u = base + s * G;
        if ((dgs[u].flags & 3u) != 1u) continue;     // a0 was even: subtract P
        Aff<F2> t;
        t.x = F2::load(tab[8 * u].x); t.y = -F2::load(tab[8 * u].y); t.inf = false;
        acc = jac_add_mixed(acc, t);
    }
    return acc;
}
template <class F2>
TCB_HD void task_g2_msm_acc(size_t w, size_t m, size_t G, const AffStore<F2> *tab, const Gls4Digits *dgs, JacStore<F2> *out) {
    size_t item = w / G, g = w % G;
    size_t cnt = (m - g + G - 1) / G;          // shares g, g + G, ... of this item (g < G <= m)
    Jac<F2> acc = g2_msm_walk<F2>(item * m + g, G, cnt, tab, dgs);
    acc.x.store(out[w].x); acc.y.store(out[w].y); acc.z.store(out[w].z);
}
// "Spill" layout for batches that leave part of the GPU's resident units idle with one unit per item (2^14 items on 18 944 unit
// slots): unit w < n adds the first m - 1 shares of item w, the spare units each add the LAST share of q items, one after the other,
// so that every unit walks less than the one-unit-per-item plan and all slots are used.  Two partial sums per item: out[2 item],
// out[2 item + 1] (then task_g2_sum with two terms).
template <class F2>
TCB_HD void task_g2_msm_acc_spill(size_t w, size_t n, size_t m, size_t q, const AffStore<F2> *tab, const Gls4Digits *dgs, JacStore<F2> *out) {
    if (w < n) {
        Jac<F2> acc = g2_msm_walk<F2>(w * m, 1, m - 1, tab, dgs);
        acc.x.store(out[2 * w].x); acc.y.store(out[2 * w].y); acc.z.store(out[2 * w].z);
        return;
    }
    size_t lo = (w - n) * q, hi = lo + q < n ? lo + q : n;
    for (size_t item = lo; item < hi; item++) {
        Jac<F2> acc = g2_msm_walk<F2>(item * m + m - 1, 1, 1, tab, dgs);
        acc.x.store(out[2 * item + 1].x); acc.y.store(out[2 * item + 1].y); acc.z.store(out[2 * item + 1].z);
    }
}

// t == 0 shortcut of interpolate (src/lib.rs:735-737): the first sample is returned unchanged
template <class F2>
TCB_HD void task_g2_copy(size_t i, const u8 *in, u8 *out) {
    if (is_writer<F2>()) for (int k = 0; k < 192; k++) out[192 * i + k] = in[192 * i + k];
}

// ---- G1 tasks (always one item per thread)
struct Jac1Store { Fp x, y, z; };
TCB_HD void store_jac(Jac1Store &o, const Jac<Fp> &p) { o.x = p.x; o.y = p.y; o.z = p.z; }
template <class F2> TCB_HD void store_jac(JacStore<F2> &o, const Jac<F2> &p) { p.x.store(o.x); p.y.store(o.y); p.z.store(o.z); }
TCB_HD void task_g1_mul(size_t i, const u8 *sk, const u8 *pts_g1, u8 *out_g1) {   // a7 share / a9
    u32 k[8];
    load_scalar_le(k, sk + 32 * i);
    bool ok = true;
    Aff<Fp> p;
    if (pts_g1) p = load_g1(pts_g1 + 96 * i, ok);
    else { p.x = CONSTS().g1x; p.y = CONSTS().g1y; p.inf = false; }
    store_g1(out_g1 + 96 * i, jac_to_aff(jac_mul_glv2(p, k)));
}
TCB_HD void task_g1_mul_store(size_t i, const u32 *k_limbs, const u8 *pts_g1, Jac1Store *out, u8 *status, size_t per_item) {
    bool ok = true;
    Aff<Fp> p = load_g1(pts_g1 + 96 * i, ok);
    Jac<Fp> r = jac_mul_glv2(p, k_limbs + 8 * i);
    out[i].x = r.x; out[i].y = r.y; out[i].z = r.z;
    if (!ok) status[i / per_item] = 3;
}
TCB_HD Aff<Fp> g1_sum(size_t i, size_t m, const Jac1Store *terms) {
    Jac<Fp> acc = jac_inf<Fp>();
    for (size_t k = 0; k < m; k++) {
        Jac<Fp> p;
        p.x = terms[i * m + k].x; p.y = terms[i * m + k].y; p.z = terms[i * m + k].z;
        acc = jac_add(acc, p);
    }
    return jac_to_aff(acc);
}
// G1 version of the shared-doubling multi-scalar multiplication (decrypt, combine_g1, g1_lincomb): GLV2
// recoding read two positions at a time (tower.cuh, glv2_table8: 8 affine entries per share, one inversion),
// 130 shared doublings + 65 mixed additions per share.
struct Aff1Store { Fp x, y; };
TCB_HD void task_g1_msm_prep(size_t u, const u32 *k_limbs, const u8 *pts_g1, Aff1Store *tab, Glv2Digits *dgs, u8 *status, size_t per_item) {
    bool ok = true;
    Aff<Fp> p = load_g1(pts_g1 + 96 * u, ok);
    Glv2Digits dg;
    glv2_recode(k_limbs + 8 * u, dg);
    if (p.inf) dg.flags |= 2u;
    else {
        Aff<Fp> E[8];
        glv2_table8(p, E);
        for (int e = 0; e < 8; e++) { tab[8 * u + e].x = E[e].x; tab[8 * u + e].y = E[e].y; }
    }
    dgs[u] = dg;
    if (!ok) status[u / per_item] = 3;
}
TCB_HD Aff<Fp> g1_msm_fetch(const Aff1Store *tab, const Glv2Digits *dgs, size_t u, int k) {
    u32 nib = glv2_wdigit(dgs[u], k);
    const Aff1Store &e = tab[8 * u + glv2_windex(nib)];
    Aff<Fp> t;
    t.x = e.x; t.y = e.y;
    t.inf = (dgs[u].flags & 2u) != 0;
    if (glv2_wneg(nib)) t.y = -t.y;
    return t;
}
TCB_HD void task_g1_msm_acc(size_t w, size_t m, size_t G, const Aff1Store *tab, const Glv2Digits *dgs, Jac1Store *out) {
    size_t item = w / G, g = w % G;
    size_t cnt = (m - g + G - 1) / G;
    size_t base = item * m + g;
    Jac<Fp> acc = jac_inf<Fp>();
    Aff<Fp> nxt = g1_msm_fetch(tab, dgs, base, GLV2_W);
    int j = GLV2_W;
    size_t si = 0;
    for (;;) {
        Aff<Fp> cur = nxt;
        bool row_start = si == 0;
        size_t si2 = si + 1;
        int j2 = j;
        if (si2 == cnt) { si2 = 0; j2 = j - 1; }
        bool more = j2 >= 0;
        if (more) nxt = g1_msm_fetch(tab, dgs, base + si2 * G, j2);
        if (row_start) acc = jac_dbl(jac_dbl(acc));
        acc = jac_add_mixed(acc, cur);
        if (!more) break;
        si = si2; j = j2;
    }
    for (size_t s = 0; s < cnt; s++) {
        size_t u = base + s * G;
        if ((dgs[u].flags & 3u) != 1u) continue;
        Aff<Fp> t;
        t.x = tab[8 * u + 4].x; t.y = -tab[8 * u + 4].y; t.inf = false;      // entry 4 is P itself
        acc = jac_add_mixed(acc, t);
    }
    out[w].x = acc.x; out[w].y = acc.y; out[w].z = acc.z;
}
// ---- batch-affine accumulation for the multi-scalar multiplication.
// sum_s k_s P_s = sum_j 2^j S_j with S_j = sum_s (+-)T_s[d_{s,j}] a sum of AFFINE table entries.  All S_j are
// formed by a pairwise tree over the shares; the additions of one tree level (all digit positions at once)
// are AFFINE additions sharing ONE inversion (Montgomery's trick): 5M + 1S per addition instead of the
// 7M + 4S of a mixed Jacobian addition, and the doublings happen only in the final Horner pass over the
// per-position sums.  Working arrays (two ping-pong point buffers and the prefix products) live in a
// scratch slab per unit.  Exceptional pairs (an operand at infinity, P + P, P + (-P)) are classified the same
// way in both passes and contribute 1 to the running product.
template <class F2> struct MsmG2 {
    typedef F2 F;
    typedef AffStore<F2> PS;
    typedef Fp2c FS;
    typedef Gls4Digits DG;
    static constexpr int L = GLS4_L, TAB = 8, P_ENTRY = 0, DBL = 1;    // positions 0..L, table entries, entry holding P, doublings per position
    TCB_HD static F fs_load(const FS &c) { return F2::load(c); }
    TCB_HD static void fs_store(FS &c, const F &v) { v.store(c); }
    TCB_HD static Aff<F> fetch(const PS *tab, const DG *dgs, size_t u, int j) { return g2_msm_fetch<F2>(tab, dgs, u, j); }
};
struct MsmG1 {
    typedef Fp F;
    typedef Aff1Store PS;
    typedef Fp FS;
    typedef Glv2Digits DG;
    static constexpr int L = GLV2_W, TAB = 8, P_ENTRY = 4, DBL = 2;
    TCB_HD static F fs_load(const FS &c) { return c; }
    TCB_HD static void fs_store(FS &c, const F &v) { c = v; }
    TCB_HD static Aff<F> fetch(const PS *tab, const DG *dgs, size_t u, int j) { return g1_msm_fetch(tab, dgs, u, j); }
};
template <class M> TCB_HD Aff<typename M::F> ps_load(const typename M::PS &e) {
    Aff<typename M::F> t;
    t.x = M::fs_load(e.x); t.y = M::fs_load(e.y);
    t.inf = is_zero(t.x) && is_zero(t.y);           // (0, 0) is not on either curve: the infinity marker
    return t;
}
template <class M> TCB_HD void ps_store(typename M::PS &e, const Aff<typename M::F> &t) {
    typedef typename M::F F;
    M::fs_store(e.x, t.inf ? FieldOne<F>::zero() : t.x);
    M::fs_store(e.y, t.inf ? FieldOne<F>::zero() : t.y);
}
// classification of a pair; den = the element that enters the shared inversion
enum { BA_REGULAR = 0, BA_TAKE_B = 1, BA_TAKE_A = 2, BA_DOUBLE = 3, BA_INF = 4 };
template <class F> TCB_HD int ba_classify(const Aff<F> &a, const Aff<F> &b, F &den) {
    den = FieldOne<F>::one();
    if (a.inf) return BA_TAKE_B;
    if (b.inf) return BA_TAKE_A;
    if (eq(a.x, b.x)) {
        if (eq(a.y, b.y)) { den = dbl(a.y); return BA_DOUBLE; }
        return BA_INF;
    }
    den = b.x - a.x;
    return BA_REGULAR;
}
template <class F> TCB_HD Aff<F> ba_finish(int kind, const Aff<F> &a, const Aff<F> &b, const F &dinv) {
    if (kind == BA_TAKE_B) return b;
    if (kind == BA_TAKE_A) return a;
    Aff<F> r;
    if (kind == BA_INF) { r.x = FieldOne<F>::zero(); r.y = FieldOne<F>::zero(); r.inf = true; return r; }
    F lam;
    if (kind == BA_DOUBLE) { F xx = sqr(a.x); lam = (dbl(xx) + xx) * dinv; }
    else lam = (b.y - a.y) * dinv;
    r.x = sqr(lam) - a.x - b.x;
    r.y = lam * (a.x - r.x) - a.y;
    r.inf = false;
    return r;
}
// One tree level over all digit positions: in-points (j, i), i < n_in  ->  out-points (j, i), i < (n_in + 1) / 2.
// `get(j, i)` reads an input point.  Pairs are enumerated position-major (k = j * pairs + i); both passes
// fetch the operands of the NEXT pair before working on the current one (the kernels run 2-3 warps per
// scheduler, so a global-memory round trip is not hidden by other warps: profiles/r1s2b_*: 26-39 % long_sb).
template <class M, class Get>
TCB_HD void ba_level(size_t n_in, Get get, typename M::PS *out, size_t out_stride, typename M::FS *prefix) {
    typedef typename M::F F;
    const size_t pairs = n_in / 2, n_out = (n_in + 1) / 2;
    const int NP = M::L + 1;
    if (pairs) {
        const size_t K = (size_t)NP * pairs;
        F run = FieldOne<F>::one();
        {
            Aff<F> a = get(0, 0), b = get(0, 1);
            for (size_t k = 0; k < K; k++) {
                Aff<F> an = a, bn = b;
                if (k + 1 < K) { size_t jn = (k + 1) / pairs, in = (k + 1) % pairs; an = get((int)jn, 2 * in); bn = get((int)jn, 2 * in + 1); }
                F den;
                ba_classify(a, b, den);
                M::fs_store(prefix[k], run);          // product of the denominators BEFORE this pair
                run = run * den;
                a = an; b = bn;
            }
        }
        F rinv = inv(run);
        {
            size_t jl = (K - 1) / pairs, il = (K - 1) % pairs;
            Aff<F> a = get((int)jl, 2 * il), b = get((int)jl, 2 * il + 1);
            F pf = M::fs_load(prefix[K - 1]);
            for (size_t k = K; k-- > 0;) {
                Aff<F> an = a, bn = b;
                F pfn = pf;
                if (k > 0) { size_t jn = (k - 1) / pairs, in = (k - 1) % pairs; an = get((int)jn, 2 * in); bn = get((int)jn, 2 * in + 1); pfn = M::fs_load(prefix[k - 1]); }
                F den;
                int kind = ba_classify(a, b, den);
                F dinv = rinv * pf;
                rinv = rinv * den;
                ps_store<M>(out[(k / pairs) * out_stride + (k % pairs)], ba_finish(kind, a, b, dinv));
                a = an; b = bn; pf = pfn;
            }
        }
    }
    if (n_in & 1)
        for (int j = 0; j < NP; j++) ps_store<M>(out[(size_t)j * out_stride + n_out - 1], get(j, n_in - 1));
}
// scratch per unit: two point buffers of (L + 1) * ceil(cnt_max / 2) entries and (L + 1) * floor(cnt_max / 2) prefix products
template <class M> TCB_HD size_t ba_points_per_unit(size_t cnt_max) { return (size_t)(M::L + 1) * ((cnt_max + 1) / 2); }
template <class M> TCB_HD size_t ba_prefix_per_unit(size_t cnt_max) { size_t h = cnt_max / 2; return (size_t)(M::L + 1) * (h ? h : 1); }
template <class M, class JS /* Jacobian store */>
TCB_HD void task_msm_acc_ba(size_t w, size_t m, size_t G, const typename M::PS *tab, const typename M::DG *dgs,
                            typename M::PS *buf_a, typename M::PS *buf_b, typename M::FS *prefix_all, size_t cnt_max, JS *out) {
    typedef typename M::F F;
    typedef typename M::PS PS;
    size_t item = w / G, g = w % G;
    size_t cnt = (m - g + G - 1) / G;
    size_t base = item * m + g;
    const size_t stride = (cnt_max + 1) / 2;
    PS *A = buf_a + w * ba_points_per_unit<M>(cnt_max), *B = buf_b + w * ba_points_per_unit<M>(cnt_max);
    typename M::FS *prefix = prefix_all + w * ba_prefix_per_unit<M>(cnt_max);
    // level 0 reads the per-share tables through the recoded digits
    size_t n = cnt;
    ba_level<M>(n, [&](int j, size_t i) { return M::fetch(tab, dgs, base + i * G, j); }, A, stride, prefix);
    n = (n + 1) / 2;
    PS *cur = A, *nxt = B;
    while (n > 1) {
        const PS *src = cur;
        ba_level<M>(n, [&](int j, size_t i) { return ps_load<M>(src[(size_t)j * stride + i]); }, nxt, stride, prefix);
        n = (n + 1) / 2;
        PS *t = cur; cur = nxt; nxt = t;
    }
    // Horner over the per-position sums
    Jac<F> acc = jac_inf<F>();
    Aff<F> sj = ps_load<M>(cur[(size_t)M::L * stride]);
    for (int j = M::L; j >= 0; j--) {
        Aff<F> sn = sj;
        if (j > 0) sn = ps_load<M>(cur[(size_t)(j - 1) * stride]);
        for (int d = 0; d < M::DBL; d++) acc = jac_dbl(acc);
        acc = jac_add_mixed(acc, sj);
        sj = sn;
    }
    for (size_t s = 0; s < cnt; s++) {
        size_t u = base + s * G;
        if ((dgs[u].flags & 3u) != 1u) continue;     // the first mini-scalar was even: subtract P
        Aff<F> t = ps_load<M>(tab[(size_t)M::TAB * u + M::P_ENTRY]);
        t.y = -t.y;
        acc = jac_add_mixed(acc, t);
    }
    store_jac(out[w], acc);
}

// a8: Commitment::evaluate (src/poly.rs:497-508): Horner, acc = acc * x + C_k.  The coefficient table is
// decoded once (affine, Montgomery form; the point at infinity is stored as (0, 0), which is not on the
// curve) so that "+ C_k" is a mixed addition; "acc * x" skips the leading zero bits of x (the indices
// are small integers in the reference's use: public_key_share(i) = evaluate(i + 1), src/lib.rs:596-600).
TCB_HD Aff<Fp> aff1_load(const Aff1Store &e) {
    Aff<Fp> t;
    t.x = e.x; t.y = e.y;
    t.inf = e.x.is_zero() && e.y.is_zero();
    return t;
}
// index of the leading one of a 256-bit scalar (-1 for zero)
TCB_HD int scalar_top_bit(const u32 *k) {
    for (int b = 255; b >= 0; b--)
        if ((k[b >> 5] >> (b & 31)) & 1) return b;
    return -1;
}
// one Horner step acc <- acc * x + C
TCB_HD Jac<Fp> commit_eval_step(const Jac<Fp> &acc_in, const u32 *k, int top, const Aff<Fp> &c) {
    Jac<Fp> acc = acc_in;
    if (top < 0) acc = jac_inf<Fp>();
    else {
        Jac<Fp> base = acc;
        Fp bz2 = sqr(base.z), bz3 = bz2 * base.z;       // shared by every "+ base" of this step
        for (int b = top - 1; b >= 0; b--) {
            acc = jac_dbl(acc);
            if ((k[b >> 5] >> (b & 31)) & 1) acc = jac_add_cached(acc, base, bz2, bz3);
        }
    }
    return jac_add_mixed(acc, c);
}
TCB_HD void task_commit_eval(size_t i, size_t deg, const Aff1Store *coeff, const u8 *x_fr, u8 *out_g1) {
    u32 k[8];
    load_scalar_le(k, x_fr + 32 * i);
    int top = scalar_top_bit(k);
    Jac<Fp> acc = jac_from_aff(aff1_load(coeff[deg]));
    for (size_t c = deg; c-- > 0;) acc = commit_eval_step(acc, k, top, aff1_load(coeff[c]));
    store_g1(out_g1 + 96 * i, jac_to_aff(acc));
}
// The same evaluation with B units per point (small batches: one thread per point walks deg dependent Horner steps and
// leaves most of the GPU idle).  Unit (i, b) evaluates the coefficient block [b L, (b+1) L) by Horner and multiplies its partial
// sum by x^(b L) mod r (Fr power, then one 255-bit double-and-add), so that sum_b part_{i,b} is the value; the B terms are added by
// k_g1_sum.  Same group element, hence the same output bytes, as the single-unit walk.
TCB_HD void task_commit_eval_part(size_t u, size_t B, size_t L, size_t deg, const Aff1Store *coeff, const u8 *x_fr, Jac1Store *out) {
    size_t i = u / B, b = u % B;
    size_t lo = b * L, hi = lo + L < deg + 1 ? lo + L : deg + 1;       // coefficients [lo, hi)
    Jac<Fp> acc = jac_inf<Fp>();
    if (lo <= deg) {
        u32 k[8];
        load_scalar_le(k, x_fr + 32 * i);
        int top = scalar_top_bit(k);
        acc = jac_from_aff(aff1_load(coeff[hi - 1]));
        for (size_t c = hi - 1; c-- > lo;) acc = commit_eval_step(acc, k, top, aff1_load(coeff[c]));
        if (b) {
            bool ok = true;
            Fr xm = fr_load_le(x_fr + 32 * i, ok), s = fr_one();
            int tb = 63;
            while (tb > 0 && !((lo >> tb) & 1)) tb--;
            for (int j = tb; j >= 0; j--) {
                s = s * s;
                if ((lo >> j) & 1) s = s * xm;
            }
            Fr sc = from_mont<FrParams>(s);
            acc = jac_mul_jac<Fp, 8>(acc, sc.l);
        }
    }
    store_jac(out[u], acc);
}
// §8(f) row 2, PublicKey::encrypt_with_rng (src/lib.rs:128-137) with the random scalar r supplied by
// the caller: u = g1 * r, v = xor_with_hash(pk * r, msg).  (w = hash_g1_g2(u, v) * r is produced by
// k_hash_g1_g2 + k_sign afterwards.)
TCB_HD void task_encrypt_uv(size_t i, const u8 *pk_g1, const u8 *r_fr, const u8 *msgs, const u64 *off, u8 *u_out, u8 *v_out) {
    u32 k[8];
    load_scalar_le(k, r_fr + 32 * i);
    bool ok = true;
    Aff<Fp> g;
    g.x = CONSTS().g1x; g.y = CONSTS().g1y; g.inf = false;
    store_g1(u_out + 96 * i, jac_to_aff(jac_mul_glv2(g, k)));
    Aff<Fp> pk = load_g1(pk_g1 + 96 * i, ok);
    Aff<Fp> s = jac_to_aff(jac_mul_glv2(pk, k));
    xor_with_hash(v_out + off[i], s, msgs + off[i], (size_t)(off[i + 1] - off[i]));
}
TCB_HD void task_g1_decode(size_t i, const u8 *pts_g1, Aff1Store *out) {
    bool ok = true;
    Aff<Fp> p = load_g1(pts_g1 + 96 * i, ok);
    out[i].x = p.inf ? Fp::zero() : p.x;
    out[i].y = p.inf ? Fp::zero() : p.y;
}

// ----------------------------------------------------------------------------- §8(f) row 1: batched point (de)compression
// Encodings of SURVEY App. B (EXTERNAL pairing 0.16 EncodedPoint): checked decoding = flag bits,
// x < p, on the curve, AND in the r-order subgroup (PublicKey::from_bytes src/lib.rs:140-146,
// Signature::from_bytes :246-252, serde projective::deserialize src/serde_impl.rs:187-218).
template <class F2>
TCB_HD void g2_compress(u8 *out, const Aff<F2> &a) {
    if (a.inf) {
        if (is_writer<F2>()) { for (int i = 0; i < 96; i++) out[i] = 0; out[0] = 0xc0; }
        return;
    }
    bool greatest = fp2_cmp(a.y, -a.y) > 0;
    store_f2_be<F2>(out, a.x);
    // flag bits live in the first byte, which belongs to the c1 half (role 1 in a sliced engine)
    if (!F2::SLICED || my_role<F2>() == 1) out[0] |= greatest ? 0xa0 : 0x80;
}
TCB_HD void task_g1_compress(size_t i, const u8 *unc, u8 *out48) {
    bool ok = true;
    g1_compress(out48 + 48 * i, load_g1(unc + 96 * i, ok));
}
template <class F2>
TCB_HD void task_g2_compress(size_t i, const u8 *unc, u8 *out96) {
    bool ok = true;
    g2_compress<F2>(out96 + 96 * i, load_g2<F2>(unc + 192 * i, ok));
}
// status: 0 ok, 3 invalid.  Output: uncompressed affine (all-zero with the infinity flag if invalid).
TCB_HD void task_g1_decompress(size_t i, const u8 *in48, u8 *out96, u8 *status) {
    const u8 *in = in48 + 48 * i;
    u8 *out = out96 + 96 * i;
    Aff<Fp> a;
    a.inf = true; a.x = Fp::zero(); a.y = Fp::zero();
    bool ok = (in[0] & 0x80) != 0;
    if (ok && (in[0] & 0x40)) {          // infinity: every other bit must be clear
        ok = in[0] == 0xc0;
        for (int k = 1; k < 48; k++) ok = ok && in[k] == 0;
    } else if (ok) {
        u8 tmp[48];
        for (int k = 0; k < 48; k++) tmp[k] = in[k];
        tmp[0] &= 0x1f;
        Fp x = load_fp_be(tmp, ok);
        if (ok) {
            Fp rhs = sqr(x) * x + CONSTS().b1;
            Fp y = fp_pow<ExpPp1d4>(rhs);
            ok = sqr(y) == rhs;
            if (ok) {
                bool want_greatest = (in[0] & 0x20) != 0;
                Fp ny = -y;
                if ((fp_cmp(y, ny) > 0) != want_greatest) y = ny;
                a.x = x; a.y = y; a.inf = false;
                ok = jac_is_inf(jac_mul_const<Fp, ExpR>(a));       // subgroup check: r * P == O
            }
        }
    }
    if (!ok) { a.inf = true; }
    store_g1(out, a);
    if (!ok) out[0] = 0;                 // invalid: all-zero output
    status[i] = ok ? 0 : 3;
}
template <class F2>
TCB_HD void task_g2_decompress(size_t i, const u8 *in96, u8 *out192, u8 *status) {
    const u8 *in = in96 + 96 * i;
    u8 *out = out192 + 192 * i;
    Aff<F2> a;
    a.inf = true; a.x = F2::zero(); a.y = F2::zero();
    bool ok = (in[0] & 0x80) != 0;
    if (ok && (in[0] & 0x40)) {
        ok = in[0] == 0xc0;
        for (int k = 1; k < 96; k++) ok = ok && in[k] == 0;
    } else if (ok) {
        u8 tmp[96];
        for (int k = 0; k < 96; k++) tmp[k] = in[k];
        tmp[0] &= 0x1f;
        bool dec = true;
        F2 x = load_f2_be<F2>(tmp, dec);
        ok = unit_and<F2>(dec);          // a sliced engine validates its own half only
        if (ok) {
            const Consts &C = CONSTS();
            F2 rhs = sqr(x) * x + F2::from_halves(C.b1, C.b1);
            F2 y;
            ok = fp2_sqrt(y, rhs);
            if (ok) {
                bool want_greatest = (in[0] & 0x20) != 0;
                F2 ny = -y;
                if ((fp2_cmp(y, ny) > 0) != want_greatest) y = ny;
                a.x = x; a.y = y; a.inf = false;
                ok = jac_is_inf(jac_mul_const<F2, ExpR>(a));
            }
        }
    }
    if (!ok) a.inf = true;
    store_g2<F2>(out, a);
    if (is_writer<F2>()) { if (!ok) out[0] = 0; status[i] = ok ? 0 : 3; }
}

// ----------------------------------------------------------------------------- constants builder (host side; runs at tcb_init)
// Computes every derived constant from p, r and the generator coordinates using the same
// field code, so nothing but the curve definition is hard-coded.
inline void limbs_from_hex(u32 *l, int n, const char *hex) {
    for (int i = 0; i < n; i++) l[i] = 0;
    int len = 0;
    while (hex[len]) len++;
    for (int i = 0; i < len; i++) {
        char c = hex[len - 1 - i];
        u32 v = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
        l[i / 8] |= v << (4 * (i % 8));
    }
}
template <class P>
inline Mont<P> pow2_mod(int bits) {
    Mont<P> r = Mont<P>::zero();
    r.l[0] = 1;
    for (int i = 0; i < bits; i++) r = madd<P>(r, r);
    return r;
}
inline void build_consts(Consts &C) {
    C.r1 = pow2_mod<FpParams>(384);
    C.r2 = pow2_mod<FpParams>(768);
    C.fr_r1 = pow2_mod<FrParams>(256);
    C.fr_r2 = pow2_mod<FrParams>(512);
    C.r3 = C.r2 * C.r2;   // R^2 * R^2 * R^-1
    h_consts.r1 = C.r1; h_consts.r2 = C.r2; h_consts.r3 = C.r3; h_consts.fr_r1 = C.fr_r1; h_consts.fr_r2 = C.fr_r2;
    Fp four = Fp::zero();
    four.l[0] = 4;
    C.b1 = fp_to_mont(four);
    Fp t;
    limbs_from_hex(t.l, 12, "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"); C.g1x = fp_to_mont(t);
    limbs_from_hex(t.l, 12, "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1"); C.g1y = fp_to_mont(t);
    limbs_from_hex(t.l, 12, "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"); C.g2x.c0 = fp_to_mont(t);
    limbs_from_hex(t.l, 12, "13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"); C.g2x.c1 = fp_to_mont(t);
    limbs_from_hex(t.l, 12, "0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801"); C.g2y.c0 = fp_to_mont(t);
    limbs_from_hex(t.l, 12, "0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be"); C.g2y.c1 = fp_to_mont(t);
    // Frobenius: g1 = xi^((p-1)/6) (computed by repeated squaring over the exponent bits),
    // g2 = g1 * conj(g1), g3 = g2 * g1; frob[k][m] = gk^m
    u32 e[12];
    limbs_from_hex(e, 12, "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaaa");
    {   // e = (p - 1) / 6
        u64 rem = 0;
        for (int i = 11; i >= 0; i--) { u64 cur = (rem << 32) | e[i]; e[i] = (u32)(cur / 6); rem = cur % 6; }
    }
    Fp2 xi = Fp2::from_halves(C.r1, C.r1), g[4];
    g[1] = Fp2::one();
    bool started = false;
    for (int i = 383; i >= 0; i--) {
        if (started) g[1] = sqr(g[1]);
        if ((e[i >> 5] >> (i & 31)) & 1) { g[1] = started ? g[1] * xi : xi; started = true; }
    }
    g[2] = g[1] * conj(g[1]);
    g[3] = g[2] * g[1];
    for (int k = 1; k <= 3; k++) {
        Fp2 acc = Fp2::one();
        for (int m = 0; m < 6; m++) { acc.store(C.frob[k][m]); acc = acc * g[k]; }
    }
    for (int m = 0; m < 6; m++) Fp2::one().store(C.frob[0][m]);
    h_consts = C;
    // psi(x, y) = (conj(x) / xi^((p-1)/3), conj(y) / xi^((p-1)/2))
    inv(Fp2::load(C.frob[1][2])).store(C.psi_x);
    inv(Fp2::load(C.frob[1][3])).store(C.psi_y);
    {   // psi^i coefficients: c_0 = 1, c_{i+1} = conj(c_i) * psi
        Fp2 cx = Fp2::one(), cy = Fp2::one();
        for (int i = 0; i < 4; i++) {
            cx.store(C.psi_cx[i]); cy.store(C.psi_cy[i]);
            cx = conj(cx) * Fp2::load(C.psi_x); cy = conj(cy) * Fp2::load(C.psi_y);
        }
    }
    h_consts = C;
    {   // beta: the cube root of unity in Fp with (beta x, y) = [-x^2](x, y) on G1
        u32 e[12];
        limbs_from_hex(e, 12, "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaaa");
        u64 rem = 0;
        for (int i = 11; i >= 0; i--) { u64 cur = (rem << 32) | e[i]; e[i] = (u32)(cur / 3); rem = cur % 3; }   // (p - 1) / 3
        Fp beta = C.r1;
        for (u32 g = 2; beta == C.r1; g++) {
            Fp base = Fp::zero();
            base.l[0] = g;
            base = fp_to_mont(base);
            Fp acc = C.r1;
            for (int i = 383; i >= 0; i--) { acc = sqr(acc); if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * base; }
            beta = acc;
        }
        Aff<Fp> G; G.x = C.g1x; G.y = C.g1y; G.inf = false;
        const u32 mu[8] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u, 0, 0, 0, 0};   // x^2
        Aff<Fp> m = jac_to_aff(jac_mul_aff<Fp, 8>(G, mu));       // [x^2] G = -phi(G)
        if (!(m.x == G.x * beta)) beta = sqr(beta);
        C.beta = beta;
        u32 c3[8];                                               // 3 (x^2 - 1) < 2^130
        u64 cy = 0, bw = 1;
        for (int i = 0; i < 8; i++) { u64 d = (u64)mu[i] - bw; c3[i] = (u32)d; bw = (d >> 63) & 1; }
        for (int i = 0; i < 8; i++) { u64 v = 3ull * c3[i] + cy; c3[i] = (u32)v; cy = v >> 32; }
        Aff<Fp> gc = jac_to_aff(jac_mul_aff<Fp, 8>(G, c3));
        C.g1cx = gc.x; C.g1cy = gc.y;
    }
    h_consts = C;
}

}  // namespace tcb
