/*
 * tc_oracle.c — CPU restatement of the threshold_crypto hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (threshold_crypto_b200/) never links or calls it.
 *
 * PARITY STATUS: **unpinned**.  The reference crate (threshold_crypto 0.4.0, /root/reference)
 * has no BLS12-381 known-answer vectors, cannot be built here (no Rust toolchain), and its
 * arithmetic lives in the absent crates pairing 0.16.0, ff 0.6.0, group 0.6.0, rand 0.7.3,
 * rand_chacha 0.2.2, tiny-keccak 2.0.1 (Cargo.toml:22-33).  This file restates their
 * published algorithms (6x64-bit Montgomery Fq, 2-3-2 tower, Jacobian G1/G2, MSB-first
 * double-and-add, optimal-ate Miller loop with M-twist lines through mul_by_014, the
 * easy-part/hard-part final exponentiation with generic pow(|x|)) so that its timing is a
 * fair stand-in for the `pairing`-backed CPU path.  It is cross-checked against the
 * independent big-int model oracle/pyref.py and the golden vectors in tests/golden/.
 *
 * Reference call sites followed (relative to /root/reference):
 *   verify_g2 src/lib.rs:108-110 | verify :115-117 | verify_decryption_share :182-186
 *   sign_g2 :372-374 | sign :379-381 | decrypt_share_no_verify :460-462
 *   Ciphertext::verify :508-512 | combine_signatures :608-615 | decrypt :618-626
 *   hash_g2 :691-694 | hash_g1_g2 :697-707 | xor_with_hash :710-715
 *   interpolate :719-767 | into_fr_plus_1 :769-773
 *   Commitment::evaluate src/poly.rs:497-508 | Poly::commitment src/poly.rs:372-377
 *   sha3_256 src/util.rs:3-9 | encodings src/serde_impl.rs:174-218, src/lib.rs:140-153
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;
typedef unsigned __int128 u128;

#define INL static inline __attribute__((always_inline))

/* ------------------------------------------------------------------ op counter (per thread) */
static __thread u64 g_fp_mul_count;
u64 orc_fp_mul_count(void) { return g_fp_mul_count; }
void orc_fp_mul_count_reset(void) { g_fp_mul_count = 0; }

/* ------------------------------------------------------------------ generic limb helpers */
INL int mp_geq(const u64 *a, const u64 *b, int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
INL int mp_cmp(const u64 *a, const u64 *b, int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return -1;
    }
    return 0;
}
INL int mp_is_zero(const u64 *a, int n) {
    u64 t = 0;
    for (int i = 0; i < n; i++) t |= a[i];
    return t == 0;
}
INL u64 mp_add(u64 *r, const u64 *a, const u64 *b, int n) {
    u64 c = 0;
    for (int i = 0; i < n; i++) {
        u128 t = (u128)a[i] + b[i] + c;
        r[i] = (u64)t;
        c = (u64)(t >> 64);
    }
    return c;
}
INL u64 mp_sub(u64 *r, const u64 *a, const u64 *b, int n) {
    u64 bw = 0;
    for (int i = 0; i < n; i++) {
        u128 t = (u128)a[i] - b[i] - bw;
        r[i] = (u64)t;
        bw = (u64)(t >> 64) & 1;
    }
    return bw;
}
INL void mp_add_mod(u64 *r, const u64 *a, const u64 *b, const u64 *m, int n) {
    u64 c = mp_add(r, a, b, n);
    if (c || mp_geq(r, m, n)) mp_sub(r, r, m, n);
}
INL void mp_sub_mod(u64 *r, const u64 *a, const u64 *b, const u64 *m, int n) {
    if (mp_sub(r, a, b, n)) mp_add(r, r, m, n);
}
INL void mp_neg_mod(u64 *r, const u64 *a, const u64 *m, int n) {
    if (mp_is_zero(a, n)) { for (int i = 0; i < n; i++) r[i] = 0; }
    else mp_sub(r, m, a, n);
}
/* Montgomery CIOS multiplication, R = 2^(64 n) */
INL void mp_mont_mul(u64 *r, const u64 *a, const u64 *b, const u64 *m, u64 inv, int n) {
    u64 t[8] = {0};
    u64 t_hi = 0, t_hi2 = 0;
    for (int i = 0; i < n; i++) {
        u64 c = 0;
        for (int j = 0; j < n; j++) {
            u128 p = (u128)a[j] * b[i] + t[j] + c;
            t[j] = (u64)p;
            c = (u64)(p >> 64);
        }
        u128 s = (u128)t_hi + c;
        t_hi = (u64)s;
        t_hi2 = (u64)(s >> 64);
        u64 mm = t[0] * inv;
        u128 p = (u128)mm * m[0] + t[0];
        c = (u64)(p >> 64);
        for (int j = 1; j < n; j++) {
            p = (u128)mm * m[j] + t[j] + c;
            t[j - 1] = (u64)p;
            c = (u64)(p >> 64);
        }
        s = (u128)t_hi + c;
        t[n - 1] = (u64)s;
        t_hi = t_hi2 + (u64)(s >> 64);
    }
    if (t_hi || mp_geq(t, m, n)) mp_sub(r, t, m, n);
    else for (int i = 0; i < n; i++) r[i] = t[i];
}

/* ------------------------------------------------------------------ Fp (381 bits) */
typedef struct { u64 l[6]; } fp;
static const fp FP_P = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                         0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static u64 FP_INV;          /* -p^-1 mod 2^64 */
static fp FP_R1, FP_R2;     /* R mod p, R^2 mod p */
static fp FP_ZERO;
static u64 FP_EXP_PM2[6], FP_EXP_SQRT[6], FP_EXP_PM3D4[6], FP_EXP_PM1D2[6], FP_EXP_PM1D6[6];

INL void fp_add(fp *r, const fp *a, const fp *b) { mp_add_mod(r->l, a->l, b->l, FP_P.l, 6); }
INL void fp_sub(fp *r, const fp *a, const fp *b) { mp_sub_mod(r->l, a->l, b->l, FP_P.l, 6); }
INL void fp_neg(fp *r, const fp *a) { mp_neg_mod(r->l, a->l, FP_P.l, 6); }
INL void fp_dbl(fp *r, const fp *a) { fp_add(r, a, a); }
INL void fp_mul(fp *r, const fp *a, const fp *b) {
    g_fp_mul_count++;
    mp_mont_mul(r->l, a->l, b->l, FP_P.l, FP_INV, 6);
}
INL void fp_sqr(fp *r, const fp *a) { fp_mul(r, a, a); }
INL int fp_is_zero(const fp *a) { return mp_is_zero(a->l, 6); }
INL int fp_eq(const fp *a, const fp *b) { return mp_cmp(a->l, b->l, 6) == 0; }
static void fp_pow(fp *r, const fp *a, const u64 *e, int n) {
    fp acc = FP_R1, base = *a;
    int started = 0;
    for (int i = n * 64 - 1; i >= 0; i--) {
        if (started) fp_sqr(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) {
            if (started) fp_mul(&acc, &acc, &base); else { acc = base; started = 1; }
        }
    }
    *r = acc;
}
static void fp_inv(fp *r, const fp *a) { fp_pow(r, a, FP_EXP_PM2, 6); }
static void fp_from_mont(fp *r, const fp *a) {
    fp one = {{1, 0, 0, 0, 0, 0}};
    mp_mont_mul(r->l, a->l, one.l, FP_P.l, FP_INV, 6);
}
static void fp_to_mont(fp *r, const fp *a) { mp_mont_mul(r->l, a->l, FP_R2.l, FP_P.l, FP_INV, 6); }
/* canonical-integer comparison (EXTERNAL ff: Ord on into_repr()) */
static int fp_cmp(const fp *a, const fp *b) {
    fp ca, cb;
    fp_from_mont(&ca, a);
    fp_from_mont(&cb, b);
    return mp_cmp(ca.l, cb.l, 6);
}
static int fp_sqrt(fp *r, const fp *a) {
    fp s, chk;
    fp_pow(&s, a, FP_EXP_SQRT, 6);
    fp_sqr(&chk, &s);
    if (!fp_eq(&chk, a)) return 0;
    *r = s;
    return 1;
}
/* 48-byte big-endian canonical <-> Montgomery; returns 0 if value >= p */
static int fp_from_be(fp *r, const u8 *b) {
    fp t;
    for (int i = 0; i < 6; i++) {
        u64 v = 0;
        for (int k = 0; k < 8; k++) v = (v << 8) | b[(5 - i) * 8 + k];
        t.l[i] = v;
    }
    if (mp_geq(t.l, FP_P.l, 6)) return 0;
    fp_to_mont(r, &t);
    return 1;
}
static void fp_to_be(u8 *b, const fp *a) {
    fp t;
    fp_from_mont(&t, a);
    for (int i = 0; i < 6; i++)
        for (int k = 0; k < 8; k++) b[(5 - i) * 8 + k] = (u8)(t.l[i] >> (56 - 8 * k));
}

/* ------------------------------------------------------------------ Fr (255 bits) */
typedef struct { u64 l[4]; } fr;
static const fr FR_R = {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}};
static u64 FR_INV;
static fr FR_R1, FR_R2;
static u64 FR_EXP_RM2[4];
INL void fr_add(fr *r, const fr *a, const fr *b) { mp_add_mod(r->l, a->l, b->l, FR_R.l, 4); }
INL void fr_sub(fr *r, const fr *a, const fr *b) { mp_sub_mod(r->l, a->l, b->l, FR_R.l, 4); }
INL void fr_mul(fr *r, const fr *a, const fr *b) { mp_mont_mul(r->l, a->l, b->l, FR_R.l, FR_INV, 4); }
INL int fr_is_zero(const fr *a) { return mp_is_zero(a->l, 4); }
INL int fr_eq(const fr *a, const fr *b) { return mp_cmp(a->l, b->l, 4) == 0; }
static int fr_inv(fr *r, const fr *a) {
    if (fr_is_zero(a)) return 0;
    fr acc = FR_R1;
    for (int i = 255; i >= 0; i--) {
        fr_mul(&acc, &acc, &acc);
        if ((FR_EXP_RM2[i / 64] >> (i % 64)) & 1) fr_mul(&acc, &acc, a);
    }
    *r = acc;
    return 1;
}
static void fr_from_mont(fr *r, const fr *a) {
    fr one = {{1, 0, 0, 0}};
    mp_mont_mul(r->l, a->l, one.l, FR_R.l, FR_INV, 4);
}
/* 32-byte little-endian canonical (serde_impl.rs:109) -> Montgomery; value reduced mod r if >= r is rejected */
static int fr_from_le(fr *r, const u8 *b) {
    fr t;
    for (int i = 0; i < 4; i++) {
        u64 v = 0;
        for (int k = 7; k >= 0; k--) v = (v << 8) | b[i * 8 + k];
        t.l[i] = v;
    }
    if (mp_geq(t.l, FR_R.l, 4)) return 0;
    mp_mont_mul(r->l, t.l, FR_R2.l, FR_R.l, FR_INV, 4);
    return 1;
}
static void fr_canon_le(u64 *out, const u8 *b) {
    for (int i = 0; i < 4; i++) {
        u64 v = 0;
        for (int k = 7; k >= 0; k--) v = (v << 8) | b[i * 8 + k];
        out[i] = v;
    }
}

/* ------------------------------------------------------------------ Fp2 = Fp[u]/(u^2+1) */
typedef struct { fp c0, c1; } fp2;
static fp2 FP2_ZERO, FP2_ONE;
INL void fp2_add(fp2 *r, const fp2 *a, const fp2 *b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
INL void fp2_sub(fp2 *r, const fp2 *a, const fp2 *b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
INL void fp2_neg(fp2 *r, const fp2 *a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
INL void fp2_dbl(fp2 *r, const fp2 *a) { fp2_add(r, a, a); }
INL void fp2_conj(fp2 *r, const fp2 *a) { r->c0 = a->c0; fp_neg(&r->c1, &a->c1); }
INL int fp2_is_zero(const fp2 *a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
INL int fp2_eq(const fp2 *a, const fp2 *b) { return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1); }
static void fp2_mul(fp2 *r, const fp2 *a, const fp2 *b) {
    fp aa, bb, s, t;
    fp_mul(&aa, &a->c0, &b->c0);
    fp_mul(&bb, &a->c1, &b->c1);
    fp_add(&s, &a->c0, &a->c1);
    fp_add(&t, &b->c0, &b->c1);
    fp_mul(&s, &s, &t);
    fp_sub(&s, &s, &aa);
    fp_sub(&r->c1, &s, &bb);
    fp_sub(&r->c0, &aa, &bb);
}
static void fp2_sqr(fp2 *r, const fp2 *a) {
    fp s, d, m;
    fp_add(&s, &a->c0, &a->c1);
    fp_sub(&d, &a->c0, &a->c1);
    fp_mul(&m, &a->c0, &a->c1);
    fp_mul(&r->c0, &s, &d);
    fp_dbl(&r->c1, &m);
}
INL void fp2_mul_fp(fp2 *r, const fp2 *a, const fp *k) { fp_mul(&r->c0, &a->c0, k); fp_mul(&r->c1, &a->c1, k); }
INL void fp2_mul_xi(fp2 *r, const fp2 *a) { /* * (1 + u) */
    fp t0;
    fp_sub(&t0, &a->c0, &a->c1);
    fp_add(&r->c1, &a->c0, &a->c1);
    r->c0 = t0;
}
static void fp2_inv(fp2 *r, const fp2 *a) {
    fp n, t;
    fp_sqr(&n, &a->c0);
    fp_sqr(&t, &a->c1);
    fp_add(&n, &n, &t);
    fp_inv(&n, &n);
    fp_mul(&r->c0, &a->c0, &n);
    fp_mul(&t, &a->c1, &n);
    fp_neg(&r->c1, &t);
}
static void fp2_pow(fp2 *r, const fp2 *a, const u64 *e, int n) {
    fp2 acc = FP2_ONE;
    int started = 0;
    for (int i = n * 64 - 1; i >= 0; i--) {
        if (started) fp2_sqr(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) {
            if (started) fp2_mul(&acc, &acc, a); else { acc = *a; started = 1; }
        }
    }
    *r = acc;
}
/* EXTERNAL pairing Fq2 Ord: by c1, then c0, on canonical integers */
static int fp2_cmp(const fp2 *a, const fp2 *b) {
    int c = fp_cmp(&a->c1, &b->c1);
    return c ? c : fp_cmp(&a->c0, &b->c0);
}
/* Algorithm 9 of eprint 2012/685 (as EXTERNAL pairing Fq2::sqrt) */
static int fp2_sqrt(fp2 *r, const fp2 *a) {
    if (fp2_is_zero(a)) { *r = FP2_ZERO; return 1; }
    fp2 a1, alpha, a0, neg1, t;
    fp2_pow(&a1, a, FP_EXP_PM3D4, 6);
    fp2_sqr(&alpha, &a1);
    fp2_mul(&alpha, &alpha, a);
    fp2_conj(&a0, &alpha);             /* frobenius_map(1) */
    fp2_mul(&a0, &a0, &alpha);
    neg1 = FP2_ZERO;
    fp_neg(&neg1.c0, &FP_R1);
    if (fp2_eq(&a0, &neg1)) return 0;
    fp2_mul(&a1, &a1, a);
    if (fp2_eq(&alpha, &neg1)) {
        t.c0 = FP_ZERO; t.c1 = FP_R1;
        fp2_mul(&a1, &a1, &t);
    } else {
        fp2_add(&alpha, &alpha, &FP2_ONE);
        fp2_pow(&alpha, &alpha, FP_EXP_PM1D2, 6);
        fp2_mul(&a1, &a1, &alpha);
    }
    *r = a1;
    return 1;
}

/* ------------------------------------------------------------------ Fp6 = Fp2[v]/(v^3 - xi) */
typedef struct { fp2 c0, c1, c2; } fp6;
static void fp6_add(fp6 *r, const fp6 *a, const fp6 *b) { fp2_add(&r->c0, &a->c0, &b->c0); fp2_add(&r->c1, &a->c1, &b->c1); fp2_add(&r->c2, &a->c2, &b->c2); }
static void fp6_sub(fp6 *r, const fp6 *a, const fp6 *b) { fp2_sub(&r->c0, &a->c0, &b->c0); fp2_sub(&r->c1, &a->c1, &b->c1); fp2_sub(&r->c2, &a->c2, &b->c2); }
static void fp6_neg(fp6 *r, const fp6 *a) { fp2_neg(&r->c0, &a->c0); fp2_neg(&r->c1, &a->c1); fp2_neg(&r->c2, &a->c2); }
static void fp6_mul_v(fp6 *r, const fp6 *a) { /* * v */
    fp2 t;
    fp2_mul_xi(&t, &a->c2);
    r->c2 = a->c1;
    r->c1 = a->c0;
    r->c0 = t;
}
static void fp6_mul(fp6 *r, const fp6 *a, const fp6 *b) {
    fp2 v0, v1, v2, s, t, c0, c1, c2;
    fp2_mul(&v0, &a->c0, &b->c0);
    fp2_mul(&v1, &a->c1, &b->c1);
    fp2_mul(&v2, &a->c2, &b->c2);
    fp2_add(&s, &a->c1, &a->c2); fp2_add(&t, &b->c1, &b->c2);
    fp2_mul(&c0, &s, &t); fp2_sub(&c0, &c0, &v1); fp2_sub(&c0, &c0, &v2);
    fp2_mul_xi(&c0, &c0); fp2_add(&c0, &c0, &v0);
    fp2_add(&s, &a->c0, &a->c1); fp2_add(&t, &b->c0, &b->c1);
    fp2_mul(&c1, &s, &t); fp2_sub(&c1, &c1, &v0); fp2_sub(&c1, &c1, &v1);
    fp2_mul_xi(&s, &v2); fp2_add(&c1, &c1, &s);
    fp2_add(&s, &a->c0, &a->c2); fp2_add(&t, &b->c0, &b->c2);
    fp2_mul(&c2, &s, &t); fp2_sub(&c2, &c2, &v0); fp2_sub(&c2, &c2, &v2); fp2_add(&c2, &c2, &v1);
    r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
/* self * (c1 v) */
static void fp6_mul_by_1(fp6 *r, const fp6 *a, const fp2 *c1) {
    fp2 b_b, t1, t2, s;
    fp2_mul(&b_b, &a->c1, c1);
    fp2_add(&s, &a->c1, &a->c2);
    fp2_mul(&t1, &s, c1); fp2_sub(&t1, &t1, &b_b); fp2_mul_xi(&t1, &t1);
    fp2_add(&s, &a->c0, &a->c1);
    fp2_mul(&t2, &s, c1); fp2_sub(&t2, &t2, &b_b);
    r->c0 = t1; r->c1 = t2; r->c2 = b_b;
}
/* self * (c0 + c1 v) */
static void fp6_mul_by_01(fp6 *r, const fp6 *a, const fp2 *c0, const fp2 *c1) {
    fp2 a_a, b_b, t1, t2, t3, s, t;
    fp2_mul(&a_a, &a->c0, c0);
    fp2_mul(&b_b, &a->c1, c1);
    fp2_add(&s, &a->c1, &a->c2);
    fp2_mul(&t1, &s, c1); fp2_sub(&t1, &t1, &b_b); fp2_mul_xi(&t1, &t1); fp2_add(&t1, &t1, &a_a);
    fp2_add(&s, &a->c0, &a->c2);
    fp2_mul(&t3, &s, c0); fp2_sub(&t3, &t3, &a_a); fp2_add(&t3, &t3, &b_b);
    fp2_add(&s, &a->c0, &a->c1); fp2_add(&t, c0, c1);
    fp2_mul(&t2, &s, &t); fp2_sub(&t2, &t2, &a_a); fp2_sub(&t2, &t2, &b_b);
    r->c0 = t1; r->c1 = t2; r->c2 = t3;
}
static void fp6_inv(fp6 *r, const fp6 *a) {
    fp2 t0, t1, t2, s, d;
    fp2_sqr(&t0, &a->c0); fp2_mul(&s, &a->c1, &a->c2); fp2_mul_xi(&s, &s); fp2_sub(&t0, &t0, &s);
    fp2_sqr(&t1, &a->c2); fp2_mul_xi(&t1, &t1); fp2_mul(&s, &a->c0, &a->c1); fp2_sub(&t1, &t1, &s);
    fp2_sqr(&t2, &a->c1); fp2_mul(&s, &a->c0, &a->c2); fp2_sub(&t2, &t2, &s);
    fp2_mul(&d, &a->c2, &t1); fp2_mul(&s, &a->c1, &t2); fp2_add(&d, &d, &s); fp2_mul_xi(&d, &d);
    fp2_mul(&s, &a->c0, &t0); fp2_add(&d, &d, &s);
    fp2_inv(&d, &d);
    fp2_mul(&r->c0, &t0, &d); fp2_mul(&r->c1, &t1, &d); fp2_mul(&r->c2, &t2, &d);
}

/* ------------------------------------------------------------------ Fp12 = Fp6[w]/(w^2 - v) */
typedef struct { fp6 c0, c1; } fp12;
static fp12 FP12_ONE;
static fp2 FROB[4][6];      /* FROB[k][m] = xi^(m (p^k - 1)/6) */
static void fp12_mul(fp12 *r, const fp12 *a, const fp12 *b) {
    fp6 aa, bb, s, t, c1;
    fp6_mul(&aa, &a->c0, &b->c0);
    fp6_mul(&bb, &a->c1, &b->c1);
    fp6_add(&s, &a->c0, &a->c1); fp6_add(&t, &b->c0, &b->c1);
    fp6_mul(&c1, &s, &t); fp6_sub(&c1, &c1, &aa); fp6_sub(&c1, &c1, &bb);
    fp6_mul_v(&t, &bb); fp6_add(&r->c0, &aa, &t);
    r->c1 = c1;
}
static void fp12_sqr(fp12 *r, const fp12 *a) {
    fp6 ab, s, t, c0;
    fp6_mul(&ab, &a->c0, &a->c1);
    fp6_add(&s, &a->c0, &a->c1);
    fp6_mul_v(&t, &a->c1); fp6_add(&t, &t, &a->c0);
    fp6_mul(&c0, &s, &t); fp6_sub(&c0, &c0, &ab);
    fp6_mul_v(&t, &ab); fp6_sub(&c0, &c0, &t);
    r->c0 = c0;
    fp6_add(&r->c1, &ab, &ab);
}
static void fp12_conj(fp12 *r, const fp12 *a) { r->c0 = a->c0; fp6_neg(&r->c1, &a->c1); }
static void fp12_inv(fp12 *r, const fp12 *a) {
    fp6 t0, t1;
    fp6_mul(&t0, &a->c0, &a->c0);
    fp6_mul(&t1, &a->c1, &a->c1);
    fp6_mul_v(&t1, &t1);
    fp6_sub(&t0, &t0, &t1);
    fp6_inv(&t0, &t0);
    fp6_mul(&r->c0, &a->c0, &t0);
    fp6_mul(&t1, &a->c1, &t0);
    fp6_neg(&r->c1, &t1);
}
static int fp12_eq(const fp12 *a, const fp12 *b) {
    const fp *x = (const fp *)a, *y = (const fp *)b;
    for (int i = 0; i < 12; i++) if (!fp_eq(&x[i], &y[i])) return 0;
    return 1;
}
/* coefficient of v^i w^j has w-degree m = 2i + j */
static void fp12_frob(fp12 *r, const fp12 *a, int k) {
    const fp2 *src[6] = {&a->c0.c0, &a->c1.c0, &a->c0.c1, &a->c1.c1, &a->c0.c2, &a->c1.c2};
    fp2 *dst[6] = {&r->c0.c0, &r->c1.c0, &r->c0.c1, &r->c1.c1, &r->c0.c2, &r->c1.c2};
    for (int m = 0; m < 6; m++) {
        fp2 t = *src[m];
        if (k & 1) fp2_conj(&t, &t);
        fp2_mul(dst[m], &t, &FROB[k][m]);
    }
}
static void fp12_mul_by_014(fp12 *f, const fp2 *c0, const fp2 *c1, const fp2 *c4) {
    fp6 aa, bb, s;
    fp2 o;
    fp6_mul_by_01(&aa, &f->c0, c0, c1);
    fp6_mul_by_1(&bb, &f->c1, c4);
    fp2_add(&o, c1, c4);
    fp6_add(&s, &f->c1, &f->c0);
    fp6_mul_by_01(&s, &s, c0, &o);
    fp6_sub(&s, &s, &aa); fp6_sub(&f->c1, &s, &bb);
    fp6_mul_v(&s, &bb); fp6_add(&f->c0, &s, &aa);
}
static void fp12_pow_u64(fp12 *r, const fp12 *a, u64 e) {
    fp12 acc = FP12_ONE, base = *a;
    int started = 0;
    for (int i = 63; i >= 0; i--) {
        if (started) fp12_sqr(&acc, &acc);
        if ((e >> i) & 1) {
            if (started) fp12_mul(&acc, &acc, &base); else { acc = base; started = 1; }
        }
    }
    *r = acc;
}

/* ------------------------------------------------------------------ curves: Jacobian G1 (over Fp), G2 (over Fp2) */
#define BLS_X 0xd201000000010000ULL
static const u64 G2_COFACTOR[8] = {0xcf1c38e31c7238e5ULL, 0x1616ec6e786f0c70ULL, 0x21537e293a6691aeULL, 0xa628f1cb4d9e82efULL,
                                   0xa68a205b2e5a7ddfULL, 0xcd91de4547085abaULL, 0x091d50792876a202ULL, 0x05d543a95414e7f1ULL};

#define DEFINE_CURVE(NAME, F, F_ADD, F_SUB, F_MUL, F_SQR, F_DBL, F_NEG, F_INV, F_ISZ, F_EQ, F_ONE, F_ZERO) \
typedef struct { F x, y; int inf; } NAME##_aff;                                                     \
typedef struct { F x, y, z; } NAME##_jac;   /* z == 0 <=> infinity */                               \
static void NAME##_set_inf(NAME##_jac *r) { r->x = F_ZERO; r->y = F_ONE; r->z = F_ZERO; }           \
static int NAME##_is_inf(const NAME##_jac *a) { return F_ISZ(&a->z); }                              \
static void NAME##_from_aff(NAME##_jac *r, const NAME##_aff *a) {                                   \
    if (a->inf) NAME##_set_inf(r); else { r->x = a->x; r->y = a->y; r->z = F_ONE; } }               \
static void NAME##_to_aff(NAME##_aff *r, const NAME##_jac *a) {                                     \
    if (NAME##_is_inf(a)) { r->inf = 1; r->x = F_ZERO; r->y = F_ZERO; return; }                     \
    F zi, zi2, zi3; F_INV(&zi, &a->z); F_SQR(&zi2, &zi); F_MUL(&zi3, &zi2, &zi);                    \
    F_MUL(&r->x, &a->x, &zi2); F_MUL(&r->y, &a->y, &zi3); r->inf = 0; }                             \
static void NAME##_double(NAME##_jac *r, const NAME##_jac *p) { /* dbl-2009-l */                    \
    if (NAME##_is_inf(p)) { *r = *p; return; }                                                      \
    F a, b, c, d, e, f, t, x3, y3, z3;                                                              \
    F_SQR(&a, &p->x); F_SQR(&b, &p->y); F_SQR(&c, &b);                                              \
    F_ADD(&d, &p->x, &b); F_SQR(&d, &d); F_SUB(&d, &d, &a); F_SUB(&d, &d, &c); F_DBL(&d, &d);       \
    F_DBL(&e, &a); F_ADD(&e, &e, &a);                                                               \
    F_SQR(&f, &e);                                                                                  \
    F_MUL(&z3, &p->y, &p->z); F_DBL(&z3, &z3);                                                      \
    F_SUB(&x3, &f, &d); F_SUB(&x3, &x3, &d);                                                        \
    F_SUB(&t, &d, &x3); F_MUL(&y3, &e, &t);                                                         \
    F_DBL(&c, &c); F_DBL(&c, &c); F_DBL(&c, &c); F_SUB(&y3, &y3, &c);                               \
    r->x = x3; r->y = y3; r->z = z3; }                                                              \
static void NAME##_add(NAME##_jac *r, const NAME##_jac *p, const NAME##_jac *q) { /* add-2007-bl */ \
    if (NAME##_is_inf(p)) { *r = *q; return; }                                                      \
    if (NAME##_is_inf(q)) { *r = *p; return; }                                                      \
    F z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;                                    \
    F_SQR(&z1z1, &p->z); F_SQR(&z2z2, &q->z);                                                       \
    F_MUL(&u1, &p->x, &z2z2); F_MUL(&u2, &q->x, &z1z1);                                             \
    F_MUL(&s1, &p->y, &q->z); F_MUL(&s1, &s1, &z2z2);                                               \
    F_MUL(&s2, &q->y, &p->z); F_MUL(&s2, &s2, &z1z1);                                               \
    if (F_EQ(&u1, &u2)) { if (F_EQ(&s1, &s2)) { NAME##_double(r, p); } else NAME##_set_inf(r); return; } \
    F_SUB(&h, &u2, &u1); F_DBL(&i, &h); F_SQR(&i, &i); F_MUL(&j, &h, &i);                           \
    F_SUB(&rr, &s2, &s1); F_DBL(&rr, &rr); F_MUL(&v, &u1, &i);                                      \
    F_SQR(&x3, &rr); F_SUB(&x3, &x3, &j); F_SUB(&x3, &x3, &v); F_SUB(&x3, &x3, &v);                 \
    F_SUB(&t, &v, &x3); F_MUL(&y3, &rr, &t); F_MUL(&t, &s1, &j); F_DBL(&t, &t); F_SUB(&y3, &y3, &t);\
    F_ADD(&z3, &p->z, &q->z); F_SQR(&z3, &z3); F_SUB(&z3, &z3, &z1z1); F_SUB(&z3, &z3, &z2z2);      \
    F_MUL(&z3, &z3, &h);                                                                            \
    r->x = x3; r->y = y3; r->z = z3; }                                                              \
static void NAME##_add_mixed(NAME##_jac *r, const NAME##_jac *p, const NAME##_aff *q) { /* madd-2007-bl */ \
    if (q->inf) { *r = *p; return; }                                                                \
    if (NAME##_is_inf(p)) { NAME##_from_aff(r, q); return; }                                        \
    F z1z1, u2, s2, h, hh, i, j, rr, v, t, x3, y3, z3;                                              \
    F_SQR(&z1z1, &p->z); F_MUL(&u2, &q->x, &z1z1);                                                  \
    F_MUL(&s2, &q->y, &p->z); F_MUL(&s2, &s2, &z1z1);                                               \
    if (F_EQ(&p->x, &u2)) { if (F_EQ(&p->y, &s2)) { NAME##_double(r, p); } else NAME##_set_inf(r); return; } \
    F_SUB(&h, &u2, &p->x); F_SQR(&hh, &h); F_DBL(&i, &hh); F_DBL(&i, &i); F_MUL(&j, &h, &i);        \
    F_SUB(&rr, &s2, &p->y); F_DBL(&rr, &rr); F_MUL(&v, &p->x, &i);                                  \
    F_SQR(&x3, &rr); F_SUB(&x3, &x3, &j); F_SUB(&x3, &x3, &v); F_SUB(&x3, &x3, &v);                 \
    F_SUB(&t, &v, &x3); F_MUL(&y3, &rr, &t); F_MUL(&t, &p->y, &j); F_DBL(&t, &t); F_SUB(&y3, &y3, &t); \
    F_ADD(&z3, &p->z, &h); F_SQR(&z3, &z3); F_SUB(&z3, &z3, &z1z1); F_SUB(&z3, &z3, &hh);           \
    r->x = x3; r->y = y3; r->z = z3; }                                                              \
/* EXTERNAL CurveAffine::mul -> mul_bits: MSB-first, double then conditional mixed add */           \
static void NAME##_mul_aff(NAME##_jac *r, const NAME##_aff *p, const u64 *k, int nlimbs) {          \
    NAME##_jac acc; NAME##_set_inf(&acc);                                                           \
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {                                                    \
        NAME##_double(&acc, &acc);                                                                  \
        if ((k[i / 64] >> (i % 64)) & 1) NAME##_add_mixed(&acc, &acc, p); }                         \
    *r = acc; }                                                                                     \
/* EXTERNAL CurveProjective::mul_assign: leading zeros skipped, full Jacobian add */                \
static void NAME##_mul_jac(NAME##_jac *r, const NAME##_jac *p, const u64 *k, int nlimbs) {          \
    NAME##_jac acc, base = *p; NAME##_set_inf(&acc); int found = 0;                                 \
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {                                                    \
        int bit = (k[i / 64] >> (i % 64)) & 1;                                                      \
        if (found) NAME##_double(&acc, &acc); else found = bit;                                     \
        if (bit) NAME##_add(&acc, &acc, &base); }                                                   \
    *r = acc; }                                                                                     \
static int NAME##_on_curve(const NAME##_aff *a, const F *b) {                                       \
    if (a->inf) return 1;                                                                           \
    F l, rr; F_SQR(&l, &a->y); F_SQR(&rr, &a->x); F_MUL(&rr, &rr, &a->x); F_ADD(&rr, &rr, b);       \
    return F_EQ(&l, &rr); }

DEFINE_CURVE(g1, fp, fp_add, fp_sub, fp_mul, fp_sqr, fp_dbl, fp_neg, fp_inv, fp_is_zero, fp_eq, FP_R1, FP_ZERO)
DEFINE_CURVE(g2, fp2, fp2_add, fp2_sub, fp2_mul, fp2_sqr, fp2_dbl, fp2_neg, fp2_inv, fp2_is_zero, fp2_eq, FP2_ONE, FP2_ZERO)

static g1_aff G1_GEN;
static g2_aff G2_GEN;
static fp G1_B;
static fp2 G2_B;

/* ------------------------------------------------------------------ encodings */
static void g1_aff_to_unc(u8 *out, const g1_aff *a) {
    if (a->inf) { memset(out, 0, 96); out[0] = 0x40; return; }
    fp_to_be(out, &a->x); fp_to_be(out + 48, &a->y);
}
static void g2_aff_to_unc(u8 *out, const g2_aff *a) {
    if (a->inf) { memset(out, 0, 192); out[0] = 0x40; return; }
    fp_to_be(out, &a->x.c1); fp_to_be(out + 48, &a->x.c0);
    fp_to_be(out + 96, &a->y.c1); fp_to_be(out + 144, &a->y.c0);
}
/* unchecked-on-subgroup decode of an uncompressed point; returns 0 on malformed field elements */
static int g1_aff_from_unc(g1_aff *a, const u8 *in) {
    if (in[0] & 0x40) { a->inf = 1; a->x = FP_ZERO; a->y = FP_ZERO; return 1; }
    a->inf = 0;
    return fp_from_be(&a->x, in) && fp_from_be(&a->y, in + 48);
}
static int g2_aff_from_unc(g2_aff *a, const u8 *in) {
    if (in[0] & 0x40) { a->inf = 1; a->x = FP2_ZERO; a->y = FP2_ZERO; return 1; }
    a->inf = 0;
    return fp_from_be(&a->x.c1, in) && fp_from_be(&a->x.c0, in + 48) &&
           fp_from_be(&a->y.c1, in + 96) && fp_from_be(&a->y.c0, in + 144);
}
static void g1_aff_compress(u8 *out, const g1_aff *a) {
    if (a->inf) { memset(out, 0, 48); out[0] = 0xc0; return; }
    fp ny;
    fp_to_be(out, &a->x);
    fp_neg(&ny, &a->y);
    out[0] |= 0x80;
    if (fp_cmp(&a->y, &ny) > 0) out[0] |= 0x20;
}
static void g2_aff_compress(u8 *out, const g2_aff *a) {
    if (a->inf) { memset(out, 0, 96); out[0] = 0xc0; return; }
    fp2 ny;
    fp_to_be(out, &a->x.c1); fp_to_be(out + 48, &a->x.c0);
    fp2_neg(&ny, &a->y);
    out[0] |= 0x80;
    if (fp2_cmp(&a->y, &ny) > 0) out[0] |= 0x20;
}
/* checked decode (on curve AND in the r-subgroup), as EncodedPoint::into_affine; 0 = invalid */
static int g1_aff_decompress(g1_aff *a, const u8 *in) {
    if (!(in[0] & 0x80)) return 0;
    if (in[0] & 0x40) {
        if (in[0] != 0xc0) return 0;
        for (int i = 1; i < 48; i++) if (in[i]) return 0;
        a->inf = 1; a->x = FP_ZERO; a->y = FP_ZERO; return 1;
    }
    u8 tmp[48]; memcpy(tmp, in, 48); tmp[0] &= 0x1f;
    fp x, y, ny, t;
    if (!fp_from_be(&x, tmp)) return 0;
    fp_sqr(&t, &x); fp_mul(&t, &t, &x); fp_add(&t, &t, &G1_B);
    if (!fp_sqrt(&y, &t)) return 0;
    fp_neg(&ny, &y);
    int greatest = (in[0] & 0x20) != 0;
    if ((fp_cmp(&y, &ny) > 0) != greatest) y = ny;
    a->x = x; a->y = y; a->inf = 0;
    g1_jac chk; g1_mul_aff(&chk, a, FR_R.l, 4);
    return g1_is_inf(&chk);
}
static int g2_aff_decompress(g2_aff *a, const u8 *in) {
    if (!(in[0] & 0x80)) return 0;
    if (in[0] & 0x40) {
        if (in[0] != 0xc0) return 0;
        for (int i = 1; i < 96; i++) if (in[i]) return 0;
        a->inf = 1; a->x = FP2_ZERO; a->y = FP2_ZERO; return 1;
    }
    u8 tmp[48]; memcpy(tmp, in, 48); tmp[0] &= 0x1f;
    fp2 x, y, ny, t;
    if (!fp_from_be(&x.c1, tmp) || !fp_from_be(&x.c0, in + 48)) return 0;
    fp2_sqr(&t, &x); fp2_mul(&t, &t, &x); fp2_add(&t, &t, &G2_B);
    if (!fp2_sqrt(&y, &t)) return 0;
    fp2_neg(&ny, &y);
    int greatest = (in[0] & 0x20) != 0;
    if ((fp2_cmp(&y, &ny) > 0) != greatest) y = ny;
    a->x = x; a->y = y; a->inf = 0;
    g2_jac chk; g2_mul_aff(&chk, a, FR_R.l, 4);
    return g2_is_inf(&chk);
}

/* ------------------------------------------------------------------ pairing (as EXTERNAL pairing 0.16 bls12_381) */
typedef struct { fp2 a, b, c; } line_t;
static void doubling_step(line_t *l, g2_jac *r) {
    fp2 tmp0, tmp1, tmp2, tmp3, tmp4, tmp5, tmp6, zsq, t;
    fp2_sqr(&tmp0, &r->x);
    fp2_sqr(&tmp1, &r->y);
    fp2_sqr(&tmp2, &tmp1);
    fp2_add(&tmp3, &tmp1, &r->x); fp2_sqr(&tmp3, &tmp3); fp2_sub(&tmp3, &tmp3, &tmp0); fp2_sub(&tmp3, &tmp3, &tmp2);
    fp2_dbl(&tmp3, &tmp3);
    fp2_dbl(&tmp4, &tmp0); fp2_add(&tmp4, &tmp4, &tmp0);
    fp2_add(&tmp6, &r->x, &tmp4);
    fp2_sqr(&tmp5, &tmp4);
    fp2_sqr(&zsq, &r->z);
    fp2_sub(&r->x, &tmp5, &tmp3); fp2_sub(&r->x, &r->x, &tmp3);
    fp2_add(&r->z, &r->z, &r->y); fp2_sqr(&r->z, &r->z); fp2_sub(&r->z, &r->z, &tmp1); fp2_sub(&r->z, &r->z, &zsq);
    fp2_sub(&r->y, &tmp3, &r->x); fp2_mul(&r->y, &r->y, &tmp4);
    fp2_dbl(&tmp2, &tmp2); fp2_dbl(&tmp2, &tmp2); fp2_dbl(&tmp2, &tmp2);
    fp2_sub(&r->y, &r->y, &tmp2);
    fp2_mul(&tmp3, &tmp4, &zsq); fp2_dbl(&tmp3, &tmp3); fp2_neg(&tmp3, &tmp3);
    fp2_sqr(&tmp6, &tmp6); fp2_sub(&tmp6, &tmp6, &tmp0); fp2_sub(&tmp6, &tmp6, &tmp5);
    fp2_dbl(&t, &tmp1); fp2_dbl(&t, &t); fp2_sub(&tmp6, &tmp6, &t);
    fp2_mul(&tmp0, &r->z, &zsq); fp2_dbl(&tmp0, &tmp0);
    l->a = tmp0; l->b = tmp3; l->c = tmp6;
}
static void addition_step(line_t *l, g2_jac *r, const g2_aff *q) {
    fp2 zsq, ysq, t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, zz;
    fp2_sqr(&zsq, &r->z);
    fp2_sqr(&ysq, &q->y);
    fp2_mul(&t0, &zsq, &q->x);
    fp2_add(&t1, &q->y, &r->z); fp2_sqr(&t1, &t1); fp2_sub(&t1, &t1, &ysq); fp2_sub(&t1, &t1, &zsq); fp2_mul(&t1, &t1, &zsq);
    fp2_sub(&t2, &t0, &r->x);
    fp2_sqr(&t3, &t2);
    fp2_dbl(&t4, &t3); fp2_dbl(&t4, &t4);
    fp2_mul(&t5, &t4, &t2);
    fp2_sub(&t6, &t1, &r->y); fp2_sub(&t6, &t6, &r->y);
    fp2_mul(&t9, &t6, &q->x);
    fp2_mul(&t7, &t4, &r->x);
    fp2_sqr(&r->x, &t6); fp2_sub(&r->x, &r->x, &t5); fp2_sub(&r->x, &r->x, &t7); fp2_sub(&r->x, &r->x, &t7);
    fp2_add(&r->z, &r->z, &t2); fp2_sqr(&r->z, &r->z); fp2_sub(&r->z, &r->z, &zsq); fp2_sub(&r->z, &r->z, &t3);
    fp2_add(&t10, &q->y, &r->z);
    fp2_sub(&t8, &t7, &r->x); fp2_mul(&t8, &t8, &t6);
    fp2_mul(&t0, &r->y, &t5); fp2_dbl(&t0, &t0);
    fp2_sub(&r->y, &t8, &t0);
    fp2_sqr(&t10, &t10); fp2_sub(&t10, &t10, &ysq);
    fp2_sqr(&zz, &r->z); fp2_sub(&t10, &t10, &zz);
    fp2_dbl(&t9, &t9); fp2_sub(&t9, &t9, &t10);
    fp2_dbl(&t10, &r->z);
    fp2_neg(&t6, &t6);
    fp2_dbl(&t1, &t6);
    l->a = t10; l->b = t1; l->c = t9;
}
static void ell(fp12 *f, const line_t *l, const g1_aff *p) {
    fp2 c0, c1;
    fp2_mul_fp(&c0, &l->a, &p->y);
    fp2_mul_fp(&c1, &l->b, &p->x);
    fp12_mul_by_014(f, &l->c, &c1, &c0);
}
/* multi-pair Miller loop; pairs with an infinity operand are skipped (A8) */
static void miller_loop(fp12 *out, const g1_aff *ps, const g2_aff *qs, int n) {
    fp12 f = FP12_ONE;
    g2_jac r[2];
    int act[2];
    line_t l;
    if (n > 2) n = 2;
    for (int k = 0; k < n; k++) {
        act[k] = !(ps[k].inf || qs[k].inf);
        if (act[k]) g2_from_aff(&r[k], &qs[k]);
    }
    const u64 xs = BLS_X >> 1;
    int found = 0;
    for (int i = 63; i >= 0; i--) {
        int bit = (xs >> i) & 1;
        if (!found) { found = bit; continue; }
        for (int k = 0; k < n; k++) if (act[k]) { doubling_step(&l, &r[k]); ell(&f, &l, &ps[k]); }
        if (bit) for (int k = 0; k < n; k++) if (act[k]) { addition_step(&l, &r[k], &qs[k]); ell(&f, &l, &ps[k]); }
        fp12_sqr(&f, &f);
    }
    for (int k = 0; k < n; k++) if (act[k]) { doubling_step(&l, &r[k]); ell(&f, &l, &ps[k]); }
    fp12_conj(out, &f);   /* x < 0 */
}
static void exp_by_x(fp12 *f, u64 x) { fp12_pow_u64(f, f, x); fp12_conj(f, f); }
static void final_exponentiation(fp12 *out, const fp12 *in) {
    fp12 f1, f2, r, y0, y1, y2, y3;
    fp12_conj(&f1, in);
    fp12_inv(&f2, in);
    fp12_mul(&r, &f1, &f2);
    f2 = r;
    fp12_frob(&r, &r, 2);
    fp12_mul(&r, &r, &f2);
    u64 x = BLS_X;
    fp12_sqr(&y0, &r);
    y1 = y0; exp_by_x(&y1, x);
    x >>= 1;
    y2 = y1; exp_by_x(&y2, x);
    x <<= 1;
    fp12_conj(&y3, &r);
    fp12_mul(&y1, &y1, &y3);
    fp12_conj(&y1, &y1);
    fp12_mul(&y1, &y1, &y2);
    y2 = y1; exp_by_x(&y2, x);
    y3 = y2; exp_by_x(&y3, x);
    fp12_conj(&y1, &y1);
    fp12_mul(&y3, &y3, &y1);
    fp12_conj(&y1, &y1);
    fp12_frob(&y1, &y1, 3);
    fp12_frob(&y2, &y2, 2);
    fp12_mul(&y1, &y1, &y2);
    y2 = y3; exp_by_x(&y2, x);
    fp12_mul(&y2, &y2, &y0);
    fp12_mul(&y2, &y2, &r);
    fp12_mul(&y1, &y1, &y2);
    fp12_frob(&y2, &y3, 1);
    fp12_mul(out, &y1, &y2);
}
static void pairing(fp12 *out, const g1_aff *p, const g2_aff *q) {
    fp12 f;
    miller_loop(&f, p, q, 1);
    final_exponentiation(out, &f);
}

/* ------------------------------------------------------------------ SHA3-256 (FIPS 202), ChaCha20, BlockRng word stream */
static const u64 KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
static void keccak_f(u64 s[25]) {
    for (int rnd = 0; rnd < 24; rnd++) {
        u64 c[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) {
            u64 d = c[(x + 4) % 5] ^ ((c[(x + 1) % 5] << 1) | (c[(x + 1) % 5] >> 63));
            for (int y = 0; y < 25; y += 5) s[y + x] ^= d;
        }
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) {
                int rot = KECCAK_ROT[x + 5 * y];
                u64 v = s[x + 5 * y];
                b[y + 5 * ((2 * x + 3 * y) % 5)] = rot ? ((v << rot) | (v >> (64 - rot))) : v;
            }
        for (int y = 0; y < 25; y += 5)
            for (int x = 0; x < 5; x++) s[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
        s[0] ^= KECCAK_RC[rnd];
    }
}
void orc_sha3_256(const u8 *msg, size_t len, u8 out[32]) {
    u64 s[25] = {0};
    const size_t rate = 136;
    u8 blk[136];
    while (len >= rate) {
        for (size_t i = 0; i < rate / 8; i++) { u64 v; memcpy(&v, msg + 8 * i, 8); s[i] ^= v; }
        keccak_f(s);
        msg += rate; len -= rate;
    }
    memset(blk, 0, rate);
    memcpy(blk, msg, len);
    blk[len] ^= 0x06;
    blk[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; i++) { u64 v; memcpy(&v, blk + 8 * i, 8); s[i] ^= v; }
    keccak_f(s);
    memcpy(out, s, 32);
}
typedef struct { u32 key[8]; u64 ctr; u32 buf[16]; int idx; } chacha_rng;
#define ROTL32(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define QR(a, b, c, d) a += b; d ^= a; d = ROTL32(d, 16); c += d; b ^= c; b = ROTL32(b, 12); a += b; d ^= a; d = ROTL32(d, 8); c += d; b ^= c; b = ROTL32(b, 7);
static void chacha_block(chacha_rng *g) {
    u32 s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; i++) s[4 + i] = g->key[i];
    s[12] = (u32)g->ctr; s[13] = (u32)(g->ctr >> 32); s[14] = 0; s[15] = 0;
    u32 w[16];
    memcpy(w, s, sizeof w);
    for (int i = 0; i < 10; i++) {
        QR(w[0], w[4], w[8], w[12]) QR(w[1], w[5], w[9], w[13]) QR(w[2], w[6], w[10], w[14]) QR(w[3], w[7], w[11], w[15])
        QR(w[0], w[5], w[10], w[15]) QR(w[1], w[6], w[11], w[12]) QR(w[2], w[7], w[8], w[13]) QR(w[3], w[4], w[9], w[14])
    }
    for (int i = 0; i < 16; i++) g->buf[i] = w[i] + s[i];
    g->ctr++;
    g->idx = 0;
}
static void rng_seed(chacha_rng *g, const u8 seed[32]) { memcpy(g->key, seed, 32); g->ctr = 0; g->idx = 16; }
static u32 rng_u32(chacha_rng *g) { if (g->idx >= 16) chacha_block(g); return g->buf[g->idx++]; }
static u64 rng_u64(chacha_rng *g) { u64 lo = rng_u32(g); u64 hi = rng_u32(g); return lo | (hi << 32); }

/* ff_derive 0.6 Field::random (A1): raw limbs ARE the Montgomery representation */
static void fp_random(fp *r, chacha_rng *g) {
    for (;;) {
        for (int i = 0; i < 6; i++) r->l[i] = rng_u64(g);
        r->l[5] &= 0xffffffffffffffffULL >> 3;
        if (!mp_geq(r->l, FP_P.l, 6)) return;
    }
}
static void fr_random(fr *r, chacha_rng *g) {
    for (;;) {
        for (int i = 0; i < 4; i++) r->l[i] = rng_u64(g);
        r->l[3] &= 0xffffffffffffffffULL >> 1;
        if (!mp_geq(r->l, FR_R.l, 4)) return;
    }
}
/* EXTERNAL pairing G2::random (A2, A3) */
static void g2_random(g2_jac *out, chacha_rng *g) {
    for (;;) {
        fp2 x, y, ny, t;
        fp_random(&x.c0, g);
        fp_random(&x.c1, g);
        int greatest = (rng_u32(g) % 2) != 0;
        fp2_sqr(&t, &x); fp2_mul(&t, &t, &x); fp2_add(&t, &t, &G2_B);
        if (!fp2_sqrt(&y, &t)) continue;
        fp2_neg(&ny, &y);
        g2_aff a;
        a.x = x; a.inf = 0;
        a.y = ((fp2_cmp(&y, &ny) < 0) ^ greatest) ? y : ny;
        g2_mul_aff(out, &a, G2_COFACTOR, 8);
        if (!g2_is_inf(out)) return;
    }
}
static void hash_g2(g2_jac *out, const u8 *msg, size_t len) {   /* src/lib.rs:691-694 */
    u8 d[32];
    chacha_rng g;
    orc_sha3_256(msg, len, d);
    rng_seed(&g, d);
    g2_random(out, &g);
}
static void hash_g1_g2(g2_jac *out, const g1_aff *g1, const u8 *msg, size_t len) {   /* src/lib.rs:697-707 */
    u8 buf[64 + 48];
    size_t n;
    if (len > 64) { orc_sha3_256(msg, len, buf); n = 32; }
    else { memcpy(buf, msg, len); n = len; }
    g1_aff_compress(buf + n, g1);
    hash_g2(out, buf, n + 48);
}
static void xor_with_hash(u8 *out, const g1_aff *g1, const u8 *in, size_t len) {   /* src/lib.rs:710-715 */
    u8 c[48], d[32];
    chacha_rng g;
    g1_aff_compress(c, g1);
    orc_sha3_256(c, 48, d);
    rng_seed(&g, d);
    for (size_t i = 0; i < len; i++) out[i] = (u8)rng_u32(&g) ^ in[i];
}

/* ------------------------------------------------------------------ init */
static void limbs_sub_small(u64 *r, const u64 *a, u64 k, int n) {
    u64 bw = k;
    for (int i = 0; i < n; i++) { u64 t = a[i] - bw; bw = a[i] < bw; r[i] = t; }
}
static void limbs_div_small(u64 *r, const u64 *a, u64 d, int n) {
    u128 rem = 0;
    for (int i = n - 1; i >= 0; i--) { u128 cur = (rem << 64) | a[i]; r[i] = (u64)(cur / d); rem = cur % d; }
}
static u64 neg_inv64(u64 m0) {
    u64 x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - m0 * x;   /* Newton: x = m0^-1 mod 2^64 */
    return (u64)0 - x;
}
static void pow2_mod(u64 *r, int bits, const u64 *m, int n) {
    u64 one[8] = {1};
    memcpy(r, one, n * 8);
    for (int i = 0; i < bits; i++) mp_add_mod(r, r, r, m, n);
}
static void fp_set_hex_be(fp *r, const char *hex) {   /* 96 hex digits canonical -> Montgomery */
    u8 b[48];
    for (int i = 0; i < 48; i++) {
        unsigned v = 0;
        for (int k = 0; k < 2; k++) {
            char c = hex[2 * i + k];
            v = v * 16 + (c <= '9' ? c - '0' : (c | 32) - 'a' + 10);
        }
        b[i] = (u8)v;
    }
    fp_from_be(r, b);
}
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void orc_init_impl(void) {
    FP_INV = neg_inv64(FP_P.l[0]);
    pow2_mod(FP_R1.l, 384, FP_P.l, 6);
    pow2_mod(FP_R2.l, 768, FP_P.l, 6);
    FR_INV = neg_inv64(FR_R.l[0]);
    pow2_mod(FR_R1.l, 256, FR_R.l, 4);
    pow2_mod(FR_R2.l, 512, FR_R.l, 4);
    memset(&FP_ZERO, 0, sizeof FP_ZERO);
    memset(&FP2_ZERO, 0, sizeof FP2_ZERO);
    FP2_ONE = FP2_ZERO; FP2_ONE.c0 = FP_R1;
    memset(&FP12_ONE, 0, sizeof FP12_ONE); FP12_ONE.c0.c0 = FP2_ONE;
    u64 t[6];
    limbs_sub_small(FP_EXP_PM2, FP_P.l, 2, 6);
    limbs_sub_small(t, FP_P.l, 3, 6); limbs_div_small(FP_EXP_PM3D4, t, 4, 6);
    limbs_sub_small(t, FP_P.l, 1, 6); limbs_div_small(FP_EXP_PM1D2, t, 2, 6);
    limbs_div_small(FP_EXP_PM1D6, t, 6, 6);
    /* (p+1)/4 = (p-3)/4 + 1 */
    memcpy(FP_EXP_SQRT, FP_EXP_PM3D4, sizeof t); FP_EXP_SQRT[0] += 1;
    limbs_sub_small(FR_EXP_RM2, FR_R.l, 2, 4);
    /* curve constants */
    fp four = {{4, 0, 0, 0, 0, 0}};
    fp_to_mont(&G1_B, &four);
    G2_B.c0 = G1_B; G2_B.c1 = G1_B;
    fp_set_hex_be(&G1_GEN.x, "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb");
    fp_set_hex_be(&G1_GEN.y, "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1");
    G1_GEN.inf = 0;
    fp_set_hex_be(&G2_GEN.x.c0, "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8");
    fp_set_hex_be(&G2_GEN.x.c1, "13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e");
    fp_set_hex_be(&G2_GEN.y.c0, "0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801");
    fp_set_hex_be(&G2_GEN.y.c1, "0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be");
    G2_GEN.inf = 0;
    /* Frobenius constants: g1 = xi^((p-1)/6); g2 = g1 * conj(g1); g3 = g1^2 * conj(g1); FROB[k][m] = gk^m */
    fp2 xi, g[4], cj;
    xi.c0 = FP_R1; xi.c1 = FP_R1;
    fp2_pow(&g[1], &xi, FP_EXP_PM1D6, 6);
    fp2_conj(&cj, &g[1]);
    fp2_mul(&g[2], &g[1], &cj);
    fp2_mul(&g[3], &g[2], &g[1]);
    for (int k = 1; k <= 3; k++) {
        FROB[k][0] = FP2_ONE;
        for (int m = 1; m < 6; m++) fp2_mul(&FROB[k][m], &FROB[k][m - 1], &g[k]);
    }
}
void orc_init(void) { pthread_once(&g_once, orc_init_impl); }

/* ------------------------------------------------------------------ parallel-for over items */
typedef void (*item_fn)(size_t i, void *arg);
typedef struct { item_fn fn; void *arg; size_t lo, hi; } job_t;
static void *job_main(void *p) {
    job_t *j = (job_t *)p;
    for (size_t i = j->lo; i < j->hi; i++) j->fn(i, j->arg);
    return NULL;
}
static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
static void par_for(size_t n, item_fn fn, void *arg) {
    int nt = g_threads;
    if ((size_t)nt > n) nt = n ? (int)n : 1;
    if (nt <= 1) { for (size_t i = 0; i < n; i++) fn(i, arg); return; }
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < nt; t++) {
        jobs[t].fn = fn; jobs[t].arg = arg;
        jobs[t].lo = n * t / nt; jobs[t].hi = n * (t + 1) / nt;
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------ batched API (same byte formats as include/tcb200.h) */
/* e(a,b) == e(c,d) computed as the reference does: two full pairings compared in Fq12 (src/lib.rs:108-110) */
static int pairing_eq(const g1_aff *a, const g2_aff *b, const g1_aff *c, const g2_aff *d) {
    fp12 l, r;
    pairing(&l, a, b);
    pairing(&r, c, d);
    return fp12_eq(&l, &r);
}
typedef struct { const u8 *a, *b, *c, *d; u8 *ok; } vg2_args;
static void vg2_item(size_t i, void *p) {
    vg2_args *g = (vg2_args *)p;
    g1_aff a, c; g2_aff b, d;
    int good = g1_aff_from_unc(&a, g->a + 96 * i) & g2_aff_from_unc(&b, g->b + 192 * i) & g2_aff_from_unc(&d, g->d + 192 * i);
    if (g->c) good &= g1_aff_from_unc(&c, g->c + 96 * i); else c = G1_GEN;
    g->ok[i] = good ? (u8)pairing_eq(&a, &b, &c, &d) : 0;
}
int orc_verify_g2_batch(size_t n, const u8 *a_g1, const u8 *b_g2, const u8 *c_g1, const u8 *d_g2, u8 *ok) {
    orc_init();
    vg2_args g = {a_g1, b_g2, c_g1, d_g2, ok};
    par_for(n, vg2_item, &g);
    return 0;
}
typedef struct { const u8 *msgs; const u64 *off; u8 *out; } hg2_args;
static void hg2_item(size_t i, void *p) {
    hg2_args *g = (hg2_args *)p;
    g2_jac h; g2_aff ha;
    hash_g2(&h, g->msgs + g->off[i], (size_t)(g->off[i + 1] - g->off[i]));
    g2_to_aff(&ha, &h);
    g2_aff_to_unc(g->out + 192 * i, &ha);
}
int orc_hash_g2_batch(size_t n, const u8 *msgs, const u64 *off, u8 *out_g2) {
    orc_init();
    hg2_args g = {msgs, off, out_g2};
    par_for(n, hg2_item, &g);
    return 0;
}
typedef struct { const u8 *pk, *sig, *msgs; const u64 *off; u8 *ok; } v_args;
static void v_item(size_t i, void *p) {   /* PublicKey::verify, src/lib.rs:115-117 */
    v_args *g = (v_args *)p;
    g1_aff pk; g2_aff sig, ha; g2_jac h;
    int good = g1_aff_from_unc(&pk, g->pk + 96 * i) & g2_aff_from_unc(&sig, g->sig + 192 * i);
    hash_g2(&h, g->msgs + g->off[i], (size_t)(g->off[i + 1] - g->off[i]));
    g2_to_aff(&ha, &h);
    g->ok[i] = good ? (u8)pairing_eq(&pk, &ha, &G1_GEN, &sig) : 0;
}
int orc_verify_batch(size_t n, const u8 *pk, const u8 *sig, const u8 *msgs, const u64 *off, u8 *ok) {
    orc_init();
    v_args g = {pk, sig, msgs, off, ok};
    par_for(n, v_item, &g);
    return 0;
}
typedef struct { const u8 *sk, *msgs; const u64 *off; const u8 *h; u8 *out; } s_args;
static void s_item(size_t i, void *p) {   /* SecretKey::sign, src/lib.rs:372-381 */
    s_args *g = (s_args *)p;
    g2_jac h, s; g2_aff ha, sa;
    u64 k[4];
    if (g->h) g2_aff_from_unc(&ha, g->h + 192 * i);
    else { hash_g2(&h, g->msgs + g->off[i], (size_t)(g->off[i + 1] - g->off[i])); g2_to_aff(&ha, &h); }
    fr_canon_le(k, g->sk + 32 * i);
    g2_mul_aff(&s, &ha, k, 4);
    g2_to_aff(&sa, &s);
    g2_aff_to_unc(g->out + 192 * i, &sa);
}
int orc_sign_batch(size_t n, const u8 *sk, const u8 *msgs, const u64 *off, u8 *out_g2) {
    orc_init();
    s_args g = {sk, msgs, off, NULL, out_g2};
    par_for(n, s_item, &g);
    return 0;
}
int orc_sign_g2_batch(size_t n, const u8 *sk, const u8 *h_g2, u8 *out_g2) {
    orc_init();
    s_args g = {sk, NULL, NULL, h_g2, out_g2};
    par_for(n, s_item, &g);
    return 0;
}

/* Lagrange coefficients at 0 exactly as src/lib.rs:739-765 (prefix/suffix products; by-value filter) */
static int lagrange_coeffs(fr *l0, const fr *xs, size_t m) {
    size_t t = m - 1;
    fr tmp = FR_R1;
    l0[0] = tmp;
    for (size_t i = 0; i < t; i++) { fr_mul(&tmp, &tmp, &xs[i]); l0[i + 1] = tmp; }
    tmp = FR_R1;
    for (size_t i = t; i-- > 0;) { fr_mul(&tmp, &tmp, &xs[i + 1]); fr_mul(&l0[i], &l0[i], &tmp); }
    for (size_t i = 0; i < m; i++) {
        fr denom = FR_R1, diff, inv;
        for (size_t j = 0; j < m; j++) {
            if (fr_eq(&xs[j], &xs[i])) continue;
            fr_sub(&diff, &xs[j], &xs[i]);
            fr_mul(&denom, &denom, &diff);
        }
        if (!fr_inv(&inv, &denom)) return 0;    /* Error::DuplicateEntry */
        fr_mul(&l0[i], &l0[i], &inv);
    }
    return 1;
}
typedef struct { size_t t; const u8 *x, *shares, *v; const u64 *voff; u8 *out, *status; int mode; } c_args;
/* mode 0: G2 combine; 1: G1 combine; 2: G1 combine + xor_with_hash (decrypt) */
static void c_item(size_t i, void *p) {
    c_args *g = (c_args *)p;
    size_t m = g->t + 1;
    fr *xs = (fr *)malloc(sizeof(fr) * m * 2), *l0 = xs + m;
    int bad = 0;
    for (size_t k = 0; k < m; k++) if (!fr_from_le(&xs[k], g->x + 32 * (i * m + k))) bad = 1;
    g->status[i] = 0;
    if (bad) { g->status[i] = 3; free(xs); return; }
    if (g->mode == 0) {
        g2_jac acc, term; g2_aff s, ra;
        g2_set_inf(&acc);
        if (g->t == 0) {   /* src/lib.rs:735-737 */
            if (!g2_aff_from_unc(&s, g->shares + 192 * i)) { g->status[i] = 3; free(xs); return; }
            g2_from_aff(&acc, &s);
        } else {
            if (!lagrange_coeffs(l0, xs, m)) { g->status[i] = 2; free(xs); return; }
            for (size_t k = 0; k < m; k++) {
                fr c; fr_from_mont(&c, &l0[k]);
                if (!g2_aff_from_unc(&s, g->shares + 192 * (i * m + k))) { g->status[i] = 3; free(xs); return; }
                g2_mul_aff(&term, &s, c.l, 4);
                g2_add(&acc, &acc, &term);
            }
        }
        g2_to_aff(&ra, &acc);
        g2_aff_to_unc(g->out + 192 * i, &ra);
    } else {
        g1_jac acc, term; g1_aff s, ra;
        g1_set_inf(&acc);
        if (g->t == 0) {
            if (!g1_aff_from_unc(&s, g->shares + 96 * i)) { g->status[i] = 3; free(xs); return; }
            g1_from_aff(&acc, &s);
        } else {
            if (!lagrange_coeffs(l0, xs, m)) { g->status[i] = 2; free(xs); return; }
            for (size_t k = 0; k < m; k++) {
                fr c; fr_from_mont(&c, &l0[k]);
                if (!g1_aff_from_unc(&s, g->shares + 96 * (i * m + k))) { g->status[i] = 3; free(xs); return; }
                g1_mul_aff(&term, &s, c.l, 4);
                g1_add(&acc, &acc, &term);
            }
        }
        g1_to_aff(&ra, &acc);
        if (g->mode == 1) g1_aff_to_unc(g->out + 96 * i, &ra);
        else xor_with_hash(g->out + g->voff[i], &ra, g->v + g->voff[i], (size_t)(g->voff[i + 1] - g->voff[i]));
    }
    free(xs);
}
int orc_combine_g2_batch(size_t n, size_t t, const u8 *x_fr, const u8 *shares, u8 *out_g2, u8 *status) {
    orc_init();
    c_args g = {t, x_fr, shares, NULL, NULL, out_g2, status, 0};
    par_for(n, c_item, &g);
    return 0;
}
int orc_combine_g1_batch(size_t n, size_t t, const u8 *x_fr, const u8 *shares, u8 *out_g1, u8 *status) {
    orc_init();
    c_args g = {t, x_fr, shares, NULL, NULL, out_g1, status, 1};
    par_for(n, c_item, &g);
    return 0;
}
int orc_decrypt_batch(size_t n, size_t t, const u8 *x_fr, const u8 *shares_g1, const u8 *v, const u64 *v_off, u8 *out, u8 *status) {
    orc_init();
    c_args g = {t, x_fr, shares_g1, v, v_off, out, status, 2};
    par_for(n, c_item, &g);
    return 0;
}
typedef struct { const u8 *sk, *pt; u8 *out; int fixed_base; } m1_args;
static void m1_item(size_t i, void *p) {   /* decrypt_share_no_verify src/lib.rs:460-462; public_key :367-369 */
    m1_args *g = (m1_args *)p;
    g1_aff a, ra; g1_jac r;
    u64 k[4];
    if (g->fixed_base) a = G1_GEN; else g1_aff_from_unc(&a, g->pt + 96 * i);
    fr_canon_le(k, g->sk + 32 * i);
    g1_mul_aff(&r, &a, k, 4);
    g1_to_aff(&ra, &r);
    g1_aff_to_unc(g->out + 96 * i, &ra);
}
int orc_decrypt_share_batch(size_t n, const u8 *sk, const u8 *u_g1, u8 *out_g1) {
    orc_init();
    m1_args g = {sk, u_g1, out_g1, 0};
    par_for(n, m1_item, &g);
    return 0;
}
int orc_g1_mul_gen_batch(size_t n, const u8 *sk, u8 *out_g1) {
    orc_init();
    m1_args g = {sk, NULL, out_g1, 1};
    par_for(n, m1_item, &g);
    return 0;
}
typedef struct { size_t deg; const g1_jac *coeff; const u8 *x; u8 *out; } ce_args;
static void ce_item(size_t i, void *p) {   /* Commitment::evaluate src/poly.rs:497-508 */
    ce_args *g = (ce_args *)p;
    u64 k[4];
    fr_canon_le(k, g->x + 32 * i);
    g1_jac acc = g->coeff[g->deg];
    for (size_t c = g->deg; c-- > 0;) {
        g1_mul_jac(&acc, &acc, k, 4);
        g1_add(&acc, &acc, &g->coeff[c]);
    }
    g1_aff ra;
    g1_to_aff(&ra, &acc);
    g1_aff_to_unc(g->out + 96 * i, &ra);
}
int orc_commitment_eval_batch(size_t deg, const u8 *coeff_g1, size_t n, const u8 *x_fr, u8 *out_g1) {
    orc_init();
    g1_jac *cj = (g1_jac *)malloc(sizeof(g1_jac) * (deg + 1));
    for (size_t c = 0; c <= deg; c++) { g1_aff a; g1_aff_from_unc(&a, coeff_g1 + 96 * c); g1_from_aff(&cj[c], &a); }
    ce_args g = {deg, cj, x_fr, out_g1};
    par_for(n, ce_item, &g);
    free(cj);
    return 0;
}

/* ------------------------------------------------------------------ helpers for tests / vector generation */
void orc_g1_generator(u8 out[96]) { orc_init(); g1_aff_to_unc(out, &G1_GEN); }
void orc_g2_generator(u8 out[192]) { orc_init(); g2_aff_to_unc(out, &G2_GEN); }
int orc_g1_compress(size_t n, const u8 *unc, u8 *out48) {
    orc_init();
    for (size_t i = 0; i < n; i++) { g1_aff a; if (!g1_aff_from_unc(&a, unc + 96 * i)) return -1; g1_aff_compress(out48 + 48 * i, &a); }
    return 0;
}
int orc_g2_compress(size_t n, const u8 *unc, u8 *out96) {
    orc_init();
    for (size_t i = 0; i < n; i++) { g2_aff a; if (!g2_aff_from_unc(&a, unc + 192 * i)) return -1; g2_aff_compress(out96 + 96 * i, &a); }
    return 0;
}
/* status: 0 ok, 3 invalid */
int orc_g1_decompress(size_t n, const u8 *in48, u8 *unc, u8 *status) {
    orc_init();
    for (size_t i = 0; i < n; i++) {
        g1_aff a;
        if (g1_aff_decompress(&a, in48 + 48 * i)) { g1_aff_to_unc(unc + 96 * i, &a); status[i] = 0; }
        else { memset(unc + 96 * i, 0, 96); status[i] = 3; }
    }
    return 0;
}
int orc_g2_decompress(size_t n, const u8 *in96, u8 *unc, u8 *status) {
    orc_init();
    for (size_t i = 0; i < n; i++) {
        g2_aff a;
        if (g2_aff_decompress(&a, in96 + 96 * i)) { g2_aff_to_unc(unc + 192 * i, &a); status[i] = 0; }
        else { memset(unc + 192 * i, 0, 192); status[i] = 3; }
    }
    return 0;
}
/* GT value of one pairing as 12 x 48-byte big-endian canonical Fp, in the order
 * c0.c0.c0, c0.c0.c1, c0.c1.c0, ... (Fp12 -> Fp6 -> Fp2 -> Fp) — used only to cross-check pyref */
void orc_pairing_gt(const u8 *p_g1, const u8 *q_g2, u8 *out576) {
    orc_init();
    g1_aff p; g2_aff q; fp12 e;
    g1_aff_from_unc(&p, p_g1); g2_aff_from_unc(&q, q_g2);
    pairing(&e, &p, &q);
    const fp *c = (const fp *)&e;
    for (int i = 0; i < 12; i++) fp_to_be(out576 + 48 * i, &c[i]);
}
void orc_xor_with_hash(const u8 *g1_unc, const u8 *in, size_t len, u8 *out) {
    orc_init();
    g1_aff a; g1_aff_from_unc(&a, g1_unc);
    xor_with_hash(out, &a, in, len);
}
void orc_hash_g1_g2(const u8 *g1_unc, const u8 *msg, size_t len, u8 *out_g2) {
    orc_init();
    g1_aff a; g2_jac h; g2_aff ha;
    g1_aff_from_unc(&a, g1_unc);
    hash_g1_g2(&h, &a, msg, len);
    g2_to_aff(&ha, &h);
    g2_aff_to_unc(out_g2, &ha);
}
/* Fr::random stream (rule A1) from a 32-byte ChaCha seed -> n canonical LE scalars (synthetic data) */
void orc_fr_random_stream(const u8 seed[32], size_t n, u8 *out) {
    orc_init();
    chacha_rng g; rng_seed(&g, seed);
    for (size_t i = 0; i < n; i++) {
        fr m, c; fr_random(&m, &g); fr_from_mont(&c, &m);
        memcpy(out + 32 * i, c.l, 32);
    }
}
/* Poly::evaluate (src/poly.rs:358-369) on canonical LE scalars: out = sum coeff[k] x^k */
void orc_poly_eval(size_t ncoeff, const u8 *coeff, size_t n, const u8 *x, u8 *out) {
    orc_init();
    fr *c = (fr *)malloc(sizeof(fr) * (ncoeff ? ncoeff : 1));
    for (size_t k = 0; k < ncoeff; k++) fr_from_le(&c[k], coeff + 32 * k);
    for (size_t i = 0; i < n; i++) {
        fr xv, acc, o;
        fr_from_le(&xv, x + 32 * i);
        memset(&acc, 0, sizeof acc);
        for (size_t k = ncoeff; k-- > 0;) { fr_mul(&acc, &acc, &xv); fr_add(&acc, &acc, &c[k]); }
        fr_from_mont(&o, &acc);
        memcpy(out + 32 * i, o.l, 32);
    }
    free(c);
}
/* encrypt_with_rng (src/lib.rs:128-137) with r supplied (canonical LE): u = g1*r, v = xor(pk*r, msg), w = H(u,v)*r */
void orc_encrypt(const u8 *pk_g1, const u8 *r32, const u8 *msg, size_t len, u8 *u_out, u8 *v_out, u8 *w_out) {
    orc_init();
    g1_aff pk, ua, ga; g1_jac u, gj; g2_jac h, w; g2_aff ha, wa;
    u64 k[4];
    g1_aff_from_unc(&pk, pk_g1);
    fr_canon_le(k, r32);
    g1_mul_aff(&u, &G1_GEN, k, 4); g1_to_aff(&ua, &u);
    g1_mul_aff(&gj, &pk, k, 4); g1_to_aff(&ga, &gj);
    xor_with_hash(v_out, &ga, msg, len);
    hash_g1_g2(&h, &ua, v_out, len); g2_to_aff(&ha, &h);
    g2_mul_aff(&w, &ha, k, 4); g2_to_aff(&wa, &w);
    g1_aff_to_unc(u_out, &ua);
    g2_aff_to_unc(w_out, &wa);
}
