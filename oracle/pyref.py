"""Independent big-int model of the threshold_crypto hot path (TEST INFRASTRUCTURE ONLY).

This file is a slow, textbook restatement (Python ints, affine coordinates, the pairing
computed over the untwisted curve E(Fp12) with a plain big-exponent final exponentiation)
of what the reference computes on its hot path.  It exists to cross-check the C oracle
(`oracle/tc_oracle.c`) and to generate the golden vectors under `tests/golden/`.
Nothing in the product path (`threshold_crypto_b200/`) may import it.

PARITY STATUS: **unpinned**.  The reference (`/root/reference`, crate threshold_crypto
0.4.0) holds no BLS12-381 known-answer vectors and its arithmetic lives in the absent
third-party crates pairing 0.16.0 / ff 0.6.0 / group 0.6.0 / rand 0.7.3 /
rand_chacha 0.2.2 / tiny-keccak 2.0.1 (Cargo.toml:22-33).  What is restated here is
their published algorithm; group-theoretic results (points, booleans, plaintexts) are
unique, sampling conventions (hash_g2, xor_with_hash) follow SURVEY.md §8c A1-A8.

Reference call sites followed (file:line relative to /root/reference):
  hash_g2          src/lib.rs:691-694      hash_g1_g2       src/lib.rs:697-707
  xor_with_hash    src/lib.rs:710-715      interpolate      src/lib.rs:719-767
  into_fr_plus_1   src/lib.rs:769-773      verify_g2        src/lib.rs:108-110
  sign_g2          src/lib.rs:372-374      decrypt_share    src/lib.rs:460-462
  Commitment::evaluate src/poly.rs:497-508 Poly::evaluate   src/poly.rs:358-369
  encrypt_with_rng src/lib.rs:128-137      Ciphertext::verify src/lib.rs:508-512
"""
import hashlib
import struct

# ----------------------------------------------------------------------------- constants
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
X_ABS = 0xd201000000010000  # the curve parameter is x = -X_ABS
H2 = 0x5d543a95414e7f1091d50792876a202cd91de4547085abaa68a205b2e5a7ddfa628f1cb4d9e82ef21537e293a6691ae1616ec6e786f0c70cf1c38e31c7238e5
H1 = 0x396c8c005555e1568c00aaab0000aaab
FP_MONT_R = 1 << 384
FR_MONT_R = 1 << 256

G1_GEN = (
    0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
    0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,
)
G2_GEN = (
    (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
     0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
    (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
     0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be),
)

# ----------------------------------------------------------------------------- Fp2 / Fp6 / Fp12
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a): return f2_mul(a, a)
def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], P - 2, P)
    return (a[0] * n % P, (-a[1]) * n % P)
def f2_conj(a): return (a[0], (-a[1]) % P)
F2_ZERO, F2_ONE = (0, 0), (1, 0)
XI = (1, 1)  # Fp6 = Fp2[v]/(v^3 - XI), Fp12 = Fp6[w]/(w^2 - v)

def f2_pow(a, e):
    r = F2_ONE
    for bit in bin(e)[2:]:
        r = f2_sqr(r)
        if bit == '1':
            r = f2_mul(r, a)
    return r

def f2_sqrt(a):
    """Any square root of a in Fp2 or None (brute-force-simple: p^2 = 9 mod 16 is awkward,
    so use the norm method: independent from the Algorithm-9 code in the C oracle)."""
    if a == F2_ZERO:
        return F2_ZERO
    a0, a1 = a
    if a1 == 0:
        s = pow(a0, (P + 1) // 4, P)
        if s * s % P == a0:
            return (s, 0)
        # a0 is a non-residue in Fp: sqrt is purely imaginary: (t*u)^2 = -t^2
        t = pow((-a0) % P, (P + 1) // 4, P)
        assert t * t % P == (-a0) % P
        return (0, t)
    n = (a0 * a0 + a1 * a1) % P
    s = pow(n, (P + 1) // 4, P)
    if s * s % P != n:
        return None
    inv2 = pow(2, P - 2, P)
    for sign in (1, -1):
        t = (a0 + sign * s) * inv2 % P
        x0 = pow(t, (P + 1) // 4, P)
        if x0 * x0 % P == t and x0 != 0:
            x1 = a1 * pow(2 * x0, P - 2, P) % P
            cand = (x0, x1)
            if f2_sqr(cand) == a:
                return cand
    return None

# Fp12 as a degree-12 extension is modelled directly as polynomials in w over Fp2 with
# w^6 = XI: element = list of 6 Fp2 coefficients (w^0..w^5).  This is deliberately a
# different tower shape from the 2-3-2 tower in the C oracle / CUDA code.
def f12_one(): return [F2_ONE] + [F2_ZERO] * 5
def f12_mul(a, b):
    t = [F2_ZERO] * 11
    for i in range(6):
        if a[i] == F2_ZERO:
            continue
        for j in range(6):
            t[i + j] = f2_add(t[i + j], f2_mul(a[i], b[j]))
    for k in range(10, 5, -1):
        t[k - 6] = f2_add(t[k - 6], f2_mul(t[k], XI))
    return t[:6]
def f12_sqr(a): return f12_mul(a, a)
def f12_pow(a, e):
    r = f12_one()
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == '1':
            r = f12_mul(r, a)
    return r
def f12_conj(a):  # the p^6 Frobenius: w -> -w
    return [a[i] if i % 2 == 0 else f2_neg(a[i]) for i in range(6)]
def f12_inv(a):
    # a^-1 = a^(p^12 - 2): slow but independent
    return f12_pow(a, P ** 12 - 2)
def f12_from_fp(c): return [(c % P, 0)] + [F2_ZERO] * 5

# ----------------------------------------------------------------------------- curves (affine, None = infinity)
class Curve:
    def __init__(self, add, sub, mul, sqr, inv, neg, zero, b, scal):
        self.add, self.sub, self.mul, self.sqr, self.inv, self.neg = add, sub, mul, sqr, inv, neg
        self.zero, self.b, self.scal = zero, b, scal

    def on_curve(self, pt):
        if pt is None:
            return True
        x, y = pt
        return self.sqr(y) == self.add(self.mul(self.sqr(x), x), self.b)

    def negate(self, pt):
        return None if pt is None else (pt[0], self.neg(pt[1]))

    def padd(self, a, b):
        if a is None: return b
        if b is None: return a
        if a[0] == b[0]:
            if a[1] == b[1]:
                if a[1] == self.zero:
                    return None
                lam = self.mul(self.scal(self.sqr(a[0]), 3), self.inv(self.scal(a[1], 2)))
            else:
                return None
        else:
            lam = self.mul(self.sub(b[1], a[1]), self.inv(self.sub(b[0], a[0])))
        x3 = self.sub(self.sub(self.sqr(lam), a[0]), b[0])
        y3 = self.sub(self.mul(lam, self.sub(a[0], x3)), a[1])
        return (x3, y3)

    def pmul(self, pt, k):
        if k < 0:
            return self.pmul(self.negate(pt), -k)
        acc = None
        for bit in bin(k)[2:] if k else '':
            acc = self.padd(acc, acc)
            if bit == '1':
                acc = self.padd(acc, pt)
        return acc

E1 = Curve(lambda a, b: (a + b) % P, lambda a, b: (a - b) % P, lambda a, b: a * b % P,
           lambda a: a * a % P, lambda a: pow(a, P - 2, P), lambda a: (-a) % P, 0, 4,
           lambda a, k: a * k % P)
E2 = Curve(f2_add, f2_sub, f2_mul, f2_sqr, f2_inv, f2_neg, F2_ZERO, (4, 4),
           lambda a, k: (a[0] * k % P, a[1] * k % P))
def _f12_scal(a, k): return [(c[0] * k % P, c[1] * k % P) for c in a]
E12 = Curve(lambda a, b: [f2_add(x, y) for x, y in zip(a, b)],
            lambda a, b: [f2_sub(x, y) for x, y in zip(a, b)],
            f12_mul, f12_sqr, f12_inv, lambda a: [f2_neg(x) for x in a],
            [F2_ZERO] * 6, f12_from_fp(4), _f12_scal)

# ----------------------------------------------------------------------------- pairing (textbook)
_W2_INV = None
_W3_INV = None
def _untwist(q):
    """E'(Fp2) -> E(Fp12): (x, y) -> (x / w^2, y / w^3)   (M-type twist, w^6 = XI)."""
    global _W2_INV, _W3_INV
    if _W2_INV is None:
        w = [F2_ZERO, F2_ONE] + [F2_ZERO] * 4
        w2 = f12_mul(w, w)
        w3 = f12_mul(w2, w)
        _W2_INV, _W3_INV = f12_inv(w2), f12_inv(w3)
    x = f12_mul([q[0]] + [F2_ZERO] * 5, _W2_INV)
    y = f12_mul([q[1]] + [F2_ZERO] * 5, _W3_INV)
    return (x, y)

def _line(t, q, p):
    """Value at p of the line through t and q (points of E(Fp12), affine), p in E(Fp)."""
    px, py = f12_from_fp(p[0]), f12_from_fp(p[1])
    if t[0] != q[0]:
        lam = f12_mul(E12.sub(q[1], t[1]), f12_inv(E12.sub(q[0], t[0])))
    elif t[1] == q[1]:
        lam = f12_mul(_f12_scal(f12_sqr(t[0]), 3), f12_inv(_f12_scal(t[1], 2)))
    else:
        return E12.sub(px, t[0])
    return E12.sub(E12.sub(py, t[1]), f12_mul(lam, E12.sub(px, t[0])))

def miller_loop(p, q):
    """f_{|x|,Q}(P), conjugated because x < 0 (optimal ate, as EXTERNAL pairing 0.16)."""
    if p is None or q is None:
        return f12_one()
    qq = _untwist(q)
    t = qq
    f = f12_one()
    for bit in bin(X_ABS)[3:]:
        f = f12_mul(f12_sqr(f), _line(t, t, p))
        t = E12.padd(t, t)
        if bit == '1':
            f = f12_mul(f, _line(t, qq, p))
            t = E12.padd(t, qq)
    return f12_conj(f)

FINAL_EXP = (P ** 12 - 1) // R
def final_exp(f): return f12_pow(f, FINAL_EXP)
def pairing(p, q): return final_exp(miller_loop(p, q))
def pairing_eq(a, b, c, d):
    """e(a,b) == e(c,d)  <=>  (ML(a,b) * ML(-c,d))^FINAL_EXP == 1   (src/lib.rs:108-110)."""
    f = f12_mul(miller_loop(a, b), miller_loop(E1.negate(c), d))
    return final_exp(f) == f12_one()

# ----------------------------------------------------------------------------- SHA3 / ChaCha20 / RNG conventions
def sha3_256(data): return hashlib.sha3_256(bytes(data)).digest()  # src/util.rs:3-9 (A6)

def _rotl(v, n): return ((v << n) & 0xffffffff) | (v >> (32 - n))
def chacha20_block(key_words, counter):
    """Standard ChaCha20 block, 64-bit block counter in words 12-13, nonce (stream) 0."""
    s = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574] + list(key_words) + \
        [counter & 0xffffffff, (counter >> 32) & 0xffffffff, 0, 0]
    w = list(s)
    def qr(a, b, c, d):
        w[a] = (w[a] + w[b]) & 0xffffffff; w[d] = _rotl(w[d] ^ w[a], 16)
        w[c] = (w[c] + w[d]) & 0xffffffff; w[b] = _rotl(w[b] ^ w[c], 12)
        w[a] = (w[a] + w[b]) & 0xffffffff; w[d] = _rotl(w[d] ^ w[a], 8)
        w[c] = (w[c] + w[d]) & 0xffffffff; w[b] = _rotl(w[b] ^ w[c], 7)
    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(w[i] + s[i]) & 0xffffffff for i in range(16)]

class ChaChaRng:
    """rand_chacha 0.2 ChaChaRng::from_seed + rand_core BlockRng word-stream semantics (A4)."""
    def __init__(self, seed32):
        self.key = struct.unpack('<8I', bytes(seed32))
        self.ctr = 0
        self.buf = []
    def next_u32(self):
        if not self.buf:
            self.buf = chacha20_block(self.key, self.ctr)
            self.ctr += 1
        return self.buf.pop(0)
    def next_u64(self):
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

def fq_random_mont(rng):
    """ff_derive 0.6 Field::random for Fq (A1): 6 LE u64 limbs, top limb >> 3, reject >= p.
    Returns the *raw limbs integer*, which the crate uses as the Montgomery representation."""
    while True:
        limbs = [rng.next_u64() for _ in range(6)]
        limbs[5] &= 0xffffffffffffffff >> 3
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < P:
            return v

def fr_random_mont(rng):
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= 0xffffffffffffffff >> 1
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < R:
            return v

_FP_RINV = pow(FP_MONT_R, P - 2, P)
_FR_RINV = pow(FR_MONT_R, R - 2, R)
def fq_random(rng): return fq_random_mont(rng) * _FP_RINV % P
def fr_random(rng): return fr_random_mont(rng) * _FR_RINV % R

def f2_lt(a, b):
    """Fq2 ordering of EXTERNAL pairing: by c1 then c0 on canonical integers (A3)."""
    return (a[1], a[0]) < (b[1], b[0])

def g2_random(rng):
    """EXTERNAL pairing 0.16 G2::random (A2, A3): x = Fq2::random (c0 then c1);
    greatest = next_u32() % 2 != 0; y = sqrt(x^3+b) choosing by (y < -y) ^ greatest;
    multiply by the exact cofactor h2; retry on no root / zero."""
    while True:
        c0 = fq_random(rng)
        c1 = fq_random(rng)
        x = (c0, c1)
        greatest = rng.next_u32() % 2 != 0
        y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), E2.b))
        if y is None:
            continue
        negy = f2_neg(y)
        ysel = y if (f2_lt(y, negy) ^ greatest) else negy
        pt = E2.pmul((x, ysel), H2)
        if pt is not None:
            return pt

def hash_g2(msg):  # src/lib.rs:691-694
    return g2_random(ChaChaRng(sha3_256(msg)))

def hash_g1_g2(g1, msg):  # src/lib.rs:697-707
    msg = bytes(msg)
    m = sha3_256(msg) if len(msg) > 64 else msg
    return hash_g2(m + g1_compress(g1))

def xor_with_hash(g1, data):  # src/lib.rs:710-715, one u32 keystream word per byte (A5)
    rng = ChaChaRng(sha3_256(g1_compress(g1)))
    return bytes((rng.next_u32() & 0xff) ^ b for b in bytes(data))

# ----------------------------------------------------------------------------- encodings (App. B)
def g1_uncompressed(pt):
    if pt is None:
        return bytes([0x40]) + bytes(95)
    return pt[0].to_bytes(48, 'big') + pt[1].to_bytes(48, 'big')

def g1_compress(pt):
    if pt is None:
        return bytes([0xc0]) + bytes(47)
    b = bytearray(pt[0].to_bytes(48, 'big'))
    b[0] |= 0x80
    if pt[1] > (-pt[1]) % P:
        b[0] |= 0x20
    return bytes(b)

def g2_uncompressed(pt):
    if pt is None:
        return bytes([0x40]) + bytes(191)
    (x0, x1), (y0, y1) = pt
    return b''.join(v.to_bytes(48, 'big') for v in (x1, x0, y1, y0))

def g2_compress(pt):
    if pt is None:
        return bytes([0xc0]) + bytes(95)
    (x0, x1), y = pt
    b = bytearray(x1.to_bytes(48, 'big') + x0.to_bytes(48, 'big'))
    b[0] |= 0x80
    if f2_lt(f2_neg(y), y):
        b[0] |= 0x20
    return bytes(b)

def g1_from_uncompressed(b):
    b = bytes(b)
    if b[0] & 0x40:
        return None
    return (int.from_bytes(b[:48], 'big'), int.from_bytes(b[48:], 'big'))

def g2_from_uncompressed(b):
    b = bytes(b)
    if b[0] & 0x40:
        return None
    v = [int.from_bytes(b[i * 48:(i + 1) * 48], 'big') for i in range(4)]
    return ((v[1], v[0]), (v[3], v[2]))

def g1_decompress(b):
    """Checked decode (on-curve and in-subgroup) or None."""
    b = bytes(b)
    if len(b) != 48 or not b[0] & 0x80:
        return 'invalid'
    if b[0] & 0x40:
        return None if (b[0] & 0x3f) == 0 and not any(b[1:]) and not b[0] & 0x20 else 'invalid'
    x = int.from_bytes(bytes([b[0] & 0x1f]) + b[1:], 'big')
    if x >= P:
        return 'invalid'
    y2 = (x * x * x + 4) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        return 'invalid'
    if (y > (-y) % P) != bool(b[0] & 0x20):
        y = (-y) % P
    pt = (x, y)
    return pt if E1.pmul(pt, R) is None else 'invalid'

def g2_decompress(b):
    b = bytes(b)
    if len(b) != 96 or not b[0] & 0x80:
        return 'invalid'
    if b[0] & 0x40:
        return None if (b[0] & 0x3f) == 0 and not any(b[1:]) and not b[0] & 0x20 else 'invalid'
    x1 = int.from_bytes(bytes([b[0] & 0x1f]) + b[1:48], 'big')
    x0 = int.from_bytes(b[48:], 'big')
    if x0 >= P or x1 >= P:
        return 'invalid'
    x = (x0, x1)
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), E2.b))
    if y is None:
        return 'invalid'
    if f2_lt(f2_neg(y), y) != bool(b[0] & 0x20):
        y = f2_neg(y)
    pt = (x, y)
    return pt if E2.pmul(pt, R) is None else 'invalid'

def fr_to_bytes(v): return (v % R).to_bytes(32, 'little')   # [u64;4] LE canonical (serde_impl.rs:109)
def fr_from_bytes(b): return int.from_bytes(bytes(b), 'little')

# ----------------------------------------------------------------------------- scheme (src/lib.rs)
def into_fr_plus_1(i): return (i + 1) % R            # src/lib.rs:769-773 (negative i handled by % R)

def lagrange_at_zero(xs):
    """Coefficients exactly as src/lib.rs:739-765 computes them, including the by-value
    filter `x0 != x` (so duplicate x never yields a zero denominator)."""
    n = len(xs)
    out = []
    for i in range(n):
        num = 1
        for j in range(n):
            if j != i:
                num = num * xs[j] % R
        den = 1
        for j in range(n):
            if xs[j] != xs[i]:
                den = den * ((xs[j] - xs[i]) % R) % R
        if den == 0:
            return None  # Error::DuplicateEntry (unreachable given the filter; kept for form)
        out.append(num * pow(den, R - 2, R) % R)
    return out

def interpolate(curve, t, samples):
    """samples: list of (index, point).  Returns point or the strings 'NotEnoughShares'/'DuplicateEntry'."""
    s = [(into_fr_plus_1(i), pt) for i, pt in samples[:t + 1]]
    if len(s) <= t:
        return 'NotEnoughShares'
    if t == 0:
        return s[0][1]
    lam = lagrange_at_zero([x for x, _ in s])
    if lam is None:
        return 'DuplicateEntry'
    acc = None
    for l, (_, pt) in zip(lam, s):
        acc = curve.padd(acc, curve.pmul(pt, l))
    return acc

def poly_eval(coeff, x):  # src/poly.rs:358-369
    acc = 0
    for c in reversed(coeff):
        acc = (acc * x + c) % R
    return acc

def commitment(coeff): return [E1.pmul(G1_GEN, c) for c in coeff]  # src/poly.rs:372-377

def commitment_eval(comm, x):  # src/poly.rs:497-508
    if not comm:
        return None
    acc = comm[-1]
    for c in reversed(comm[:-1]):
        acc = E1.padd(E1.pmul(acc, x), c)
    return acc

def public_key(sk): return E1.pmul(G1_GEN, sk)
def sign_g2(sk, h): return E2.pmul(h, sk)
def sign(sk, msg): return sign_g2(sk, hash_g2(msg))
def verify_g2(pk, sig, h): return pairing_eq(pk, h, G1_GEN, sig)
def verify(pk, sig, msg): return verify_g2(pk, sig, hash_g2(msg))

def encrypt_with_rng(pk, rng, msg):  # src/lib.rs:128-137
    r = fr_random(rng)
    u = E1.pmul(G1_GEN, r)
    v = xor_with_hash(E1.pmul(pk, r), msg)
    w = E2.pmul(hash_g1_g2(u, v), r)
    return (u, v, w)

def ciphertext_verify(ct):  # src/lib.rs:508-512
    u, v, w = ct
    return pairing_eq(G1_GEN, w, u, hash_g1_g2(u, v))

def decrypt_share(sk_i, ct): return E1.pmul(ct[0], sk_i)  # src/lib.rs:460-462
def threshold_decrypt(t, shares, ct):  # src/lib.rs:618-626
    g = interpolate(E1, t, shares)
    if isinstance(g, str):
        return g
    return xor_with_hash(g, ct[1])
