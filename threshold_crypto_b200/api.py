"""Host-side mirror of the reference API for the hot path (names, argument meaning and error
behaviour of /root/reference/src/lib.rs and src/poly.rs), backed by the C ABI of libtcb200.so.

The reference's toolchain (Rust) is absent from this image, so this mirror is Python; the Rust
binding a maintainer would write is in INTEGRATION.md.  Every arithmetic operation on curve
points goes to the GPU through `Engine` (no CPU fallback); only the cheap Fr polynomial algebra
(`Poly.evaluate`, index -> x = i + 1, `take(t+1)`, error mapping) stays on the host, exactly the
split SURVEY.md §8b describes.  Single-item methods are batches of one; the `*_batch` functions
are what a throughput-minded caller uses.
"""
import numpy as np

from ._lib import Engine

R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
PK_SIZE, SIG_SIZE = 48, 96          # compressed sizes (src/lib.rs:71,75); this mirror holds uncompressed bytes

_engine = None


def engine():
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine


def set_engine(e):
    global _engine
    _engine = e


class Error(Exception):
    """src/error.rs:7-20"""


class NotEnoughShares(Error):
    pass


class DuplicateEntry(Error):
    pass


def into_fr(i):
    """IntoFr (src/into_fr.rs:16-50): ints (negative ones negate), or an Fr given as int."""
    return int(i) % R


def into_fr_plus_1(i):   # src/lib.rs:769-773
    return (into_fr(i) + 1) % R


def _fr(v):
    return np.frombuffer((int(v) % R).to_bytes(32, "little"), np.uint8)


def _frs(vals):
    return np.frombuffer(b"".join((int(v) % R).to_bytes(32, "little") for v in vals), np.uint8).copy()


class FromBytesError(Error):
    """src/error.rs:36-44 (FromBytesError::Invalid)"""


class _Point:
    """Holds the UNCOMPRESSED affine encoding the C ABI works on.  `raw` must come from the engine or from `from_bytes`
    (the checked decoder): like the reference, where a PublicKey / Signature can only be built through a checked decode
    (src/lib.rs:140-146, 246-252), the arithmetic assumes a point of the r-order subgroup."""
    SIZE = 0          # uncompressed size
    COMPRESSED = 0    # wire size (PK_SIZE / SIG_SIZE)

    def __init__(self, raw):
        raw = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
        if raw.size != self.SIZE:
            raise ValueError(f"{type(self).__name__}: expected {self.SIZE} bytes of uncompressed encoding, got {raw.size}")
        self.raw = raw

    @classmethod
    def from_bytes(cls, data):
        """Checked decode of the compressed wire format (flags, x < p, on the curve, in the subgroup): src/lib.rs:140-146,246-252."""
        data = np.frombuffer(bytes(data), np.uint8)
        if data.size != cls.COMPRESSED:
            raise FromBytesError(f"{cls.__name__}: expected {cls.COMPRESSED} bytes")
        dec = engine().g1_decompress_batch if cls.COMPRESSED == PK_SIZE else engine().g2_decompress_batch
        out, st = dec(data)
        if st[0]:
            raise FromBytesError(f"{cls.__name__}: invalid encoding")
        return cls(out[0])

    def to_bytes(self):
        """Compressed wire format (src/lib.rs:149-153, 255-259)."""
        comp = engine().g1_compress_batch if self.COMPRESSED == PK_SIZE else engine().g2_compress_batch
        return comp(self.raw)[0].tobytes()

    def __eq__(self, o):
        return type(self) is type(o) and np.array_equal(self.raw, o.raw)

    def __hash__(self):
        return hash(self.raw.tobytes())

    def __repr__(self):
        return f"{type(self).__name__}({self.raw.tobytes().hex()[:10]}..)"


class Signature(_Point):      # Signature(G2), src/lib.rs:202
    SIZE, COMPRESSED = 192, SIG_SIZE


class SignatureShare(Signature):   # src/lib.rs:266
    pass


class DecryptionShare(_Point):     # DecryptionShare(G1), src/lib.rs:517
    SIZE, COMPRESSED = 96, PK_SIZE


class PublicKey(_Point):           # PublicKey(G1), src/lib.rs:79
    SIZE, COMPRESSED = 96, PK_SIZE

    def verify_g2(self, sig, hash_g2_point):          # src/lib.rs:108-110
        return bool(engine().verify_g2_batch(self.raw, np.asarray(hash_g2_point, np.uint8), None, sig.raw)[0])

    def verify(self, sig, msg):                        # src/lib.rs:115-117
        return bool(engine().verify_batch(self.raw, sig.raw, [bytes(msg)])[0])

    def encrypt_with_rng(self, rng, msg):              # src/lib.rs:128-137; Fr::random stays on the host
        r = int.from_bytes(rng.bytes(40), "little") % R
        u, v, w = engine().encrypt_batch(self.raw, _fr(r), [bytes(msg)])
        return Ciphertext(u[0], v[0], w[0])


class PublicKeyShare(PublicKey):   # src/lib.rs:160

    def verify_decryption_share(self, share, ct):      # src/lib.rs:182-186: e(share, H(U,V)) == e(pk_i, W)
        h = engine().hash_g1_g2_batch(ct.u, [ct.v])
        return bool(engine().verify_g2_batch(share.raw, h[0], self.raw, ct.w)[0])


class Ciphertext:                  # Ciphertext(G1, Vec<u8>, G2), src/lib.rs:474-478
    def __init__(self, u, v, w):
        self.u, self.v, self.w = np.asarray(u, np.uint8), bytes(v), np.asarray(w, np.uint8)

    def verify(self):                                  # src/lib.rs:508-512: e(g1, W) == e(U, H(U,V))
        h = engine().hash_g1_g2_batch(self.u, [self.v])
        # (a, b, c, d) = (U, H, g1, W): e(U, H) == e(g1, W)
        return bool(engine().verify_g2_batch(self.u, h[0], None, self.w)[0])


class SecretKey:                   # SecretKey(Fr), src/lib.rs:302
    def __init__(self, fr):
        self.fr = int(fr) % R

    def public_key(self):                              # src/lib.rs:367-369
        return PublicKey(engine().g1_mul_gen_batch(_fr(self.fr))[0])

    def sign_g2(self, hash_g2_point):                  # src/lib.rs:372-374
        return Signature(engine().sign_g2_batch(_fr(self.fr), np.asarray(hash_g2_point, np.uint8))[0])

    def sign(self, msg):                               # src/lib.rs:379-381
        return Signature(engine().sign_batch(_fr(self.fr), [bytes(msg)])[0])

    def decrypt(self, ct):                             # src/lib.rs:384-391: None if the ciphertext is invalid
        if not ct.verify():
            return None
        g = engine().decrypt_share_batch(_fr(self.fr), ct.u)
        out, _ = engine().decrypt_batch(1, 0, _fr(1), g, [ct.v])     # t = 0: xor_with_hash(g, v)
        return out[0]

    def __eq__(self, o):
        return isinstance(o, SecretKey) and self.fr == o.fr

    def __repr__(self):
        return "SecretKey(...)"                        # Debug redaction, src/lib.rs:335-339


class SecretKeyShare(SecretKey):   # src/lib.rs:412
    def public_key_share(self):
        return PublicKeyShare(self.public_key().raw)

    def sign(self, msg):
        return SignatureShare(super().sign(msg).raw)

    def sign_g2(self, h):
        return SignatureShare(super().sign_g2(h).raw)

    def decrypt_share_no_verify(self, ct):             # src/lib.rs:460-462
        return DecryptionShare(engine().decrypt_share_batch(_fr(self.fr), ct.u)[0])


class Poly:                        # poly::Poly, src/poly.rs:40-44 (Fr algebra stays on the host)
    def __init__(self, coeff):
        self.coeff = [int(c) % R for c in coeff]

    @staticmethod
    def random(degree, rng):
        return Poly([int.from_bytes(rng.bytes(40), "little") % R for _ in range(degree + 1)])

    def degree(self):
        return len(self.coeff) - 1

    def evaluate(self, i):                             # src/poly.rs:358-369
        x, acc = into_fr(i), 0
        for c in reversed(self.coeff):
            acc = (acc * x + c) % R
        return acc

    def evaluate_batch(self, idx):                     # Poly::evaluate at many points in one GPU call (Fr Horner per thread)
        if not self.coeff:
            return [0] * len(list(idx))
        out = engine().poly_eval_batch(_frs(self.coeff), _frs([into_fr(i) for i in idx]))
        return [int.from_bytes(bytes(o), "little") for o in out]

    def mul_gpu(self, o):                              # impl Mul for Poly (src/poly.rs:173-194) on the GPU
        if not self.coeff or not o.coeff:
            return Poly([])
        out = engine().poly_mul_batch(1, _frs(self.coeff), _frs(o.coeff))[0]
        return Poly([int.from_bytes(bytes(c), "little") for c in out])._trim()

    def commitment(self):                              # src/poly.rs:372-377: g1 * c_k on the GPU
        return Commitment(engine().g1_mul_gen_batch(_frs(self.coeff)))

    # Fr-side algebra (src/poly.rs:63-194): cheap 255-bit work, stays on the host
    def __add__(self, o):
        o = o if isinstance(o, Poly) else Poly([o])
        n = max(len(self.coeff), len(o.coeff))
        return Poly([(a + b) % R for a, b in zip(self.coeff + [0] * (n - len(self.coeff)), o.coeff + [0] * (n - len(o.coeff)))])._trim()

    def __sub__(self, o):
        o = o if isinstance(o, Poly) else Poly([o])
        return self + Poly([(-c) % R for c in o.coeff])

    def __mul__(self, o):
        if not isinstance(o, Poly):
            return Poly([c * int(o) % R for c in self.coeff])._trim()
        if not self.coeff or not o.coeff:
            return Poly([])
        out = [0] * (len(self.coeff) + len(o.coeff) - 1)
        for i, a in enumerate(self.coeff):
            for j, b in enumerate(o.coeff):
                out[i + j] = (out[i + j] + a * b) % R
        return Poly(out)._trim()

    def _trim(self):                                   # remove_zeros, src/poly.rs:380-384
        while self.coeff and self.coeff[-1] == 0:
            self.coeff.pop()
        return self

    def __eq__(self, o):
        return isinstance(o, Poly) and self.coeff == o.coeff

    @staticmethod
    def monomial(degree):
        return Poly([0] * degree + [1])

    @staticmethod
    def interpolate(samples):                          # src/poly.rs:388-417 (Lagrange through distinct points)
        pts = [(into_fr(x), into_fr(y)) for x, y in samples]
        if len({x for x, _ in pts}) != len(pts):
            raise ValueError("interpolate: sample points must be distinct")     # the reference panics (poly.rs:407)
        res = Poly([])
        for i, (xi, yi) in enumerate(pts):
            num, den = Poly([1]), 1
            for j, (xj, _) in enumerate(pts):
                if j != i:
                    num = num * Poly([(-xj) % R, 1])
                    den = den * (xi - xj) % R
            res = res + num * (yi * pow(den, R - 2, R) % R)
        return res


class Commitment:                  # poly::Commitment, src/poly.rs:429-433
    def __init__(self, coeff_g1):
        self.coeff = np.ascontiguousarray(coeff_g1, dtype=np.uint8).reshape(-1, 96)

    def degree(self):
        return self.coeff.shape[0] - 1

    def evaluate(self, i):                             # src/poly.rs:497-508
        return self.evaluate_batch([i])[0]

    def evaluate_batch(self, idx):
        idx = list(idx)
        if self.coeff.shape[0] == 0:       # the empty commitment evaluates to G1::zero() (src/poly.rs:497-508)
            inf = np.zeros(96, np.uint8); inf[0] = 0x40
            return np.tile(inf, (len(idx), 1))
        return engine().commitment_eval_batch(self.coeff, _frs([into_fr(i) for i in idx]))

    def __add__(self, o):            # src/poly.rs:436-470: coefficient-wise G1 addition
        n = max(self.coeff.shape[0], o.coeff.shape[0])
        inf = np.zeros(96, np.uint8); inf[0] = 0x40
        a = np.concatenate([self.coeff, np.tile(inf, (n - self.coeff.shape[0], 1))]) if self.coeff.shape[0] < n else self.coeff
        b = np.concatenate([o.coeff, np.tile(inf, (n - o.coeff.shape[0], 1))]) if o.coeff.shape[0] < n else o.coeff
        pts = np.stack([a, b], axis=1).reshape(-1, 96)
        return Commitment(engine().g1_lincomb_batch(n, 2, _frs([1] * (2 * n)), pts))

    def __eq__(self, o):
        return isinstance(o, Commitment) and np.array_equal(self.coeff, o.coeff)


def coeff_pos(i, j):                # src/poly.rs:744-748
    i, j = (i, j) if j >= i else (j, i)
    return i + j * (j + 1) // 2


def _powers(x, degree):             # src/poly.rs:729-738
    x = into_fr(x)
    out, cur = [], 1
    for _ in range(degree + 1):
        out.append(cur)
        cur = cur * x % R
    return out


class BivarPoly:                    # poly::BivarPoly, src/poly.rs:513-650 (symmetric, Fr only: host side)
    def __init__(self, degree, coeff):
        self.degree, self.coeff = degree, [int(c) % R for c in coeff]
        if len(self.coeff) != coeff_pos(degree, degree) + 1:
            raise ValueError("BivarPoly: wrong number of coefficients for the degree")

    @staticmethod
    def random(degree, rng):
        return BivarPoly(degree, [int.from_bytes(rng.bytes(40), "little") % R for _ in range(coeff_pos(degree, degree) + 1)])

    def evaluate(self, x, y):
        xp, yp = _powers(x, self.degree), _powers(y, self.degree)
        return sum(self.coeff[coeff_pos(i, j)] * xp[i] * yp[j] for i in range(self.degree + 1) for j in range(self.degree + 1)) % R

    def row(self, x):
        xp = _powers(x, self.degree)
        return Poly([sum(self.coeff[coeff_pos(i, j)] * xp[j] for j in range(self.degree + 1)) % R for i in range(self.degree + 1)])

    def commitment(self):            # src/poly.rs:627-633: g1 * c on the GPU
        return BivarCommitment(self.degree, engine().g1_mul_gen_batch(_frs(self.coeff)))


class BivarCommitment:              # poly::BivarCommitment, src/poly.rs:652-726 — G1 sweeps on the GPU
    def __init__(self, degree, coeff_g1):
        self.degree = degree
        self.coeff = np.ascontiguousarray(coeff_g1, dtype=np.uint8).reshape(-1, 96)
        if self.coeff.shape[0] != coeff_pos(degree, degree) + 1:        # serde validation, src/serde_impl.rs:150-161
            raise ValueError("BivarCommitment: wrong number of coefficients for the degree")

    def evaluate(self, x, y):        # src/poly.rs:693-709
        d = self.degree
        xp, yp = _powers(x, d), _powers(y, d)
        idx = [coeff_pos(i, j) for i in range(d + 1) for j in range(d + 1)]
        sc = _frs([xp[i] * yp[j] % R for i in range(d + 1) for j in range(d + 1)])
        return engine().g1_lincomb_batch(1, (d + 1) ** 2, sc, self.coeff[idx])[0]

    def row(self, x):                # src/poly.rs:712-726
        d = self.degree
        xp = _powers(x, d)
        idx = [coeff_pos(i, j) for i in range(d + 1) for j in range(d + 1)]
        sc = _frs([xp[j] for i in range(d + 1) for j in range(d + 1)])
        return Commitment(engine().g1_lincomb_batch(d + 1, d + 1, sc, self.coeff[idx]))


class SecretKeySet:                # src/lib.rs:630-688
    def __init__(self, poly):
        self.poly = poly

    @staticmethod
    def random(threshold, rng):
        return SecretKeySet(Poly.random(threshold, rng))

    def threshold(self):
        return self.poly.degree()

    def secret_key_share(self, i):                     # src/lib.rs:669-673
        return SecretKeyShare(self.poly.evaluate(into_fr_plus_1(i)))

    def public_keys(self):
        return PublicKeySet(self.poly.commitment())

    def secret_key(self):
        return SecretKey(self.poly.evaluate(0))


def _samples(t, shares):
    """`take(t + 1)`, NotEnoughShares (src/lib.rs:726-733); shares: mapping or iterable of (i, share)."""
    items = list(shares.items()) if hasattr(shares, "items") else list(shares)
    items = items[: t + 1]
    if len(items) <= t:
        raise NotEnoughShares()
    return items


def _raise(status):
    if status == 2:
        raise DuplicateEntry()
    if status:
        raise Error(f"invalid encoding (status {status})")


class PublicKeySet:                # src/lib.rs:539-626
    def __init__(self, commit):
        self.commit = commit

    def threshold(self):
        return self.commit.degree()

    def public_key(self):
        return PublicKey(self.commit.coeff[0])

    def public_key_share(self, i):                     # src/lib.rs:570-573
        return PublicKeyShare(self.commit.evaluate(into_fr_plus_1(i)))

    def combine_signatures(self, shares):              # src/lib.rs:608-615
        return combine_signatures_batch(self, [shares])[0]

    def decrypt(self, shares, ct):                     # src/lib.rs:618-626
        t = self.threshold()
        items = _samples(t, shares)
        xs = _frs([into_fr_plus_1(i) for i, _ in items])
        pts = np.stack([s.raw for _, s in items])
        out, st = engine().decrypt_batch(1, t, xs, pts, [ct.v])
        _raise(int(st[0]))
        return out[0]


# ---- batched entry points (the reason this engine exists)
def verify_batch(pks, sigs, msgs):
    """[pk.verify(sig, msg)] for many triples in one GPU call."""
    pk = np.stack([p.raw for p in pks])
    sg = np.stack([s.raw for s in sigs])
    return [bool(b) for b in engine().verify_batch(pk, sg, [bytes(m) for m in msgs])]


def sign_batch(sks, msgs):
    return [Signature(s) for s in engine().sign_batch(_frs([k.fr for k in sks]), [bytes(m) for m in msgs])]


def combine_signatures_batch(pk_set, batches):
    """PublicKeySet::combine_signatures for many messages signed under one key set.
    Raises the first per-item error like the reference would for that item."""
    t = pk_set.threshold()
    items = [_samples(t, b) for b in batches]
    xs = _frs([into_fr_plus_1(i) for it in items for i, _ in it])
    pts = np.stack([s.raw for it in items for _, s in it])
    out, st = engine().combine_g2_batch(len(items), t, xs, pts)
    for s in st:
        _raise(int(s))
    return [Signature(o) for o in out]
