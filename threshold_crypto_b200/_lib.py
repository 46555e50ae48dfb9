"""ctypes binding of the C ABI declared in include/tcb200.h.

`Engine()` loads the CUDA library `csrc/libtcb200.so` (built by `__graft_entry__.build()` /
`csrc/build.sh`) and FAILS LOUDLY when it is missing or when no CUDA device is present —
there is no CPU fallback in this package.  Tests may point `Engine(path)` at the
host-emulation build under tests/hostemu (test infrastructure, same symbols) to check logic
on a GPU-less box; that library is never loaded by default.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TCB200_LIB selects an experiment build of the SAME CUDA library (e.g. another launch-bounds variant)
DEFAULT_LIB = os.environ.get("TCB200_LIB") or os.path.join(_HERE, "csrc", "libtcb200.so")

ENGINE_QUAD_REG = 1      # round-1 fused register-engine pairing kernel (self-test reference, A/B measurement)
ENGINE_QUAD_SMEM = 2     # default: Miller loop and final exponentiation on shared-memory cells (k_miller_quad + k_final_exp_sm)
ENGINE_QUAD_SMEM_REGFE = 3  # shared-memory Miller loop + round 1's register-engine final exponentiation (A/B measurement)


class TcbError(RuntimeError):
    pass


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=np.uint8)


def _need(name, arr, size):
    """Every buffer length is checked before it crosses the C ABI (the library trusts its callers' sizes)."""
    if arr is None or arr.size != size:
        raise ValueError(f"{name}: expected {size} bytes, got {None if arr is None else arr.size}")


def pack_msgs(msgs):
    """list of bytes -> (concatenated uint8 buffer, uint64 offsets[n+1])"""
    off = np.zeros(len(msgs) + 1, dtype=np.uint64)
    if len(msgs):
        off[1:] = np.cumsum([len(m) for m in msgs])
    buf = np.frombuffer(b"".join(bytes(m) for m in msgs) or b"\0", dtype=np.uint8).copy()
    return buf, off


class Engine:
    """One tcb_ctx.  Methods mirror include/tcb200.h one to one (numpy uint8 arrays in/out)."""

    def __init__(self, lib_path=None, devices=None):
        path = lib_path or DEFAULT_LIB
        if not os.path.exists(path):
            raise TcbError(
                f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "threshold_crypto_b200 has no CPU fallback.")
        self.lib = C.CDLL(path)
        self.lib.tcb_last_error.restype = C.c_char_p
        self.lib.tcb_launch_count.restype = C.c_uint64
        self.ctx = C.c_void_p()
        if devices is None:
            rc = self.lib.tcb_init(C.byref(self.ctx), None, 0)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.tcb_init(C.byref(self.ctx), arr, len(devices))
        if rc != 0:
            raise TcbError(f"tcb_init failed (rc={rc}): no usable CUDA device; there is no CPU fallback")
        self.path = path

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.tcb_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise TcbError(f"tcb call failed rc={rc}: {self.lib.tcb_last_error(self.ctx).decode()}")

    def set_engine(self, engine):
        self._ck(self.lib.tcb_set_engine(self.ctx, int(engine)))

    def set_verify_hash(self, exact):
        """verify_batch: 0 (default) pairs [3(x^2-1)]H(m) with [3(x^2-1)]g1 (the hash kernel skips the last third of the cofactor
        clearing), 1 the exact H(m) with g1.  Same booleans."""
        self._ck(self.lib.tcb_set_verify_hash(self.ctx, int(exact)))

    def set_hash_algo(self, algo):
        """hash_g2: 0 (default) point kernel (one thread per item) + cofactor-clearing kernel (lane pairs), 1 the one-kernel version."""
        self._ck(self.lib.tcb_set_hash_algo(self.ctx, int(algo)))

    def set_msm_groups(self, groups):
        """Partial sums per item of the shared-doubling multi-scalar multiplication (0 = auto)."""
        self._ck(self.lib.tcb_set_msm_groups(self.ctx, C.c_size_t(int(groups))))

    def set_msm_algo(self, algo):
        """0 Straus / shared doublings (default), 1 batch-affine tree, 2 one multiplication per share, 3 G2 accumulation per thread,
        4 G2 accumulation on shared-memory cells, 5 Straus without the spill layout, 6 Straus with the spill layout forced."""
        self._ck(self.lib.tcb_set_msm_algo(self.ctx, int(algo)))

    def set_eval_split(self, units):
        """Units per point of Commitment::evaluate: 0 auto, 1 never split, k forced."""
        self._ck(self.lib.tcb_set_eval_split(self.ctx, C.c_size_t(int(units))))

    def verifier_generator(self):
        """96-byte G1 point that verify_batch pairs with the signature: [3(x^2-1)] g1 by default, g1 after set_verify_hash(1)."""
        out = np.zeros(96, np.uint8)
        self._ck(self.lib.tcb_verifier_generator(self.ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def launch_count(self):
        return int(self.lib.tcb_launch_count(self.ctx))

    # ---- host-buffer API
    def verify_g2_batch(self, a_g1, b_g2, c_g1, d_g2):
        a, b, c, d = _u8(a_g1), _u8(b_g2), _u8(c_g1), _u8(d_g2)
        n = a.size // 96
        _need("a_g1", a, 96 * n); _need("b_g2", b, 192 * n); _need("d_g2", d, 192 * n)
        if c is not None:
            _need("c_g1", c, 96 * n)
        ok = np.zeros(n, np.uint8)
        self._ck(self.lib.tcb_verify_g2_batch(self.ctx, C.c_size_t(n), _p(a), _p(b), _p(c), _p(d), _p(ok)))
        return ok

    def hash_g2_batch(self, msgs):
        buf, off = pack_msgs(msgs)
        out = np.zeros((len(msgs), 192), np.uint8)
        self._ck(self.lib.tcb_hash_g2_batch(self.ctx, C.c_size_t(len(msgs)), _p(buf), _p(off), _p(out)))
        return out

    def hash_g1_g2_batch(self, g1, msgs):
        g = _u8(g1)
        _need("g1", g, 96 * len(msgs))
        buf, off = pack_msgs(msgs)
        out = np.zeros((len(msgs), 192), np.uint8)
        self._ck(self.lib.tcb_hash_g1_g2_batch(self.ctx, C.c_size_t(len(msgs)), _p(g), _p(buf), _p(off), _p(out)))
        return out

    def verify_batch(self, pk_g1, sig_g2, msgs):
        pk, sig = _u8(pk_g1), _u8(sig_g2)
        _need("pk_g1", pk, 96 * len(msgs)); _need("sig_g2", sig, 192 * len(msgs))
        buf, off = pack_msgs(msgs)
        ok = np.zeros(len(msgs), np.uint8)
        self._ck(self.lib.tcb_verify_batch(self.ctx, C.c_size_t(len(msgs)), _p(pk), _p(sig), _p(buf), _p(off), _p(ok)))
        return ok

    def sign_batch(self, sk, msgs):
        sk = _u8(sk)
        _need("sk", sk, 32 * len(msgs))
        buf, off = pack_msgs(msgs)
        out = np.zeros((len(msgs), 192), np.uint8)
        self._ck(self.lib.tcb_sign_batch(self.ctx, C.c_size_t(len(msgs)), _p(sk), _p(buf), _p(off), _p(out)))
        return out

    def sign_g2_batch(self, sk, h_g2):
        sk, h = _u8(sk), _u8(h_g2)
        n = sk.size // 32
        _need("sk", sk, 32 * n); _need("h_g2", h, 192 * n)
        out = np.zeros((n, 192), np.uint8)
        self._ck(self.lib.tcb_sign_g2_batch(self.ctx, C.c_size_t(n), _p(sk), _p(h), _p(out)))
        return out

    def combine_g2_batch(self, n, t, x_fr, shares_g2):
        x, s = _u8(x_fr), _u8(shares_g2)
        _need("x_fr", x, n * (t + 1) * 32); _need("shares_g2", s, n * (t + 1) * 192)
        out = np.zeros((n, 192), np.uint8)
        st = np.zeros(n, np.uint8)
        self._ck(self.lib.tcb_combine_g2_batch(self.ctx, C.c_size_t(n), C.c_size_t(t), _p(x), _p(s), _p(out), _p(st)))
        return out, st

    def combine_g1_batch(self, n, t, x_fr, shares_g1):
        x, s = _u8(x_fr), _u8(shares_g1)
        _need("x_fr", x, n * (t + 1) * 32); _need("shares_g1", s, n * (t + 1) * 96)
        out = np.zeros((n, 96), np.uint8)
        st = np.zeros(n, np.uint8)
        self._ck(self.lib.tcb_combine_g1_batch(self.ctx, C.c_size_t(n), C.c_size_t(t), _p(x), _p(s), _p(out), _p(st)))
        return out, st

    def decrypt_share_batch(self, sk, u_g1):
        sk, u = _u8(sk), _u8(u_g1)
        n = sk.size // 32
        _need("sk", sk, 32 * n); _need("u_g1", u, 96 * n)
        out = np.zeros((n, 96), np.uint8)
        self._ck(self.lib.tcb_decrypt_share_batch(self.ctx, C.c_size_t(n), _p(sk), _p(u), _p(out)))
        return out

    def decrypt_batch(self, n, t, x_fr, shares_g1, vs):
        x, s = _u8(x_fr), _u8(shares_g1)
        _need("x_fr", x, n * (t + 1) * 32); _need("shares_g1", s, n * (t + 1) * 96)
        if len(vs) != n:
            raise ValueError(f"decrypt_batch: {n} items but {len(vs)} ciphertext bodies")
        buf, off = pack_msgs(vs)
        out = np.zeros(buf.size, np.uint8)
        st = np.zeros(n, np.uint8)
        self._ck(self.lib.tcb_decrypt_batch(self.ctx, C.c_size_t(n), C.c_size_t(t), _p(x), _p(s), _p(buf), _p(off), _p(out), _p(st)))
        return [bytes(out[int(off[i]):int(off[i + 1])]) for i in range(n)], st

    def commitment_eval_batch(self, coeff_g1, x_fr):
        c, x = _u8(coeff_g1), _u8(x_fr)
        n = x.size // 32
        _need("x_fr", x, 32 * n)
        if c.size == 0 or c.size % 96:
            raise ValueError("commitment_eval_batch: the commitment must hold at least one 96-byte coefficient "
                             "(the empty commitment evaluates to G1::zero(): api.Commitment handles it)")
        deg = c.size // 96 - 1
        out = np.zeros((n, 96), np.uint8)
        self._ck(self.lib.tcb_commitment_eval_batch(self.ctx, C.c_size_t(deg), _p(c), C.c_size_t(n), _p(x), _p(out)))
        return out

    def g1_mul_gen_batch(self, sk):
        sk = _u8(sk)
        n = sk.size // 32
        out = np.zeros((n, 96), np.uint8)
        self._ck(self.lib.tcb_g1_mul_gen_batch(self.ctx, C.c_size_t(n), _p(sk), _p(out)))
        return out

    def g1_lincomb_batch(self, n, m, scalars, pts):
        sc, p = _u8(scalars), _u8(pts)
        _need("scalars", sc, n * m * 32); _need("pts_g1", p, n * m * 96)
        out = np.zeros((n, 96), np.uint8)
        self._ck(self.lib.tcb_g1_lincomb_batch(self.ctx, C.c_size_t(n), C.c_size_t(m), _p(sc), _p(p), _p(out)))
        return out

    def g2_lincomb_batch(self, n, m, scalars, pts):
        sc, p = _u8(scalars), _u8(pts)
        _need("scalars", sc, n * m * 32); _need("pts_g2", p, n * m * 192)
        out = np.zeros((n, 192), np.uint8)
        self._ck(self.lib.tcb_g2_lincomb_batch(self.ctx, C.c_size_t(n), C.c_size_t(m), _p(sc), _p(p), _p(out)))
        return out

    # ---- Fr-side Poly algebra
    def poly_eval_batch(self, coeff_fr, x_fr):
        """Poly::evaluate of one polynomial (canonical LE coefficients, constant term first) at every x"""
        c, x = _u8(coeff_fr), _u8(x_fr)
        if c.size == 0 or c.size % 32 or x.size % 32:
            raise ValueError("poly_eval_batch: coefficients and points are 32-byte scalars, at least one coefficient")
        n = x.size // 32
        out = np.zeros((n, 32), np.uint8)
        self._ck(self.lib.tcb_poly_eval_batch(self.ctx, C.c_size_t(c.size // 32 - 1), _p(c), C.c_size_t(n), _p(x), _p(out)))
        return out

    def poly_mul_batch(self, n, a_fr, b_fr):
        """n products of a polynomial of degree da with one of degree db -> (n, da + db + 1, 32)"""
        a, b = _u8(a_fr), _u8(b_fr)
        if n == 0 or a.size % (32 * n) or b.size % (32 * n) or a.size == 0 or b.size == 0:
            raise ValueError("poly_mul_batch: n items of (da + 1) and (db + 1) 32-byte coefficients")
        da, db = a.size // (32 * n) - 1, b.size // (32 * n) - 1
        out = np.zeros((n, da + db + 1, 32), np.uint8)
        self._ck(self.lib.tcb_poly_mul_batch(self.ctx, C.c_size_t(n), C.c_size_t(da), _p(a), C.c_size_t(db), _p(b), _p(out)))
        return out

    def encrypt_batch(self, pk_g1, r_fr, msgs):
        pk, r = _u8(pk_g1), _u8(r_fr)
        buf, off = pack_msgs(msgs)
        n = len(msgs)
        _need("pk_g1", pk, 96 * n); _need("r_fr", r, 32 * n)
        u = np.zeros((n, 96), np.uint8); v = np.zeros(buf.size, np.uint8); w = np.zeros((n, 192), np.uint8)
        self._ck(self.lib.tcb_encrypt_batch(self.ctx, C.c_size_t(n), _p(pk), _p(r), _p(buf), _p(off), _p(u), _p(v), _p(w)))
        return u, [bytes(v[int(off[i]):int(off[i + 1])]) for i in range(n)], w

    # ---- wire-format codecs
    def _codec(self, name, data, in_w, out_w, with_status):
        a = _u8(data)
        n = a.size // in_w
        _need(name, a, in_w * n)
        out = np.zeros((n, out_w), np.uint8)
        if with_status:
            st = np.zeros(n, np.uint8)
            self._ck(getattr(self.lib, name)(self.ctx, C.c_size_t(n), _p(a), _p(out), _p(st)))
            return out, st
        self._ck(getattr(self.lib, name)(self.ctx, C.c_size_t(n), _p(a), _p(out)))
        return out

    def g1_compress_batch(self, unc):
        return self._codec("tcb_g1_compress_batch", unc, 96, 48, False)

    def g2_compress_batch(self, unc):
        return self._codec("tcb_g2_compress_batch", unc, 192, 96, False)

    def g1_decompress_batch(self, comp):
        return self._codec("tcb_g1_decompress_batch", comp, 48, 96, True)

    def g2_decompress_batch(self, comp):
        return self._codec("tcb_g2_decompress_batch", comp, 96, 192, True)

    # ---- self-test / probes (CUDA library only)
    def selftest_fp(self, n=1 << 16, seed=1):
        rc = self.lib.tcb_selftest_fp(self.ctx, C.c_size_t(n), C.c_uint64(seed))
        if rc < 0:
            self._ck(rc)
        return rc

    def selftest_miller(self, a_g1, b_g2, c_g1, d_g2):
        """differing words between the Miller-loop values of the shared-memory and the register engine (0 = identical)"""
        a, b, c, d = _u8(a_g1), _u8(b_g2), _u8(c_g1), _u8(d_g2)
        rc = self.lib.tcb_selftest_miller(self.ctx, C.c_size_t(a.size // 96), _p(a), _p(b), _p(c), _p(d))
        if rc < 0:
            self._ck(rc)
        return rc

    def probe_imad(self):
        v = C.c_double()
        self._ck(self.lib.tcb_probe_imad(self.ctx, C.byref(v)))
        return v.value

    def probe_fpmul(self):
        v = C.c_double()
        self._ck(self.lib.tcb_probe_fpmul(self.ctx, C.byref(v)))
        return v.value

    # ---- device-pointer API (ints are raw device addresses, stream is a cudaStream_t as int)
    def dev_call(self, name, stream, *args):
        fn = getattr(self.lib, name)
        cargs = [self.ctx, C.c_void_p(stream)]
        for a in args:
            if isinstance(a, tuple):      # ("size", value)
                cargs.append(C.c_size_t(a[1]) if a[0] == "size" else C.c_uint64(a[1]))
            elif a is None:
                cargs.append(None)
            else:
                cargs.append(C.c_void_p(int(a)))
        self._ck(fn(*cargs))
