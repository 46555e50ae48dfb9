"""ctypes wrapper over oracle/libtc_oracle.so (TEST INFRASTRUCTURE ONLY — see tc_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtc_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "tc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_fp_mul_count.restype = C.c_uint64
        _lib.orc_init()
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(C.c_void_p)


def set_threads(n):
    lib().orc_set_threads(int(n))


def pack_msgs(msgs):
    off = np.zeros(len(msgs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(m) for m in msgs])
    buf = np.frombuffer(b"".join(bytes(m) for m in msgs) or b"\0", dtype=np.uint8).copy()
    return buf, off


def g1_generator():
    out = np.zeros(96, np.uint8); lib().orc_g1_generator(out.ctypes.data_as(C.c_void_p)); return out


def g2_generator():
    out = np.zeros(192, np.uint8); lib().orc_g2_generator(out.ctypes.data_as(C.c_void_p)); return out


def verify_g2_batch(a, b, c, d):
    a, pa = _u8(a); b, pb = _u8(b); d, pd = _u8(d)
    n = a.size // 96
    pc = None
    if c is not None:
        c, pc = _u8(c)
    ok = np.zeros(n, np.uint8)
    lib().orc_verify_g2_batch(C.c_size_t(n), pa, pb, pc, pd, ok.ctypes.data_as(C.c_void_p))
    return ok


def hash_g2_batch(msgs):
    buf, off = pack_msgs(msgs)
    out = np.zeros((len(msgs), 192), np.uint8)
    lib().orc_hash_g2_batch(C.c_size_t(len(msgs)), buf.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                            out.ctypes.data_as(C.c_void_p))
    return out


def verify_batch(pk, sig, msgs):
    pk, ppk = _u8(pk); sig, psig = _u8(sig)
    buf, off = pack_msgs(msgs)
    ok = np.zeros(len(msgs), np.uint8)
    lib().orc_verify_batch(C.c_size_t(len(msgs)), ppk, psig, buf.ctypes.data_as(C.c_void_p),
                           off.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p))
    return ok


def sign_batch(sk, msgs):
    sk, psk = _u8(sk)
    buf, off = pack_msgs(msgs)
    out = np.zeros((len(msgs), 192), np.uint8)
    lib().orc_sign_batch(C.c_size_t(len(msgs)), psk, buf.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                         out.ctypes.data_as(C.c_void_p))
    return out


def sign_g2_batch(sk, h):
    sk, psk = _u8(sk); h, ph = _u8(h)
    n = sk.size // 32
    out = np.zeros((n, 192), np.uint8)
    lib().orc_sign_g2_batch(C.c_size_t(n), psk, ph, out.ctypes.data_as(C.c_void_p))
    return out


def _combine(fn, width, n, t, x, shares):
    x, px = _u8(x); shares, ps = _u8(shares)
    out = np.zeros((n, width), np.uint8)
    status = np.zeros(n, np.uint8)
    fn(C.c_size_t(n), C.c_size_t(t), px, ps, out.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p))
    return out, status


def combine_g2_batch(n, t, x, shares):
    return _combine(lib().orc_combine_g2_batch, 192, n, t, x, shares)


def combine_g1_batch(n, t, x, shares):
    return _combine(lib().orc_combine_g1_batch, 96, n, t, x, shares)


def decrypt_batch(n, t, x, shares, vs):
    x, px = _u8(x); shares, ps = _u8(shares)
    buf, off = pack_msgs(vs)
    out = np.zeros(buf.size, np.uint8)
    status = np.zeros(n, np.uint8)
    lib().orc_decrypt_batch(C.c_size_t(n), C.c_size_t(t), px, ps, buf.ctypes.data_as(C.c_void_p),
                            off.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                            status.ctypes.data_as(C.c_void_p))
    return [bytes(out[int(off[i]):int(off[i + 1])]) for i in range(n)], status


def decrypt_share_batch(sk, u):
    sk, psk = _u8(sk); u, pu = _u8(u)
    n = sk.size // 32
    out = np.zeros((n, 96), np.uint8)
    lib().orc_decrypt_share_batch(C.c_size_t(n), psk, pu, out.ctypes.data_as(C.c_void_p))
    return out


def g1_mul_gen_batch(sk):
    sk, psk = _u8(sk)
    n = sk.size // 32
    out = np.zeros((n, 96), np.uint8)
    lib().orc_g1_mul_gen_batch(C.c_size_t(n), psk, out.ctypes.data_as(C.c_void_p))
    return out


def commitment_eval_batch(coeff, x):
    coeff, pc = _u8(coeff); x, px = _u8(x)
    deg = coeff.size // 96 - 1
    n = x.size // 32
    out = np.zeros((n, 96), np.uint8)
    lib().orc_commitment_eval_batch(C.c_size_t(deg), pc, C.c_size_t(n), px, out.ctypes.data_as(C.c_void_p))
    return out


def g1_compress(unc):
    unc, p = _u8(unc); n = unc.size // 96
    out = np.zeros((n, 48), np.uint8)
    assert lib().orc_g1_compress(C.c_size_t(n), p, out.ctypes.data_as(C.c_void_p)) == 0
    return out


def g2_compress(unc):
    unc, p = _u8(unc); n = unc.size // 192
    out = np.zeros((n, 96), np.uint8)
    assert lib().orc_g2_compress(C.c_size_t(n), p, out.ctypes.data_as(C.c_void_p)) == 0
    return out


def g1_decompress(comp):
    comp, p = _u8(comp); n = comp.size // 48
    out = np.zeros((n, 96), np.uint8); st = np.zeros(n, np.uint8)
    lib().orc_g1_decompress(C.c_size_t(n), p, out.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
    return out, st


def g2_decompress(comp):
    comp, p = _u8(comp); n = comp.size // 96
    out = np.zeros((n, 192), np.uint8); st = np.zeros(n, np.uint8)
    lib().orc_g2_decompress(C.c_size_t(n), p, out.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
    return out, st


def pairing_gt(p, q):
    p, pp = _u8(p); q, pq = _u8(q)
    out = np.zeros(576, np.uint8)
    lib().orc_pairing_gt(pp, pq, out.ctypes.data_as(C.c_void_p))
    return out


def xor_with_hash(g1_unc, data):
    g, pg = _u8(g1_unc)
    d = np.frombuffer(bytes(data) or b"\0", np.uint8).copy()
    out = np.zeros(max(len(data), 1), np.uint8)
    lib().orc_xor_with_hash(pg, d.ctypes.data_as(C.c_void_p), C.c_size_t(len(data)), out.ctypes.data_as(C.c_void_p))
    return bytes(out[:len(data)])


def hash_g1_g2(g1_unc, msg):
    g, pg = _u8(g1_unc)
    d = np.frombuffer(bytes(msg) or b"\0", np.uint8).copy()
    out = np.zeros(192, np.uint8)
    lib().orc_hash_g1_g2(pg, d.ctypes.data_as(C.c_void_p), C.c_size_t(len(msg)), out.ctypes.data_as(C.c_void_p))
    return out


def fr_random_stream(seed32, n):
    s = np.frombuffer(bytes(seed32), np.uint8).copy()
    out = np.zeros((n, 32), np.uint8)
    lib().orc_fr_random_stream(s.ctypes.data_as(C.c_void_p), C.c_size_t(n), out.ctypes.data_as(C.c_void_p))
    return out


def poly_eval(coeff, x):
    coeff, pc = _u8(coeff); x, px = _u8(x)
    n = x.size // 32
    out = np.zeros((n, 32), np.uint8)
    lib().orc_poly_eval(C.c_size_t(coeff.size // 32), pc, C.c_size_t(n), px, out.ctypes.data_as(C.c_void_p))
    return out


def encrypt(pk, r32, msg):
    pk, ppk = _u8(pk); r, pr = _u8(r32)
    d = np.frombuffer(bytes(msg) or b"\0", np.uint8).copy()
    u = np.zeros(96, np.uint8); v = np.zeros(max(len(msg), 1), np.uint8); w = np.zeros(192, np.uint8)
    lib().orc_encrypt(ppk, pr, d.ctypes.data_as(C.c_void_p), C.c_size_t(len(msg)), u.ctypes.data_as(C.c_void_p),
                      v.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
    return u, bytes(v[:len(msg)]), w


def fp_mul_count(reset=False):
    if reset:
        lib().orc_fp_mul_count_reset()
        return 0
    return int(lib().orc_fp_mul_count())
