"""Kernel-level timing of every BASELINE config on ONE GPU for one build of the library (TCB200_LIB selects an
experiment variant).  Inputs are made on the GPU with the same library, every output is checked through a
size-independent identity (so a wrong variant is caught), times are CUDA events on torch's current stream with
device-resident inputs (the `_dev` entry points), best and median of REPS.

    TCB200_LIB=threshold_crypto_b200/csrc/libtcb200_x.so python tools/kbench.py [tag] [which,...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import fr_bytes, rand_fr
from threshold_crypto_b200._lib import Engine, pack_msgs

tag = sys.argv[1] if len(sys.argv) > 1 else "lib"
which = set((sys.argv[2] if len(sys.argv) > 2 else "verify,combine,decrypt,eval,g1mul,sign").split(","))
REPS = int(os.environ.get("REPS", "5"))
E = Engine(devices=[0])
if "HASH_ALGO" in os.environ:
    E.set_hash_algo(int(os.environ["HASH_ALGO"]))
if "VERIFY_HASH" in os.environ:
    E.set_verify_hash(int(os.environ["VERIFY_HASH"]))
if "MSM_ALGO" in os.environ:
    E.set_msm_algo(int(os.environ["MSM_ALGO"]))
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(7)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"tag": tag, "lib": os.path.basename(E.path), "hash_algo": os.environ.get("HASH_ALGO", "default"), "verify_hash": os.environ.get("VERIFY_HASH", "default")}


def D(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def timeit(fn, reps=REPS):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for k in range(reps):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return {"best_ms": min(ts), "median_ms": float(np.median(ts))}


if "verify" in which:
    n = 1 << 16
    sk = rand_fr(rng, n)
    msgs = [i.to_bytes(8, "little") * 4 for i in range(n)]
    pk = E.g1_mul_gen_batch(sk)
    sig = E.sign_batch(sk, msgs)
    bad = np.arange(n) % 16 == 5
    sig[bad] = np.roll(sig, 1, axis=0)[bad]
    mbuf, moff = pack_msgs(msgs)
    d_pk, d_sig, d_msg, d_off = D(pk.reshape(-1)), D(sig.reshape(-1)), D(mbuf), D(moff.view(np.int64))
    d_h = torch.zeros(n * 192, dtype=torch.uint8, device=dev)
    d_ok = torch.zeros(n, dtype=torch.uint8, device=dev)
    res["hash_g2_2^16"] = timeit(lambda: E.dev_call("tcb_hash_g2_batch_dev", st, ("size", n), d_msg.data_ptr(), d_off.data_ptr(), d_h.data_ptr()))
    res["pairing_2^16"] = timeit(lambda: E.dev_call("tcb_verify_g2_batch_dev", st, ("size", n), d_pk.data_ptr(), d_h.data_ptr(), 0, d_sig.data_ptr(), d_ok.data_ptr()))
    assert np.array_equal(d_ok.cpu().numpy().astype(bool), ~bad), "pairing output wrong"
    d_ok.zero_()
    res["verify_2^16"] = timeit(lambda: E.dev_call("tcb_verify_batch_dev", st, ("size", n), d_pk.data_ptr(), d_sig.data_ptr(), d_msg.data_ptr(), d_off.data_ptr(), d_ok.data_ptr()))
    assert np.array_equal(d_ok.cpu().numpy().astype(bool), ~bad), "verify output wrong"
    res["verify_2^16"]["per_s"] = n / (res["verify_2^16"]["best_ms"] * 1e-3)

if "sign" in which:
    n = 1 << 14
    sk = rand_fr(rng, n)
    h = E.hash_g2_batch([b"s%d" % i for i in range(n)])
    d_sk, d_h2 = D(sk), D(h.reshape(-1))
    d_o = torch.zeros(n * 192, dtype=torch.uint8, device=dev)
    res["sign_g2_2^14"] = timeit(lambda: E.dev_call("tcb_sign_batch_dev", st, ("size", n), d_sk.data_ptr(), 0, 0, d_h2.data_ptr(), d_o.data_ptr()))

if "combine" in which:
    n, t = 1 << 14, 10
    m = t + 1
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    poly = rand_fr(rng, m)
    coeffs = [int.from_bytes(bytes(poly[32 * k:32 * k + 32]), "little") for k in range(m)]
    skj = []
    for j in range(32):
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * (j + 1) + c) % R
        skj.append(acc.to_bytes(32, "little"))
    idx = np.stack([np.sort(rng.choice(32, size=m, replace=False)) for _ in range(n)])
    xs = fr_bytes([int(j) + 1 for j in idx.reshape(-1)])
    hm = E.hash_g2_batch([b"c%d" % i for i in range(n)])
    sk_rep = np.frombuffer(b"".join(skj[j] for j in idx.reshape(-1)), np.uint8).copy()
    shares = E.sign_g2_batch(sk_rep, np.repeat(hm, m, axis=0))
    master = E.sign_g2_batch(np.tile(poly[:32], n), hm)
    d_x, d_sh = D(xs), D(shares.reshape(-1))
    d_out = torch.zeros(n * 192, dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    res["combine_t10_2^14"] = timeit(lambda: E.dev_call("tcb_combine_g2_batch_dev", st, ("size", n), ("size", t), d_x.data_ptr(), d_sh.data_ptr(), d_out.data_ptr(), d_st.data_ptr()))
    assert np.array_equal(d_out.cpu().numpy().reshape(n, 192), master), "combine output wrong"
    for sub in (2048,):
        res[f"combine_t10_{sub}"] = timeit(lambda: E.dev_call("tcb_combine_g2_batch_dev", st, ("size", sub), ("size", t), d_x.data_ptr(), d_sh.data_ptr(), d_out.data_ptr(), d_st.data_ptr()))

if "decrypt" in which:
    n, t = 1 << 12, 64
    m = t + 1
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    poly = rand_fr(rng, m)
    coeffs = [int.from_bytes(bytes(poly[32 * k:32 * k + 32]), "little") for k in range(m)]
    sks = []
    for j in range(m):
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * (j + 1) + c) % R
        sks.append(acc.to_bytes(32, "little"))
    sk_shares = np.frombuffer(b"".join(sks), np.uint8).copy()
    pkm = E.g1_mul_gen_batch(poly[:32])
    rs = rand_fr(rng, n)
    plains = [bytes([i & 0xff, (i >> 8) & 0xff]) * 32 for i in range(n)]
    u, v, w = E.encrypt_batch(np.tile(pkm[0], (n, 1)), rs, plains)
    sk_rep = np.tile(sk_shares, n)
    u_rep = np.repeat(u, m, axis=0)
    d_skr, d_ur = D(sk_rep), D(u_rep.reshape(-1))
    d_sh1 = torch.zeros(n * m * 96, dtype=torch.uint8, device=dev)
    res["decrypt_shares_65x2^12"] = timeit(lambda: E.dev_call("tcb_g1_mul_batch_dev", st, ("size", n * m), d_skr.data_ptr(), d_ur.data_ptr(), d_sh1.data_ptr()), 3)
    xs = np.tile(fr_bytes([i + 1 for i in range(m)]), n)
    vbuf, voff = pack_msgs(v)
    d_x1, d_v, d_voff = D(xs), D(vbuf), D(voff.view(np.int64))
    d_pl = torch.zeros(vbuf.size, dtype=torch.uint8, device=dev)
    d_st1 = torch.zeros(n, dtype=torch.uint8, device=dev)
    res["decrypt_t64_2^12"] = timeit(lambda: E.dev_call("tcb_decrypt_batch_dev", st, ("size", n), ("size", t), d_x1.data_ptr(), d_sh1.data_ptr(), d_v.data_ptr(),
                                                         d_voff.data_ptr(), ("u64", int(vbuf.size)), d_pl.data_ptr(), d_st1.data_ptr()))
    assert d_pl.cpu().numpy().tobytes() == b"".join(plains), "decrypt output wrong"
    res["decrypt_t64_512"] = timeit(lambda: E.dev_call("tcb_decrypt_batch_dev", st, ("size", 512), ("size", t), d_x1.data_ptr(), d_sh1.data_ptr(), d_v.data_ptr(),
                                                        d_voff.data_ptr(), ("u64", int(vbuf.size)), d_pl.data_ptr(), d_st1.data_ptr()))

if "eval" in which:
    deg = int(os.environ.get("DEG", "1023"))
    coeff = rand_fr(rng, deg + 1)
    comm = E.g1_mul_gen_batch(coeff)
    n = 1 << 16
    xs = fr_bytes([i + 1 for i in range(n)])
    d_c, d_x5 = D(comm.reshape(-1)), D(xs)
    d_o5 = torch.zeros(n * 96, dtype=torch.uint8, device=dev)
    res[f"commit_eval_deg{deg}_2^16"] = timeit(lambda: E.dev_call("tcb_commitment_eval_batch_dev", st, ("size", deg), d_c.data_ptr(), ("size", n), d_x5.data_ptr(), d_o5.data_ptr()), 2)
    out = d_o5.cpu().numpy().reshape(n, 96)
    sel = np.array([0, 1, 2, 4095, 4096, 65534, 65535])
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    cs = [int.from_bytes(bytes(coeff[32 * k:32 * k + 32]), "little") for k in range(deg + 1)]
    ys = []
    for i in sel:
        acc = 0
        for c in reversed(cs):
            acc = (acc * (int(i) + 1) + c) % R
        ys.append(acc)
    assert np.array_equal(out[sel], E.g1_mul_gen_batch(fr_bytes(ys))), "commit_eval output wrong"
    res[f"commit_eval_deg{deg}_8192"] = timeit(lambda: E.dev_call("tcb_commitment_eval_batch_dev", st, ("size", deg), d_c.data_ptr(), ("size", 8192), d_x5.data_ptr(), d_o5.data_ptr()), 2)
    assert np.array_equal(d_o5.cpu().numpy().reshape(n, 96)[:8192], out[:8192])

if "g1mul" in which:
    n = 1 << 18
    sk = rand_fr(rng, n)
    d_sk1 = D(sk)
    d_o1 = torch.zeros(n * 96, dtype=torch.uint8, device=dev)
    res["g1_mul_gen_2^18"] = timeit(lambda: E.dev_call("tcb_g1_mul_batch_dev", st, ("size", n), d_sk1.data_ptr(), 0, d_o1.data_ptr()), 3)

print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"kbench_{tag}.json"), "w") as f:
    json.dump(res, f, indent=1)
