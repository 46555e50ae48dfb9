"""CPU: the N > 1 data plane (threshold_crypto_b200/dist.py) with world_size 2 over gloo, each rank computing on the
host-emulation engine; results must equal the oracle's.  Covers the partitioning (ragged 2 + 3 split, a rank with
no items), ragged messages scattered as bytes + offsets, and every sharded operation (verify, combine in G2,
decrypt, Commitment::evaluate with the broadcast table)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    import cases
    import conftest
    from threshold_crypto_b200._lib import Engine
    from threshold_crypto_b200.dist import ShardedEngine
    S = ShardedEngine(Engine(conftest.build_hostemu()))
    root = rank == 0
    res = {}
    n = 5                                    # ragged split 2 + 3, messages of different lengths (1..29 bytes)
    pk = sig = msgs = None
    if root:
        sk, pk, sig, msgs = cases.make_sig_batch(O, n, 31, corrupt_every=3)
    ok = S.verify_batch(n, pk, sig, msgs)
    if root:
        exp = O.verify_batch(pk, sig, msgs)
        res["verify"] = bool(np.array_equal(ok, exp)) and 0 < exp.sum() < n and len({len(m) for m in msgs}) > 1
        res["timing_keys"] = sorted(S.last_timing) == sorted(["h2d", "scatter", "compute", "gather", "d2h", "total"])
    # one item only: rank 0 gets nothing (shard_bounds(1, 0, 2) == (0, 0)), rank 1 the item
    ok1 = S.verify_batch(1, pk[:1] if root else None, sig[:1] if root else None, msgs[:1] if root else None)
    if root:
        res["verify_one"] = bool(np.array_equal(ok1, exp[:1]))
    t = 2
    xs = sh = None
    if root:
        xs, sh, master = cases.make_combine_batch(O, 3, t, 32, group=2)
    out, st = S.combine_g2_batch(3, t, xs, sh)
    if root:
        res["combine"] = bool(np.array_equal(out, master)) and not st.any()
    x1 = s1 = vs = None
    if root:
        x1, s1, _ = cases.make_combine_batch(O, 3, t, 33, group=1)
        vs = [b"", bytes(range(70)), b"abc"]
    dec, dst = S.decrypt_batch(3, t, x1, s1, vs)
    if root:
        odec, ost = O.decrypt_batch(3, t, x1, s1, vs)
        res["decrypt"] = dec == odec and bool(np.array_equal(dst, ost))
    comm = xe = None
    deg, ne = 4, 7
    if root:
        rng = np.random.default_rng(3)
        comm = O.g1_mul_gen_batch(conftest.rand_fr(rng, deg + 1))
        xe = conftest.fr_bytes([i + 1 for i in range(ne - 1)] + [int.from_bytes(rng.bytes(40), "little")])
    ev = S.commitment_eval_batch(ne, deg, comm, xe)
    if root:
        res["eval"] = bool(np.array_equal(ev, O.commitment_eval_batch(comm, xe)))
        q.put(res)
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=60)
    assert res == {k: True for k in ("verify", "timing_keys", "verify_one", "combine", "decrypt", "eval")}, res


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from threshold_crypto_b200.dist import shard_bounds
    for n in (0, 1, 5, 16, 65537):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
