"""CPU: the N > 1 sharding path (threshold_crypto_b200/dist.py) with world_size 2 over gloo,
each rank computing on the host-emulation engine; results must equal the oracle's."""
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
    from threshold_crypto_b200 import dist as tdist
    E = Engine(conftest.build_hostemu())
    n = 5                                    # ragged split 2 + 3
    pk = sig = msgs = None
    if rank == 0:
        sk, pk, sig, msgs = cases.make_sig_batch(O, n, 31, corrupt_every=3)
        msgs = [m.ljust(32, b"\0")[:32] for m in msgs]
        sig = O.sign_batch(sk, msgs)
        sig[1] = sig[2]
    ok = tdist.verify_batch_sharded(E, n, pk, sig, msgs, msg_len=32)
    t = 2
    xs = sh = None
    if rank == 0:
        xs, sh, master = cases.make_combine_batch(O, 3, t, 32, group=2)
    out, st = tdist.combine_g2_batch_sharded(E, 3, t, xs, sh)
    if rank == 0:
        exp = O.verify_batch(pk, sig, msgs)
        q.put((bool(np.array_equal(ok, exp)) and 0 < exp.sum() < n, bool(np.array_equal(out, master)) and not st.any()))
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
    assert res == (True, True)


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from threshold_crypto_b200.dist import shard_bounds
    for n in (0, 1, 5, 16, 65537):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
