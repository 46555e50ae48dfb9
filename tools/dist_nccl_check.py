"""Run under torchrun on the GPU box: the scatter -> per-rank verify/combine -> gather path of
threshold_crypto_b200/dist.py over NCCL, checked against the oracle on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
from threshold_crypto_b200._lib import Engine
from threshold_crypto_b200 import dist as tdist
E = Engine(devices=[local])
n = 1000 + 3          # ragged over the ranks
pk = sig = msgs = xs = sh = None
if rank == 0:
    import oracle as O
    import cases
    O.set_threads(16)
    sk, pk, sig, msgs = cases.make_sig_batch(O, n, 5, corrupt_every=7)
    msgs = [m.ljust(32, b"\0")[:32] for m in msgs]
    sig = O.sign_batch(sk, msgs)
    sig[1::7] = np.roll(sig, 1, axis=0)[1::7]
    xs, sh, master = cases.make_combine_batch(O, 101, 4, 6, group=2)
ok = tdist.verify_batch_sharded(E, n, pk, sig, msgs, msg_len=32)
out, st = tdist.combine_g2_batch_sharded(E, 101, 4, xs, sh)
if rank == 0:
    exp = O.verify_batch(pk, sig, msgs)
    assert np.array_equal(ok, exp) and 0 < exp.sum() < n
    assert np.array_equal(out, master) and not st.any()
    print(f"dist NCCL check ok: world={world}, {int(exp.sum())}/{n} valid, combine over {world} ranks matches", flush=True)
dist.destroy_process_group()
