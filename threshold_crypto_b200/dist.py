"""Multi-GPU plumbing for the hot path: every item is independent, so a batch
held by one rank is split into contiguous slices — ONE scatter of the inputs and ONE gather of
the results over torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
There is no data-path collective inside the computation itself; bench.py therefore measures
weak scaling with rank-local batches and uses no collective at all.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    return n * rank // world, n * (rank + 1) // world


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _scatter_rows(arr, width, n, root):
    """Scatter the rows of a (n, width) uint8 array held by `root`; returns this rank's rows."""
    rank, world = dist.get_rank(), dist.get_world_size()
    per = (n + world - 1) // world
    dev = _device()
    out = torch.zeros(per * width, dtype=torch.uint8, device=dev)
    chunks = None
    if rank == root:
        a = np.zeros((per * world, width), np.uint8)
        flat = np.ascontiguousarray(arr, dtype=np.uint8).reshape(n, width)
        for r in range(world):
            lo, hi = shard_bounds(n, r, world)
            a[r * per: r * per + (hi - lo)] = flat[lo:hi]
        chunks = [torch.from_numpy(a[r * per:(r + 1) * per].reshape(-1)).to(dev) for r in range(world)]
    dist.scatter(out, chunks, src=root)
    lo, hi = shard_bounds(n, rank, world)
    return out.cpu().numpy().reshape(per, width)[: hi - lo]


def _gather_rows(local, width, n, root):
    rank, world = dist.get_rank(), dist.get_world_size()
    per = (n + world - 1) // world
    dev = _device()
    buf = np.zeros((per, width), np.uint8)
    buf[: local.shape[0]] = np.asarray(local, np.uint8).reshape(-1, width)
    t = torch.from_numpy(buf.reshape(-1)).to(dev)
    outs = [torch.zeros_like(t) for _ in range(world)] if rank == root else None
    dist.gather(t, outs, dst=root)
    if rank != root:
        return None
    res = np.zeros((n, width), np.uint8)
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        res[lo:hi] = outs[r].cpu().numpy().reshape(per, width)[: hi - lo]
    return res


def verify_batch_sharded(engine, n, pk=None, sig=None, msgs=None, msg_len=32, root=0):
    """PublicKey::verify over a batch held by `root` (fixed-length messages of msg_len bytes):
    scatter -> each rank verifies its slice on its own device -> gather on root."""
    m = None
    if dist.get_rank() == root:
        m = np.frombuffer(b"".join(bytes(x).ljust(msg_len, b"\0")[:msg_len] for x in msgs), np.uint8).reshape(n, msg_len)
    my_pk = _scatter_rows(pk, 96, n, root)
    my_sig = _scatter_rows(sig, 192, n, root)
    my_msg = _scatter_rows(m, msg_len, n, root)
    ok = engine.verify_batch(my_pk, my_sig, [bytes(r) for r in my_msg]) if len(my_pk) else np.zeros(0, np.uint8)
    res = _gather_rows(ok.reshape(-1, 1), 1, n, root)
    return None if res is None else res.reshape(-1)


def combine_g2_batch_sharded(engine, n, t, x_fr=None, shares=None, root=0):
    m = t + 1
    my_x = _scatter_rows(x_fr, 32 * m, n, root)
    my_s = _scatter_rows(shares, 192 * m, n, root)
    if len(my_x):
        out, st = engine.combine_g2_batch(len(my_x), t, my_x.reshape(-1), my_s.reshape(-1))
    else:
        out, st = np.zeros((0, 192), np.uint8), np.zeros(0, np.uint8)
    res = _gather_rows(out, 192, n, root)
    stat = _gather_rows(st.reshape(-1, 1), 1, n, root)
    return (None, None) if res is None else (res, stat.reshape(-1))
