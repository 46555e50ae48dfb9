"""Multi-GPU data plane of the hot path: every item is independent, so a batch held by ONE
rank (the root: pinned host memory) is split into contiguous slices and processed on all ranks' GPUs:

    root H2D  ->  one group of NCCL send/recv over NVLink (scatter)  ->  tcb_*_batch_dev on every rank
              ->  one group of send/recv (gather)  ->  root D2H

Everything between the two host copies stays in HBM: the received slices are device tensors whose pointers go
straight into the `_dev` entry points of the C ABI (no host bounce).  The commitment table of
Commitment::evaluate is broadcast once (96 KB).  There is no exchange step inside the computation, hence no
all-reduce.  With the gloo backend (CPU tests: tests/test_dist_gloo.py) the "device" is the host and the engine is the
host-buffer API of whatever library the caller passes (the tests pass the host-emulation build) — same code path
for the partitioning, the offsets and the collectives.

Each `*_sharded` call returns the result on the root (None elsewhere) and, in `.last_timing`, the milliseconds
of its phases (h2d, scatter, compute, gather, d2h) measured with CUDA events on the current stream.
"""
import ctypes as C
import time

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    return n * rank // world, n * (rank + 1) // world


class _Phases:
    """Phase timer: CUDA events on the current stream (nccl) or perf_counter (gloo)."""

    def __init__(self, cuda):
        self.cuda, self.marks = cuda, []
        self.mark("start")

    def mark(self, name):
        if self.cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))
        else:
            self.marks.append((name, time.perf_counter()))

    def result(self):
        if self.cuda:
            torch.cuda.synchronize()
            out = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(self.marks, self.marks[1:])}
        else:
            out = {b[0]: 1e3 * (b[1] - a[1]) for a, b in zip(self.marks, self.marks[1:])}
        out["total"] = sum(out.values())
        return out


class ShardedEngine:
    """Scatter -> compute -> gather around one `Engine` per rank.  `group` defaults to the world."""

    def __init__(self, engine, root=0, group=None):
        self.E, self.root, self.group = engine, root, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.cuda = dist.get_backend(group) == "nccl"
        self.dev = torch.device("cuda", torch.cuda.current_device()) if self.cuda else torch.device("cpu")
        self.last_timing = None
        if self.cuda and not hasattr(engine.lib, "tcb_verify_batch_dev"):
            raise RuntimeError("the NCCL data plane needs the CUDA library (device entry points)")

    # ---- plumbing
    def _is_root(self):
        return self.rank == self.root

    def _to_dev(self, host):
        """root: host numpy array -> flat uint8 tensor on the compute device (async H2D on the current stream)"""
        if isinstance(host, torch.Tensor):
            t = host.contiguous().view(torch.uint8).reshape(-1)
        else:
            t = torch.from_numpy(np.ascontiguousarray(host).view(np.uint8).reshape(-1))
        if not self.cuda:
            return t
        return t.to(self.dev, non_blocking=True)

    def _scatter(self, specs, n):
        """specs: list of (root device tensor or None, bytes per item).  Every rank gets its contiguous rows
        [lo, hi) of each array as a device tensor.  ONE batch of point-to-point operations for all arrays."""
        lo, hi = shard_bounds(n, self.rank, self.world)
        ops, outs = [], []
        for k, (src, width) in enumerate(specs):
            if self._is_root():
                outs.append(src[lo * width: hi * width])
                for r in range(self.world):
                    if r == self.root:
                        continue
                    rl, rh = shard_bounds(n, r, self.world)
                    if rh > rl:
                        ops.append(dist.P2POp(dist.isend, src[rl * width: rh * width], self._global(r), self.group))
            else:
                buf = torch.empty((hi - lo) * width, dtype=torch.uint8, device=self.dev)
                outs.append(buf)
                if hi > lo:
                    ops.append(dist.P2POp(dist.irecv, buf, self._global(self.root), self.group))
        self._run(ops)
        return outs

    def _gather(self, specs, n):
        """specs: list of (this rank's device tensor, bytes per item).  Root returns full device tensors."""
        lo, hi = shard_bounds(n, self.rank, self.world)
        ops, outs = [], []
        for part, width in specs:
            if self._is_root():
                full = torch.empty(n * width, dtype=torch.uint8, device=self.dev)
                full[lo * width: hi * width] = part[: (hi - lo) * width]
                outs.append(full)
                for r in range(self.world):
                    if r == self.root:
                        continue
                    rl, rh = shard_bounds(n, r, self.world)
                    if rh > rl:
                        ops.append(dist.P2POp(dist.irecv, full[rl * width: rh * width], self._global(r), self.group))
            else:
                outs.append(None)
                if hi > lo:
                    ops.append(dist.P2POp(dist.isend, part[: (hi - lo) * width], self._global(self.root), self.group))
        self._run(ops)
        return outs

    def _global(self, r):
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def _run(self, ops):
        if not ops:
            return
        for w in dist.batch_isend_irecv(ops):
            w.wait()       # nccl: the current stream waits for the transfer; gloo: blocks

    def _bcast(self, t):
        dist.broadcast(t, self._global(self.root), group=self.group)
        return t

    def _ptr(self, t):
        return t.data_ptr() if t is not None and t.numel() else 0

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream if self.cuda else 0

    def _call(self, dev_name, host_name, args):
        """args: list of tensors / ("size", v) / None.  nccl: the `_dev` entry point on the current stream;
        gloo: the host-buffer entry point of the same name on the CPU tensors' memory."""
        E = self.E
        if self.cuda:
            E.dev_call(dev_name, self._stream(), *[a if isinstance(a, tuple) or a is None else self._ptr(a) for a in args])
            return
        cargs = [E.ctx]
        for a in args:
            if isinstance(a, tuple):
                if a[0] == "size":
                    cargs.append(C.c_size_t(a[1]))
                # ("u64", total bytes) only exists in the _dev signature of decrypt
            elif a is None:
                cargs.append(None)
            else:
                cargs.append(C.c_void_p(a.data_ptr()))
        E._ck(getattr(E.lib, host_name)(*cargs))

    def _scatter_ragged(self, n, bufs, offs):
        """Ragged byte strings (messages / ciphertext bodies): root holds the concatenation `bufs` (uint8) and the
        n+1 uint64 offsets; every rank gets its slice of the bytes and offsets rebased to its slice."""
        lo, hi = shard_bounds(n, self.rank, self.world)
        sizes = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        if self._is_root():
            o = np.asarray(offs, np.uint64)
            per = [int(o[shard_bounds(n, r, self.world)[1]] - o[shard_bounds(n, r, self.world)[0]]) for r in range(self.world)]
            sizes = torch.tensor(per, dtype=torch.int64).to(self.dev)
        self._bcast(sizes)
        sizes = [int(x) for x in sizes.cpu()]
        ops = []
        if self._is_root():
            o = np.asarray(offs, np.uint64)
            d_buf = self._to_dev(np.asarray(bufs, np.uint8))
            # rebased offsets of every rank, laid out back to back: rank r's block has (cnt_r + 1) entries
            blocks = []
            for r in range(self.world):
                rl, rh = shard_bounds(n, r, self.world)
                blocks.append((o[rl: rh + 1] - o[rl]).astype(np.uint64))
            d_offs = [self._to_dev(b) for b in blocks]
            my_buf = d_buf[int(o[lo]): int(o[hi])]
            my_off = d_offs[self.root]
            for r in range(self.world):
                if r == self.root:
                    continue
                rl, rh = shard_bounds(n, r, self.world)
                if rh > rl:
                    ops.append(dist.P2POp(dist.isend, d_offs[r], self._global(r), self.group))
                    if sizes[r]:
                        ops.append(dist.P2POp(dist.isend, d_buf[int(o[rl]): int(o[rh])], self._global(r), self.group))
        else:
            my_off = torch.zeros((hi - lo + 1) * 8, dtype=torch.uint8, device=self.dev)
            my_buf = torch.empty(max(sizes[self.rank], 1), dtype=torch.uint8, device=self.dev)
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, my_off, self._global(self.root), self.group))
                if sizes[self.rank]:
                    ops.append(dist.P2POp(dist.irecv, my_buf, self._global(self.root), self.group))
        self._run(ops)
        if my_buf.numel() == 0:
            my_buf = torch.zeros(1, dtype=torch.uint8, device=self.dev)
        return my_buf, my_off, sizes

    def _host(self, t):
        """root: device tensor -> numpy (the D2H copy of the result)"""
        return t.cpu().numpy() if self.cuda else t.numpy()

    # ---- the sharded operations (root passes the arrays, the other ranks pass None)
    def verify_batch(self, n, pk=None, sig=None, msgs=None):
        """PublicKey::verify (src/lib.rs:115-117) over a batch held by the root.  `msgs` is a list of byte strings of
        any lengths (scattered as bytes + offsets, the format tcb_verify_batch uses) or a (buffer, offsets) pair."""
        ph = _Phases(self.cuda)
        d_pk = d_sig = mbuf = moff = None
        if self._is_root():
            if pk is None or sig is None or msgs is None:
                raise ValueError("the root rank must pass pk, sig and msgs")
            if isinstance(msgs, tuple):
                mbuf, moff = msgs
            else:
                from ._lib import pack_msgs
                mbuf, moff = pack_msgs(msgs)
            pk, sig = np.ascontiguousarray(pk, np.uint8).reshape(-1), np.ascontiguousarray(sig, np.uint8).reshape(-1)
            if pk.size != 96 * n or sig.size != 192 * n or len(moff) != n + 1:
                raise ValueError("verify_batch: array sizes do not match n")
            d_pk, d_sig = self._to_dev(pk), self._to_dev(sig)
        ph.mark("h2d")
        my_pk, my_sig = self._scatter([(d_pk, 96), (d_sig, 192)], n)
        my_msg, my_off, _ = self._scatter_ragged(n, mbuf, moff)
        ph.mark("scatter")
        lo, hi = shard_bounds(n, self.rank, self.world)
        cnt = hi - lo
        ok = torch.zeros(max(cnt, 1), dtype=torch.uint8, device=self.dev)
        if cnt:
            self._call("tcb_verify_batch_dev", "tcb_verify_batch", [("size", cnt), my_pk, my_sig, my_msg, my_off, ok])
        ph.mark("compute")
        (full,) = self._gather([(ok, 1)], n)
        ph.mark("gather")
        res = self._host(full) if self._is_root() else None
        ph.mark("d2h")
        self.last_timing = ph.result()
        return res

    def combine_g2_batch(self, n, t, x_fr=None, shares=None):
        """PublicKeySet::combine_signatures (src/lib.rs:608-615): n items of t+1 (x, share) pairs held by the root."""
        return self._combine(n, t, x_fr, shares, 192, "tcb_combine_g2_batch")

    def combine_g1_batch(self, n, t, x_fr=None, shares=None):
        return self._combine(n, t, x_fr, shares, 96, "tcb_combine_g1_batch")

    def _combine(self, n, t, x_fr, shares, pw, name):
        m = t + 1
        ph = _Phases(self.cuda)
        d_x = d_s = None
        if self._is_root():
            if x_fr is None or shares is None:
                raise ValueError("the root rank must pass x_fr and shares")
            x_fr, shares = np.ascontiguousarray(x_fr, np.uint8).reshape(-1), np.ascontiguousarray(shares, np.uint8).reshape(-1)
            if x_fr.size != n * m * 32 or shares.size != n * m * pw:
                raise ValueError("combine: array sizes do not match n and t")
            d_x, d_s = self._to_dev(x_fr), self._to_dev(shares)
        ph.mark("h2d")
        my_x, my_s = self._scatter([(d_x, 32 * m), (d_s, pw * m)], n)
        ph.mark("scatter")
        lo, hi = shard_bounds(n, self.rank, self.world)
        cnt = hi - lo
        out = torch.zeros(max(cnt, 1) * pw, dtype=torch.uint8, device=self.dev)
        st = torch.zeros(max(cnt, 1), dtype=torch.uint8, device=self.dev)
        if cnt:
            self._call(name + "_dev", name, [("size", cnt), ("size", t), my_x, my_s, out, st])
        ph.mark("compute")
        full, fst = self._gather([(out, pw), (st, 1)], n)
        ph.mark("gather")
        res = (self._host(full).reshape(n, pw), self._host(fst)) if self._is_root() else (None, None)
        ph.mark("d2h")
        self.last_timing = ph.result()
        return res

    def decrypt_batch(self, n, t, x_fr=None, shares=None, vs=None):
        """PublicKeySet::decrypt (src/lib.rs:618-626): returns (list of plaintext byte strings, status) on the root."""
        m = t + 1
        ph = _Phases(self.cuda)
        d_x = d_s = vbuf = voff = None
        if self._is_root():
            from ._lib import pack_msgs
            vbuf, voff = vs if isinstance(vs, tuple) else pack_msgs(vs)
            x_fr, shares = np.ascontiguousarray(x_fr, np.uint8).reshape(-1), np.ascontiguousarray(shares, np.uint8).reshape(-1)
            if x_fr.size != n * m * 32 or shares.size != n * m * 96 or len(voff) != n + 1:
                raise ValueError("decrypt: array sizes do not match n and t")
            d_x, d_s = self._to_dev(x_fr), self._to_dev(shares)
        ph.mark("h2d")
        my_x, my_s = self._scatter([(d_x, 32 * m), (d_s, 96 * m)], n)
        my_v, my_off, sizes = self._scatter_ragged(n, vbuf, voff)
        ph.mark("scatter")
        lo, hi = shard_bounds(n, self.rank, self.world)
        cnt = hi - lo
        out = torch.zeros(max(sizes[self.rank], 1), dtype=torch.uint8, device=self.dev)
        st = torch.zeros(max(cnt, 1), dtype=torch.uint8, device=self.dev)
        if cnt:
            self._call("tcb_decrypt_batch_dev", "tcb_decrypt_batch",
                       [("size", cnt), ("size", t), my_x, my_s, my_v, my_off, ("u64", sizes[self.rank]), out, st])
        ph.mark("compute")
        (fst,) = self._gather([(st, 1)], n)
        # plaintext bytes: ragged gather (sizes are known to every rank)
        ops, full = [], None
        if self._is_root():
            starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            full = torch.empty(max(int(starts[-1]), 1), dtype=torch.uint8, device=self.dev)
            full[int(starts[self.root]): int(starts[self.root + 1])] = out[: sizes[self.root]]
            for r in range(self.world):
                if r != self.root and sizes[r]:
                    ops.append(dist.P2POp(dist.irecv, full[int(starts[r]): int(starts[r + 1])], self._global(r), self.group))
        elif sizes[self.rank]:
            ops.append(dist.P2POp(dist.isend, out[: sizes[self.rank]], self._global(self.root), self.group))
        self._run(ops)
        ph.mark("gather")
        res = (None, None)
        if self._is_root():
            flat = self._host(full)
            o = np.asarray(voff, np.uint64)
            res = ([bytes(flat[int(o[i]): int(o[i + 1])]) for i in range(n)], self._host(fst))
        ph.mark("d2h")
        self.last_timing = ph.result()
        return res

    def commitment_eval_batch(self, n, deg, coeff_g1=None, x_fr=None):
        """Commitment::evaluate (src/poly.rs:497-508) of ONE commitment of degree `deg` (every rank passes n and deg) at the
        root's n points: the (deg+1) x 96 B table is broadcast, the points are scattered, the results gathered."""
        ph = _Phases(self.cuda)
        d_x = None
        d_c = torch.empty(96 * (deg + 1), dtype=torch.uint8, device=self.dev)
        # The cost of a point grows with the bit length of x (Horner skips leading zeros) and callers pass increasing indices
        # (public_key_share(i) -> x = i + 1), so contiguous slices would leave the last rank with the most expensive points: the
        # points are dealt round-robin instead (rank r gets i = r, r + world, ...), a host-side permutation of 32 B per point.
        perm = np.concatenate([np.arange(r, n, self.world) for r in range(self.world)]) if self.world > 1 else None
        if self._is_root():
            coeff_g1, x_fr = np.ascontiguousarray(coeff_g1, np.uint8).reshape(-1), np.ascontiguousarray(x_fr, np.uint8).reshape(-1)
            if coeff_g1.size != 96 * (deg + 1) or x_fr.size != 32 * n:
                raise ValueError("commitment_eval: array sizes do not match n and deg")
            if perm is not None:
                x_fr = torch.from_numpy(x_fr.reshape(n, 32)[perm].reshape(-1))
                if self.cuda:
                    x_fr = x_fr.pin_memory()
            d_c, d_x = self._to_dev(coeff_g1), self._to_dev(x_fr)
        ph.mark("h2d")
        self._bcast(d_c)
        (my_x,) = self._scatter([(d_x, 32)], n)
        ph.mark("scatter")
        lo, hi = shard_bounds(n, self.rank, self.world)
        cnt = hi - lo
        out = torch.zeros(max(cnt, 1) * 96, dtype=torch.uint8, device=self.dev)
        if cnt:
            if self.cuda:
                self._call("tcb_commitment_eval_batch_dev", None, [("size", deg), d_c, ("size", cnt), my_x, out])
            else:
                self._call(None, "tcb_commitment_eval_batch", [("size", deg), d_c, ("size", cnt), my_x, out])
        ph.mark("compute")
        (full,) = self._gather([(out, 96)], n)
        ph.mark("gather")
        res = None
        if self._is_root():
            res = self._host(full).reshape(n, 96)
            if perm is not None:
                unperm = np.empty_like(res)
                unperm[perm] = res
                res = unperm
        ph.mark("d2h")
        self.last_timing = ph.result()
        return res


# ---- functional wrappers kept from round 1 (tests/test_dist_gloo.py, tools/dist_nccl_check.py)
def verify_batch_sharded(engine, n, pk=None, sig=None, msgs=None, root=0):
    return ShardedEngine(engine, root).verify_batch(n, pk, sig, msgs)


def combine_g2_batch_sharded(engine, n, t, x_fr=None, shares=None, root=0):
    return ShardedEngine(engine, root).combine_g2_batch(n, t, x_fr, shares)
