#!/usr/bin/env python
"""bench.py — BLS verifies/sec and (t+1)-share combines/sec on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # one process; N>1 via torchrun
    python bench.py --impl reference ...                   # the CPU stand-in for the `pairing` path

A "step" is one pass of the hot path over one synthetic batch:
  workload (BASELINE configs[1]): 2^16 independent PublicKey::verify (src/lib.rs:115-117:
  hash_g2 on device + pairing equality) per GPU; items with i % 16 == 5 carry a wrong signature
  so the output is not constant.  N > 1 = weak scaling: every rank verifies its own 2^16 batch,
  no data-path collective (items are independent).
  `value`  : verifies/s with inputs resident in HBM, CUDA events on the launching stream,
             L2 flushed between steps (outside the events), max over ranks.
  `e2e`    : the same metric through the reference-facing C-ABI call tcb_verify_batch with
             pinned HOST buffers (H2D + kernels + D2H inside the timed region).
  `combine`: combines/s for BASELINE configs[2] (combine_signatures, t=10, 2^14 messages).
  `roofline`: integer-MAC roofline of the dominant kernel (k_verify): algorithmic 32x32->64 MACs
             per launch / event time vs the IMAD.WIDE ceiling measured live by tcb_probe_imad,
             plus the (tiny, by design) HBM fraction vs MEASURED_PEAKS.json.
  `cpu_baseline`: the oracle (kind "port" — the Rust reference cannot be built in this image)
             on the host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VERIFY = 1 << 16
N_COMBINE = 1 << 14
T_COMBINE = 10
MSG_LEN = 32
# Algorithmic 32x32->64 MACs per item of the product algorithms, counted by the instrumented
# host emulation of the same code (tests/hostemu, see DESIGN.md "op counts").
MACS_PER_VERIFY = None      # filled from profiles/op_counts.json if present
MACS_PER_PAIRING = None
MACS_PER_HASH = None
MACS_PER_COMBINE = None


def load_op_counts():
    global MACS_PER_VERIFY, MACS_PER_COMBINE, MACS_PER_PAIRING, MACS_PER_HASH
    p = os.path.join(ROOT, "profiles", "op_counts.json")
    if os.path.exists(p):
        d = json.load(open(p))
        MACS_PER_VERIFY = d.get("verify_macs_per_item")
        MACS_PER_PAIRING = d.get("verify_g2_macs_per_item")
        MACS_PER_HASH = d.get("hash_g2_macs_per_item")
        MACS_PER_COMBINE = d.get("combine_g2_t10_macs_per_item")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """Threads for the CPU baseline: the container's CPU quota (cgroup cpu.max), not the number
    of visible CPUs — on the GPU box 128 CPUs are visible but the quota is 16."""
    n = os.cpu_count() or 1
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(round(int(q) / int(p)))))
    except Exception:
        pass
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return n


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference's own CPU path cannot be built here (Rust + absent crates), so this arm times
    the oracle running the reference's algorithms (two full pairings per verify, MSB-first
    double-and-add) on all host cores — "oracle-as-`pairing` stand-in" (BASELINE.md §2)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    import cases
    cores = host_threads()
    O.set_threads(cores)
    n = max(cores * 64, 256)        # bounded sample of the 2^16-verify workload per step (~6 ms of CPU per verify)
    # the same synthetic shape as the GPU arm: 32-byte messages, every 16th signature (i % 16 == 5) replaced by its neighbour's
    rng = np.random.default_rng(2024)
    Rr = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    sk = np.frombuffer(b"".join((int.from_bytes(rng.bytes(40), "little") % Rr).to_bytes(32, "little") for _ in range(n)), np.uint8).copy()
    msgs = [(0).to_bytes(4, "little") + i.to_bytes(8, "little") + b"\x5a" * (MSG_LEN - 12) for i in range(n)]
    pk = O.g1_mul_gen_batch(sk)
    sig = O.sign_batch(sk, msgs)
    bad = np.arange(n) % 16 == 5
    sig[bad] = np.roll(sig, 1, axis=0)[bad]
    for _ in range(max(args.warmup, 1)):
        O.verify_batch(pk, sig, msgs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ok = O.verify_batch(pk, sig, msgs)
    dt = time.perf_counter() - t0
    assert np.array_equal(ok.astype(bool), ~bad), "reference arm: unexpected verify results"
    vps = n * args.steps / dt
    # combines
    nc = max(cores, 8)
    xs, sh, master = cases.make_combine_batch(O, nc, T_COMBINE, 77, group=2, extra=21)
    t0 = time.perf_counter()
    out, st = O.combine_g2_batch(nc, T_COMBINE, xs, sh)
    dtc = time.perf_counter() - t0
    assert np.array_equal(out, master)
    line = {
        "impl": "reference", "metric": "bls_verifies_per_sec", "value": vps, "unit": "verifies/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (6x64 Montgomery, int)",
        "data": "synthetic", "config": {"workload": "PublicKey::verify (hash_g2 + pairing equality), BASELINE configs[1]", "items_per_gpu": N_VERIFY,
                                        "msg_len": MSG_LEN, "corrupted": "i % 16 == 5", "sample_items_per_step": n,
                                        "engine": "oracle/tc_oracle.c on the host cores (the reference crate needs Rust: not buildable here)"},
        "cpu_baseline": {"value": vps, "unit": "verifies/s", "cores": cores, "kind": "port",
                         "sample": f"{n} verifies/step x {args.steps} steps, oracle/tc_oracle.c with {cores} threads"},
        "combine": {"value": nc / dtc, "unit": "combines/s", "t": T_COMBINE, "sample": nc},
        "e2e": {"value": vps, "unit": "verifies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"WORLD_SIZE={world} != --gpus {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: threshold_crypto_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from threshold_crypto_b200._lib import Engine, pack_msgs
    load_op_counts()
    E = Engine(devices=[local])
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- synthetic batch, generated with the engine itself (parity with the oracle is what tests/ establish)
    n = args.items
    rng = np.random.default_rng(1000 + rank)
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    sk = np.frombuffer(b"".join((int.from_bytes(rng.bytes(40), "little") % R).to_bytes(32, "little") for _ in range(n)), np.uint8).copy()
    msgs = [(rank.to_bytes(4, "little") + i.to_bytes(8, "little")).ljust(MSG_LEN, b"\x5a") for i in range(n)]
    pk = E.g1_mul_gen_batch(sk)
    sig = E.sign_batch(sk, msgs)
    bad = np.arange(n) % 16 == 5
    sig[bad] = np.roll(sig, 1, axis=0)[bad]
    expect = (~bad).astype(np.uint8)
    mbuf, moff = pack_msgs(msgs)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t
    h_pk, h_sig, h_msg, h_off = pinned(pk.reshape(-1)), pinned(sig.reshape(-1)), pinned(mbuf), pinned(moff.view(np.int64))
    h_ok = torch.zeros(n, dtype=torch.uint8).pin_memory()
    d_pk, d_sig, d_msg, d_off = (t.to(dev) for t in (h_pk, h_sig, h_msg, h_off))
    d_ok = torch.zeros(n, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step_dev():
        E.dev_call("tcb_verify_batch_dev", stream, ("size", n), d_pk.data_ptr(), d_sig.data_ptr(), d_msg.data_ptr(),
                   d_off.data_ptr(), d_ok.data_ptr())

    def step_e2e():
        import ctypes as C
        rc = E.lib.tcb_verify_batch(E.ctx, C.c_size_t(n), C.c_void_p(h_pk.data_ptr()), C.c_void_p(h_sig.data_ptr()),
                                    C.c_void_p(h_msg.data_ptr()), C.c_void_p(h_off.data_ptr()), C.c_void_p(h_ok.data_ptr()))
        if rc:
            raise RuntimeError(E.lib.tcb_last_error(E.ctx))

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    torch.cuda.synchronize()
    assert np.array_equal(d_ok.cpu().numpy(), expect), "GPU verify output wrong"

    # ---- timed region: K steps, device events per step, L2 flushed between steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = E.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        ev[k][0].record()
        step_dev()
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = E.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    assert np.array_equal(d_ok.cpu().numpy(), expect)

    # ---- per-kernel split of the step (rank 0): the two kernels behind tcb_verify_batch timed on their own entry points
    # with the same inputs (k_hash_g2 -> H in HBM, then k_verify_g2_quad on (pk, H, g1, sig)); L2 flushed as above
    kern_ms = None
    if rank == 0:
        d_h = torch.zeros(n * 192, dtype=torch.uint8, device=dev)

        def step_hash():
            E.dev_call("tcb_hash_g2_batch_dev", stream, ("size", n), d_msg.data_ptr(), d_off.data_ptr(), d_h.data_ptr())

        def step_pair():
            E.dev_call("tcb_verify_g2_batch_dev", stream, ("size", n), d_pk.data_ptr(), d_h.data_ptr(), 0, d_sig.data_ptr(), d_ok.data_ptr())
        kern_ms = {}
        for name, fn in (("k_hash_g2", step_hash), ("k_verify_g2_quad", step_pair)):
            fn()
            torch.cuda.synchronize()
            kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for k in range(args.steps):
                flush.fill_(k & 0xff)
                kev[k][0].record(); fn(); kev[k][1].record()
            torch.cuda.synchronize()
            kern_ms[name] = sum(a.elapsed_time(b) for a, b in kev) / args.steps
        assert np.array_equal(d_ok.cpu().numpy(), expect), "split verify output wrong"

    # ---- e2e through the host-buffer C-ABI call
    for _ in range(2):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(h_ok.numpy(), expect)
    clocks = sampler.stop() if rank == 0 else None

    tot = torch.tensor([total_ms, e2e_s, wall], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, wall = (float(x) for x in tot.cpu())

    # ---- combine_signatures (config #3), N=1 semantics per rank
    comb = None
    if not args.no_combine:
        nc, t = args.combine_items, T_COMBINE
        m = t + 1
        poly = np.frombuffer(b"".join((int.from_bytes(rng.bytes(40), "little") % R).to_bytes(32, "little") for _ in range(m)), np.uint8).copy()
        # shares of item i: sk_j * H(msg_i) for j in a per-item subset of 0..31; generated on the GPU
        idx = np.stack([rng.choice(32, size=m, replace=False) for _ in range(nc)])
        idx.sort(axis=1)
        xs = np.frombuffer(b"".join(int(j + 1).to_bytes(32, "little") for j in idx.reshape(-1)), np.uint8).copy()
        # sk_j = poly(j+1) for j in 0..31 (host big-int Horner: 32 evaluations)
        coeffs = [int.from_bytes(bytes(poly[32 * k:32 * k + 32]), "little") for k in range(m)]
        skj = []
        for j in range(32):
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * (j + 1) + c) % R
            skj.append(acc.to_bytes(32, "little"))
        cm = [b"combine" + rank.to_bytes(4, "little") + i.to_bytes(8, "little") for i in range(nc)]
        hm = E.hash_g2_batch(cm)
        sk_rep = np.frombuffer(b"".join(skj[j] for j in idx.reshape(-1)), np.uint8).copy()
        shares = E.sign_g2_batch(sk_rep, np.repeat(hm, m, axis=0))
        master = E.sign_g2_batch(np.tile(poly[:32], nc), hm)
        d_x, d_sh = torch.from_numpy(xs).to(dev), torch.from_numpy(shares.reshape(-1)).to(dev)
        d_out = torch.zeros(nc * 192, dtype=torch.uint8, device=dev)
        d_st = torch.zeros(nc, dtype=torch.uint8, device=dev)

        def step_comb():
            E.dev_call("tcb_combine_g2_batch_dev", stream, ("size", nc), ("size", t), d_x.data_ptr(), d_sh.data_ptr(),
                       d_out.data_ptr(), d_st.data_ptr())
        for _ in range(3):
            step_comb()
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy().reshape(nc, 192), master), "combine output wrong"
        csteps = max(3, args.steps)
        cev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(csteps)]
        barrier()
        for k in range(csteps):
            flush.fill_(k & 0xff)
            cev[k][0].record(); step_comb(); cev[k][1].record()
        barrier()
        cms = sum(a.elapsed_time(b) for a, b in cev)
        ct = torch.tensor([cms], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        cms = float(ct.cpu()[0])
        comb = {"value": nc * world * csteps / (cms * 1e-3), "unit": "combines/s", "t": t, "items_per_gpu": nc,
                "ms_per_step": cms / csteps, "workload": "PublicKeySet::combine_signatures t=10, distinct messages"}
        if MACS_PER_COMBINE:
            comb["macs_per_item"] = MACS_PER_COMBINE

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_verify = the whole step) and the CPU baseline (rank 0, N=1 only)
    peaks, peak_kind = measured_peaks()
    imad_peak = E.probe_imad()            # MAC/s, measured live on this GPU
    fpmul_rate = E.probe_fpmul()
    per_launch_ms = total_ms / args.steps
    alg_bytes = n * (96 + 192 + MSG_LEN + 8 + 1)
    roof = {"bound": "int_mac", "unit": "GMAC/s", "peak": imad_peak / 1e9,
            "peak_source": "tcb_probe_imad: plain IMAD.WIDE.U32, 16 independent accumulators per thread, measured live",
            "achieved": None, "frac": None, "traffic": None,
            # the carry-chained form (IMAD.WIDE.U32.X, predicate carries) that a Montgomery multiply needs issues at half
            # that rate: tcb_probe_fpmul x 300 MACs is the ceiling of the CURRENT multiplier design (profiles/r1j_*)
            "carry_chain_ceiling": fpmul_rate * 300 / 1e9, "frac_of_carry_chain_ceiling": None,
            "fpmul_per_s": fpmul_rate,
            "hbm": {"achieved_gbs": alg_bytes / (per_launch_ms * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"], "peak_kind": peak_kind,
                    "frac": alg_bytes / (per_launch_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": alg_bytes}}
    if MACS_PER_VERIFY:
        ach = n * MACS_PER_VERIFY / (per_launch_ms * 1e-3)
        roof.update({"achieved": ach / 1e9, "frac": ach / imad_peak, "macs_per_item": MACS_PER_VERIFY,
                     "frac_of_carry_chain_ceiling": ach / (fpmul_rate * 300),
                     "scope": "whole step = k_hash_g2 + k_verify_g2_quad (the per-kernel numbers are in roofline.kernels)"})
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(tp)) if os.path.exists(tp) else {}
    if kern_ms:
        # the dominant kernel on its own: algorithmic MACs of one launch / its average launch duration (CUDA events above)
        ks = {}
        for name, macs in (("k_verify_g2_quad", MACS_PER_PAIRING), ("k_hash_g2", MACS_PER_HASH)):
            ms = kern_ms[name]
            k = {"ms_per_launch": ms, "share_of_step": ms / per_launch_ms, "dram_bytes_per_launch_ncu": traffic.get(name + "_dram_bytes_per_launch")}
            if macs:
                a = n * macs / (ms * 1e-3)
                k.update({"macs_per_item": macs, "achieved": a / 1e9, "frac": a / imad_peak, "frac_of_carry_chain_ceiling": a / (fpmul_rate * 300)})
            ks[name] = k
        roof["kernels"] = ks
        roof["dominant_kernel"] = "k_verify_g2_quad"
        roof["traffic"] = traffic.get("k_verify_g2_quad_dram_bytes_per_launch")

    cpu = None
    if world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        cores = host_threads()
        O.set_threads(cores)
        ns = max(cores * 32, 128)
        t0 = time.perf_counter()
        ok = O.verify_batch(pk[:ns], sig[:ns], msgs[:ns])
        dt = time.perf_counter() - t0
        assert np.array_equal(ok, expect[:ns]), "oracle disagrees with the GPU output"
        cpu = {"value": ns / dt, "unit": "verifies/s", "cores": cores, "kind": "port",
               "sample": f"first {ns} items of the same batch, oracle/tc_oracle.c (reference algorithms), {cores} threads, {dt:.1f} s"}

    value = n * world * args.steps / (total_ms * 1e-3)
    line = {
        "metric": "bls_verifies_per_sec", "value": value, "unit": "verifies/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (12x32 Montgomery, IMAD.WIDE integer)", "data": "synthetic",
        "config": {"workload": "PublicKey::verify (hash_g2 + pairing equality), BASELINE configs[1]", "items_per_gpu": n,
                   "msg_len": MSG_LEN, "corrupted": "i % 16 == 5", "l2": "flushed between steps (256 MiB fill, outside the events)",
                   "engine": "quad (pairing) + lane-pair (hash_g2)"},
        "e2e": {"value": n * world * e2e_steps / e2e_s, "unit": "verifies/s", "h2d_bytes_per_step": int(h_pk.numel() + h_sig.numel() + h_msg.numel() + 8 * h_off.numel()),
                "d2h_bytes_per_step": int(n), "steps": e2e_steps, "api": "tcb_verify_batch (host buffers, pinned)"},
        "gpu_launches": int(launches), "wall_s_timed_region": wall, "clocks": clocks,
        "roofline": roof, "cpu_baseline": cpu, "combine": comb,
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tcb200", choices=["tcb200", "reference"])
    ap.add_argument("--items", type=int, default=N_VERIFY)
    ap.add_argument("--combine-items", type=int, default=N_COMBINE)
    ap.add_argument("--no-combine", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
