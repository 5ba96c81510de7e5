#!/usr/bin/env python
"""bench.py -- headline benchmark of the hexl-fpga B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): NTT/s at N=16384.  Workload = BASELINE configs[1]:
forward + inverse negacyclic NTT, N=16384, one 52-bit prime, batch 4096
polynomials per GPU (512 MiB, four times the 126 MB L2, so every step streams
from HBM).  One step = one forward NTT launch + one inverse NTT launch over the
whole batch = 8192 transforms.  With N GPUs every rank runs its own batch
(independent polynomials, no data-path collective -> weak scaling); NCCL is
used only to replicate the twiddle tables from rank 0 and to agree on the
time (max over ranks).

Also reported on the same line (N=1: measured after the headline region):
keyswitch (configs[3]: N=16384, decomp 7, key 8, batch 1024) and dyadic
multiply (configs[2]) device-resident throughput with their own roofline
fractions, the end-to-end numbers through the reference's host-pointer API,
and a CPU baseline timed on the host cores.

--impl reference times the reference's own CPU code for this path (its scalar
NTT, tests/test_utils/ntt.cpp compiled unmodified into oracle/_ref; intel-hexl
itself is an unvendored dependency) on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

N = 16384
Q52 = 2251799814045697            # GeneratePrimes(1, 51, 16384)[0], a true 52-bit prime
BATCH = 4096
NTT_BYTES = 2 * N * 8             # algorithmic HBM bytes per transform (read + write)
WORKLOAD = ("fwd+inv negacyclic NTT, N=16384, one 52-bit prime, batch 4096 per GPU "
            "(BASELINE configs[1]); step = 1 fwd + 1 inv pass over the batch = 8192 transforms per GPU")
KS_D, KS_K, KS_BATCH = 7, 8, 1024
KS_SHARD = 4096                   # configs[4]: 32768 items over 8 GPUs
KS_BYTES = (KS_D + 2 * 2 * KS_D) * N * 8          # t_target + result read + result write
DY_N, DY_M, DY_BATCH = 8192, 4, 8192
DY_BYTES = (2 * 2 + 3) * DY_M * DY_N * 8


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """per-launch DRAM bytes of `kernel` from the committed ncu summary, or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is
    polled from a thread every ~2 ms (the timed region is tens of milliseconds,
    far below nvidia-smi's loop granularity); nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.nvml, self.h = index, [], False, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        bits = [("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)]
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.time(), mhz, [nm for nm, b in bits if r & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            f = [x.strip() for x in out.strip().split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return float(f[0]), float(f[1]), [nm for nm, v in zip(names, f[3:7]) if v.lower().startswith("active")]
        except Exception:
            return None

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.nvml is None or not self.rows:
            one = self._smi_once()
            if one is None:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock query unavailable"]}
            return {"sm_mhz": one[0], "sm_max_mhz": one[1], "samples": 1, "reasons": one[2],
                    "source": "nvidia-smi right after the timed region"}
        self.t.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        reasons = sorted({x for r in rows for x in r[2]})
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max, "samples": len(rows),
                "reasons": reasons, "source": "NVML polled every ~2 ms inside the timed region"}


# --------------------------------------------------------------------------
# reference arm / CPU baseline
# --------------------------------------------------------------------------
def cpu_ntt_rate(polys, threads, reps=1):
    """fwd + inv NTT over `polys` polynomials on `threads` host threads.
    Uses oracle/_ref (the reference's own scalar NTT) when it was built, else
    the oracle port.  Returns (transforms/s, kind)."""
    import oracle_binding as ob

    t = ob.Tables(N, Q52)
    a = np.stack([ob.splitmix(N, 1234 + i, Q52) for i in range(min(polys, 64))])
    a = np.ascontiguousarray(np.resize(a, (polys, N)))
    r = ob.ref()
    best = None
    for _ in range(reps + 1):          # first pass warms caches / thread pool
        t0 = time.perf_counter()
        if r is not None:
            r.ref_fwd_ntt_batch(ob.P(a), polys, N, Q52, ob.P(t.roots), ob.P(t.precon), threads)
            r.ref_inv_ntt_batch(ob.P(a), polys, N, Q52, ob.P(t.inv_roots), ob.P(t.precon_inv), threads)
        else:
            o = ob.oracle()
            o.ho_fwd_ntt_batch(ob.P(a), polys, N, Q52, ob.P(t.roots), ob.P(t.precon), threads)
            o.ho_inv_ntt_batch(ob.P(a), polys, N, Q52, ob.P(t.inv_roots), ob.P(t.precon_inv), t.inv_n,
                               t.inv_n_w, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return 2 * polys / best, ("reference" if r is not None else "port")


def host_threads():
    """Host cores this process may use.  Deliberately NOT omp_get_max_threads(): torchrun
    exports OMP_NUM_THREADS=1, and the CPU arms pass their thread count explicitly
    (num_threads clause), so the baseline keeps all cores under torchrun too."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args, rank):
    if rank != 0:
        return
    threads = host_threads()
    polys = BATCH                                          # the same batch per step as the GPU arm
    import oracle_binding as ob

    t = ob.Tables(N, Q52)
    a = np.ascontiguousarray(np.resize(np.stack([ob.splitmix(N, 1234 + i, Q52) for i in range(16)]), (polys, N)))
    r = ob.ref()
    kind = "reference" if r is not None else "port"

    def step():
        if r is not None:
            r.ref_fwd_ntt_batch(ob.P(a), polys, N, Q52, ob.P(t.roots), ob.P(t.precon), threads)
            r.ref_inv_ntt_batch(ob.P(a), polys, N, Q52, ob.P(t.inv_roots), ob.P(t.precon_inv), threads)
        else:
            o = ob.oracle()
            o.ho_fwd_ntt_batch(ob.P(a), polys, N, Q52, ob.P(t.roots), ob.P(t.precon), threads)
            o.ho_inv_ntt_batch(ob.P(a), polys, N, Q52, ob.P(t.inv_roots), ob.P(t.precon_inv), t.inv_n,
                               t.inv_n_w, threads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = 2 * polys * args.steps / dt
    sample = f"{polys} polynomials fwd+inv per step ({threads} threads), scalar Harvey NTT"
    print(json.dumps({
        "impl": "reference", "metric": "NTT/s (N=16384)", "value": value, "unit": "NTT/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": N, "modulus": Q52, "batch_per_gpu": BATCH},
        "cpu_baseline": {"value": value, "unit": "NTT/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "NTT/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU path: tests/test_utils/ntt.cpp (intel-hexl's scalar NTT as vendored by the "
                "reference) compiled unmodified into oracle/_ref; intel-hexl AVX-512 itself is unvendored"}))


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def gpu_tensor(a, dev):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(dev)


def event_pair():
    import torch

    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _claim_stdout():
    """Route everything that prints to fd 1 (NCCL's version banner, library chatter) to
    stderr and return a file object on the real stdout for the one JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    json_out = _claim_stdout()

    import hexl_b200 as hb
    import oracle_binding as ob

    hb.lib()                                   # fails loudly if the CUDA library is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = peaks()

    # twiddle tables: rank 0 computes, NCCL broadcast replicates (the only collective)
    tab_np = None
    if rank == 0:
        t = ob.Tables(N, Q52)
        tab_np = np.stack([t.roots, t.precon, t.inv_roots, t.precon_inv])
    tabs = gpu_tensor(tab_np, dev) if rank == 0 else torch.empty((4, N), dtype=torch.int64, device=dev)
    scal = torch.tensor([t.inv_n, t.inv_n_w] if rank == 0 else [0, 0], dtype=torch.int64, device=dev)
    from sharding import replicate

    replicate([tabs, scal])        # NCCL broadcast over NVLink when world > 1
    roots, precon, inv_roots, precon_inv = tabs[0], tabs[1], tabs[2], tabs[3]
    inv_n, inv_n_w = int(scal[0]), int(scal[1])

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randint(0, Q52, (BATCH, N), dtype=torch.int64, device=dev, generator=g)
    x0 = x[:2].clone()

    def step(timing=None):
        if timing is not None:
            e0, e1 = event_pair()
            e0.record()
        hb.ntt_fwd(x, roots, precon, Q52, N)
        if timing is not None:
            e1.record()
            timing.append((e0, e1))
        hb.ntt_inv(x, inv_roots, precon_inv, Q52, inv_n, inv_n_w, N)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    hb.reset_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    fwd_events = []
    s0, s1 = event_pair()
    barrier()
    w0 = time.time()
    s0.record()
    for _ in range(args.steps):
        step(fwd_events)
    s1.record()
    barrier()
    w1 = time.time()
    elapsed = s0.elapsed_time(s1) * 1e-3
    launches = hb.get_stats()["kernel_launches"]
    clocks = sampler.stop(w0, w1) if sampler else None
    if world > 1:
        tt = torch.tensor([elapsed], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt[0])
        ll = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(ll)
        launches = int(ll[0])
    # the data went through steps x (fwd, inv): must be back where it started
    assert torch.equal(x[:2], x0), "fwd/inv round trip changed the data"
    value = world * 2 * BATCH * args.steps / elapsed
    call_s = float(np.mean([a.elapsed_time(b) for a, b in fwd_events])) * 1e-3   # pack + transform + deferred-list pass
    # roofline leg: the same steps once more with the library recording CUDA events on its stream directly around
    # every kernel launch (option "time_kernels"), which gives the forward kernel's own duration; the bracket
    # around the whole call above also holds the twiddle pack and the (empty) deferred-list pass
    hb.set_option("time_kernels", 1)
    fwd_k, inv_k = [], []
    try:
        hb.kernel_times()
        for _ in range(args.steps):
            hb.ntt_fwd(x, roots, precon, Q52, N)
            fwd_k.append(float(hb.kernel_times().max()))
            hb.ntt_inv(x, inv_roots, precon_inv, Q52, inv_n, inv_n_w, N)
            inv_k.append(float(hb.kernel_times().max()))
    finally:
        hb.set_option("time_kernels", 0)
    assert torch.equal(x[:2], x0), "fwd/inv round trip changed the data"
    fwd_s = float(np.mean(fwd_k)) * 1e-3
    inv_s = float(np.mean(inv_k)) * 1e-3
    achieved = BATCH * NTT_BYTES / fwd_s / 1e9

    line = {
        "metric": "NTT/s (N=16384)", "value": value, "unit": "NTT/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "n": N, "modulus": Q52, "batch_per_gpu": BATCH, "sharding": f"batch x{world}, no collective",
                   "l2": "working set 512 MiB per GPU > 126 MB L2 (no flush needed)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_ntt_fwd<.., FP64> (forward NTT, one CTA per polynomial, butterflies on the FP64 pipe)",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "peak_source": peak_src, "traffic": ncu_traffic("ntt_fwd"),
                     "traffic_source": "committed ncu --set full capture of this kernel (profiles/ncu_traffic.json), "
                                       "not measured in this run",
                     "algorithmic_bytes_per_launch": BATCH * NTT_BYTES, "launch_s": fwd_s,
                     "launch_s_source": "CUDA events recorded by the library on its stream directly around the forward "
                                        "kernel's launch (option time_kernels), mean over `steps` launches right "
                                        "after the timed region, same data",
                     "call_s": call_s,
                     "call_s_note": "events around the whole hexl_b200_ntt_fwd call inside the timed region: twiddle "
                                    "pack + forward kernel + deferred-list pass (3 launches)",
                     "inverse_kernel": {"launch_s": inv_s, "achieved": BATCH * NTT_BYTES / inv_s / 1e9,
                                        "frac": BATCH * NTT_BYTES / inv_s / 1e9 / hbm_peak}},
        "clocks": clocks,
        "kernel_variant": "persistent TMA-fed CTAs, 32 words/thread, FP64-pipe butterflies on centred integer-valued doubles (bit-exact), head twiddles in shared / tail twiddles in tensor memory, range vote with out-of-contract polynomials deferred to the exact kernel",
    }
    del x
    torch.cuda.empty_cache()
    # ---- BASELINE configs[4]: the keyswitch batch sharded over the ranks (every rank) ----
    ks_sh = keyswitch_sharded(hb, ob, dev, rank, world, barrier)

    def rank_max(v):
        if world == 1:
            return v
        tt = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0])

    ks_s = rank_max(ks_sh["s"])
    line["keyswitch_sharded"] = {
        "metric": "KeySwitch/s (N=16384, decomp=7, key=8)", "value": world * KS_SHARD * ks_sh["steps"] / ks_s,
        "unit": "KeySwitch/s", "n_gpus": world, "total_batch": world * KS_SHARD, "batch_per_gpu": KS_SHARD,
        "steps": ks_sh["steps"], "ms_per_step": ks_s / ks_sh["steps"] * 1e3, "scaling": "weak",
        "burst": {"value": world * KS_SHARD / ks_sh["first_steps_s"], "unit": "KeySwitch/s",
                  "note": "rank 0's first two steps (~70 ms): 340 ms of this FP64-heavy work run into the board's "
                          "power cap (sw_power_cap, SM clock 1965 -> ~1780 MHz), see clocks"},
        "clocks": ks_sh["clocks"],
        "sharding": "contiguous shards (sharding.shard), keys + tables replicated, no data-path collective",
        "checked": "first and last item of every rank's shard bit-exact vs the oracle",
        "workload": "BASELINE configs[4]: 32768 items over 8 GPUs = 4096 per GPU; the same shard size at every N",
        "roofline": {"bound": "hbm", "achieved": KS_SHARD * ks_sh["steps"] * KS_BYTES / ks_s / 1e9, "peak": hbm_peak,
                     "unit": "GB/s per GPU", "frac": KS_SHARD * ks_sh["steps"] * KS_BYTES / ks_s / 1e9 / hbm_peak}}
    if rank == 0:
        line.update(extras(args, hb, ob, dev, hbm_peak, world))
    # ---- end-to-end through the reference's host-pointer API (every rank) ----
    ceiling = rank_max(copy_ceiling(dev))
    e2e = {}
    for kind in ("pinned", "pageable", "registered"):
        r = e2e_ntt(args, hb, ob, dev, world, kind)
        r["s"] = rank_max(r["s"])
        e2e[kind] = r
    pin, pag, reg = e2e["pinned"], e2e["pageable"], e2e["registered"]
    step_bytes = 4 * BATCH * N * 8                      # fwd + inv, in + out
    line["e2e"] = {"value": world * 2 * pin["batch"] * pin["steps"] / pin["s"], "unit": "NTT/s",
                   "h2d_bytes_per_step": pin["h2d"], "d2h_bytes_per_step": pin["d2h"],
                   "api": "hexl_b200_host_{set_worksize_,}ntt/intt(+completed) on pinned host buffers",
                   "batch_per_gpu": pin["batch"], "steps": pin["steps"],
                   "pageable": {"value": world * 2 * pag["batch"] * pag["steps"] / pag["s"], "unit": "NTT/s",
                                "note": "the same calls on pageable (numpy) buffers: staged through the runtime's "
                                        "pinned ring by its copy threads",
                                "ratio_to_pinned": pin["s"] / pag["s"] * pag["steps"] / pin["steps"]},
                   "pageable_registered": {"value": world * 2 * reg["batch"] * reg["steps"] / reg["s"], "unit": "NTT/s",
                                           "note": "the same pageable buffer registered once in place with "
                                                   "hexl_b200_host_pin_buffer (what an integration does for its "
                                                   "ciphertext pool): no staging copies",
                                           "ratio_to_pinned": pin["s"] / reg["s"] * reg["steps"] / pin["steps"]},
                   "copy_ceiling": {"seconds_per_step": ceiling, "GBps_per_gpu_each_way": step_bytes / 2 / ceiling / 1e9,
                                    "note": "plain cudaMemcpyAsync of one step's bytes (H2D and D2H on two streams, "
                                            "pinned buffers, all ranks at once): the PCIe / host-memory ceiling of "
                                            "this box for this traffic",
                                    "e2e_fraction_of_ceiling": ceiling * pin["steps"] / pin["s"]}}
    if rank == 0:
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def copy_ceiling(dev):
    """seconds for one step's worth of plain copies: 2 x (H2D + D2H) of the batch, both directions at once"""
    import torch

    h_in = torch.empty((BATCH, N), dtype=torch.int64).pin_memory()
    h_out = torch.empty((BATCH, N), dtype=torch.int64).pin_memory()
    d_a = torch.empty((BATCH, N), dtype=torch.int64, device=dev)
    d_b = torch.zeros((BATCH, N), dtype=torch.int64, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for _ in range(3):
        if torch.distributed.is_initialized():
            torch.distributed.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):                                  # fwd pass + inv pass
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def keyswitch_sharded(hb, ob, dev, rank, world, barrier):
    """BASELINE configs[4]: world x 4096 keyswitch items, this rank's contiguous shard, device resident."""
    import torch

    from ks_util import KsProblem
    from sharding import shard

    total = world * KS_SHARD
    start, count = shard(total, world, rank)
    assert count == KS_SHARD
    chk = KsProblem(N, KS_D, KS_K, 2, 51, seed=77 + rank)          # items whose answer the oracle knows
    plan = hb.KsPlan(N, KS_D, KS_K, KS_D + 1, 2, chk.moduli, chk.keys, chk.msf)
    g = torch.Generator(device=dev).manual_seed(4242 + rank)
    res = torch.empty((count, 2 * KS_D * N), dtype=torch.int64, device=dev)
    tt = torch.empty((count, KS_D * N), dtype=torch.int64, device=dev)
    for j in range(KS_D):
        q = int(chk.moduli[j])
        tt[:, j * N:(j + 1) * N] = torch.randint(0, q, (count, N), dtype=torch.int64, device=dev, generator=g)
        for c in range(2):
            o = (c * KS_D + j) * N
            res[:, o:o + N] = torch.randint(0, q, (count, N), dtype=torch.int64, device=dev, generator=g)
    for pos, b in ((0, 0), (count - 1, 1)):
        tt[pos] = gpu_tensor(chk.t_target[b], dev)
        res[pos] = gpu_tensor(chk.result[b], dev)
    plan.keyswitch(res, tt, count)                                   # warm-up, and the checked pass
    want = chk.expected()
    for pos, b in ((0, 0), (count - 1, 1)):
        got = res[pos].cpu().numpy().view(np.uint64)
        assert np.array_equal(got, want[b]), f"rank {rank}: keyswitch item {start + pos} differs from the oracle"
    steps = 10
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    evs = [event_pair()[0] for _ in range(steps + 1)]
    barrier()
    w0 = time.time()
    evs[0].record()
    for k in range(steps):
        plan.keyswitch(res, tt, count)
        evs[k + 1].record()
    barrier()
    w1 = time.time()
    s = evs[0].elapsed_time(evs[steps]) * 1e-3
    first = evs[0].elapsed_time(evs[2]) * 1e-3 / 2          # the first two steps: before the power cap bites
    clocks = sampler.stop(w0, w1) if sampler else None
    del res, tt
    plan.close()
    torch.cuda.empty_cache()
    return {"s": s, "steps": steps, "first_steps_s": first, "clocks": clocks}


def e2e_ntt(args, hb, ob, dev, world, kind="pinned"):
    """Same workload through the host API: host buffers in, host buffers out."""
    import torch

    t = ob.Tables(N, Q52)
    hb.acquire_FPGA_resources()
    try:
        host = torch.randint(0, Q52, (BATCH, N), dtype=torch.int64)
        if kind == "pinned":
            host = host.pin_memory()
        elif kind == "registered":
            hb.pin_buffer(host.numpy())                 # released with the library below
        ref0 = host[:1].clone()
        ptr = host.data_ptr()
        steps = max(2, min(args.steps, 5))

        def step():
            hb.set_worksize_NTT(BATCH)
            hb.NTT_many(ptr, N, BATCH, t.roots, t.precon, Q52, N)
            hb.NTTCompleted()
            hb.set_worksize_INTT(BATCH)
            hb.INTT_many(ptr, N, BATCH, t.inv_roots, t.precon_inv, Q52, t.inv_n, t.inv_n_w, N)
            hb.INTTCompleted()

        step()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        hb.reset_stats()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = hb.get_stats()
        assert torch.equal(host[:1], ref0)
        return {"s": dt, "steps": steps, "batch": BATCH, "h2d": st["h2d_bytes"] // steps,
                "d2h": st["d2h_bytes"] // steps}
    finally:
        hb.release_FPGA_resources()


def extras(args, hb, ob, dev, hbm_peak, world):
    """keyswitch + dyadic device-resident numbers and the CPU baseline (rank 0)."""
    import torch

    from ks_util import KsProblem

    out = {}
    # -- keyswitch, BASELINE configs[3] --
    p = KsProblem(N, KS_D, KS_K, 1, 51)
    plan = hb.KsPlan(N, KS_D, KS_K, KS_D + 1, 2, p.moduli, p.keys, p.msf)
    g = torch.Generator(device=dev).manual_seed(99)
    res = torch.empty((KS_BATCH, 2 * KS_D * N), dtype=torch.int64, device=dev)
    tt = torch.empty((KS_BATCH, KS_D * N), dtype=torch.int64, device=dev)
    for j in range(KS_D):
        q = int(p.moduli[j])
        tt[:, j * N:(j + 1) * N] = torch.randint(0, q, (KS_BATCH, N), dtype=torch.int64, device=dev, generator=g)
        for c in range(2):
            o = (c * KS_D + j) * N
            res[:, o:o + N] = torch.randint(0, q, (KS_BATCH, N), dtype=torch.int64, device=dev, generator=g)
    plan.keyswitch(res, tt, KS_BATCH)
    torch.cuda.synchronize()
    hb.reset_stats()
    ksteps = 10
    ks_sampler = ClockSampler(torch.cuda.current_device())
    kev = [event_pair()[0] for _ in range(ksteps + 1)]
    kw0 = time.time()
    kev[0].record()
    for k in range(ksteps):
        plan.keyswitch(res, tt, KS_BATCH)
        kev[k + 1].record()
    kev[ksteps].synchronize()
    kw1 = time.time()
    ks_s = kev[0].elapsed_time(kev[ksteps]) * 1e-3 / ksteps
    ks_first = kev[0].elapsed_time(kev[3]) * 1e-3 / 3
    ks_clocks = ks_sampler.stop(kw0, kw1)
    out["keyswitch"] = {
        "metric": "KeySwitch/s (N=16384, decomp=7, key=8)", "value": KS_BATCH / ks_s, "unit": "KeySwitch/s",
        "batch": KS_BATCH, "ms_per_step": ks_s * 1e3, "gpu_launches": hb.get_stats()["kernel_launches"],
        "roofline": {"bound": "hbm", "achieved": KS_BATCH * KS_BYTES / ks_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": KS_BATCH * KS_BYTES / ks_s / 1e9 / hbm_peak,
                     "note": "compute-bound by construction: 72 NTT-equivalents per 4.4 MiB"},
        "steps": ksteps, "clocks": ks_clocks,
        "burst": {"value": KS_BATCH / ks_first, "unit": "KeySwitch/s", "note": "the first three steps (~25 ms)"}}
    del res, tt
    plan.close()
    # keyswitch end to end through the reference-facing host API (pinned host buffers)
    eb = 256
    h_res = torch.empty((eb, 2 * KS_D * N), dtype=torch.int64).pin_memory()
    h_t = torch.empty((eb, KS_D * N), dtype=torch.int64).pin_memory()
    for j in range(KS_D):
        q = int(p.moduli[j])
        h_t[:, j * N:(j + 1) * N] = torch.randint(0, q, (eb, N), dtype=torch.int64)
        for c in range(2):
            o = (c * KS_D + j) * N
            h_res[:, o:o + N] = torch.randint(0, q, (eb, N), dtype=torch.int64)
    keys = hb.KeyArray(p.keys)
    hb.acquire_FPGA_resources()
    try:
        def ks_step():
            hb.set_worksize_KeySwitch(eb)
            hb.KeySwitch_many(h_res.data_ptr(), h_t.data_ptr(), eb, N, KS_D, KS_K, KS_D + 1, 2, p.moduli, keys, p.msf)
            hb.KeySwitchCompleted()
        ks_step()
        hb.reset_stats()
        t0 = time.perf_counter()
        for _ in range(2):
            ks_step()
        ks_e2e_s = (time.perf_counter() - t0) / 2
        st = hb.get_stats()
    finally:
        hb.release_FPGA_resources()
    out["keyswitch"]["e2e"] = {"value": eb / ks_e2e_s, "unit": "KeySwitch/s", "batch": eb,
                               "h2d_bytes_per_step": st["h2d_bytes"] // 2, "d2h_bytes_per_step": st["d2h_bytes"] // 2,
                               "api": "hexl_b200_host_keyswitch (+set_worksize/completed), pinned host buffers"}
    if world == 1:      # CPU baselines are an N=1 item
        # CPU baseline for keyswitch: the oracle port (intel-hexl's KeySwitch is unvendored)
        threads = host_threads()
        cb = 4 * max(2, threads)
        c_res = np.ascontiguousarray(np.resize(p.result, (cb, 2 * KS_D * N)))
        c_t = np.ascontiguousarray(np.resize(p.t_target, (cb, KS_D * N)))
        best = None
        for alt in (True, True, False):          # the cheaper restatement twice (first pass warms up), then the other
            t0 = time.perf_counter()
            ob.keyswitch(c_res.reshape(-1), c_t.reshape(-1), N, KS_D, KS_K, p.moduli, p.keys, p.msf, cb, alt=alt,
                         threads=threads)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out["keyswitch"]["cpu_baseline"] = {
            "value": cb / best, "unit": "KeySwitch/s", "cores": threads, "kind": "port",
            "sample": f"{cb} items on {threads} threads, best of the two scalar oracle restatements (hexl order: digit "
                      "reuse + lazy transforms); the 8 modulus tables are built once per call and amortised over the "
                      "items; intel-hexl's AVX-512 KeySwitch is unvendored"}
    # -- the reference's own NTT benchmark modulus (benchmark/bench_fwd_ntt.cpp:29-30: q = 136314881,
    #    28 bits): q < 2^30 takes the uint32 kernels --
    q28 = 136314881
    t28 = ob.Tables(N, q28)
    x28 = torch.randint(0, q28, (BATCH, N), dtype=torch.int64, device=dev)
    r28, p28 = gpu_tensor(t28.roots, dev), gpu_tensor(t28.precon, dev)
    ir28, ip28 = gpu_tensor(t28.inv_roots, dev), gpu_tensor(t28.precon_inv, dev)
    for _ in range(3):
        hb.ntt_fwd(x28, r28, p28, q28, N)
        hb.ntt_inv(x28, ir28, ip28, q28, t28.inv_n, t28.inv_n_w, N)
    ef, ei = [], []
    for _ in range(10):
        a0, a1 = event_pair(); a0.record(); hb.ntt_fwd(x28, r28, p28, q28, N); a1.record(); ef.append((a0, a1))
        b0, b1 = event_pair(); b0.record(); hb.ntt_inv(x28, ir28, ip28, q28, t28.inv_n, t28.inv_n_w, N); b1.record(); ei.append((b0, b1))
    torch.cuda.synchronize()
    f_s = float(np.mean([a.elapsed_time(b) for a, b in ef])) * 1e-3
    i_s = float(np.mean([a.elapsed_time(b) for a, b in ei])) * 1e-3
    out["ntt_28bit_modulus"] = {
        "note": "same workload with the reference benchmark's modulus q=136314881 (benchmark/bench_fwd_ntt.cpp:30); "
                "q < 2^30 runs the uint32 small-modulus kernels",
        "fwd_per_s": BATCH / f_s, "inv_per_s": BATCH / i_s, "value": 2 * BATCH / (f_s + i_s), "unit": "NTT/s",
        "roofline": {"bound": "hbm", "kernel": "k_ntt_small (forward)", "achieved": BATCH * NTT_BYTES / f_s / 1e9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": BATCH * NTT_BYTES / f_s / 1e9 / hbm_peak,
                     "frac_inverse": BATCH * NTT_BYTES / i_s / 1e9 / hbm_peak}}
    del x28
    # -- dyadic multiply, BASELINE configs[2] --
    moduli = np.array(ob.primes(DY_M, 51, DY_N), dtype=np.uint64)
    op1 = torch.randint(0, int(moduli[0]), (DY_BATCH, 2 * DY_M * DY_N), dtype=torch.int64, device=dev)
    op2 = torch.randint(0, int(moduli[0]), (DY_BATCH, 2 * DY_M * DY_N), dtype=torch.int64, device=dev)
    rs = torch.empty((DY_BATCH, 3 * DY_M * DY_N), dtype=torch.int64, device=dev)
    dm = gpu_tensor(moduli, dev)
    hb.dyadic_multiply(rs, op1, op2, DY_N, dm, DY_M, DY_BATCH)
    e0, e1 = event_pair()
    e0.record()
    for _ in range(3):
        hb.dyadic_multiply(rs, op1, op2, DY_N, dm, DY_M, DY_BATCH)
    e1.record()
    e1.synchronize()
    dy_s = e0.elapsed_time(e1) * 1e-3 / 3
    out["dyadic_multiply"] = {
        "metric": "DyadicMultiply/s (n=8192, 4 moduli)", "value": DY_BATCH / dy_s, "unit": "multiplies/s",
        "batch": DY_BATCH, "roofline": {"bound": "hbm", "achieved": DY_BATCH * DY_BYTES / dy_s / 1e9,
                                        "peak": hbm_peak, "unit": "GB/s",
                                        "frac": DY_BATCH * DY_BYTES / dy_s / 1e9 / hbm_peak}}
    del op1, op2, rs
    # -- fused polynomial multiply: a*b mod (x^N+1, q), N=16384, 52-bit prime --
    pb = 2048
    t52 = ob.Tables(N, Q52)
    pa = torch.randint(0, Q52, (pb, N), dtype=torch.int64, device=dev)
    pbb = torch.randint(0, Q52, (pb, N), dtype=torch.int64, device=dev)
    pr = torch.empty_like(pa)
    tw = [gpu_tensor(x, dev) for x in (t52.roots, t52.precon, t52.inv_roots, t52.precon_inv)]
    hb.poly_multiply(pr, pa, pbb, *tw, Q52, t52.inv_n, t52.inv_n_w, N)
    e0, e1 = event_pair()
    e0.record()
    for _ in range(3):
        hb.poly_multiply(pr, pa, pbb, *tw, Q52, t52.inv_n, t52.inv_n_w, N)
    e1.record()
    e1.synchronize()
    pm_s = e0.elapsed_time(e1) * 1e-3 / 3
    out["poly_multiply"] = {
        "metric": "negacyclic polynomial multiplies/s (N=16384, 52-bit prime)", "value": pb / pm_s,
        "unit": "multiplies/s", "batch": pb,
        "note": "2 forward NTTs + inverse NTT with the dyadic product fused into its first pass (3 transforms per "
                "multiply, compute-bound like the NTT itself)",
        "roofline": {"bound": "hbm", "achieved": pb * 3 * N * 8 / pm_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": pb * 3 * N * 8 / pm_s / 1e9 / hbm_peak,
                     "algorithmic_bytes_per_multiply": 3 * N * 8}}
    del pa, pbb, pr
    if world == 1:
        # -- CPU baseline (bounded sample of the headline workload) --
        threads = host_threads()
        polys = 128 * threads
        rate, kind = cpu_ntt_rate(polys, threads)
        out["cpu_baseline"] = {"value": rate, "unit": "NTT/s", "cores": threads, "kind": kind,
                               "sample": f"{polys} polynomials fwd+inv, N=16384, 52-bit prime, {threads} threads; "
                                         "reference tests/test_utils/ntt.cpp scalar Harvey NTT "
                                         "(intel-hexl AVX-512 is unvendored)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
               os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup",
               str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
