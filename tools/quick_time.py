"""Development timing script (not the contract bench): device-resident timings of
every primitive at the BASELINE shapes, CUDA events, L2 flushed between reps by
working on buffers far larger than L2."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb
import oracle_binding as ob
from ks_util import KsProblem

PEAK = 6546.6e9


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    out = []
    if os.environ.get("QT_PDL"):
        hb.set_option("pdl", int(os.environ["QT_PDL"]))
    N, B = 16384, 4096
    for q in (2251799814045697, 136314881):
        t = ob.Tables(N, q)
        x = torch.randint(0, q, (B, N), dtype=torch.int64, device="cuda")
        r, p, ir, ip = gpu(t.roots), gpu(t.precon), gpu(t.inv_roots), gpu(t.precon_inv)
        # (variant, small_path, fp64_path)
        # fp64 = 2: FP64 kernels without the warp-dealt tail rows
        for variant, small, fp64 in (((1, 1, 1), (3, 1, 1)) if q < 2**30 else ((1, 1, 1), (3, 1, 1), (1, 1, 2), (1, 1, 0))):
            hb.set_option("warp_tail", 0 if fp64 == 2 else 1)
            hb.set_option("ntt_variant", variant)
            hb.set_option("small_path", small)
            hb.set_option("fp64_path", fp64)
            med, best = timeit(lambda: hb.ntt_fwd(x, r, p, q, N))
            out.append({"op": "ntt_fwd", "q": q, "variant": variant, "small_path": small, "fp64_path": fp64, "batch": B, "s": med, "best_s": best,
                        "per_s": B / med, "frac_hbm": B * 262144 / med / PEAK})
            x %= q
            med, best = timeit(lambda: hb.ntt_inv(x, ir, ip, q, t.inv_n, t.inv_n_w, N))
            out.append({"op": "ntt_inv", "q": q, "variant": variant, "fp64_path": fp64, "batch": B, "s": med, "best_s": best,
                        "per_s": B / med, "frac_hbm": B * 262144 / med / PEAK})
            print(json.dumps(out[-2]), flush=True)
            print(json.dumps(out[-1]), flush=True)
        hb.set_option("ntt_variant", 1)
        hb.set_option("small_path", 1)
        hb.set_option("fp64_path", 1)
        hb.set_option("warp_tail", 1)
        del x
    # dyadic config 3
    n, M, Bd = 8192, 4, 8192
    moduli = np.array(ob.primes(M, 51, n), dtype=np.uint64)
    op1 = torch.randint(0, int(moduli[0]), (Bd, 2 * M * n), dtype=torch.int64, device="cuda")
    op2 = torch.randint(0, int(moduli[0]), (Bd, 2 * M * n), dtype=torch.int64, device="cuda")
    res = torch.empty((Bd, 3 * M * n), dtype=torch.int64, device="cuda")
    dm = gpu(moduli)
    med, best = timeit(lambda: hb.dyadic_multiply(res, op1, op2, n, dm, M, Bd), reps=5)
    out.append({"op": "dyadic", "batch": Bd, "s": med, "best_s": best, "per_s": Bd / med,
                "frac_hbm": Bd * 1835008 / med / PEAK})
    print(json.dumps(out[-1]), flush=True)
    del op1, op2, res
    # keyswitch config 4
    n, D, K, Bk = 16384, 7, 8, 1024
    p = KsProblem(n, D, K, 1, 51)
    plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
    resk = gpu(p.result).repeat(Bk, 1).contiguous()
    tt = gpu(p.t_target).repeat(Bk, 1).contiguous()
    med, best = timeit(lambda: plan.keyswitch(resk, tt, Bk), reps=3, warm=1)
    out.append({"op": "keyswitch", "batch": Bk, "s": med, "best_s": best, "per_s": Bk / med,
                "frac_hbm": Bk * 4587520 / med / PEAK})
    print(json.dumps(out[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "quick_time.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
