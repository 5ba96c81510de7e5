"""Latency of single synchronous calls through the reference-facing host API (worksize 1),
the way SEAL's switch_key_inplace drives it: python tools/latency_host_api.py"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem

N, Q = 16384, 2251799814045697
t = ob.Tables(N, Q)
p = KsProblem(N, 7, 8, 1, 51)
keys = hb.KeyArray(p.keys)
hb.acquire_FPGA_resources()
try:
    for pinned in (True, False):
        x = torch.randint(0, Q, (N,), dtype=torch.int64)
        res = torch.from_numpy(p.result[0].view(np.int64).copy())
        tt = torch.from_numpy(p.t_target[0].view(np.int64).copy())
        if pinned:
            x, res, tt = x.pin_memory(), res.pin_memory(), tt.pin_memory()
        xn, rn, tn = x.numpy().view(np.uint64), res.numpy().view(np.uint64), tt.numpy().view(np.uint64)
        for name, fn in (("NTT", lambda: (hb.NTT(xn, t.roots, t.precon, Q, N), hb.NTTCompleted())),
                         ("KeySwitch", lambda: (hb.KeySwitch(rn, tn, N, 7, 8, 8, 2, p.moduli, keys, p.msf),
                                                hb.KeySwitchCompleted()))):
            for _ in range(5):
                fn()
            ts = []
            for _ in range(50):
                t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
            print(json.dumps({"op": name, "pinned": pinned, "median_us": float(np.median(ts)) * 1e6,
                              "min_us": float(np.min(ts)) * 1e6}), flush=True)
finally:
    hb.release_FPGA_resources()
