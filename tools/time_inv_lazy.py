"""A/B of the inverse butterflies: python tools/time_inv_lazy.py"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from quick_time import timeit, gpu
N, B, q = 16384, 4096, 2251799814045697
t = ob.Tables(N, q)
x = torch.randint(0, q, (B, N), dtype=torch.int64, device="cuda")
ir, ip = gpu(t.inv_roots), gpu(t.precon_inv)
for lazy in (1, 0, 1, 0):
    hb.set_option("inv_lazy", lazy)
    x %= q
    med, best = timeit(lambda: hb.ntt_inv(x, ir, ip, q, t.inv_n, t.inv_n_w, N))
    print(json.dumps({"inv_lazy": lazy, "s": med, "per_s": B / med}), flush=True)
