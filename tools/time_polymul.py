import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import hexl_b200 as hb, oracle_binding as ob
from quick_time import timeit, gpu
N, q, B = 16384, 2251799814045697, 2048
hb.set_option("polymul_fused", int(os.environ.get("PM_FUSED", "1")))
t = ob.Tables(N, q)
a = torch.randint(0, q, (B, N), dtype=torch.int64, device="cuda")
b = torch.randint(0, q, (B, N), dtype=torch.int64, device="cuda")
r = torch.empty_like(a)
tw = [gpu(x) for x in (t.roots, t.precon, t.inv_roots, t.precon_inv)]
med, best = timeit(lambda: hb.poly_multiply(r, a, b, *tw, q, t.inv_n, t.inv_n_w, N))
print(json.dumps({"op": "poly_multiply", "batch": B, "s": med, "per_s": B / med}))
