"""Forward / inverse NTT timing per kernel variant: python tools/time_variants.py"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from quick_time import timeit, gpu
N, B, q = 16384, 4096, 2251799814045697
t = ob.Tables(N, q)
x = torch.randint(0, q, (B, N), dtype=torch.int64, device="cuda")
r, p, ir, ip = gpu(t.roots), gpu(t.precon), gpu(t.inv_roots), gpu(t.precon_inv)
for variant in (0, 2, 1, 3):
    hb.set_option("ntt_variant", variant)
    x %= q
    med, _ = timeit(lambda: hb.ntt_fwd(x, r, p, q, N))
    x %= q
    medi, _ = timeit(lambda: hb.ntt_inv(x, ir, ip, q, t.inv_n, t.inv_n_w, N))
    print(json.dumps({"variant": variant, "fwd_per_s": B / med, "inv_per_s": B / medi}), flush=True)
hb.set_option("ntt_variant", 1)
