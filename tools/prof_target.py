"""Small launch target for ncu: python tools/prof_target.py {ntt|dyadic|keyswitch} [variant]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb
import oracle_binding as ob
from ks_util import KsProblem


def gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


for kv in os.environ.get("HB_OPTS", "").split(","):      # option=value[,option=value]: applied before anything runs
    if kv:
        hb.set_option(kv.split("=")[0], int(kv.split("=")[1]))
op = sys.argv[1] if len(sys.argv) > 1 else "ntt"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
if op in ("ntt", "ntt28"):
    N, q = 16384, (2251799814045697 if op == "ntt" else 136314881)
    t = ob.Tables(N, q)
    x = torch.randint(0, q, (batch, N), dtype=torch.int64, device="cuda")
    r, p, ir, ip = gpu(t.roots), gpu(t.precon), gpu(t.inv_roots), gpu(t.precon_inv)
    hb.set_option("ntt_variant", variant)
    if os.environ.get("SMALL_PATH"):
        hb.set_option("small_path", int(os.environ["SMALL_PATH"]))
    for _ in range(3):
        hb.ntt_fwd(x, r, p, q, N)
        hb.ntt_inv(x, ir, ip, q, t.inv_n, t.inv_n_w, N)
elif op == "dyadic":
    n, M, B = 8192, 4, 2048
    moduli = np.array(ob.primes(M, 51, n), dtype=np.uint64)
    op1 = torch.randint(0, int(moduli[0]), (B, 2 * M * n), dtype=torch.int64, device="cuda")
    op2 = torch.randint(0, int(moduli[0]), (B, 2 * M * n), dtype=torch.int64, device="cuda")
    res = torch.empty((B, 3 * M * n), dtype=torch.int64, device="cuda")
    for _ in range(3):
        hb.dyadic_multiply(res, op1, op2, n, gpu(moduli), M, B)
elif op == "polymul":
    N, q = 16384, 2251799814045697
    t = ob.Tables(N, q)
    hb.set_option("polymul_fused", variant)
    a = torch.randint(0, q, (batch, N), dtype=torch.int64, device="cuda")
    b = torch.randint(0, q, (batch, N), dtype=torch.int64, device="cuda")
    res = torch.empty_like(a)
    tw = [gpu(x) for x in (t.roots, t.precon, t.inv_roots, t.precon_inv)]
    for _ in range(2):
        hb.poly_multiply(res, a, b, *tw, q, t.inv_n, t.inv_n_w, N)
elif op == "keyswitch_fused":
    n, D, K, B = 16384, 7, 8, batch
    hb.set_option("ks_fused", 1)
    p = KsProblem(n, D, K, 1, 51)
    plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
    res = gpu(p.result).repeat(B, 1).contiguous()
    tt = gpu(p.t_target).repeat(B, 1).contiguous()
    for _ in range(2):
        plan.keyswitch(res, tt, B)
elif op == "keyswitch":
    for kv in os.environ.get("KS_OPTS", "").split(","):
        if kv:
            hb.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    n, D, K, B = 16384, 7, 8, batch if len(sys.argv) > 3 else 64
    p = KsProblem(n, D, K, 1, 51)
    plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
    res = gpu(p.result).repeat(B, 1).contiguous()
    tt = gpu(p.t_target).repeat(B, 1).contiguous()
    for _ in range(2):
        plan.keyswitch(res, tt, B)
torch.cuda.synchronize()
print("done", op)
