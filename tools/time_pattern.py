"""Forward / inverse NTT call times (N=16384, 52-bit, batch 4096) in two launch patterns: the same call repeated,
   and forward / inverse alternating as in bench.py.  HB_LIB=<path> times another build of libhexl_b200.so."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
if os.environ.get("HB_LIB"):
    hb.LIB_PATH = os.environ["HB_LIB"]
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
N, q, B = 16384, 2251799814045697, 4096
t = ob.Tables(N, q)
x = torch.randint(0, q, (B, N), dtype=torch.int64, device="cuda")
r, p, ir, ip = gpu(t.roots), gpu(t.precon), gpu(t.inv_roots), gpu(t.precon_inv)
fwd = lambda: hb.ntt_fwd(x, r, p, q, N)
inv = lambda: hb.ntt_inv(x, ir, ip, q, t.inv_n, t.inv_n_w, N)
def pair():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); fwd(); e[1].record(); inv(); e[2].record()
    return e
for _ in range(5): fwd(); inv()
torch.cuda.synchronize()
ev = [pair() for _ in range(20)]
torch.cuda.synchronize()
alt_f = float(np.mean([e[0].elapsed_time(e[1]) for e in ev])) * 1e3
alt_i = float(np.mean([e[1].elapsed_time(e[2]) for e in ev])) * 1e3
def rep(fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 20
rf = rep(fwd)
x %= q
ri = rep(inv)
print(json.dumps({"lib": os.environ.get("HB_LIB", "current"), "alternating_fwd_us": alt_f, "alternating_inv_us": alt_i,
                  "repeated_fwd_us": rf, "repeated_inv_us": ri}))
