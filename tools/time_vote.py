"""Where does the range vote's cost go?  forward / inverse NTT (N=16384, 52-bit, batch 4096) with
   (a) vote + deferred-list pass (default), (b) vote, list pass skipped, (c) no vote (trust)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
N, q, B = 16384, 2251799814045697, 4096
t = ob.Tables(N, q)
x = torch.randint(0, q, (B, N), dtype=torch.int64, device="cuda")
r, p, ir, ip = gpu(t.roots), gpu(t.precon), gpu(t.inv_roots), gpu(t.precon_inv)
def timeit(fn, reps=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps
for name, variant, skip in (("vote+list", 1, 0), ("vote, no list pass", 1, 1), ("trust", 3, 0)):
    hb.set_option("ntt_variant", variant); hb.set_option("debug_skip_list", skip)
    f = timeit(lambda: hb.ntt_fwd(x, r, p, q, N))
    x %= q
    i = timeit(lambda: hb.ntt_inv(x, ir, ip, q, t.inv_n, t.inv_n_w, N))
    print(json.dumps({"mode": name, "fwd_us": f * 1e6, "inv_us": i * 1e6, "fwd_per_s": B / f, "inv_per_s": B / i}), flush=True)
hb.set_option("ntt_variant", 1); hb.set_option("debug_skip_list", 0)
