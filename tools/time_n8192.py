import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
def timeit(fn, reps=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))
for N, B in ((8192, 8192), (4096, 16384)):
    q = ob.primes(1, 51, N)[0]; t = ob.Tables(N, q)
    x = torch.randint(0, q, (B, N), dtype=torch.int64, device="cuda")
    r, p, ir, ip = gpu(t.roots), gpu(t.precon), gpu(t.inv_roots), gpu(t.precon_inv)
    bf = N // 2 * int(np.log2(N))
    for variant in (0, 1, 2, 3):
        hb.set_option("ntt_variant", variant)
        s = timeit(lambda: hb.ntt_fwd(x, r, p, q, N)); x %= q
        si = timeit(lambda: hb.ntt_inv(x, ir, ip, q, t.inv_n, t.inv_n_w, N))
        print(json.dumps({"N": N, "variant": variant, "fwd_per_s": B / s, "inv_per_s": B / si,
                          "fwd_equiv16384_per_s": B / s * bf / 114688, "fwd_frac_hbm": B * N * 16 / s / 6546.6e9}), flush=True)
hb.set_option("ntt_variant", 1)
