"""A/B of one option on the same box: plain fwd / inv NTT calls (N=16384, 52-bit, batch 4096) and the keyswitch
(7/8, batch 1024), interleaved rounds.   python tools/time_ab.py option [value_a value_b]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem
opt = sys.argv[1]
va, vb = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1)
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
N, Q, B = 16384, 2251799814045697, 4096
t = ob.Tables(N, Q)
x = torch.randint(0, Q, (B, N), dtype=torch.int64, device="cuda")
r, p, ir, ip = gpu(t.roots), gpu(t.precon), gpu(t.inv_roots), gpu(t.precon_inv)
def ev(): return torch.cuda.Event(enable_timing=True)
def time_ntt(reps=20):
    for _ in range(3):
        hb.ntt_fwd(x, r, p, Q, N); hb.ntt_inv(x, ir, ip, Q, t.inv_n, t.inv_n_w, N)
    evs = []
    for _ in range(reps):           # enqueued back to back, as bench.py does; nothing waits before the end
        a, b, c = ev(), ev(), ev()
        a.record(); hb.ntt_fwd(x, r, p, Q, N); b.record(); hb.ntt_inv(x, ir, ip, Q, t.inv_n, t.inv_n_w, N); c.record()
        evs.append((a, b, c))
    torch.cuda.synchronize()
    f = [a.elapsed_time(b) for a, b, c in evs]; i = [b.elapsed_time(c) for a, b, c in evs]
    return float(np.mean(f)) * 1e3, float(np.mean(i)) * 1e3
kp = KsProblem(N, 7, 8, 1, 51)
KB = int(os.environ.get('KB', '1024'))
tt = gpu(kp.t_target).repeat(KB, 1).contiguous()
r2 = gpu(kp.result).repeat(KB, 1).contiguous()
def time_ks():
    plan = hb.KsPlan(N, 7, 8, 8, 2, kp.moduli, kp.keys, kp.msf)
    plan.keyswitch(r2, tt, KB); torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        a, b = ev(), ev(); a.record(); plan.keyswitch(r2, tt, KB); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    plan.close()
    return KB / (float(np.median(ts)) * 1e-3)
for rnd in range(5):
    for v in (va, vb):
        hb.set_option(opt, v)
        f, i = time_ntt()
        print(json.dumps({"option": opt, "value": v, "round": rnd, "fwd_call_us": f, "inv_call_us": i, "keyswitch_per_s": time_ks()}), flush=True)
hb.set_option(opt, vb)
