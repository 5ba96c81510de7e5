"""Does the rate hold over a long run?  Per-window rates of the plain NTT step and of the keyswitch with the SM clock
and the power draw sampled through NVML."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
def ev(): return torch.cuda.Event(enable_timing=True)
def nv():
    return {"sm_mhz": pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), "watts": pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
            "reasons": hex(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else hex(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))}
N, Q, B = 16384, 2251799814045697, 4096
t = ob.Tables(N, Q)
x = torch.randint(0, Q, (B, N), dtype=torch.int64, device="cuda")
r, p, ir, ip = gpu(t.roots), gpu(t.precon), gpu(t.inv_roots), gpu(t.precon_inv)
for w in range(8):                       # 8 windows of 50 steps (~37 ms each)
    a, b = ev(), ev(); a.record()
    for _ in range(50):
        hb.ntt_fwd(x, r, p, Q, N); hb.ntt_inv(x, ir, ip, Q, t.inv_n, t.inv_n_w, N)
    b.record(); s = nv(); b.synchronize()
    print(json.dumps({"op": "ntt", "window": w, "ntt_per_s": 2 * B * 50 / (a.elapsed_time(b) * 1e-3), **s}), flush=True)
kp = KsProblem(N, 7, 8, 1, 51); KB = 1024
tt = gpu(kp.t_target).repeat(KB, 1).contiguous(); r2 = gpu(kp.result).repeat(KB, 1).contiguous()
plan = hb.KsPlan(N, 7, 8, 8, 2, kp.moduli, kp.keys, kp.msf)
plan.keyswitch(r2, tt, KB); torch.cuda.synchronize()
for w in range(10):                      # 10 windows of 4 calls (~33 ms each)
    a, b = ev(), ev(); a.record()
    for _ in range(4): plan.keyswitch(r2, tt, KB)
    b.record(); s = nv(); b.synchronize()
    print(json.dumps({"op": "keyswitch", "window": w, "per_s": 4 * KB / (a.elapsed_time(b) * 1e-3), **s}), flush=True)
