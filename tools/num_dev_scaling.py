"""The library's own multi-GPU mode (NUM_DEV workers behind ONE caller, reference DevicePool fpga.cpp:1646-1673):
end-to-end NTT+INTT (pinned, 4096 polynomials) and keyswitch (7/8, 256 items) throughput for NUM_DEV = 1, 2, ...,
with the per-worker item counts.   python tools/num_dev_scaling.py [max_dev]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.join(%(root)r, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem
nd = int(os.environ["NUM_DEV"])
N, Q, B = 16384, 2251799814045697, 4096
t = ob.Tables(N, Q)
hb.acquire_FPGA_resources()
host = torch.randint(0, Q, (B, N), dtype=torch.int64).pin_memory()
ptr = host.data_ptr()
def step():
    hb.set_worksize_NTT(B); hb.NTT_many(ptr, N, B, t.roots, t.precon, Q, N); hb.NTTCompleted()
    hb.set_worksize_INTT(B); hb.INTT_many(ptr, N, B, t.inv_roots, t.precon_inv, Q, t.inv_n, t.inv_n_w, N); hb.INTTCompleted()
step()
t0 = time.perf_counter()
for _ in range(4): step()
dt = (time.perf_counter() - t0) / 4
p = KsProblem(N, 7, 8, 1, 51)
KB = 256
keys = hb.KeyArray(p.keys)
res = torch.from_numpy(np.ascontiguousarray(np.repeat(p.result, KB, axis=0)).view(np.int64)).pin_memory()
tt = torch.from_numpy(np.ascontiguousarray(np.repeat(p.t_target, KB, axis=0)).view(np.int64)).pin_memory()
def kstep():
    hb.set_worksize_KeySwitch(KB)
    hb.KeySwitch_many(res.data_ptr(), tt.data_ptr(), KB, N, 7, 8, 8, 2, p.moduli, keys, p.msf)
    hb.KeySwitchCompleted()
kstep()
t0 = time.perf_counter()
for _ in range(3): kstep()
kdt = (time.perf_counter() - t0) / 3
st = [hb.device_stats(w) for w in range(nd)]
print(json.dumps({"NUM_DEV": nd, "ntt_per_s": 2 * B / dt, "keyswitch_per_s": KB / kdt, "items_per_worker": [s["items"] for s in st]}))
hb.release_FPGA_resources()
'''
import torch
mx = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
nd = 1
while nd <= mx:
    subprocess.run([sys.executable, "-c", CODE % {"root": ROOT}], env=dict(os.environ, NUM_DEV=str(nd)))
    nd *= 2
