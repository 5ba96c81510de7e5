"""e2e NTT throughput through the host API from PAGEABLE memory for several copy-thread counts."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.join(%(root)r, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import hexl_b200 as hb, oracle_binding as ob
N, Q, B = 16384, 2251799814045697, 4096
t = ob.Tables(N, Q)
hb.acquire_FPGA_resources()
host = torch.randint(0, Q, (B, N), dtype=torch.int64)
if os.environ.get("PIN") == "1": host = host.pin_memory()
ptr = host.data_ptr()
def step():
    hb.set_worksize_NTT(B); hb.NTT_many(ptr, N, B, t.roots, t.precon, Q, N); hb.NTTCompleted()
    hb.set_worksize_INTT(B); hb.INTT_many(ptr, N, B, t.inv_roots, t.precon_inv, Q, t.inv_n, t.inv_n_w, N); hb.INTTCompleted()
step()
t0 = time.perf_counter()
for _ in range(4): step()
dt = (time.perf_counter() - t0) / 4
print(json.dumps({"copy_threads": os.environ.get("HEXL_B200_COPY_THREADS"), "slot_mb": os.environ.get("HEXL_B200_SLOT_MB"), "pinned": os.environ.get("PIN"),
                  "ms_per_step": dt * 1e3, "ntt_per_s": 2 * B / dt, "cpus": len(os.sched_getaffinity(0))}))
hb.release_FPGA_resources()
'''
CONFIGS = [("8", "32", "1"), ("4", "32", "0"), ("8", "32", "0"), ("12", "32", "0"), ("14", "32", "0"), ("16", "32", "0"),
           ("12", "16", "0"), ("12", "8", "0"), ("12", "64", "0")]
if len(sys.argv) > 1:      # threads:slot_mb:pinned ...
    CONFIGS = [tuple(a.split(":")) for a in sys.argv[1:]]
for th, mb, pin in CONFIGS:
    env = dict(os.environ, HEXL_B200_COPY_THREADS=th, HEXL_B200_SLOT_MB=mb, PIN=pin)
    subprocess.run([sys.executable, "-c", CODE % {"root": ROOT}], env=env)
