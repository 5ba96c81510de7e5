"""Where a pinned end-to-end NTT call spends its time: wall time per call (set_worksize + 4096 submissions + Completed)
against the worker's own batch time (HEXL_B200_DEBUG=1 prints it) and the plain-copy floor."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
N, Q, B = 16384, 2251799814045697, 4096
t = ob.Tables(N, Q)
hb.acquire_FPGA_resources()
host = torch.randint(0, Q, (B, N), dtype=torch.int64).pin_memory()
ptr = host.data_ptr()
def call_ntt():
    t0 = time.perf_counter(); hb.set_worksize_NTT(B); hb.NTT_many(ptr, N, B, t.roots, t.precon, Q, N)
    t1 = time.perf_counter(); hb.NTTCompleted(); t2 = time.perf_counter()
    return t1 - t0, t2 - t1
def call_intt():
    t0 = time.perf_counter(); hb.set_worksize_INTT(B); hb.INTT_many(ptr, N, B, t.inv_roots, t.precon_inv, Q, t.inv_n, t.inv_n_w, N)
    t1 = time.perf_counter(); hb.INTTCompleted(); t2 = time.perf_counter()
    return t1 - t0, t2 - t1
call_ntt(); call_intt()
rows = []
for _ in range(4):
    rows.append(("ntt",) + call_ntt()); rows.append(("intt",) + call_intt())
for op, a, b in rows:
    print(json.dumps({"op": op, "submit_ms": a * 1e3, "completed_ms": b * 1e3, "call_ms": (a + b) * 1e3}), flush=True)
# plain copies of the same bytes: one direction at a time and both at once
dev = torch.empty((B, N), dtype=torch.int64, device="cuda"); dev2 = torch.empty_like(dev); host2 = torch.empty_like(host).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return time.perf_counter() - t0
h2d = timed(lambda: dev.copy_(host, non_blocking=True))
d2h = timed(lambda: host2.copy_(dev2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): host2.copy_(dev2, non_blocking=True)
bo = timed(both)
print(json.dumps({"plain_h2d_ms": h2d * 1e3, "plain_d2h_ms": d2h * 1e3, "both_at_once_ms": bo * 1e3,
                  "GBps_h2d": B * N * 8 / h2d / 1e9, "GBps_d2h": B * N * 8 / d2h / 1e9, "GBps_each_way_both": B * N * 8 / bo / 1e9}))
hb.release_FPGA_resources()
