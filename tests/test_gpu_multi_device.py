"""NUM_DEV=2: the library's own multi-GPU mode (one worker thread per GPU popping the
shared request queue, like the reference's DevicePool, host/src/fpga.cpp:1646-1673).
Runs in a subprocess because NUM_DEV is read when the resources are acquired; skipped
on single-GPU boxes.  No BATCH_SIZE_* is set: the runtime deals every run out over the workers by itself,
and the per-worker counters (hexl_b200_host_device_stats) must show it."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(%(root)r, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem
n, q, B = 16384, 2251799814045697, 96
t = ob.Tables(n, q)
data = np.stack([ob.splitmix(n, 500 + i, q) for i in range(B)])
want = np.stack([ob.fwd_ntt(data[i], t) for i in range(0, B, 7)])
hb.acquire_FPGA_resources()
try:
    for rep in range(3):                       # several runs: both workers must take batches
        work = data.copy()
        hb.set_worksize_NTT(B)
        for i in range(B):
            hb.NTT(work[i], t.roots, t.precon, q, n)
        hb.NTTCompleted()
        assert np.array_equal(work[::7], want), "forward NTT mismatch with NUM_DEV=2"
        hb.set_worksize_INTT(B)
        for i in range(B):
            hb.INTT(work[i], t.inv_roots, t.precon_inv, q, t.inv_n, t.inv_n_w, n)
        hb.INTTCompleted()
        assert np.array_equal(work, data), "inverse NTT mismatch with NUM_DEV=2"
    p = KsProblem(4096, 3, 4, 6, 45)
    keys = hb.KeyArray(p.keys)
    out = p.result.copy()
    hb.set_worksize_KeySwitch(6)
    for b in range(6):
        hb.KeySwitch(out[b], p.t_target[b], 4096, 3, 4, 4, 2, p.moduli, keys, p.msf)
    hb.KeySwitchCompleted()
    assert np.array_equal(out, p.expected()), "keyswitch mismatch with NUM_DEV=2"
    st = [hb.device_stats(w) for w in range(2)]
    print("DEVICE_STATS", st)
    assert st[0]["device"] != st[1]["device"]
    # no BATCH_SIZE_* knob is set: every run must have been dealt out over both workers
    assert min(s["items"] for s in st) >= 0.3 * sum(s["items"] for s in st), st
    assert min(s["batches"] for s in st) >= 6, st
finally:
    hb.release_FPGA_resources()
print("MULTI_DEVICE_OK")
'''


def test_two_worker_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, NUM_DEV="2")
    for k in ("BATCH_SIZE_NTT", "BATCH_SIZE_INTT", "BATCH_SIZE_KEYSWITCH"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, "-c", WORKER % {"root": ROOT}], env=env, capture_output=True, text=True,
                         timeout=600)
    assert "MULTI_DEVICE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
