import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
n, D, K, B = 16384, 7, 8, 1024
p = KsProblem(n, D, K, 1, 51)
plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
res = gpu(p.result).repeat(B, 1).contiguous(); tt = gpu(p.t_target).repeat(B, 1).contiguous()
exp = gpu(p.expected())
for mi in (4, 2, 8):
    for ws in (1024, 4096):
        hb.set_option("ks_mac_items", mi); hb.set_option("ks_workspace_mb", ws)
        r2 = gpu(p.result).repeat(B, 1).contiguous()
        plan.keyswitch(r2, tt, B); torch.cuda.synchronize()
        ok = bool(torch.equal(r2, exp.expand(B, -1)))
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.keyswitch(res, tt, B); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
        print(json.dumps({"mac_items": mi, "workspace_mb": ws, "per_s": B / float(np.median(ts)), "ok": ok}), flush=True)
