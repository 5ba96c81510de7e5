"""keyswitch throughput, fused (S2+S3+S4 in one kernel, sums in tensor memory) vs staged, device resident"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
n, B = 16384, int(os.environ.get("KS_B", "1024"))
for D, K in ((7, 8), (6, 7)):
    p = KsProblem(n, D, K, 2, 51)
    exp = gpu(p.expected())
    for fused in (1, 0):
        for ws in (4096,):
            hb.set_option("ks_fused", fused); hb.set_option("ks_workspace_mb", ws)
            plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
            res = gpu(p.result).repeat(B // 2, 1).contiguous(); tt = gpu(p.t_target).repeat(B // 2, 1).contiguous()
            plan.keyswitch(res, tt, B); torch.cuda.synchronize()
            ok = bool(torch.equal(res, exp.repeat(B // 2, 1)))
            ts = []
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); plan.keyswitch(res, tt, B); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
            print(json.dumps({"D": D, "K": K, "fused": fused, "workspace_mb": ws, "per_s": B / float(np.median(ts)), "ok": ok}), flush=True)
            plan.close()
hb.set_option("ks_fused", 0); hb.set_option("ks_workspace_mb", 4096)
for D, K in ((7, 8), (6, 7)):
    p = KsProblem(n, D, K, 2, 51)
    exp = gpu(p.expected())
    for sub in (0, 6, 9, 12, 15, 18, 24, 36):
        hb.set_option("ks_sub_items", sub)
        plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
        res = gpu(p.result).repeat(B // 2, 1).contiguous(); tt = gpu(p.t_target).repeat(B // 2, 1).contiguous()
        plan.keyswitch(res, tt, B); torch.cuda.synchronize()
        ok = bool(torch.equal(res, exp.repeat(B // 2, 1)))
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.keyswitch(res, tt, B); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
        print(json.dumps({"D": D, "K": K, "sub_items": sub, "per_s": B / float(np.median(ts)), "ok": ok}), flush=True)
        plan.close()
hb.set_option("ks_sub_items", 0)
