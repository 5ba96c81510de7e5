"""Device-resident keyswitch rate (N=16384, D/K = 7/8 and 6/7, batch 1024) for a list of option settings:
   python tools/time_ks2.py opt=val[,opt=val] ...   (each argument is one configuration; 'default' = no options)"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
B = 1024
for (n, D, K) in ((16384, 7, 8), (16384, 6, 7)):
    p = KsProblem(n, D, K, 1, 51)
    exp = gpu(p.expected())
    tt = gpu(p.t_target).repeat(B, 1).contiguous()
    for cfg in (sys.argv[1:] or ["default"]):
        opts = [] if cfg == "default" else [(kv.split("=")[0], int(kv.split("=")[1])) for kv in cfg.split(",")]
        for k, v in opts: hb.set_option(k, v)
        plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
        r2 = gpu(p.result).repeat(B, 1).contiguous()
        plan.keyswitch(r2, tt, B); torch.cuda.synchronize()
        ok = bool(torch.equal(r2, exp.expand(B, -1)))
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.keyswitch(r2, tt, B); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
        print(json.dumps({"shape": [n, D, K], "cfg": cfg, "per_s": B / float(np.median(ts)), "ok": ok}), flush=True)
        plan.close()
        for k, v in opts: hb.set_option(k, 1 if k in ("ks_mac_fp64",) else 0)
