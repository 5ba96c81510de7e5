"""Device-resident keyswitch rate (N=16384, D/K = 7/8) against the workspace size, i.e. the number of items per
chunk: larger chunks mean longer runs of one modulus per CTA (fewer twiddle reloads) and less wave quantisation.
   python tools/time_ks_ws.py [batch] [mb ...]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sizes = [int(x) for x in sys.argv[2:]] or [4096, 8192, 16384, 40960]
n, D, K = 16384, 7, 8
p = KsProblem(n, D, K, 1, 51)
exp = gpu(p.expected())
tt = gpu(p.t_target).repeat(B, 1).contiguous()
for mb in sizes:
    hb.set_option("ks_workspace_mb", mb)
    plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
    r2 = gpu(p.result).repeat(B, 1).contiguous()
    plan.keyswitch(r2, tt, B); torch.cuda.synchronize()
    ok = bool(torch.equal(r2, exp.expand(B, -1)))
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.keyswitch(r2, tt, B); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    print(json.dumps({"batch": B, "ks_workspace_mb": mb, "per_s": B / float(np.median(ts)), "ok": ok}), flush=True)
    plan.close()
    del r2
    torch.cuda.empty_cache()
