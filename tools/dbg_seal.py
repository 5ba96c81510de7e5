import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hexl_b200 as hb, oracle_binding as ob
from ks_util import KsProblem
def gpu(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
for (n, D, K, bits) in [(2048, 3, 4, 40), (2048, 3, 4, 45), (1024, 2, 3, 40), (4096, 5, 7, 51)]:
    p = KsProblem(n, D, K, 3, bits)
    exp = p.expected()
    plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
    res = gpu(p.result); plan.keyswitch(res, gpu(p.t_target), 3)
    got = res.cpu().numpy().view(np.uint64)
    print("device", n, D, K, bits, "mismatch words:", int((got != exp).sum()))
    hb.acquire_FPGA_resources()
    keys = hb.KeyArray(p.keys)
    outs = [p.result[b].copy() for b in range(3)]
    tts = [p.t_target[b].copy() for b in range(3)]
    hb.set_worksize_KeySwitch(3)
    for b in range(3):
        hb.KeySwitch(outs[b], tts[b], n, D, K, D + 1, 2, p.moduli, keys, p.msf)
    hb.KeySwitchCompleted()
    for b in range(3):
        bad = np.nonzero(outs[b] != exp[b])[0]
        print(" host item", b, "mismatch words:", len(bad), "first", bad[:4], "of", outs[b].size)
    hb.release_FPGA_resources()
