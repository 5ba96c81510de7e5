"""KeySwitch parity pinned on the REFERENCE'S OWN device code.

oracle/_ref/ks_ref_emul is device/keyswitch.cpp + device/keyswitch/*.hpp + device/mod_ops.hpp of the
reference compiled UNMODIFIED for the CPU (oracle/sycl_shim stands in for the oneAPI FPGA emulator;
recipe in oracle/Makefile, host glue restated in oracle/ref_ks_emul.cpp).  Its answers for seeded
problems -- the reference's own test shapes 6/7/7/2 and 5/7/6/2 at N = 16384 and 8192
(tests/test_keyswitch.cpp:148-191) and smaller ones -- are committed in
tests/golden/ks_ref_emul_golden.json (generator: tests/golden/make_keyswitch_ref_golden.py).

  CPU : the oracle restatement reproduces every golden answer (and, when the emulator binary is
        present, arbitrary fresh problems live);
  GPU : the CUDA path reproduces the same golden answers through the device API and the host API."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
import ref_emul
from ks_util import KsProblem

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "ks_ref_emul_golden.json")) as fh:
    GOLDEN = json.load(fh)["cases"]
IDS = ["n%d_D%d_b%d" % (c["n"], c["D"], c["bits"]) for c in GOLDEN]


def problem(c):
    return KsProblem(c["n"], c["D"], c["K"], c["batch"], c["bits"], seed=c["seed"])


def check(got, c):
    got = np.ascontiguousarray(got, dtype=np.uint64).reshape(-1)
    assert [int(x) for x in got[:4]] == c["head"] and [int(x) for x in got[-2:]] == c["tail"]
    assert "%016x" % ob.fnv(got) == c["fnv"]


@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_oracle_matches_reference_device_code(c):
    p = problem(c)
    check(p.expected(), c)


@pytest.mark.parametrize("c", [c for c in GOLDEN if c["n"] <= 2048], ids=[i for i, c in zip(IDS, GOLDEN) if c["n"] <= 2048])
def test_second_restatement_matches_reference_device_code(c):
    check(problem(c).expected(alt=True), c)


@pytest.mark.skipif(not ref_emul.available(), reason="oracle/_ref/ks_ref_emul not built (needs /root/reference)")
@pytest.mark.parametrize("n,D,bits,seed", [(1024, 6, 51, 501), (1024, 3, 44, 502), (2048, 5, 33, 503), (4096, 2, 51, 504)])
def test_live_reference_device_code_vs_oracle(n, D, bits, seed):
    """Fresh problems, two items, the second accumulating on top of a non-zero result."""
    p = KsProblem(n, D, 7, 2, bits, seed=seed)
    got = ref_emul.keyswitch(p.result, p.t_target, n, D, 7, p.moduli, p.keys, p.msf, 2)
    assert np.array_equal(got, p.expected().reshape(-1))


@pytest.mark.skipif(not ref_emul.available(), reason="oracle/_ref/ks_ref_emul not built (needs /root/reference)")
def test_live_reference_unreduced_modswitch_factors():
    """modswitch_factors are reduced from [0, 8q) by the host (host/src/fpga.cpp:1057-1061)."""
    n, D = 1024, 4
    p = KsProblem(n, D, 7, 1, 48, seed=77)
    msf = p.msf.copy()
    for i in range(D):
        msf[i] += np.uint64(int(p.moduli[i]) * (i % 7))
    got = ref_emul.keyswitch(p.result, p.t_target, n, D, 7, p.moduli, p.keys, msf, 1)
    assert np.array_equal(got, p.expected().reshape(-1))


def gpu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("fp64", [1, 0])
@pytest.mark.parametrize("c", GOLDEN, ids=IDS)
def test_gpu_device_api_matches_reference_device_code(hb, c, fp64):
    p = problem(c)
    hb.set_option("fp64_path", fp64)
    try:
        plan = hb.KsPlan(p.n, p.D, p.K, p.D + 1, 2, p.moduli, p.keys, p.msf)
    finally:
        hb.set_option("fp64_path", 1)
    res = gpu(p.result)
    plan.keyswitch(res, gpu(p.t_target), p.batch)
    check(res.cpu().numpy().view(np.uint64), c)
    plan.close()


@pytest.mark.gpu
@pytest.mark.parametrize("c", [GOLDEN[0], GOLDEN[1], GOLDEN[6]], ids=[IDS[0], IDS[1], IDS[6]])
def test_gpu_host_api_matches_reference_device_code(acquired, c):
    """The reference test's flow (tests/test_keyswitch.cpp:119-146) through the drop-in host API."""
    hb = acquired
    p = problem(c)
    keys = hb.KeyArray(p.keys)
    out = p.result.copy()
    hb.set_worksize_KeySwitch(p.batch)
    for b in range(p.batch):
        hb.KeySwitch(out[b], p.t_target[b], p.n, p.D, p.K, p.D + 1, 2, p.moduli, keys, p.msf)
    assert hb.KeySwitchCompleted()
    check(out, c)
