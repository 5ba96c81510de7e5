"""Forward NTT, inverse NTT and dyadic multiply pinned on the REFERENCE'S OWN device kernels.

oracle/_ref/dev_ref_emul_{ntt,intt,dyadic} are device/fwd_ntt.cpp, device/inv_ntt.cpp and
device/dyadic_multiply.cpp of the reference, compiled UNMODIFIED for the CPU (oracle/sycl_shim in place
of the oneAPI FPGA emulator; oracle/ref_dev_emul.cpp, oracle/Makefile).  Their answers for the seeded
inputs of tests/dev_cases.py -- the reference tests' stimuli, 2^64-1 garbage included, and its
composite-modulus dyadic inputs -- are committed in tests/golden/dev_ref_emul_golden.json.

  CPU: the oracle reproduces every golden answer (and, when the binaries are present, live runs);
  GPU: the CUDA kernels reproduce the same answers through the C ABI."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
import ref_emul
from dev_cases import N, dyadic_input, ntt_input

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "dev_ref_emul_golden.json")) as fh:
    GOLDEN = json.load(fh)
NTT_IDS = ["b%d_%s" % (c["bits"], c["stimulus"]) for c in GOLDEN["ntt"]]
DY_IDS = ["n%d_M%d_%s" % (c["n"], c["M"], c["kind"]) for c in GOLDEN["dyadic"]]


def h(a):
    return "%016x" % ob.fnv(np.ascontiguousarray(a, dtype=np.uint64).reshape(-1))


@pytest.mark.parametrize("c", GOLDEN["ntt"], ids=NTT_IDS)
def test_oracle_ntt_matches_reference_device_kernels(c):
    q = c["q"]
    assert ob.primes(1, c["bits"], N)[0] == q
    t = ob.Tables(N, q)
    a = ntt_input(c["stimulus"], q)
    assert h(np.stack([ob.fwd_ntt(x, t) for x in a])) == c["fwd_fnv"]
    assert h(np.stack([ob.inv_ntt(x, t) for x in a])) == c["inv_fnv"]


@pytest.mark.parametrize("c", GOLDEN["dyadic"], ids=DY_IDS)
def test_oracle_dyadic_matches_reference_device_kernel(c):
    op1, op2, mods = dyadic_input(c["n"], c["M"], c["batch"], c["kind"])
    assert h(ob.dyadic(op1.reshape(-1), op2.reshape(-1), c["n"], mods.reshape(-1), c["batch"], per_item=True)) == c["fnv"]


@pytest.mark.skipif(not ref_emul.device_available("ntt"), reason="oracle/_ref/dev_ref_emul_ntt not built")
def test_live_reference_device_ntt_roundtrip():
    q = ob.primes(3, 48, N)[2]
    t = ob.Tables(N, q)
    a = np.stack([ob.splitmix(N, 900 + i, q) for i in range(3)])
    f = ref_emul.fwd_ntt(a, q, t.roots, t.precon)
    assert all(np.array_equal(f[i], ob.fwd_ntt(a[i], t)) for i in range(3))
    assert np.array_equal(ref_emul.inv_ntt(f, q, t.inv_n, t.inv_n_w, t.inv_roots, t.precon_inv), a)


@pytest.mark.skipif(not ref_emul.device_available("dyadic"), reason="oracle/_ref/dev_ref_emul_dyadic not built")
def test_live_reference_device_dyadic():
    op1, op2, mods = dyadic_input(512, 5, 4, "reftest")
    got = ref_emul.dyadic(op1, op2, 512, mods, 4)
    assert np.array_equal(got, ob.dyadic(op1.reshape(-1), op2.reshape(-1), 512, mods.reshape(-1), 4, per_item=True))


def gpu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("c", GOLDEN["ntt"], ids=NTT_IDS)
def test_gpu_ntt_matches_reference_device_kernels(hb, c):
    q = c["q"]
    t = ob.Tables(N, q)
    a = ntt_input(c["stimulus"], q)
    d = gpu(a)
    hb.ntt_fwd(d, gpu(t.roots), gpu(t.precon), q, N)
    assert h(d.cpu().numpy().view(np.uint64)) == c["fwd_fnv"]
    d = gpu(a)
    hb.ntt_inv(d, gpu(t.inv_roots), gpu(t.precon_inv), q, t.inv_n, t.inv_n_w, N)
    assert h(d.cpu().numpy().view(np.uint64)) == c["inv_fnv"]


@pytest.mark.gpu
@pytest.mark.parametrize("c", GOLDEN["dyadic"], ids=DY_IDS)
def test_gpu_dyadic_matches_reference_device_kernel(hb, c):
    import torch

    op1, op2, mods = dyadic_input(c["n"], c["M"], c["batch"], c["kind"])
    res = torch.zeros(c["batch"] * 3 * c["M"] * c["n"], dtype=torch.int64, device="cuda")
    hb.dyadic_multiply(res, gpu(op1), gpu(op2), c["n"], gpu(mods), c["M"], c["batch"], moduli_per_item=True)
    assert h(res.cpu().numpy().view(np.uint64)) == c["fnv"]
