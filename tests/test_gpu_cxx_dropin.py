"""The C++ drop-in boundary, EXECUTED: tests/cpp/dropin_driver.cpp is a caller written against the
reference's own unmodified header (host/inc/hexl-fpga.h, -I/root/reference/host/inc at build time) and
linked to our libhexl-fpga.so.  It replays the call sequences of the reference's tests
(tests/test_fwd_ntt.cpp:97-117, test_inv_ntt.cpp:97-125, test_dyadic_multiply.cpp:88-109,
test_keyswitch.cpp:119-146); its output files are compared with the oracle here."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_binding as ob
from dev_cases import dyadic_input
from ks_util import KsProblem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "cpp", "_build", "dropin_driver")


def run(mode, header, arrays, check=True):
    parts = [np.array(header, dtype=np.uint64)] + [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1) for a in arrays]
    with tempfile.TemporaryDirectory() as d:
        pin, pout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        np.concatenate(parts).tofile(pin)
        r = subprocess.run([DRIVER, mode, pin, pout], capture_output=True, text=True, timeout=600)
        if check:
            assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
            return np.fromfile(pout, dtype=np.uint64)
        return r


def test_driver_is_built_against_the_reference_header():
    if not os.path.exists(DRIVER):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    assert os.path.exists(DRIVER)
    with open(os.path.join(os.path.dirname(DRIVER), "header_used.txt")) as fh:
        used = fh.read().strip()
    # in the build container the header is the reference's; elsewhere the prebuilt binary is used
    if os.path.exists("/root/reference/host/inc/hexl-fpga.h"):
        assert used == "/root/reference/host/inc"
    out = subprocess.run(["nm", "-D", "--undefined-only", "-C", DRIVER], capture_output=True, text=True).stdout
    for sym in ("intel::hexl::_NTT(", "intel::hexl::_INTT(", "intel::hexl::DyadicMultiply(", "intel::hexl::KeySwitch(",
                "intel::hexl::acquire_FPGA_resources()", "intel::hexl::KeySwitchCompleted()"):
        assert sym in out, sym


def test_driver_fails_loudly_without_a_gpu():
    """No CPU fallback behind the drop-in: without a CUDA device acquire_FPGA_resources() aborts."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run("ntt", [1, 1024, 12289], [np.zeros(3 * 1024, dtype=np.uint64)], check=False)
    assert r.returncode != 0 and "acquire_FPGA_resources" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n,bits,batch", [(16384, 51, 5), (16384, 28, 3), (4096, 60, 2)])
def test_ntt_and_intt_flow(n, bits, batch):
    q = ob.primes(1, bits, n)[0]
    t = ob.Tables(n, q)
    a = np.stack([ob.splitmix(n, 70 + i, q) for i in range(batch)])
    got = run("ntt", [batch, n, q], [t.roots, t.precon, a]).reshape(batch, n)
    assert all(np.array_equal(got[i], ob.fwd_ntt(a[i], t)) for i in range(batch))
    back = run("intt", [batch, n, q, t.inv_n, t.inv_n_w], [t.inv_roots, t.precon_inv, got]).reshape(batch, n)
    assert np.array_equal(back, a)


@pytest.mark.gpu
@pytest.mark.parametrize("n,M,batch,kind", [(4096, 4, 3, "reftest"), (8192, 7, 2, "reftest"), (8192, 4, 2, "prime51")])
def test_dyadic_flow(n, M, batch, kind):
    op1, op2, mods = dyadic_input(n, M, batch, kind)
    got = run("dyadic", [batch, n, M], [mods, op1, op2])
    assert np.array_equal(got, ob.dyadic(op1.reshape(-1), op2.reshape(-1), n, mods.reshape(-1), batch, per_item=True))


@pytest.mark.gpu
@pytest.mark.parametrize("n,D,K,batch", [(16384, 6, 7, 3), (8192, 5, 7, 2), (16384, 7, 8, 2)])
def test_keyswitch_flow(n, D, K, batch):
    p = KsProblem(n, D, K, batch, 51, seed=31)
    got = run("keyswitch", [batch, n, D, K], [p.moduli, p.msf] + p.keys + [p.t_target, p.result])
    assert np.array_equal(got, p.expected().reshape(-1))
