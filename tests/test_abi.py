"""The C-ABI library loads, exports every symbol include/hexl_b200.h declares,
validates arguments and fails loudly (never silently computes) without a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hexl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hexl_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(hb):
    lib = hb.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    # and the Python binding knows every one of them
    assert sorted(hb.exported_symbols()) == names


def test_cxx_dropin_exports_reference_symbols(hb):
    """libhexl-fpga.so exports the mangled intel::hexl::* names of the reference
    header (host/inc/hexl-fpga.h:15-161) and its intel::hexl::fpga twin."""
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", hb.CXX_LIB_PATH], capture_output=True, text=True,
                         check=True).stdout
    for sym in ["intel::hexl::acquire_FPGA_resources()", "intel::hexl::release_FPGA_resources()",
                "intel::hexl::set_worksize_DyadicMultiply(unsigned long)",
                "intel::hexl::DyadicMultiply(unsigned long*, unsigned long const*, unsigned long const*, unsigned long, unsigned long const*, unsigned long)",
                "intel::hexl::DyadicMultiplyCompleted()", "intel::hexl::set_worksize_KeySwitch(unsigned long)",
                "intel::hexl::KeySwitch(unsigned long*, unsigned long const*, unsigned long, unsigned long, unsigned long, unsigned long, unsigned long, unsigned long const*, unsigned long const**, unsigned long const*, unsigned long const*)",
                "intel::hexl::KeySwitchCompleted()", "intel::hexl::_set_worksize_NTT(unsigned long)",
                "intel::hexl::_NTT(unsigned long*, unsigned long const*, unsigned long const*, unsigned long, unsigned long)",
                "intel::hexl::_NTTCompleted()", "intel::hexl::_set_worksize_INTT(unsigned long)",
                "intel::hexl::_INTT(unsigned long*, unsigned long const*, unsigned long const*, unsigned long, unsigned long, unsigned long, unsigned long)",
                "intel::hexl::_INTTCompleted()", "intel::hexl::fpga::NTT(", "intel::hexl::fpga::INTT(",
                "intel::hexl::fpga::KeySwitch(", "intel::hexl::fpga::DyadicMultiply("]:
        assert sym in out, sym


def test_product_library_does_not_link_the_oracle(hb):
    out = subprocess.run(["ldd", hb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "hexl_ref" not in out
    dyn = subprocess.run(["nm", "-D", hb.LIB_PATH], capture_output=True, text=True).stdout
    assert "ho_" not in dyn


def test_argument_validation(hb):
    lib = hb.lib()
    assert lib.hexl_b200_ntt_fwd(None, None, None, 97, 1000, 1, None) == -1
    assert "unsupported" in hb.last_error()
    assert lib.hexl_b200_ntt_fwd(16, 16, 16, 1 << 63, 16384, 1, None) == -1
    assert lib.hexl_b200_dyadic_multiply(16, 16, 16, 7, 16, 1, 1, 0, None) == -1     # odd n
    assert lib.hexl_b200_keyswitch(None, 16, 16, 1, None) == -1
    assert lib.hexl_b200_set_option(b"no_such_option", 1) == -1
    h = ctypes.c_void_p()
    m = np.array([97, 193], dtype=np.uint64)
    assert lib.hexl_b200_ks_plan_create(ctypes.byref(h), 1024, 1, 2, 2, 3, m.ctypes.data, m.ctypes.data,
                                        m.ctypes.data, None) == -1   # key_component_count != 2


def test_host_api_requires_acquire(hb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    a = np.zeros(16384, dtype=np.uint64)
    with pytest.raises(hb.HexlB200Error):
        hb.NTT(a, a, a, 97, 16384)              # not acquired
    with pytest.raises(hb.HexlB200Error, match="no CUDA device"):
        hb.acquire_FPGA_resources()             # no GPU here: must fail, not fall back


@pytest.mark.parametrize("n,bits", [(1024, 30), (4096, 51), (16384, 51), (16384, 61)])
def test_product_twiddles_equal_oracle_tables(hb, n, bits):
    """host/src/number_theory.cpp (product) vs the oracle's tables."""
    q = ob.primes(1, bits, n)[0]
    t = ob.Tables(n, q)
    roots, precon, inv_roots, precon_inv, inv_n, inv_n_w = hb.compute_twiddles(n, q)
    assert np.array_equal(roots, t.roots) and np.array_equal(precon, t.precon)
    assert np.array_equal(inv_roots, t.inv_roots) and np.array_equal(precon_inv, t.precon_inv)
    assert (inv_n, inv_n_w) == (t.inv_n, t.inv_n_w)


def test_composite_modulus_is_rejected_quickly(hb):
    """a composite modulus = 1 mod 2n has no use (and once sent the root search on a 2^40-step walk)"""
    import time

    q = 1099511687169           # = 1 mod 2048, composite
    t0 = time.time()
    with pytest.raises(hb.HexlB200Error):
        hb.compute_twiddles(1024, q)
    assert time.time() - t0 < 5
