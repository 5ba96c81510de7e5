"""GPU parity of the fused polynomial multiply:
result = a * b mod (x^n + 1, q), computed as INTT(NTT(a) (.) NTT(b)) with the
dyadic product fused into the inverse transform's first pass.  Checked against
(i) a schoolbook negacyclic convolution and (ii) the oracle's NTT pipeline."""
import numpy as np
import pytest

import oracle_binding as ob

pytestmark = pytest.mark.gpu


def gpu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_np(t):
    return t.cpu().numpy().view(np.uint64)


def tables(hb, t):
    return gpu(t.roots), gpu(t.precon), gpu(t.inv_roots), gpu(t.precon_inv)


def run(hb, a, b, t, alias=False):
    import torch

    da, db = gpu(a), gpu(b)
    res = da if alias else torch.zeros_like(da)
    r, p, ir, ip = tables(hb, t)
    hb.poly_multiply(res, da, db, r, p, ir, ip, t.q, t.inv_n, t.inv_n_w, t.n)
    return to_np(res)


def schoolbook(a, b, q):
    """negacyclic convolution with int64 arithmetic (q < 2^20, n <= 4096)"""
    n = len(a)
    c = np.convolve(a.astype(np.int64), b.astype(np.int64))
    out = c[:n].copy()
    out[:n - 1] -= c[n:]
    return (out % q).astype(np.uint64)


def oracle_pipeline(a, b, t):
    fa, fb = ob.fwd_ntt(a, t), ob.fwd_ntt(b, t)
    prod = np.array([(int(x) % t.q) * (int(y) % t.q) % t.q for x, y in zip(fa, fb)], dtype=np.uint64)
    return ob.inv_ntt(prod, t)


@pytest.mark.parametrize("n", [1024, 4096])
def test_against_schoolbook(hb, n):
    q = ob.primes(1, 18, n)[0]
    t = ob.Tables(n, q)
    a = np.stack([ob.splitmix(n, 10 + i, q) for i in range(5)])
    b = np.stack([ob.splitmix(n, 90 + i, q) for i in range(5)])
    got = run(hb, a, b, t)
    for i in range(5):
        assert np.array_equal(got[i], schoolbook(a[i], b[i], q)), i


@pytest.mark.parametrize("n,bits", [(16384, 51), (16384, 27), (8192, 45), (2048, 60)])
def test_against_oracle_pipeline(hb, n, bits):
    q = ob.primes(1, bits, n)[0]
    t = ob.Tables(n, q)
    a = np.stack([ob.splitmix(n, 3 + i, q) for i in range(3)])
    b = np.stack([ob.splitmix(n, 7 + i, q) for i in range(3)])
    b[1, :] = 0
    b[1, 1] = 1                                   # multiply by x: negacyclic shift
    a[2] = ob.splitmix(n, 55, 0)                  # out-of-contract words: the exact forward kernel
    got = run(hb, a, b, t)
    for i in range(3):
        assert np.array_equal(got[i], oracle_pipeline(a[i], b[i], t)), i
    shifted = np.empty(n, dtype=np.uint64)
    shifted[0] = (q - int(a[1, -1])) % q
    shifted[1:] = a[1, :-1]
    assert np.array_equal(got[1], shifted)
    # result may alias the first operand
    assert np.array_equal(run(hb, a, b, t, alias=True), got)


def test_chunked_batch_commutes_and_has_identity(hb):
    """batch above the 2048-polynomial scratch chunk; a*b == b*a and a*1 == a at scale"""
    import torch

    n, B = 1024, 2500
    q = ob.primes(1, 51, n)[0]
    t = ob.Tables(n, q)
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(0, q, (B, n), dtype=torch.int64, device="cuda", generator=g)
    b = torch.randint(0, q, (B, n), dtype=torch.int64, device="cuda", generator=g)
    one = torch.zeros_like(a)
    one[:, 0] = 1
    r, p, ir, ip = tables(hb, t)
    ab, ba, a1 = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
    hb.poly_multiply(ab, a, b, r, p, ir, ip, q, t.inv_n, t.inv_n_w, n)
    hb.poly_multiply(ba, b, a, r, p, ir, ip, q, t.inv_n, t.inv_n_w, n)
    hb.poly_multiply(a1, a, one, r, p, ir, ip, q, t.inv_n, t.inv_n_w, n)
    assert torch.equal(ab, ba)
    assert torch.equal(a1, a)
    for i in (0, 2047, 2048, B - 1):
        assert np.array_equal(to_np(ab[i]), oracle_pipeline(to_np(a[i]), to_np(b[i]), t)), i


def test_bad_arguments(hb):
    import torch

    n, q = 1024, ob.primes(1, 30, 1024)[0]
    t = ob.Tables(n, q)
    x = torch.zeros((1, n), dtype=torch.int64, device="cuda")
    r, p, ir, ip = tables(hb, t)
    f = hb.lib().hexl_b200_poly_multiply
    args = lambda res, nn: (res, x.data_ptr(), x.data_ptr(), r.data_ptr(), p.data_ptr(), ir.data_ptr(),
                            ip.data_ptr(), q, t.inv_n, t.inv_n_w, nn, 1, None)
    assert f(*args(x.data_ptr(), n)) == 0
    assert f(*args(x.data_ptr() + 8, n)) == -1      # misaligned
    assert f(*args(x.data_ptr(), 1000)) == -1       # not a power of two
    assert f(*args(None, n)) == -1


@pytest.mark.parametrize("bits", [51, 40, 36])
def test_single_launch_kernel_matches_three_launch_path(hb, bits):
    """N = 16384, FP64-contract modulus: one launch per chunk (NTT(a) parked in tensor memory, NTT(b), product,
    INTT -- polymul_fused.cu) vs the three-launch version, vs the oracle pipeline; items with out-of-contract
    words in a, in b, or in both go to the exact kernels through the deferred list; aliasing result = a."""
    n = 16384
    q = ob.primes(1, bits, n)[0]
    t = ob.Tables(n, q)
    B = 300                                  # more items than CTAs: the persistent loop wraps
    a = np.stack([ob.splitmix(n, 1000 + i, q) for i in range(12)])
    b = np.stack([ob.splitmix(n, 2000 + i, q) for i in range(12)])
    a = np.ascontiguousarray(np.resize(a, (B, n)))
    b = np.ascontiguousarray(np.resize(b, (B, n)))
    a[:, 1] = np.arange(B, dtype=np.uint64)
    a[5, 77] = np.uint64(2**64 - 1)          # a out of contract
    b[9, 0] = np.uint64(3 * q + 1)           # b out of contract
    a[160] = np.uint64(2**63 + 11)           # both, in the second round of the persistent loop
    b[160, 5] = np.uint64(2**64 - 3)
    a[299, 16383] = np.uint64(2 * q)         # last item
    outs = {}
    for fused in (1, 0):
        hb.set_option("polymul_fused", fused)
        try:
            outs[fused] = run(hb, a, b, t)
            if fused:
                alias = run(hb, a, b, t, alias=True)
        finally:
            hb.set_option("polymul_fused", 1)
    assert np.array_equal(outs[1], outs[0])
    assert np.array_equal(alias, outs[1])
    for i in (0, 1, 147, 148, 149, 298):
        assert np.array_equal(outs[1][i], oracle_pipeline(a[i], b[i], t)), i
