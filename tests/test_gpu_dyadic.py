"""GPU parity of DyadicMultiply.  Mirrors tests/test_dyadic_multiply.cpp:32-265
of the reference: its deterministic inputs (composite even moduli (b+m+1)*10,
unreduced operands) and shapes, plus random 52-bit prime cases vs the oracle."""
import numpy as np
import pytest

import oracle_binding as ob

pytestmark = pytest.mark.gpu


def reference_io(num, M, n):
    """setup_dyadic_io of tests/test_dyadic_multiply.cpp:32-86 (closed form)."""
    moduli = np.zeros(num * M, dtype=np.uint64)
    op1 = np.zeros((num, 2, M, n), dtype=np.uint64)
    op2 = np.zeros((num, 2, M, n), dtype=np.uint64)
    i = np.arange(n, dtype=np.uint64)
    for b in range(num):
        for m in range(M):
            moduli[b * M + m] = (b + m + 1) * 10
            op1[b, 0, m] = b + i + 1 + m * n
            op2[b, 0, m] = b + i + 2 + m * n
            op1[b, 1, m] = b + i + 11 + m * n
            op2[b, 1, m] = b + i + 22 + m * n
    exp = np.zeros((num, 3, M, n), dtype=np.uint64)
    for b in range(num):
        for m in range(M):
            q = moduli[b * M + m]
            exp[b, 0, m] = (op1[b, 0, m] * op2[b, 0, m]) % q
            exp[b, 1, m] = (op1[b, 0, m] * op2[b, 1, m] + op1[b, 1, m] * op2[b, 0, m]) % q
            exp[b, 2, m] = (op1[b, 1, m] * op2[b, 1, m]) % q
    return moduli, op1, op2, exp


def gpu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


# (num_dyadic_multiply, num_moduli, coeff_count) from test_dyadic_multiply.cpp:149-265
REF_SHAPES = [(2, 1, 1024), (3, 2, 2048), (2, 4, 4096), (4, 7, 8192), (2, 14, 16384), (16, 3, 256),
              (5, 2, 32768)]


@pytest.mark.parametrize("num,M,n", REF_SHAPES)
def test_reference_inputs_device_api(hb, num, M, n):
    import torch

    moduli, op1, op2, exp = reference_io(num, M, n)
    res = torch.zeros(num * 3 * M * n, dtype=torch.int64, device="cuda")
    hb.dyadic_multiply(res, gpu(op1), gpu(op2), n, gpu(moduli), M, num, moduli_per_item=True)
    assert np.array_equal(res.cpu().numpy().view(np.uint64), exp.reshape(-1))
    # the oracle agrees with the closed form too
    assert np.array_equal(ob.dyadic(op1.reshape(-1), op2.reshape(-1), n, moduli, num, True), exp.reshape(-1))


@pytest.mark.parametrize("bits", [30, 51, 61])
def test_random_primes_vs_oracle(hb, bits):
    import torch

    n, M, B = 8192, 4, 3
    moduli = np.array(ob.primes(M, bits, n), dtype=np.uint64)
    op1 = np.stack([ob.splitmix(n, 11 + 31 * b + k, int(moduli[k % M])) for b in range(B) for k in range(2 * M)])
    op2 = np.stack([ob.splitmix(n, 977 + 31 * b + k, int(moduli[k % M])) for b in range(B) for k in range(2 * M)])
    res = torch.zeros(B * 3 * M * n, dtype=torch.int64, device="cuda")
    hb.dyadic_multiply(res, gpu(op1), gpu(op2), n, gpu(moduli), M, B)
    assert np.array_equal(res.cpu().numpy().view(np.uint64), ob.dyadic(op1.reshape(-1), op2.reshape(-1), n, moduli, B))


def test_unreduced_and_extreme_moduli(hb):
    import torch

    n, M, B = 1024, 8, 2
    # 2^63 is where the kernel switches from the lazily summed cross term to two
    # separate reductions: cover both sides of it
    moduli = np.array([1, 2, 2**63 + 29, 2**64 - 59, 10, 2**63 - 25, 2**63 - 1, 2**63], dtype=np.uint64)
    op1 = ob.splitmix(B * 2 * M * n, 5, 0)
    op2 = ob.splitmix(B * 2 * M * n, 6, 0)
    res = torch.zeros(B * 3 * M * n, dtype=torch.int64, device="cuda")
    hb.dyadic_multiply(res, gpu(op1), gpu(op2), n, gpu(moduli), M, B)
    assert np.array_equal(res.cpu().numpy().view(np.uint64), ob.dyadic(op1, op2, n, moduli, B))


def test_host_api_reference_flow(acquired):
    """test_dyadic_multiply of the reference (tests/test_dyadic_multiply.cpp:
    88-109) through set_worksize / DyadicMultiply / Completed."""
    hb = acquired
    num, M, n = 6, 4, 4096
    moduli, op1, op2, exp = reference_io(num, M, n)
    out = np.zeros((num, 3 * M * n), dtype=np.uint64)
    op1 = op1.reshape(num, -1)
    op2 = op2.reshape(num, -1)
    hb.set_worksize_DyadicMultiply(num)
    for b in range(num):
        hb.DyadicMultiply(out[b], op1[b], op2[b], n, moduli[b * M:(b + 1) * M], M)
    assert hb.DyadicMultiplyCompleted()
    assert np.array_equal(out.reshape(-1), exp.reshape(-1))
