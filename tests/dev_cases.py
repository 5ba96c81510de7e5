"""Seeded inputs shared by tests/golden/make_dev_ref_golden.py and tests/test_ref_device_emul.py.
Stimuli and shapes follow the reference's own tests: tests/test_fwd_ntt.cpp:15-62,119-170 (RANDOM, RAMP,
ALL_ZEROS, ALL_ONES, IMPULSE, ALL_MAX_VALUES; prime sizes 20/32/55/62 plus BASELINE's 52-bit class) and
tests/test_dyadic_multiply.cpp:35-52 (composite even moduli (b+m+1)*10, unreduced operands)."""
import numpy as np

import oracle_binding as ob

N = 16384
NTT_CASES = [(bits, stim) for bits in (20, 32, 51, 55, 61) for stim in ("random", "ramp", "ones", "impulse", "max")]
DYADIC_CASES = [(1024, 1, 2, "reftest"), (4096, 4, 3, "reftest"), (8192, 7, 2, "reftest"), (16384, 2, 2, "reftest"),
                (8192, 4, 2, "prime51"), (2048, 14, 1, "prime51")]


def ntt_input(stim, q, batch=2):
    a = np.zeros((batch, N), dtype=np.uint64)
    for b in range(batch):
        if stim == "random":
            a[b] = ob.splitmix(N, 41 + b, q)
        elif stim == "ramp":
            a[b] = (np.arange(N, dtype=np.uint64) + np.uint64(b)) % np.uint64(q)
        elif stim == "ones":
            a[b] = 1
        elif stim == "impulse":
            a[b, b] = 1
        elif stim == "max":
            a[b] = np.uint64(2**64 - 1)      # out of contract on purpose (tests/test_fwd_ntt.cpp:139-147)
    return a


def dyadic_input(n, M, batch, kind):
    if kind == "reftest":
        mods = np.array([[(b + m + 1) * 10 for m in range(M)] for b in range(batch)], dtype=np.uint64)
        op1 = np.zeros((batch, 2, M, n), dtype=np.uint64)
        op2 = np.zeros_like(op1)
        for b in range(batch):
            for m in range(M):
                op1[b, 0, m] = np.arange(n, dtype=np.uint64) + np.uint64(b + 1 + m * n)
                op1[b, 1, m] = np.arange(n, dtype=np.uint64) + np.uint64(b + 2 + m * n)
                op2[b, 0, m] = np.arange(n, dtype=np.uint64) + np.uint64(b + 3 + m * n)
                op2[b, 1, m] = np.arange(n, dtype=np.uint64) + np.uint64(b + 4 + m * n)
        return op1.reshape(batch, -1), op2.reshape(batch, -1), mods
    mods = np.array([ob.primes(M, 51, n)] * batch, dtype=np.uint64)
    op1 = np.stack([ob.splitmix(2 * M * n, 3 + b, int(mods[0, 0])) for b in range(batch)])
    op2 = np.stack([ob.splitmix(2 * M * n, 30 + b, int(mods[0, 0])) for b in range(batch)])
    return op1, op2, mods
