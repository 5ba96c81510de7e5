"""Generates tests/golden/ntt_golden.json from the REFERENCE's own scalar NTT
(oracle/_ref/libhexl_ref.so = /root/reference/tests/test_utils/ntt.cpp compiled
unmodified by oracle/Makefile).  Run in the build container (the reference tree
does not exist on the GPU box); the JSON is committed.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob  # noqa: E402


def stim(kind, n, q, seed):
    if kind == "random":
        return ob.splitmix(n, seed, q)
    if kind == "ramp":
        return np.arange(n, dtype=np.uint64)
    if kind == "ones":
        return np.ones(n, dtype=np.uint64)
    if kind == "impulse":
        a = np.zeros(n, dtype=np.uint64)
        a[0] = 1
        return a
    if kind == "all_max":
        return np.full(n, 2**64 - 1, dtype=np.uint64)
    if kind == "garbage":
        return ob.splitmix(n, seed, 0)
    raise ValueError(kind)


def main():
    r = ob.ref()
    assert r is not None, "build oracle/_ref first (make -C oracle)"
    cases = []
    for n in (1024, 4096, 16384):
        for bits in (20, 32, 51, 55, 61):
            primes = np.zeros(1, dtype=np.uint64)
            assert r.ref_generate_primes(ob.P(primes), 1, bits, n) == 1
            q = int(primes[0])
            w = int(r.ref_min_primitive_root(2 * n, q))
            tabs = [np.zeros(n, dtype=np.uint64) for _ in range(4)]
            r.ref_tables(n, q, *[ob.P(t) for t in tabs])
            case = {"n": n, "bits": bits, "q": q, "root": w,
                    "inv_n": int(r.ref_inverse_mod(n, q)),
                    "tables_fnv": [f"{ob.fnv(t):016x}" for t in tabs],
                    "roots_head": [int(x) for x in tabs[0][:4]],
                    "inv_roots_head": [int(x) for x in tabs[2][:4]],
                    "inv_roots_last": int(tabs[2][n - 1]),
                    "vectors": []}
            for k, kind in enumerate(["random", "ramp", "ones", "impulse", "all_max", "garbage"]):
                seed = 1000 * bits + n + k
                a = stim(kind, n, q, seed)
                f = a.copy()
                r.ref_fwd_ntt(ob.P(f), n, q)
                i = a.copy()
                r.ref_inv_ntt(ob.P(i), n, q)
                case["vectors"].append({
                    "kind": kind, "seed": seed, "in_fnv": f"{ob.fnv(a):016x}",
                    "fwd_fnv": f"{ob.fnv(f):016x}", "fwd_head": [int(x) for x in f[:4]],
                    "inv_fnv": f"{ob.fnv(i):016x}", "inv_head": [int(x) for x in i[:4]]})
            cases.append(case)
    # one tiny fully spelled-out vector (n = 16 is below the kernels' range but
    # pins the oracle word for word)
    n, q = 16, 97
    a = np.arange(1, n + 1, dtype=np.uint64)
    f = a.copy()
    r.ref_fwd_ntt(ob.P(f), n, q)
    small = {"n": n, "q": q, "in": [int(x) for x in a], "fwd": [int(x) for x in f]}
    with open(os.path.join(HERE, "ntt_golden.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py via oracle/_ref (reference tests/test_utils/ntt.cpp)",
                   "hash": "64-bit FNV-1a over little-endian words", "cases": cases, "small": small}, fh, indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
