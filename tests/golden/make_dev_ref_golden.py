"""Golden answers of the reference's OWN device kernels (device/fwd_ntt.cpp, inv_ntt.cpp,
dyadic_multiply.cpp compiled unmodified into oracle/_ref/dev_ref_emul_*, see oracle/ref_dev_emul.cpp)
for seeded inputs -> tests/golden/dev_ref_emul_golden.json.  Run in the build container:
    python tests/golden/make_dev_ref_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob  # noqa: E402
import ref_emul  # noqa: E402
from dev_cases import NTT_CASES, DYADIC_CASES, ntt_input, dyadic_input  # noqa: E402

out = {"generator": "oracle/_ref/dev_ref_emul_{ntt,intt,dyadic} (reference device code on the CPU)", "ntt": [], "dyadic": []}
for bits, stim in NTT_CASES:
    q = ob.primes(1, bits, 16384)[0]
    t = ob.Tables(16384, q)
    a = ntt_input(stim, q)
    f = ref_emul.fwd_ntt(a, q, t.roots, t.precon)
    i = ref_emul.inv_ntt(a, q, t.inv_n, t.inv_n_w, t.inv_roots, t.precon_inv)
    out["ntt"].append({"bits": bits, "stimulus": stim, "q": q, "fwd_fnv": "%016x" % ob.fnv(f), "inv_fnv": "%016x" % ob.fnv(i)})
    print(out["ntt"][-1])
for n, M, batch, kind in DYADIC_CASES:
    op1, op2, mods = dyadic_input(n, M, batch, kind)
    r = ref_emul.dyadic(op1, op2, n, mods, batch)
    out["dyadic"].append({"n": n, "M": M, "batch": batch, "kind": kind, "fnv": "%016x" % ob.fnv(r)})
    print(out["dyadic"][-1])
with open(os.path.join(HERE, "dev_ref_emul_golden.json"), "w") as fh:
    json.dump(out, fh, indent=1)
