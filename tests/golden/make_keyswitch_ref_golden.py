"""Golden keyswitch answers produced by the REFERENCE'S OWN device code (device/keyswitch.cpp and
device/keyswitch/*.hpp compiled unmodified into oracle/_ref/ks_ref_emul, see oracle/ref_ks_emul.cpp),
run here in the build container (it needs /root/reference to be built):

    python tests/golden/make_keyswitch_ref_golden.py

Writes
  tests/golden/ks_ref_emul_golden.json            per case: shape, seed, FNV-1a of the reference result and its
                                                  first words (inputs are the seeded KsProblem of tests/ks_util.py)
  tests/golden/keyswitch_1024_5_7_6_2_refemul.json.gz   one full vector in the reference's JSON format
                                                  (tests/test_keyswitch.cpp:55-104), expected_output = reference
Shapes are the bitstream's limits (K = 7, D <= 6): the reference's own test shapes 6/7/7/2 and 5/7/6/2
at N = 16384 and 8192 (tests/test_keyswitch.cpp:148-191, tests/micro_keyswitch.sh:22-34) plus small ones."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob  # noqa: E402
import ref_emul  # noqa: E402
from keyswitch_vectors import write_vector  # noqa: E402
from ks_util import KsProblem  # noqa: E402

CASES = [  # n, D, K, batch, prime bits, seed
    (16384, 6, 7, 2, 51, 1234), (16384, 5, 7, 2, 51, 4321), (8192, 6, 7, 1, 51, 11), (8192, 5, 7, 1, 45, 12),
    (4096, 4, 7, 2, 40, 13), (2048, 3, 7, 1, 30, 14), (1024, 6, 7, 3, 51, 15), (1024, 1, 7, 2, 36, 16),
    (1024, 2, 7, 1, 20, 17),
]

assert ref_emul.available(), "build oracle/_ref first (make -C oracle)"
out = []
for n, D, K, batch, bits, seed in CASES:
    p = KsProblem(n, D, K, batch, bits, seed=seed)
    got = ref_emul.keyswitch(p.result, p.t_target, n, D, K, p.moduli, p.keys, p.msf, batch)
    out.append({"n": n, "D": D, "K": K, "batch": batch, "bits": bits, "seed": seed,
                "fnv": "%016x" % ob.fnv(got), "head": [int(x) for x in got[:4]], "tail": [int(x) for x in got[-2:]]})
    print(out[-1])
with open(os.path.join(HERE, "ks_ref_emul_golden.json"), "w") as fh:
    json.dump({"generator": "oracle/_ref/ks_ref_emul (reference device/keyswitch.cpp on the CPU)", "cases": out}, fh,
              indent=1)

n, D, K = 1024, 5, 7
p = KsProblem(n, D, K, 1, 30, seed=99)
exp = ref_emul.keyswitch(p.result, p.t_target, n, D, K, p.moduli, p.keys, p.msf, 1)
write_vector(os.path.join(HERE, "keyswitch_1024_5_7_6_2_refemul.json.gz"), n, D, K, p.moduli, p.msf, p.keys,
             p.t_target[0], p.result[0], exp)
print("ok")
