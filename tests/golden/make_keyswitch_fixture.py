"""Writes tests/golden/keyswitch_1024_2_3_3_2_oracle.json.gz: a vector in the
reference's JSON format, produced by the ORACLE (not by the reference -- its
corpus is external); exercises the loader and the caller-supplied-twiddle path
(tables stored in intel-hexl's 1-based layout, as a hexl dump would have them).
    python tests/golden/make_keyswitch_fixture.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob  # noqa: E402
from keyswitch_vectors import write_vector  # noqa: E402
from ks_util import KsProblem  # noqa: E402

n, D, K = 1024, 2, 3
p = KsProblem(n, D, K, 1, 40, seed=77)
tabs = [ob.Tables(n, int(q)) for q in p.moduli]
tw = {"inv_root_of_unity_powers": [t.inv_roots for t in tabs],
      "precon64_inv_root_of_unity_powers": [t.precon_inv for t in tabs],
      "root_of_unity_powers": [t.roots for t in tabs], "precon64_root_of_unity_powers": [t.precon for t in tabs]}
write_vector(os.path.join(HERE, "keyswitch_1024_2_3_3_2_oracle.json.gz"), n, D, K, p.moduli, p.msf, p.keys,
             p.t_target[0], p.result[0], p.expected()[0], tw)
print("ok")
