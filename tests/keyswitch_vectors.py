"""Reader / writer for keyswitch test vectors in the reference's JSON format
(tests/test_keyswitch.cpp:55-104: coeff_count, decomp_modulus_size,
key_modulus_size, rns_modulus_size, key_component_count, moduli,
modswitch_factors, [inv_root_of_unity_powers, precon64_inv_root_of_unity_powers,
root_of_unity_powers, precon64_root_of_unity_powers], key_vector,
t_target_iter_ptr, input, expected_output).  The official corpus (testdata.zip,
hexl-fpga release v1.1) is not available offline; point KEYSWITCH_DATA_DIR at an
unpacked copy and the tests pick it up."""
import glob
import gzip
import json
import os

import numpy as np


class KsVector:
    def __init__(self, path):
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rt") as fh:
            js = json.load(fh)
        self.path = path
        self.n = int(js["coeff_count"])
        self.D = int(js["decomp_modulus_size"])
        self.K = int(js["key_modulus_size"])
        self.R = int(js["rns_modulus_size"])
        self.C = int(js["key_component_count"])
        u = lambda x: np.array(x, dtype=np.uint64)
        self.moduli = u(js["moduli"])
        self.msf = u(js["modswitch_factors"])
        names = ["inv_root_of_unity_powers", "precon64_inv_root_of_unity_powers", "root_of_unity_powers",
                 "precon64_root_of_unity_powers"]
        self.twiddles = None
        if all(k in js for k in names):   # [k][i] per modulus, concatenated as the reference test does
            self.twiddles = np.concatenate([np.concatenate([u(js[nm][k][:self.n]) for nm in names])
                                            for k in range(self.K)])
        self.keys = [u(js["key_vector"][k][:2 * self.K * self.n]) for k in range(self.D)]
        self.t_target = u(js["t_target_iter_ptr"])
        self.input = u(js["input"])
        self.expected = u(js["expected_output"])


def write_vector(path, n, D, K, moduli, msf, keys, t_target, inp, expected, twiddle_tables=None):
    js = {"coeff_count": n, "decomp_modulus_size": D, "key_modulus_size": K, "rns_modulus_size": D + 1,
          "key_component_count": 2, "moduli": [int(x) for x in moduli],
          "modswitch_factors": [int(x) for x in msf],
          "key_vector": [[int(x) for x in k] for k in keys],
          "t_target_iter_ptr": [int(x) for x in t_target], "input": [int(x) for x in inp],
          "expected_output": [int(x) for x in expected]}
    if twiddle_tables is not None:
        for name, tab in twiddle_tables.items():
            js[name] = [[int(x) for x in row] for row in tab]
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt") as fh:
        json.dump(js, fh)


def find_vectors():
    here = os.path.dirname(os.path.abspath(__file__))
    files = sorted(glob.glob(os.path.join(here, "golden", "keyswitch_*.json*")))
    d = os.environ.get("KEYSWITCH_DATA_DIR")
    if d:
        files += sorted(glob.glob(os.path.join(d, "*.json")))
    return files
