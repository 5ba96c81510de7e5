"""Keyswitch vectors in the reference's JSON format (tests/test_keyswitch.cpp:
55-146): the committed oracle-generated fixture, plus the official corpus when
KEYSWITCH_DATA_DIR points at it.  CPU: loader + oracle; GPU: the reference
test's exact flow through the host API."""
import numpy as np
import pytest

import oracle_binding as ob
from keyswitch_vectors import KsVector, find_vectors

FILES = find_vectors()


@pytest.mark.parametrize("path", FILES)
def test_oracle_reproduces_expected_output(path):
    v = KsVector(path)
    assert v.R == v.D + 1 and v.C == 2
    got = ob.keyswitch(v.input, v.t_target, v.n, v.D, v.K, v.moduli, v.keys, v.msf, 1)
    assert np.array_equal(got, v.expected)


@pytest.mark.gpu
def test_reference_flow_on_gpu(acquired):
    """test_KeySwitch of the reference (tests/test_keyswitch.cpp:119-146): all
    vectors of one shape in one worksize, keys / moduli / twiddles of the first."""
    hb = acquired
    vecs = [KsVector(p) for p in FILES]
    shapes = sorted({(v.n, v.D, v.K) for v in vecs})
    for shape in shapes:
        group = [v for v in vecs if (v.n, v.D, v.K) == shape]
        v0 = group[0]
        keys = hb.KeyArray(v0.keys)
        outs = [v.input.copy() for v in group]
        hb.set_worksize_KeySwitch(len(group))
        for v, o in zip(group, outs):
            hb.KeySwitch(o, v.t_target, v0.n, v0.D, v0.K, v0.R, v0.C, v0.moduli, keys, v0.msf, v0.twiddles)
        assert hb.KeySwitchCompleted()
        for v, o in zip(group, outs):
            assert np.array_equal(o, v.expected), v.path
