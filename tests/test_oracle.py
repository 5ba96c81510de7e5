"""CPU tests pinning the oracle (oracle/hexl_oracle.c):
  * SURVEY.md Appendix B known answers,
  * golden fixtures produced by the reference's own scalar NTT
    (tests/golden/ntt_golden.json, generator committed beside it),
  * live comparison with oracle/_ref when that library is present,
  * the closed-form dyadic expectation of the reference's tests,
  * keyswitch: two independent restatements agree, and an RLWE decryption
    self-test validates the formula itself (rounding, fix, sign, modswitch).
"""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
N = 16384
P8 = [2251799814045697, 2251799814799361, 2251799814930433, 2251799815094273,
      2251799815487489, 2251799815520257, 2251799816273921, 2251799816568833]


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


def test_appendix_b_constants():
    o = ob.oracle()
    assert ob.primes(8, 51, 16384) == P8
    roots = [25432709486, 280013155948, 307952950864, 26468127340, 209858860512, 72703961923,
             71747157513, 25573066223]
    inv_n = [2251662375092203, 2251662375845821, 2251662375976885, 2251662376140715,
             2251662376533907, 2251662376566673, 2251662377320291, 2251662377615185]
    for q, w, i in zip(P8, roots, inv_n):
        assert o.ho_min_primitive_root(2 * N, q) == w
        assert o.ho_inv_mod(N, q) == i
    assert ob.primes(4, 51, 8192) == [2251799814045697, 2251799814291457, 2251799814356993, 2251799814799361]
    t = ob.Tables(N, P8[0])
    assert [int(x) for x in t.roots[:4]] == [1, 1111640190223217, 678197397923777, 1057859029963613]
    assert [int(x) for x in t.inv_roots[:4]] == [1, 764937211625596, 2077271659520341, 1688001479666130]
    assert int(t.inv_roots[N - 1]) == 1140159623822480


def test_kat1_splitmix_52bit():
    t = ob.Tables(N, P8[0])
    a = ob.splitmix(N, 1, t.q)
    assert [int(x) for x in a[:3]] == [613442214742688, 1000147061265546, 1023569363416652]
    assert ob.fnv(a) == 0x1582CC4702BCF720
    f = ob.fwd_ntt(a, t)
    assert [int(x) for x in f[:3]] == [1955457978075445, 1092550199427436, 1923103082448610]
    assert ob.fnv(f) == 0x428B5C898DD187A3
    assert np.array_equal(ob.inv_ntt(f, t), a)


@pytest.mark.parametrize("bits,q,w,head,h", [
    (20, 1146881, 53, [348330, 392823, 713595, 912732], 0xBB10994907704BC4),
    (32, 4295294977, 280141, [2873566674, 2398889886, 607636391, 3061699463], 0xBD3ECAEEF0141E84),
    (52, 4503599627763713, 51902047037, [1661452251559784, 898320242551355, 838956788285548, 2067457885824797],
     0x64381823DFF21901),
    (55, 36028797019389953, 1256158037438, [2479348741857115, 10095325432057649, 24900993376251465,
                                            34421326335434761], 0xFAE12F2917FA203E),
    (62, 4611686018428010497, 57381806132760, [3618466054774408537, 3224252691126909885, 3753730892694855808,
                                               3606789851836976590], 0x158850D062B9DC57),
])
def test_kat2_ramp(bits, q, w, head, h):
    assert ob.primes(1, bits, N)[0] == q
    t = ob.Tables(N, q)
    assert t.w == w
    f = ob.fwd_ntt(np.arange(N, dtype=np.uint64), t)
    assert [int(x) for x in f[:4]] == head
    assert ob.fnv(f) == h


def test_kat3_all_max_is_not_reduced():
    t = ob.Tables(N, 4503599627763713)
    f = ob.fwd_ntt(np.full(N, 2**64 - 1, dtype=np.uint64), t)
    assert [int(x) for x in f[:2]] == [18373839436896342053, 18371929644362097091]
    assert ob.fnv(f) == 0x47244FE2BFC8E399


def test_golden_from_reference():
    with open(os.path.join(HERE, "golden", "ntt_golden.json")) as fh:
        g = json.load(fh)
    o = ob.oracle()
    for case in g["cases"]:
        n, q = case["n"], case["q"]
        assert ob.primes(1, case["bits"], n)[0] == q
        t = ob.Tables(n, q)
        assert t.w == case["root"] and t.inv_n == case["inv_n"]
        tabs = [t.roots, t.precon, t.inv_roots, t.precon_inv]
        assert [f"{ob.fnv(x):016x}" for x in tabs] == case["tables_fnv"]
        assert int(t.inv_roots[n - 1]) == case["inv_roots_last"]
        for v in case["vectors"]:
            a = stim(v["kind"], n, q, v["seed"])
            assert f"{ob.fnv(a):016x}" == v["in_fnv"]
            f = ob.fwd_ntt(a, t)
            assert f"{ob.fnv(f):016x}" == v["fwd_fnv"], (n, q, v["kind"])
            assert [int(x) for x in f[:4]] == v["fwd_head"]
            i = ob.inv_ntt(a, t)
            assert f"{ob.fnv(i):016x}" == v["inv_fnv"], (n, q, v["kind"])
            assert [int(x) for x in i[:4]] == v["inv_head"]
    s = g["small"]
    t = ob.Tables(s["n"], s["q"])
    assert [int(x) for x in ob.fwd_ntt(np.array(s["in"], dtype=np.uint64), t)] == s["fwd"]


def test_against_live_reference_build():
    r = ob.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    for n, bits in ((2048, 30), (16384, 51), (16384, 61)):
        q = ob.primes(1, bits, n)[0]
        t = ob.Tables(n, q)
        tabs = [np.zeros(n, dtype=np.uint64) for _ in range(4)]
        r.ref_tables(n, q, *[ob.P(x) for x in tabs])
        for mine, theirs in zip([t.roots, t.precon, t.inv_roots, t.precon_inv], tabs):
            assert np.array_equal(mine, theirs)
        for kind in ("random", "garbage", "all_max"):
            a = stim(kind, n, q, 5)
            f = a.copy()
            r.ref_fwd_ntt(ob.P(f), n, q)
            assert np.array_equal(ob.fwd_ntt(a, t), f)
            i = a.copy()
            r.ref_inv_ntt(ob.P(i), n, q)
            assert np.array_equal(ob.inv_ntt(a, t), i)


def test_fwd_matches_textbook_transform():
    o = ob.oracle()
    n = 4096
    q = ob.primes(1, 45, n)[0]
    t = ob.Tables(n, q)
    a = ob.splitmix(n, 3, q)
    b = a.copy()
    o.ho_fwd_ntt_reference(ob.P(b), n, q, ob.P(t.roots))
    assert np.array_equal(ob.fwd_ntt(a, t), b)


def test_keyswitch_twiddle_block_layout():
    """4-table layout of host/src/twiddle-factors.cpp:16-62 / fpga.cpp:1102-1109."""
    o = ob.oracle()
    n = 1024
    q = ob.primes(1, 40, n)[0]
    t = ob.Tables(n, q)
    blk = np.zeros(4 * n, dtype=np.uint64)
    o.ho_compute_roots_keyswitch(n, q, t.w, ob.P(blk))
    assert np.array_equal(blk[: n - 1], t.inv_roots[1:]) and blk[n - 1] == 0
    assert np.array_equal(blk[2 * n: 3 * n], t.roots)
    assert blk[3 * n] == 0 and np.array_equal(blk[3 * n + 1:], t.precon[1:])


def test_dyadic_closed_form():
    """expected values of tests/test_dyadic_multiply.cpp:54-84."""
    from test_gpu_dyadic import reference_io

    for num, M, n in [(2, 1, 64), (3, 4, 256), (2, 7, 1024)]:
        moduli, op1, op2, exp = reference_io(num, M, n)
        got = ob.dyadic(op1.reshape(-1), op2.reshape(-1), n, moduli, num, True)
        assert np.array_equal(got, exp.reshape(-1))


@pytest.mark.parametrize("n,D,K,bits", [(1024, 2, 3, 40), (1024, 6, 7, 51), (2048, 5, 7, 51), (1024, 7, 8, 60),
                                        (1024, 2, 6, 45)])
def test_keyswitch_two_restatements_agree(n, D, K, bits):
    from ks_util import KsProblem

    p = KsProblem(n, D, K, 2, bits)
    assert np.array_equal(p.expected(), p.expected(alt=True))


def negacyclic_mul(a, b, q):
    n = len(a)
    res = [0] * n
    for i in range(n):
        if a[i] == 0:
            continue
        for j in range(n):
            k = i + j
            if k < n:
                res[k] = (res[k] + a[i] * b[j]) % q
            else:
                res[k - n] = (res[k - n] - a[i] * b[j]) % q
    return res


@pytest.mark.parametrize("n,D,K", [(64, 3, 4), (64, 2, 4), (32, 6, 7), (32, 5, 7)])
def test_keyswitch_rlwe_noise(n, D, K):
    """Build a toy RLWE key-switching key from s' to s, keyswitch a random
    polynomial c, and check r0 + r1*s - c*s' is small and the SAME small
    polynomial in every RNS limb -- i.e. the formula (rounding, fix, sign of the
    subtraction, modswitch factor) is the SEAL/HEXL one, not merely
    self-consistent."""
    rng = np.random.default_rng(n * 100 + D * 10 + K)
    moduli = ob.primes(K, 28, n)
    qk = moduli[K - 1]
    tabs = [ob.Tables(n, q) for q in moduli]
    s = [int(x) for x in rng.integers(-1, 2, n)]
    s2 = [int(x) for x in rng.integers(-1, 2, n)]

    def ntt(poly, i):
        return ob.fwd_ntt(np.array([x % moduli[i] for x in poly], dtype=np.uint64), tabs[i])

    # keys[j][c][i]: c0 = -(a_j*s + e_j) + qk * s' * [i == j], c1 = a_j   (all in NTT form mod q_i)
    keys = []
    for j in range(D):
        a_big = [int(x) for x in rng.integers(0, 2**62, n)]
        e = [int(x) for x in rng.integers(-3, 4, n)]
        k = np.zeros((2, K, n), dtype=np.uint64)
        for i in range(K):
            q = moduli[i]
            a_i = [x % q for x in a_big]
            as_ = negacyclic_mul(a_i, [x % q for x in s], q)
            c0 = [(-(as_[l] + e[l])) % q for l in range(n)]
            if i == j:
                c0 = [(c0[l] + qk * s2[l]) % q for l in range(n)]
            k[0, i] = ntt(c0, i)
            k[1, i] = ntt(a_i, i)
        keys.append(k.reshape(-1))
    # target c: independent uniform residues per limb (an RNS polynomial), NTT form
    c_coeff = [[int(x) for x in rng.integers(0, moduli[j], n)] for j in range(D)]
    t_target = np.concatenate([ntt(c_coeff[j], j) for j in range(D)])
    o = ob.oracle()
    msf = np.array([o.ho_inv_mod(qk % q, q) for q in moduli], dtype=np.uint64)
    res = ob.keyswitch(np.zeros(2 * D * n, dtype=np.uint64), t_target, n, D, K, moduli, keys, msf, 1)
    res = res.reshape(2, D, n)
    noises = []
    for i in range(D):
        q = moduli[i]
        r0 = [int(x) for x in ob.inv_ntt(res[0, i], tabs[i])]
        r1 = [int(x) for x in ob.inv_ntt(res[1, i], tabs[i])]
        r1s = negacyclic_mul(r1, [x % q for x in s], q)
        # c as an integer polynomial is only defined limb-wise; use limb i of c
        # against digit decomposition: sum_j (c mod q_j lifted) * [i == j] part
        # cancels exactly in limb i, the others contribute multiples handled by
        # the key structure, so the check is per-limb: r0 + r1 s - c_i s' small.
        cs = negacyclic_mul(c_coeff[i], [x % q for x in s2], q)
        d = [(r0[l] + r1s[l] - cs[l]) % q for l in range(n)]
        d = [x - q if x > q // 2 else x for x in d]
        noises.append(d)
    bound = 64 * D * n // 8
    for d in noises:
        assert max(abs(x) for x in d) < bound, max(abs(x) for x in d)
    for d in noises[1:]:
        assert d == noises[0]


@pytest.mark.parametrize("n,bits", [(1024, 20), (4096, 51), (16384, 51), (16384, 60)])
def test_lazy_output_mod_factors_against_live_reference_build(n, bits):
    """output_mod_factor 4 (forward) / 2 (inverse): the oracle's lazy words equal the reference's own
    (tests/test_utils/ntt.cpp:442-470 through oracle/_ref), and reduce to the canonical transform."""
    r = ob.ref()
    q = ob.primes(1, bits, n)[0]
    t = ob.Tables(n, q)
    a = ob.splitmix(n, 4711, q)
    lf, li = ob.fwd_ntt_lazy(a, t), ob.inv_ntt_lazy(a, t)
    assert lf.max() < 4 * q and li.max() < 2 * q
    assert np.array_equal(lf % np.uint64(q), ob.fwd_ntt(a, t)) and np.array_equal(li % np.uint64(q), ob.inv_ntt(a, t))
    assert (lf >= q).any()                            # genuinely lazy (the inverse rarely leaves a word above q)
    if r is None or not hasattr(r, "ref_fwd_ntt_factors"):
        pytest.skip("oracle/_ref not built")
    x = a.copy()
    r.ref_fwd_ntt_factors(ob.P(x), n, q, 1, 4)
    assert np.array_equal(x, lf)
    x = a.copy()
    r.ref_inv_ntt_factors(ob.P(x), n, q, 1, 2)
    assert np.array_equal(x, li)
