"""GPU parity of KeySwitch against the CPU oracle.

The reference's own keyswitch test (tests/test_keyswitch.cpp:119-191) replays
JSON vectors from an external testdata.zip that is not available offline, so
the pins here are: the stage-by-stage oracle, its independently structured
twin (ho_keyswitch_alt) and the RLWE noise self-test in test_oracle.py.
Shapes follow the reference's (6/7/7/2 and 5/7/6/2, tests/test_keyswitch.cpp:
148-191) plus BASELINE's decomp 7 / key 8."""
import numpy as np
import pytest

import oracle_binding as ob
from conftest import set_variant
from ks_util import KsProblem

pytestmark = pytest.mark.gpu


def gpu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


SHAPES = [
    # n, D, K, batch, bits
    (1024, 2, 3, 3, 40),
    (2048, 3, 4, 2, 45),
    (4096, 5, 7, 2, 51),   # 5/7/6/2
    (8192, 6, 7, 2, 51),   # 6/7/7/2 at N=8192
    (16384, 6, 7, 2, 51),  # 6/7/7/2, the reference's largest shape
    (16384, 7, 8, 3, 51),  # BASELINE config 4 shape
    (16384, 2, 8, 1, 51),  # dropped levels: D < K-1
]


@pytest.mark.parametrize("fp64", [1, 0])
@pytest.mark.parametrize("n,D,K,batch,bits", SHAPES)
def test_device_api_vs_oracle(hb, n, D, K, batch, bits, fp64):
    """fp64 = 1: the transform stages run on the FP64 pipe (all moduli within 2^36 .. 2^53/3, the
    default); 0: the integer kernels.  The option is read when the plan is created."""
    p = KsProblem(n, D, K, batch, bits)
    hb.set_option("fp64_path", fp64)
    try:
        plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
    finally:
        hb.set_option("fp64_path", 1)
    res = gpu(p.result)
    hb_t = gpu(p.t_target)
    plan.keyswitch(res, hb_t, batch)
    got = res.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, p.expected())
    plan.close()


@pytest.mark.parametrize("n,D,K", [(16384, 7, 8), (16384, 6, 7), (8192, 5, 7)])
def test_device_api_vs_second_restatement(hb, n, D, K):
    """The independently structured CPU restatement (ho_keyswitch_alt, hexl order with lazy 128-bit
    accumulation) at the headline shapes."""
    p = KsProblem(n, D, K, 2, 51, seed=4242)
    plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
    res = gpu(p.result)
    plan.keyswitch(res, gpu(p.t_target), 2)
    assert np.array_equal(res.cpu().numpy().view(np.uint64), p.expected(alt=True))
    plan.close()


@pytest.mark.parametrize("opt,val", [("ks_fused", 1), ("ks_sub_items", 2), ("ks_sub_items", 5)])
@pytest.mark.parametrize("n,D,K,batch", [(16384, 7, 8, 5), (16384, 6, 7, 3), (16384, 2, 8, 4), (16384, 1, 2, 3), (8192, 5, 7, 3)])
def test_fused_kernel_and_l2_rounds(hb, n, D, K, batch, opt, val):
    """ks_fused = 1: stages S2 + S3 + S4 in one kernel, the sums in tensor memory (keyswitch_fused.cu; N = 16384
    only, other sizes take the staged kernels); ks_sub_items: S2 + S3 in rounds of a few items so that the
    NTT'd digits stay in L2.  Both must reproduce the oracle bit for bit."""
    p = KsProblem(n, D, K, batch, 51, seed=2024)
    hb.set_option(opt, val)
    try:
        plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
        res = gpu(p.result)
        plan.keyswitch(res, gpu(p.t_target), batch)
        got = res.cpu().numpy().view(np.uint64)
        plan.close()
    finally:
        hb.set_option(opt, 0)
    assert np.array_equal(got, p.expected())


@pytest.mark.parametrize("option", ["ks_mac_fp64", "ks_s5_fp64", "ks_u_fp64"])
@pytest.mark.parametrize("n,D,K,batch", [(16384, 7, 8, 5), (16384, 6, 7, 3), (16384, 2, 8, 4), (16384, 1, 2, 3)])
def test_integer_stages_match_the_fp64_ones(hb, n, D, K, batch, option):
    """ks_mac_fp64 = 1 (default at N = 16384 with moduli up to 2^51 (1 + 1/32)): stage S2 leaves raw doubles in V
    and stage S3 multiplies on the FP64 pipe; 0: canonical words and the integer Shoup products.  Same bits,
    also with target words the FP64 product does not take as they are (>= 2^52: reduced first) -- the digit
    under its own modulus reaches the multiply-accumulate straight from the caller's buffer.
    ks_s5_fp64 = 1 (default under the same conditions): stage S5's base conversion and modswitch / accumulate
    epilogue on the FP64 pipe, `result` through TMA; 0: the integer epilogue.  Same bits, also for `result` words
    outside [0, q) (the reference's wrap-around add_mod decides those).
    ks_u_fp64 = 1 (default with ks_mac_fp64 and same-size moduli): stage S1 hands U over as non-negative doubles
    (the exact pass behind it converts) and S2 takes them without an entry conversion; 0: canonical integers."""
    p = KsProblem(n, D, K, batch, 51, seed=77)
    t = p.t_target.reshape(batch, D, n).copy()
    t[0, 0, 3] = np.uint64((1 << 63) + 12345)          # garbage: every kernel family must agree on it
    t[batch - 1, D - 1, n - 1] = np.uint64((1 << 52) + 1)
    r_in = p.result.reshape(batch, 2, D, n).copy()
    r_in[0, 1, D - 1, 7] = np.uint64((1 << 64) - 5)
    r_in[batch - 1, 0, 0, n - 2] = np.uint64(int(p.moduli[0]) + 3)
    plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
    out = []
    for opt in (1, 0):
        hb.set_option(option, opt)
        try:
            res = gpu(p.result)
            plan.keyswitch(res, gpu(p.t_target), batch)
            clean = res.cpu().numpy().view(np.uint64).copy()
            res = gpu(r_in.reshape(batch, -1))
            plan.keyswitch(res, gpu(t.reshape(batch, -1)), batch)
            out.append((clean, res.cpu().numpy().view(np.uint64).copy()))
        finally:
            hb.set_option(option, 1)
    plan.close()
    assert np.array_equal(out[0][0], p.expected())
    assert np.array_equal(out[1][0], p.expected())
    assert np.array_equal(out[0][1], out[1][1])


def test_out_of_range_target_words_go_to_the_exact_kernel(hb):
    """t_target words in [1.25 q, 2q) are outside the FP64 arithmetic's contract but inside the
    integer kernels': the first stage's range vote must send those digits to the exact kernel, so
    both arithmetic paths produce the same words (the reference leaves such inputs undefined)."""
    n, D, K, batch = 16384, 3, 4, 2
    p = KsProblem(n, D, K, batch, 51)
    t = p.t_target.reshape(batch, D, n).copy()
    t[0, 1, 5] += np.uint64(int(p.moduli[1]) // 2)
    t[1, 2, n - 1] = np.uint64(int(p.moduli[2]) + int(p.moduli[2]) // 2)
    out = []
    for fp64 in (1, 0):
        hb.set_option("fp64_path", fp64)
        try:
            plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
        finally:
            hb.set_option("fp64_path", 1)
        res = gpu(p.result)
        plan.keyswitch(res, gpu(t.reshape(batch, -1)), batch)
        out.append(res.cpu().numpy().view(np.uint64).copy())
        plan.close()
    assert np.array_equal(out[0], out[1])


def test_caller_twiddle_tables_are_honoured(hb):
    """twiddle_factors in the reference's 4-table format
    (tests/test_keyswitch.cpp:73-90, host/src/fpga.cpp:1102-1109)."""
    n, D, K = 2048, 3, 4
    p = KsProblem(n, D, K, 2, 45)
    tw = np.zeros((K, 4 * n), dtype=np.uint64)
    o = ob.oracle()
    for i, q in enumerate(p.moduli):
        o.ho_compute_roots_keyswitch(n, int(q), o.ho_min_primitive_root(2 * n, int(q)), ob.P(tw[i]))
    plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf, twiddles=tw.reshape(-1))
    res = gpu(p.result)
    plan.keyswitch(res, gpu(p.t_target), 2)
    assert np.array_equal(res.cpu().numpy().view(np.uint64), p.expected())


def test_chunked_workspace_matches(hb):
    """A workspace too small for the batch forces several chunks."""
    n, D, K, batch = 4096, 3, 4, 7
    p = KsProblem(n, D, K, batch, 51)
    hb.set_option("ks_workspace_mb", 16)
    try:
        plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
        res = gpu(p.result)
        plan.keyswitch(res, gpu(p.t_target), batch)
        assert np.array_equal(res.cpu().numpy().view(np.uint64), p.expected())
    finally:
        hb.set_option("ks_workspace_mb", 10240)


def test_host_api_accumulates(acquired):
    """KeySwitch on host pointers: async worksize protocol, result += output
    (host/src/fpga.cpp:441-475), scattered (non-contiguous) items, and a second
    call accumulating on top of the first."""
    hb = acquired
    n, D, K, batch = 4096, 5, 7, 4
    p = KsProblem(n, D, K, batch, 51)
    keys = hb.KeyArray(p.keys)
    res = [p.result[b].copy() for b in range(batch)]       # separately allocated
    tt = [p.t_target[b].copy() for b in range(batch)]
    hb.set_worksize_KeySwitch(batch)
    for b in range(batch):
        hb.KeySwitch(res[b], tt[b], n, D, K, D + 1, 2, p.moduli, keys, p.msf)
    assert hb.KeySwitchCompleted()
    exp = p.expected()
    for b in range(batch):
        assert np.array_equal(res[b], exp[b])
    # accumulate again on top (synchronous call)
    hb.KeySwitch(res[0], tt[0], n, D, K, D + 1, 2, p.moduli, keys, p.msf)
    exp2 = ob.keyswitch(exp[0], tt[0], n, D, K, p.moduli, p.keys, p.msf, 1)
    assert np.array_equal(res[0], exp2)


def test_full_size_property(hb):
    """BASELINE config 4 size slice (batch 256 of the 1024): every item uses the
    same t_target/result, so all outputs must be identical, and equal to the
    oracle's single-item answer; result < q everywhere."""
    import torch

    n, D, K, batch = 16384, 7, 8, 256
    p = KsProblem(n, D, K, 1, 51)
    plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
    res = gpu(p.result).repeat(batch, 1).contiguous()
    tt = gpu(p.t_target).repeat(batch, 1).contiguous()
    plan.keyswitch(res, tt, batch)
    exp = gpu(p.expected())
    assert torch.equal(res, exp.expand(batch, -1))


@pytest.mark.parametrize("mac_items", [1, 2, 8])
@pytest.mark.parametrize("n,D,K,batch,bits", [(16384, 7, 8, 3, 51), (4096, 3, 4, 5, 45), (16384, 2, 8, 2, 57)])
def test_mac_variants_agree_with_oracle(hb, n, D, K, batch, bits, mac_items):
    """Stage S3 has several kernels (option ks_mac_items: 4 = Shoup products, four items per key
    load (default); 2 = 128-bit accumulators, one reduction per output; 8; 1 = register-resident
    keys): all must give the oracle's words, odd batch sizes included."""
    p = KsProblem(n, D, K, batch, bits)
    set_variant(hb, "ks_mac_items", mac_items)
    try:
        plan = hb.KsPlan(n, D, K, D + 1, 2, p.moduli, p.keys, p.msf)
        res = gpu(p.result)
        plan.keyswitch(res, gpu(p.t_target), batch)
        got = res.cpu().numpy().view(np.uint64)
        plan.close()
    finally:
        hb.set_option("ks_mac_items", 4)
    assert np.array_equal(got, p.expected())
