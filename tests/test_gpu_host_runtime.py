"""The host-pointer runtime (hexl-fpga_b200/host/src/runtime.cpp) beyond the happy path: pageable and
pinned caller memory (the reference's callers hold std::vector memory, benchmark/bench_fwd_ntt.cpp:19-21;
it stages every batch, host/src/fpga.cpp:329-413), the key-set cache (fpga.cpp:1158-1165) with its
by-value checks, calls that touch the same output twice, and error reporting."""
import numpy as np
import pytest

import oracle_binding as ob
from ks_util import KsProblem

pytestmark = pytest.mark.gpu


def test_pageable_and_pinned_callers_agree(acquired):
    """NTT + INTT over a batch larger than one staging slot (64 MiB = 512 polynomials), scattered order."""
    import torch

    hb = acquired
    n, q, B = 16384, 2251799814045697, 1200
    t = ob.Tables(n, q)
    data = np.stack([ob.splitmix(n, 3000 + i, q) for i in range(8)])
    src = np.ascontiguousarray(np.resize(data, (B, n)))
    src[:, 0] = np.arange(B, dtype=np.uint64)          # every polynomial different
    want = {i: ob.fwd_ntt(src[i], t) for i in (0, 1, 511, 512, 777, B - 1)}
    pinned = torch.from_numpy(src.view(np.int64).copy()).pin_memory()
    pageable = src.copy()
    for buf, ptr in ((pageable, None), (pinned, pinned.data_ptr())):
        arr = buf if ptr is None else buf.numpy().view(np.uint64)
        order = list(range(B))
        order = order[600:] + order[:600]               # two contiguous runs, not one
        hb.set_worksize_NTT(B)
        for i in order:
            hb.NTT(arr[i], t.roots, t.precon, q, n)
        assert hb.NTTCompleted()
        for i, w in want.items():
            assert np.array_equal(arr[i], w), i
        hb.set_worksize_INTT(B)
        for i in order:
            hb.INTT(arr[i], t.inv_roots, t.precon_inv, q, t.inv_n, t.inv_n_w, n)
        assert hb.INTTCompleted()
        assert np.array_equal(arr, src)


def test_buffer_pinned_in_place(acquired):
    """hexl_b200_host_pin_buffer: a pageable (numpy) buffer registered once is DMA'd directly -- no bytes pass
    through the staging ring's copy threads, same results -- until it is unpinned again."""
    hb = acquired
    n, q, B = 16384, 2251799814045697, 300
    t = ob.Tables(n, q)
    data = np.stack([ob.splitmix(n, 4000 + i, q) for i in range(4)])
    src = np.ascontiguousarray(np.resize(data, (B, n)))
    src[:, 1] = np.arange(B, dtype=np.uint64)
    buf = np.empty(B * n + 7, dtype=np.uint64)[7:].reshape(B, n)     # not page-aligned
    buf[:] = src
    hb.pin_buffer(buf)
    hb.pin_buffer(buf[10:20])                           # already covered: a no-op
    with pytest.raises(hb.HexlB200Error, match="overlaps"):
        hb._check(hb.lib().hexl_b200_host_pin_buffer(buf.ctypes.data + buf.nbytes - 4096, 1 << 20), "pin_buffer")
    try:
        hb.set_worksize_NTT(B)
        for i in range(B):
            hb.NTT(buf[i], t.roots, t.precon, q, n)
        assert hb.NTTCompleted()
        for i in (0, 1, 150, B - 1):
            assert np.array_equal(buf[i], ob.fwd_ntt(src[i], t)), i
        hb.set_worksize_INTT(B)
        for i in range(B):
            hb.INTT(buf[i], t.inv_roots, t.precon_inv, q, t.inv_n, t.inv_n_w, n)
        assert hb.INTTCompleted()
        assert np.array_equal(buf, src)
    finally:
        hb.unpin_buffer(buf)
    with pytest.raises(hb.HexlB200Error, match="not pinned"):
        hb.unpin_buffer(buf)
    # and staged again afterwards
    hb.set_worksize_NTT(2)
    hb.NTT(buf[0], t.roots, t.precon, q, n)
    hb.NTT(buf[1], t.roots, t.precon, q, n)
    assert hb.NTTCompleted()
    assert np.array_equal(buf[1], ob.fwd_ntt(src[1], t))


def test_bulk_submission_helpers(acquired):
    """hexl_b200_host_*_many: `count` calls in one FFI crossing, queued under one lock per stretch of free queue
    space -- more calls than the request queue holds (FPGA_BUFSIZE 4096), asynchronous and (worksize 1) synchronous."""
    hb = acquired
    n, B = 1024, 5000
    q = int(ob.primes(1, 51, n)[0])
    t = ob.Tables(n, q)
    src = np.stack([ob.splitmix(n, 5000 + i, q) for i in range(16)])
    buf = np.ascontiguousarray(np.resize(src, (B, n)))
    buf[:, 2] = np.arange(B, dtype=np.uint64)
    orig = buf.copy()
    hb.set_worksize_NTT(B)
    hb.NTT_many(buf.ctypes.data, n, B, t.roots, t.precon, q, n)
    assert hb.NTTCompleted()
    for i in (0, 1, 4095, 4096, B - 1):
        assert np.array_equal(buf[i], ob.fwd_ntt(orig[i], t)), i
    hb.set_worksize_INTT(B)
    hb.INTT_many(buf.ctypes.data, n, B, t.inv_roots, t.precon_inv, q, t.inv_n, t.inv_n_w, n)
    assert hb.INTTCompleted()
    assert np.array_equal(buf, orig)
    # no set_worksize: every call of the bulk helper is synchronous, like a single call
    hb.NTT_many(buf.ctypes.data, n, 3, t.roots, t.precon, q, n)
    for i in range(3):
        assert np.array_equal(buf[i], ob.fwd_ntt(orig[i], t)), i
    assert np.array_equal(buf[3], orig[3])
    # keyswitch and dyadic
    p = KsProblem(4096, 3, 4, 9, 47, seed=11)
    keys = hb.KeyArray(p.keys)
    out = np.ascontiguousarray(p.result.copy())
    tt = np.ascontiguousarray(p.t_target)
    hb.set_worksize_KeySwitch(p.batch)
    hb.KeySwitch_many(out.ctypes.data, tt.ctypes.data, p.batch, p.n, p.D, p.K, p.D + 1, 2, p.moduli, keys, p.msf)
    assert hb.KeySwitchCompleted()
    assert np.array_equal(out, p.expected())
    nd, M, Bd = 4096, 2, 7
    moduli = np.array(ob.primes(M, 50, nd), dtype=np.uint64)
    op1 = np.stack([ob.splitmix(2 * M * nd, 30 + b, int(moduli[0])) for b in range(Bd)])
    op2 = np.stack([ob.splitmix(2 * M * nd, 40 + b, int(moduli[0])) for b in range(Bd)])
    res = np.zeros((Bd, 3 * M * nd), dtype=np.uint64)
    hb.set_worksize_DyadicMultiply(Bd)
    hb.DyadicMultiply_many(res.ctypes.data, op1.ctypes.data, op2.ctypes.data, Bd, nd, moduli, M)
    assert hb.DyadicMultiplyCompleted()
    assert np.array_equal(res.reshape(-1), ob.dyadic(op1.reshape(-1), op2.reshape(-1), nd, moduli, Bd))


def test_pageable_keyswitch_and_dyadic(acquired):
    hb = acquired
    p = KsProblem(8192, 4, 5, 70, 48, seed=5)           # 70 items x 768 KiB > one 64 MiB slot
    keys = hb.KeyArray(p.keys)
    out = p.result.copy()
    hb.set_worksize_KeySwitch(p.batch)
    for b in range(p.batch):
        hb.KeySwitch(out[b], p.t_target[b], p.n, p.D, p.K, p.D + 1, 2, p.moduli, keys, p.msf)
    assert hb.KeySwitchCompleted()
    assert np.array_equal(out, p.expected())
    n, M, B = 4096, 3, 5
    moduli = np.array(ob.primes(M, 50, n), dtype=np.uint64)
    op1 = np.stack([ob.splitmix(2 * M * n, 10 + b, int(moduli[0])) for b in range(B)])
    op2 = np.stack([ob.splitmix(2 * M * n, 20 + b, int(moduli[0])) for b in range(B)])
    res = np.zeros((B, 3 * M * n), dtype=np.uint64)
    hb.set_worksize_DyadicMultiply(B)
    for b in range(B):
        hb.DyadicMultiply(res[b], op1[b], op2[b], n, moduli, M)
    assert hb.DyadicMultiplyCompleted()
    assert np.array_equal(res.reshape(-1), ob.dyadic(op1.reshape(-1), op2.reshape(-1), n, moduli, B))


def test_two_calls_into_the_same_result_accumulate_twice(acquired):
    """Both queued in one worksize: the reference's host loop adds them one after the other
    (host/src/fpga.cpp:441-475), so the runtime must not put them into one device batch."""
    hb = acquired
    p = KsProblem(2048, 3, 4, 2, 45, seed=8)
    keys = hb.KeyArray(p.keys)
    out = p.result[0].copy()
    hb.set_worksize_KeySwitch(2)
    hb.KeySwitch(out, p.t_target[0], p.n, p.D, p.K, p.D + 1, 2, p.moduli, keys, p.msf)
    hb.KeySwitch(out, p.t_target[1], p.n, p.D, p.K, p.D + 1, 2, p.moduli, keys, p.msf)
    assert hb.KeySwitchCompleted()
    step1 = ob.keyswitch(p.result[0], p.t_target[0], p.n, p.D, p.K, p.moduli, p.keys, p.msf, 1)
    step2 = ob.keyswitch(step1, p.t_target[1], p.n, p.D, p.K, p.moduli, p.keys, p.msf, 1)
    assert np.array_equal(out, step2)


def test_key_cache_checks_small_arrays_by_value(acquired):
    """Same key-set pointer: (a) fresh copies of moduli / factors per call hit the cache and share a batch,
    (b) new CONTENTS behind the same factor buffer are honoured (the reference rebuilds its modulus metadata
    on every fence, host/src/fpga.cpp:1049-1061)."""
    hb = acquired
    p = KsProblem(2048, 2, 3, 3, 40, seed=9)
    keys = hb.KeyArray(p.keys)
    out = p.result.copy()
    copies = []
    hb.set_worksize_KeySwitch(p.batch)
    for b in range(p.batch):
        m, f = p.moduli.copy(), p.msf.copy()
        copies.append((m, f))
        hb.KeySwitch(out[b], p.t_target[b], p.n, p.D, p.K, p.D + 1, 2, m, keys, f)
    assert hb.KeySwitchCompleted()
    assert np.array_equal(out, p.expected())
    msf = p.msf.copy()
    out1 = p.result[0].copy()
    hb.KeySwitch(out1, p.t_target[0], p.n, p.D, p.K, p.D + 1, 2, p.moduli, keys, msf)
    assert np.array_equal(out1, p.expected()[0])
    msf[0] = (int(msf[0]) * 3 + 1) % int(p.moduli[0])      # same buffer, new contents
    out2 = p.result[0].copy()
    hb.KeySwitch(out2, p.t_target[0], p.n, p.D, p.K, p.D + 1, 2, p.moduli, keys, msf)
    want = ob.keyswitch(p.result[0], p.t_target[0], p.n, p.D, p.K, p.moduli, p.keys, msf, 1)
    assert np.array_equal(out2, want) and not np.array_equal(out2, out1)


def test_many_key_sets_stay_within_the_cache_bound(acquired):
    """More key sets than HEXL_B200_PLAN_CACHE (8): old plans are evicted, results stay right."""
    hb = acquired
    n, D, K = 1024, 2, 3
    for seed in range(11):
        p = KsProblem(n, D, K, 1, 40, seed=100 + seed)
        keys = hb.KeyArray(p.keys)
        out = p.result[0].copy()
        hb.KeySwitch(out, p.t_target[0], n, D, K, D + 1, 2, p.moduli, keys, p.msf)
        assert np.array_equal(out, p.expected()[0]), seed


def test_an_error_is_reported_once(acquired):
    """A failing batch (modulus without a 2n-th root of unity) is reported to its completer; the next
    call starts clean instead of failing with the old message."""
    hb = acquired
    p = KsProblem(1024, 2, 3, 1, 40, seed=3)
    keys = hb.KeyArray(p.keys)
    bad = p.moduli.copy()
    bad[1] = np.uint64(int(bad[1]) + 2048 * 2)     # still 1 mod 2n but composite or rootless with high probability
    while ob.is_prime(int(bad[1])):
        bad[1] = np.uint64(int(bad[1]) + 2048)
    out = p.result[0].copy()
    with pytest.raises(hb.HexlB200Error):
        hb.KeySwitch(out, p.t_target[0], p.n, p.D, p.K, p.D + 1, 2, bad, keys, p.msf)
    out = p.result[0].copy()
    hb.KeySwitch(out, p.t_target[0], p.n, p.D, p.K, p.D + 1, 2, p.moduli, keys, p.msf)
    assert np.array_equal(out, p.expected()[0])
