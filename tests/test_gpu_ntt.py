"""GPU parity of forward / inverse NTT against the CPU oracle, through the C ABI.

Mirrors the reference's tests/test_fwd_ntt.cpp:97-170 and
tests/test_inv_ntt.cpp:97-178: same stimuli (RANDOM, RAMP, ALL_ZEROS, ALL_ONES,
IMPULSE, ALL_MAX_VALUES = 2^64-1) and prime sizes, bit-exact comparison.
"""
import numpy as np
import pytest

import oracle_binding as ob
from conftest import set_variant

pytestmark = pytest.mark.gpu

N = 16384
STIMULI = ["random", "ramp", "zeros", "ones", "impulse", "all_max", "garbage"]


def stimulus(kind, n, q, seed):
    if kind == "random":
        return ob.splitmix(n, seed, q)
    if kind == "ramp":
        return np.arange(n, dtype=np.uint64)
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint64)
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


def to_gpu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_np(t):
    return t.cpu().numpy().view(np.uint64)


def run_fwd(hb, polys, t):
    d = to_gpu(np.stack(polys))
    hb.ntt_fwd(d, to_gpu(t.roots), to_gpu(t.precon), t.q, t.n)
    return to_np(d)


def run_inv(hb, polys, t):
    d = to_gpu(np.stack(polys))
    hb.ntt_inv(d, to_gpu(t.inv_roots), to_gpu(t.precon_inv), t.q, t.inv_n, t.inv_n_w, t.n)
    return to_np(d)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("bits", [20, 32, 51, 55, 61])
def test_fwd_inv_all_stimuli_n16384(hb, bits, variant):
    hb.set_option("ntt_variant", variant)
    try:
        q = ob.primes(1, bits, N)[0]
        t = ob.Tables(N, q)
        polys = [stimulus(k, N, q, 100 + i) for i, k in enumerate(STIMULI)]
        got = run_fwd(hb, polys, t)
        for i, k in enumerate(STIMULI):
            assert np.array_equal(got[i], ob.fwd_ntt(polys[i], t)), f"fwd {k} bits={bits}"
        got = run_inv(hb, polys, t)
        for i, k in enumerate(STIMULI):
            assert np.array_equal(got[i], ob.inv_ntt(polys[i], t)), f"inv {k} bits={bits}"
    finally:
        hb.set_option("ntt_variant", 1)


@pytest.mark.parametrize("n", [1024, 2048, 4096, 8192])
def test_fwd_inv_other_sizes(hb, n):
    q = ob.primes(1, 51, n)[0]
    t = ob.Tables(n, q)
    polys = [stimulus(k, n, q, 7 + i) for i, k in enumerate(["random", "ramp", "all_max", "garbage"])]
    got = run_fwd(hb, polys, t)
    goti = run_inv(hb, polys, t)
    for i in range(len(polys)):
        assert np.array_equal(got[i], ob.fwd_ntt(polys[i], t))
        assert np.array_equal(goti[i], ob.inv_ntt(polys[i], t))


def test_known_answers(hb):
    """SURVEY.md Appendix B KAT-1 / KAT-2 / KAT-3 through the GPU path."""
    q = 2251799814045697
    t = ob.Tables(N, q)
    a = ob.splitmix(N, 1, q)
    out = run_fwd(hb, [a], t)[0]
    assert [int(x) for x in out[:3]] == [1955457978075445, 1092550199427436, 1923103082448610]
    assert ob.fnv(out) == 0x428B5C898DD187A3
    assert np.array_equal(run_inv(hb, [out], t)[0], a)
    ramp = {20: 0xBB10994907704BC4, 32: 0xBD3ECAEEF0141E84, 52: 0x64381823DFF21901,
            55: 0xFAE12F2917FA203E}
    for bits, h in ramp.items():
        qq = ob.primes(1, bits, N)[0]
        tt = ob.Tables(N, qq)
        assert ob.fnv(run_fwd(hb, [np.arange(N, dtype=np.uint64)], tt)[0]) == h, bits
    qq = 4503599627763713
    out = run_fwd(hb, [np.full(N, 2**64 - 1, dtype=np.uint64)], ob.Tables(N, qq))[0]
    assert int(out[0]) == 18373839436896342053 and ob.fnv(out) == 0x47244FE2BFC8E399


def test_full_size_roundtrip_and_linearity(hb):
    """BASELINE config 2 size (batch 4096): inverse(forward(x)) == x and
    NTT(a) + NTT(b) == NTT(a + b) mod q, checked on the GPU with torch."""
    import torch

    q = 2251799814045697
    t = ob.Tables(N, q)
    B = 4096
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randint(0, q, (B, N), dtype=torch.int64, device="cuda", generator=g)
    y = x.clone()
    r, p = to_gpu(t.roots), to_gpu(t.precon)
    hb.ntt_fwd(y, r, p, q, N)
    assert int(y.max()) < q and int(y.min()) >= 0
    # spot-check three polynomials against the oracle
    for i in (0, 1777, B - 1):
        assert np.array_equal(to_np(y[i]), ob.fwd_ntt(to_np(x[i]), t))
    # linearity on the first half vs second half
    s = (x[: B // 2] + x[B // 2:]) % q
    hb.ntt_fwd(s, r, p, q, N)
    assert torch.equal(s, (y[: B // 2] + y[B // 2:]) % q)
    hb.ntt_inv(y, to_gpu(t.inv_roots), to_gpu(t.precon_inv), q, t.inv_n, t.inv_n_w, N)
    assert torch.equal(x, y)


def test_host_api_ntt_intt(acquired):
    """_set_worksize_NTT / _NTT / _NTTCompleted on host pointers, batch laid out
    contiguously as the reference tests do (tests/test_fwd_ntt.cpp:109-115)."""
    hb = acquired
    q = ob.primes(1, 51, N)[0]
    t = ob.Tables(N, q)
    B = 9
    data = np.stack([ob.splitmix(N, 50 + i, q) for i in range(B)])
    work = data.copy()
    hb.set_worksize_NTT(B)
    for i in range(B):
        hb.NTT(work[i], t.roots, t.precon, q, N)
    assert hb.NTTCompleted()
    for i in range(B):
        assert np.array_equal(work[i], ob.fwd_ntt(data[i], t))
    hb.set_worksize_INTT(B)
    for i in range(B):
        hb.INTT(work[i], t.inv_roots, t.precon_inv, q, t.inv_n, t.inv_n_w, N)
    assert hb.INTTCompleted()
    assert np.array_equal(work, data)
    # synchronous mode (worksize 1, the default)
    one = data[0].copy()
    hb.NTT(one, t.roots, t.precon, q, N)
    assert np.array_equal(one, ob.fwd_ntt(data[0], t))


@pytest.mark.parametrize("n,batch,bits", [(16384, 1, 51), (16384, 149, 51), (16384, 300, 51), (1024, 1000, 51),
                                          (4096, 777, 51), (16384, 297, 27), (16384, 700, 27), (16384, 2, 27)])
def test_persistent_loop_remainders(hb, n, batch, bits):
    """Batches that are not a multiple of the persistent grid (148 CTAs x occupancy),
    a mix of in-contract and garbage polynomials (deferred exact list), checked
    against the oracle on a sample of positions and by the inverse round trip."""
    import torch

    q = ob.primes(1, bits, n)[0]          # bits=27: the uint32 small-modulus kernels (two CTAs per SM)
    t = ob.Tables(n, q)
    g = torch.Generator(device="cuda").manual_seed(batch)
    x = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    # incl. pairs that land on the same persistent CTA in consecutive iterations (grid 148 or 296)
    extra = [r for r in (batch // 2 + 1, batch // 2 + 148, batch // 2 + 296) if batch > 4 and r < batch]
    garbage_rows = sorted({0, batch // 2, batch - 1, *extra})
    for r in garbage_rows[1:] if batch > 1 else []:
        x[r] = torch.from_numpy(ob.splitmix(n, r + 1, 0).view(np.int64)).cuda()
    x0 = x.clone()
    hb.ntt_fwd(x, to_gpu(t.roots), to_gpu(t.precon), q, n)
    for r in sorted(set(garbage_rows + [1 % batch, batch // 3])):
        assert np.array_equal(to_np(x[r]), ob.fwd_ntt(to_np(x0[r]), t)), r
    clean = [r for r in range(batch) if r not in garbage_rows[1:]] if batch > 1 else [0]
    hb.ntt_inv(x, to_gpu(t.inv_roots), to_gpu(t.precon_inv), q, t.inv_n, t.inv_n_w, n)
    idx = torch.tensor(clean, device="cuda")
    assert torch.equal(x[idx], x0[idx])


def test_empty_batch_and_bad_arguments(hb):
    import torch

    n, q = 16384, 2251799814045697
    t = ob.Tables(n, q)
    x = torch.zeros((1, n), dtype=torch.int64, device="cuda")
    r, p = to_gpu(t.roots), to_gpu(t.precon)
    assert hb.lib().hexl_b200_ntt_fwd(x.data_ptr(), r.data_ptr(), p.data_ptr(), q, n, 0, None) == 0   # batch 0: no-op
    assert hb.lib().hexl_b200_ntt_fwd(x.data_ptr() + 8, r.data_ptr(), p.data_ptr(), q, n, 1, None) == -1  # misaligned
    assert hb.lib().hexl_b200_ntt_fwd(x.data_ptr(), r.data_ptr(), p.data_ptr(), q, 65536, 1, None) == -1  # n too large
    assert hb.lib().hexl_b200_ntt_inv(x.data_ptr(), r.data_ptr(), p.data_ptr(), q, q, 0, n, 1, None) == -1  # inv_n >= q


@pytest.mark.parametrize("bits", [58, 59, 61])
def test_large_moduli_take_the_exact_kernels(hb, bits):
    """q >= 2^58 has no lazy forward path, q >= 2^60 no fast inverse: kExactAll."""
    q = ob.primes(1, bits, N)[0]
    t = ob.Tables(N, q)
    polys = [stimulus(k, N, q, 3 + i) for i, k in enumerate(["random", "ramp", "garbage"])]
    got = run_fwd(hb, polys, t)
    goti = run_inv(hb, polys, t)
    for i in range(3):
        assert np.array_equal(got[i], ob.fwd_ntt(polys[i], t))
        assert np.array_equal(goti[i], ob.inv_ntt(polys[i], t))


@pytest.mark.parametrize("q", [136314881, None, 65537])
def test_small_modulus_32bit_path(hb, q):
    """q < 2^30 runs the uint32 kernels (N = 16384); same stimuli, incl. garbage
    (deferred to the exact 64-bit kernel) and the edge of the contract; results
    must be identical with the path switched off."""
    q = q or ob.primes(1, 29, N)[0]
    t = ob.Tables(N, q)
    polys = [stimulus(k, N, q, 200 + i) for i, k in enumerate(STIMULI)]
    polys.append(np.where(np.arange(N) % 2 == 0, 4 * q - 1, q - 1).astype(np.uint64))     # fwd contract edge
    polys.append(np.where(np.arange(N) % 2 == 0, 2 * q - 1, 0).astype(np.uint64))         # inv contract edge
    polys.append(np.full(N, 2**32 + 5, dtype=np.uint64))                                  # high word set
    # small_path 1: TMA landing buffer (default); 2: direct loads, two CTAs per SM; "1t": path 1 with the
    # forward epilogue through TMA stores (option small_tma_store)
    for small in (3, 2, 1, "1t", 0):      # 3: two transforms per SM sharing three 64 KiB regions
        try:
            hb.set_option("small_path", 1 if small == "1t" else small)
        except hb.HexlB200Error as e:         # 2 and 3 only exist in builds with make EXPERIMENTAL=1
            assert "EXPERIMENTAL" in str(e) and small in (2, 3)
            continue
        hb.set_option("small_tma_store", 1 if small == "1t" else 0)
        try:
            got = run_fwd(hb, polys, t)
            goti = run_inv(hb, polys, t)
        finally:
            hb.set_option("small_path", 1)
            hb.set_option("small_tma_store", 0)
        for i in range(len(polys)):
            assert np.array_equal(got[i], ob.fwd_ntt(polys[i], t)), (small, i)
            assert np.array_equal(goti[i], ob.inv_ntt(polys[i], t)), (small, i)


@pytest.mark.parametrize("n,bits", [(16384, 51), (16384, 40), (8192, 51), (4096, 51), (2048, 45), (1024, 51)])
def test_inverse_lazy_and_corrected_butterflies_agree(hb, n, bits):
    """q < 2^52 takes the correction-free inverse butterflies (one reduction half way);
    option inv_lazy=0 keeps the per-stage corrected ones.  Both must give the oracle's
    words, including on the inputs that make the lazy sums grow fastest (all 2q-1)."""
    q = ob.primes(1, bits, n)[0]
    t = ob.Tables(n, q)
    polys = [stimulus(k, n, q, 300 + i) for i, k in enumerate(STIMULI)]
    polys.append(np.full(n, 2 * q - 1, dtype=np.uint64))
    polys.append(np.where(np.arange(n) % 2 == 0, 2 * q - 1, 0).astype(np.uint64))
    want = [ob.inv_ntt(p, t) for p in polys]
    for lazy in (1, 0):
        set_variant(hb, "inv_lazy", lazy)
        try:
            got = run_inv(hb, polys, t)
        finally:
            hb.set_option("inv_lazy", 0)
        for i in range(len(polys)):
            assert np.array_equal(got[i], want[i]), (lazy, i)


@pytest.mark.parametrize("batch", [1, 2, 3, 147, 149, 297, 700])
def test_small_path_two_transforms_per_sm(hb, batch):
    """small_path=3: two 512-thread groups per CTA take turns on the shared landing half; odd and
    even item counts per CTA, garbage polynomials (deferred to the exact kernel) on both groups."""
    import torch

    n, q = 16384, 136314881
    t = ob.Tables(n, q)
    g = torch.Generator(device="cuda").manual_seed(1000 + batch)
    x = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    garbage = sorted({r for r in (batch // 2, batch // 2 + 148, batch - 1) if 0 < r < batch})
    for r in garbage:
        x[r] = torch.from_numpy(ob.splitmix(n, r + 1, 0).view(np.int64)).cuda()
    x0 = x.clone()
    set_variant(hb, "small_path", 3)
    try:
        hb.ntt_fwd(x, to_gpu(t.roots), to_gpu(t.precon), q, n)
        fwd = x.clone()
        hb.ntt_inv(x, to_gpu(t.inv_roots), to_gpu(t.precon_inv), q, t.inv_n, t.inv_n_w, n)
    finally:
        hb.set_option("small_path", 1)
    for r in sorted(set(garbage + [0, batch // 3, batch - 1])):
        assert np.array_equal(to_np(fwd[r]), ob.fwd_ntt(to_np(x0[r]), t)), r
    clean = torch.tensor([r for r in range(batch) if r not in garbage], device="cuda")
    assert torch.equal(x[clean], x0[clean])


def _largest_fp64_prime(n):
    """the last NTT-friendly prime not above 2^53 / 3 (the FP64-pipe path's upper limit)"""
    lim = 2**53 // 3
    c = (lim // (2 * n)) * (2 * n) + 1
    while c > lim or not ob.is_prime(c):
        c -= 2 * n
    return c


@pytest.mark.parametrize("n", [16384, 4096])
@pytest.mark.parametrize("which", ["b36", "b44", "b50", "b51", "top", "above"])
def test_fp64_pipe_path_matches_oracle_and_integer_path(hb, n, which):
    """Butterflies on the FP64 pipe (option fp64_path, default on, 2^36 <= q <= 2^53/3): bit-exact
    against the oracle and the integer kernels on random words, on the edges of the centred ranges
    (all q-1, alternating 0 / q-1, words in [q, 1.25q)), and on out-of-contract polynomials, which
    the range vote sends to the exact kernel."""
    if which == "top":
        q = _largest_fp64_prime(n)
    elif which == "above":                      # first prime beyond the limit: integer kernels
        q = 2**53 // 3 // (2 * n) * (2 * n) + 1
        while q <= 2**53 // 3 or not ob.is_prime(q):
            q += 2 * n
    else:
        q = ob.primes(1, int(which[1:]), n)[0]
    t = ob.Tables(n, q)
    top = q + q // 4 - 1
    idx = np.arange(n, dtype=np.uint64)
    polys = [
        ob.splitmix(n, 5, q),
        np.full(n, q - 1, dtype=np.uint64),
        np.where(idx & np.uint64(1), np.uint64(q - 1), np.uint64(0)),
        np.where(idx & np.uint64(1), np.uint64(top) - idx % np.uint64(7), np.uint64(q // 2) + idx % np.uint64(3)),
        np.full(n, top, dtype=np.uint64),
        np.full(n, q + q // 4 + 2**33, dtype=np.uint64),          # just out of the FP64 contract, inside the integer one
        ob.splitmix(n, 6, 0),                                      # garbage
        np.full(n, 2**64 - 1, dtype=np.uint64),
        ob.splitmix(n, 7, q),
    ]
    polys[8][n // 3] = 2**63 + 12345                               # one bad word in an otherwise clean polynomial
    want_f = [ob.fwd_ntt(p, t) for p in polys]
    want_i = [ob.inv_ntt(p, t) for p in polys]
    # (fp64_path, ntt_variant, warp_tail); warp_tail = 0: the FP64 kernels with three block barriers
    # per transform instead of the warp-dealt tail rows (n = 16384, 32 words per thread only)
    combos = [(1, 1, 1), (0, 1, 1)] + ([(1, 0, 1), (0, 0, 1), (1, 1, 0), (1, 3, 1)] if n == 16384 else [])
    for fp64, variant, warp_tail in combos:
        hb.set_option("fp64_path", fp64)
        hb.set_option("ntt_variant", variant)
        hb.set_option("warp_tail", warp_tail)
        trusted = (variant & 2) != 0          # no range vote: in-contract polynomials only
        try:
            sel = [i for i in range(len(polys)) if not trusted or i < 5]
            got_f = run_fwd(hb, [polys[i] for i in sel], t)
            got_i = run_inv(hb, [polys[i] for i in sel], t)
        finally:
            hb.set_option("fp64_path", 1)
            hb.set_option("ntt_variant", 1)
            hb.set_option("warp_tail", 1)
        for k, i in enumerate(sel):
            assert np.array_equal(got_f[k], want_f[i]), (which, fp64, variant, warp_tail, "fwd", i)
            assert np.array_equal(got_i[k], want_i[i]), (which, fp64, variant, warp_tail, "inv", i)


def test_fp64_pipe_path_large_batch_round_trip(hb):
    """4096 polynomials (BASELINE configs[1]) through the FP64-pipe kernels: forward spot-checked
    against the oracle, inverse(forward(x)) == x for the whole batch."""
    import torch

    q = 2251799814045697
    t = ob.Tables(N, q)
    g = torch.Generator(device="cuda").manual_seed(4321)
    x = torch.randint(0, q, (4096, N), dtype=torch.int64, device="cuda", generator=g)
    x0 = x.clone()
    hb.ntt_fwd(x, to_gpu(t.roots), to_gpu(t.precon), q, N)
    for r in (0, 147, 148, 2047, 4095):
        assert np.array_equal(to_np(x[r]), ob.fwd_ntt(to_np(x0[r]), t)), r
    hb.ntt_inv(x, to_gpu(t.inv_roots), to_gpu(t.precon_inv), q, t.inv_n, t.inv_n_w, N)
    assert torch.equal(x, x0)
    # every polynomial of the batch against the three-barrier kernels (same words expected)
    hb.ntt_fwd(x, to_gpu(t.roots), to_gpu(t.precon), q, N)
    y = x0.clone()
    hb.set_option("warp_tail", 0)
    try:
        hb.ntt_fwd(y, to_gpu(t.roots), to_gpu(t.precon), q, N)
    finally:
        hb.set_option("warp_tail", 1)
    assert torch.equal(x, y)


@pytest.mark.parametrize("n,bits", [(16384, 51), (16384, 28), (16384, 60), (4096, 45), (1024, 20)])
def test_output_mod_factors_of_the_reference_ntt_class(hb, n, bits):
    """hetest::utils::NTT::ComputeForward(..., output_mod_factor = 4) / ComputeInverse(..., 2)
    (tests/test_utils/ntt.cpp:442-470): the lazy words of the Harvey butterflies, bit for bit, incl. on
    input_mod_factor 4 / 2 inputs; factor 1 stays the fully reduced transform."""
    q = ob.primes(1, bits, n)[0]
    t = ob.Tables(n, q)
    a = np.stack([ob.splitmix(n, 60 + i, q) for i in range(3)])
    a[1] += np.uint64(q) * (np.arange(n, dtype=np.uint64) % np.uint64(4))        # forward input in [0, 4q)
    d = to_gpu(a)
    hb.ntt_fwd(d, to_gpu(t.roots), to_gpu(t.precon), q, n, input_mod_factor=4, output_mod_factor=4)
    got = to_np(d)
    for i in range(3):
        assert np.array_equal(got[i], ob.fwd_ntt_lazy(a[i], t)), i
    assert (got >= q).any()
    b = np.stack([ob.splitmix(n, 80 + i, q) for i in range(3)])
    b[2] += np.uint64(q) * (np.arange(n, dtype=np.uint64) % np.uint64(2))        # inverse input in [0, 2q)
    d = to_gpu(b)
    hb.ntt_inv(d, to_gpu(t.inv_roots), to_gpu(t.precon_inv), q, t.inv_n, t.inv_n_w, n, input_mod_factor=2,
               output_mod_factor=2)
    got = to_np(d)
    for i in range(3):
        assert np.array_equal(got[i], ob.inv_ntt_lazy(b[i], t)), i
    with pytest.raises(hb.HexlB200Error):
        hb.ntt_fwd(d, to_gpu(t.roots), to_gpu(t.precon), q, n, output_mod_factor=2)
    with pytest.raises(hb.HexlB200Error):
        hb.ntt_inv(d, to_gpu(t.inv_roots), to_gpu(t.precon_inv), q, t.inv_n, t.inv_n_w, n, output_mod_factor=4)


@pytest.mark.parametrize("bits", [51, 28, 60, 40])
def test_n32768_forward_and_inverse(hb, bits):
    """N = 32768: first / last stage as a streaming kernel, the halves on the 16384-point
    kernels (csrc/ntt_big.cu); against the oracle, which follows tests/test_utils/ntt.cpp for any power of two."""
    n = 32768
    q = ob.primes(1, bits, n)[0]
    t = ob.Tables(n, q)
    polys = [stimulus(k, n, q, 700 + i) for i, k in enumerate(STIMULI) if k not in ("all_max", "garbage")]
    polys += [ob.splitmix(n, 900 + i, q) for i in range(150)]           # more items than CTAs
    got = run_fwd(hb, polys, t)
    for i in list(range(5)) + [77, 148, len(polys) - 1]:
        assert np.array_equal(got[i], ob.fwd_ntt(polys[i], t)), ("fwd", i)
    goti = run_inv(hb, polys, t)
    for i in list(range(5)) + [77, 148, len(polys) - 1]:
        assert np.array_equal(goti[i], ob.inv_ntt(polys[i], t)), ("inv", i)
    back = run_inv(hb, list(got), t)
    assert all(np.array_equal(back[i], polys[i]) for i in range(len(polys)))


@pytest.mark.parametrize("n", [16384, 8192, 4096])
def test_forward_every_other_stage_correction_agrees(hb, n):
    """Moduli up to 2^51 (1 + 1/32) take forward FP64 butterflies that correct every other stage (option
    fp64_alt, default on; csrc/modarith.cuh fwd_bfly_fp64_a / _b).  Same words as with the option off and as
    the oracle, on random polynomials and on the words that push the bounds (all at the contract's edge)."""
    import torch

    lim = (1 << 51) + (1 << 46)
    q_edge = next(c for c in range(lim - (lim - 1) % (2 * n), 0, -2 * n) if ob.is_prime(c))
    for q in (ob.primes(1, 51, n)[0], q_edge, ob.primes(1, 40, n)[0]):
        t = ob.Tables(n, q)
        top = q + (q >> 2) - 1
        polys = [ob.splitmix(n, 4000 + i, q) for i in range(3)]
        polys.append(np.full(n, q - 1, dtype=np.uint64))
        polys.append(np.full(n, top, dtype=np.uint64))                                   # in contract for the vote, not canonical
        polys.append(np.where(np.arange(n) % 2 == 0, top, q >> 1).astype(np.uint64))
        polys.append(np.where((np.arange(n) >> (np.arange(n) % 14)) & 1, top, (q >> 1) + 1).astype(np.uint64))
        outs = []
        for alt in (1, 0):
            hb.set_option("fp64_alt", alt)
            try:
                outs.append(run_fwd(hb, polys, t))
            finally:
                hb.set_option("fp64_alt", 1)
        assert np.array_equal(outs[0], outs[1]), q
        for i in range(len(polys)):
            assert np.array_equal(outs[0][i], ob.fwd_ntt(polys[i] % np.uint64(q), t)), (q, i)
