"""ctypes binding of the TEST-ONLY CPU oracle (oracle/liboracle.so) and, when
present, of the reference's own scalar NTT (oracle/_ref/libhexl_ref.so)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
u64, vp = C.c_uint64, C.c_void_p

_o = None
_r = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def oracle():
    global _o
    if _o is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        o = C.CDLL(path)
        for name in ("ho_mul_mod", "ho_add_mod", "ho_sub_mod", "ho_pow_mod"):
            getattr(o, name).restype = u64
        o.ho_mul_mod.argtypes = o.ho_add_mod.argtypes = o.ho_sub_mod.argtypes = [u64, u64, u64]
        o.ho_pow_mod.argtypes = [u64, u64, u64]
        o.ho_inv_mod.argtypes = [u64, u64]; o.ho_inv_mod.restype = u64
        o.ho_generate_primes.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_size_t]
        o.ho_generate_primes.restype = C.c_size_t
        o.ho_min_primitive_root.argtypes = [u64, u64]; o.ho_min_primitive_root.restype = u64
        o.ho_mult_factor64.argtypes = [u64, u64]; o.ho_mult_factor64.restype = u64
        o.ho_compute_roots.argtypes = [u64, u64, u64, vp, vp, vp, vp]
        o.ho_compute_roots_keyswitch.argtypes = [u64, u64, u64, vp]
        o.ho_fwd_ntt.argtypes = [vp, u64, u64, vp, vp]
        o.ho_inv_ntt.argtypes = [vp, u64, u64, vp, vp, u64, u64]
        o.ho_fwd_ntt_lazy.argtypes = [vp, u64, u64, vp, vp]
        o.ho_inv_ntt_lazy.argtypes = [vp, u64, u64, vp, vp, u64, u64]
        o.ho_fwd_ntt_reference.argtypes = [vp, u64, u64, vp]
        o.ho_fwd_ntt_batch.argtypes = [vp, u64, u64, u64, vp, vp, C.c_int]
        o.ho_inv_ntt_batch.argtypes = [vp, u64, u64, u64, vp, vp, u64, u64, C.c_int]
        o.ho_dyadic_multiply.argtypes = [vp, vp, vp, u64, vp, u64]
        o.ho_dyadic_multiply_batch.argtypes = [vp, vp, vp, u64, vp, u64, u64, C.c_int, C.c_int]
        ks = [vp, vp, u64, u64, u64, u64, u64, vp, vp, vp]
        o.ho_keyswitch.argtypes = ks; o.ho_keyswitch.restype = C.c_int
        o.ho_keyswitch_alt.argtypes = ks; o.ho_keyswitch_alt.restype = C.c_int
        o.ho_keyswitch_batch.argtypes = [vp, vp, u64, u64, u64, u64, u64, u64, vp, vp, vp, C.c_int]
        o.ho_keyswitch_batch.restype = C.c_int
        o.ho_keyswitch_alt_batch.argtypes = [vp, vp, u64, u64, u64, u64, u64, u64, vp, vp, vp, C.c_int]
        o.ho_keyswitch_alt_batch.restype = C.c_int
        o.ho_fnv1a.argtypes = [vp, C.c_size_t]; o.ho_fnv1a.restype = u64
        o.ho_splitmix_fill.argtypes = [vp, C.c_size_t, u64, u64]; o.ho_splitmix_fill.restype = u64
        o.ho_max_threads.restype = C.c_int
        o.ho_is_prime.argtypes = [u64]; o.ho_is_prime.restype = C.c_int
        _o = o
    return _o


def ref():
    """The reference's own scalar NTT (None if oracle/_ref was never built)."""
    global _r
    if _r is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libhexl_ref.so")
        if not os.path.exists(path):
            return None
        r = C.CDLL(path)
        r.ref_generate_primes.argtypes = [vp, u64, u64, u64]; r.ref_generate_primes.restype = u64
        r.ref_min_primitive_root.argtypes = [u64, u64]; r.ref_min_primitive_root.restype = u64
        r.ref_inverse_mod.argtypes = [u64, u64]; r.ref_inverse_mod.restype = u64
        r.ref_multiply_mod.argtypes = [u64, u64, u64]; r.ref_multiply_mod.restype = u64
        r.ref_tables.argtypes = [u64, u64, vp, vp, vp, vp]
        r.ref_fwd_ntt.argtypes = [vp, u64, u64]
        r.ref_inv_ntt.argtypes = [vp, u64, u64]
        if hasattr(r, "ref_fwd_ntt_factors"):
            r.ref_fwd_ntt_factors.argtypes = [vp, u64, u64, u64, u64]
            r.ref_inv_ntt_factors.argtypes = [vp, u64, u64, u64, u64]
        r.ref_fwd_ntt_batch.argtypes = [vp, u64, u64, u64, vp, vp, C.c_int]
        r.ref_inv_ntt_batch.argtypes = [vp, u64, u64, u64, vp, vp, C.c_int]
        _r = r
    return _r


def P(a):
    return a.ctypes.data


def primes(num, bits, n):
    out = np.zeros(num, dtype=np.uint64)
    got = oracle().ho_generate_primes(P(out), num, bits, n)
    assert got == num
    return [int(x) for x in out]


def is_prime(n):
    return bool(oracle().ho_is_prime(n))


class Tables:
    """hexl-layout twiddle tables of one modulus, computed by the oracle."""

    def __init__(self, n, q):
        o = oracle()
        self.n, self.q = n, q
        self.w = o.ho_min_primitive_root(2 * n, q)
        self.roots = np.zeros(n, dtype=np.uint64)
        self.precon = np.zeros(n, dtype=np.uint64)
        self.inv_roots = np.zeros(n, dtype=np.uint64)
        self.precon_inv = np.zeros(n, dtype=np.uint64)
        o.ho_compute_roots(n, q, self.w, P(self.roots), P(self.precon), P(self.inv_roots), P(self.precon_inv))
        self.inv_n = o.ho_inv_mod(n % q, q)
        self.inv_n_w = o.ho_mul_mod(self.inv_n, int(self.inv_roots[n - 1]), q)


def splitmix(n, seed, q=0):
    out = np.zeros(n, dtype=np.uint64)
    oracle().ho_splitmix_fill(P(out), n, seed, q)
    return out


def fnv(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return oracle().ho_fnv1a(P(a), a.size)


def fwd_ntt(a, t):
    a = np.array(a, dtype=np.uint64, copy=True)
    oracle().ho_fwd_ntt(P(a), t.n, t.q, P(t.roots), P(t.precon))
    return a


def inv_ntt(a, t):
    a = np.array(a, dtype=np.uint64, copy=True)
    oracle().ho_inv_ntt(P(a), t.n, t.q, P(t.inv_roots), P(t.precon_inv), t.inv_n, t.inv_n_w)
    return a


def fwd_ntt_lazy(a, t):
    """output_mod_factor = 4: words in [0, 4q)"""
    a = np.array(a, dtype=np.uint64, copy=True)
    oracle().ho_fwd_ntt_lazy(P(a), t.n, t.q, P(t.roots), P(t.precon))
    return a


def inv_ntt_lazy(a, t):
    """output_mod_factor = 2: words in [0, 2q)"""
    a = np.array(a, dtype=np.uint64, copy=True)
    oracle().ho_inv_ntt_lazy(P(a), t.n, t.q, P(t.inv_roots), P(t.precon_inv), t.inv_n, t.inv_n_w)
    return a


def dyadic(op1, op2, n, moduli, batch=1, per_item=False):
    moduli = np.ascontiguousarray(moduli, dtype=np.uint64)
    M = moduli.size // (batch if per_item else 1)
    res = np.zeros(batch * 3 * M * n, dtype=np.uint64)
    oracle().ho_dyadic_multiply_batch(P(res), P(op1), P(op2), n, P(moduli), M, batch, int(per_item),
                                      oracle().ho_max_threads())
    return res


def keyswitch(result, t_target, n, D, K, moduli, keys, msf, batch=1, alt=False, threads=0):
    """result (accumulated in a copy) for `batch` contiguous items."""
    o = oracle()
    res = np.array(result, dtype=np.uint64, copy=True)
    moduli = np.ascontiguousarray(moduli, dtype=np.uint64)
    msf = np.ascontiguousarray(msf, dtype=np.uint64)
    keys = [np.ascontiguousarray(k, dtype=np.uint64) for k in keys]
    arr = (vp * len(keys))(*[k.ctypes.data for k in keys])
    if alt:
        rc = o.ho_keyswitch_alt_batch(P(res), P(t_target), batch, n, D, K, D + 1, 2, P(moduli), C.cast(arr, vp),
                                      P(msf), threads or o.ho_max_threads())
    else:
        rc = o.ho_keyswitch_batch(P(res), P(t_target), batch, n, D, K, D + 1, 2, P(moduli), C.cast(arr, vp),
                                  P(msf), threads or o.ho_max_threads())
    assert rc == 0
    return res
