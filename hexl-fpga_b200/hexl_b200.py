"""hexl_b200 -- ctypes binding of lib/libhexl_b200.so (include/hexl_b200.h).

Two groups, mirroring the C ABI:

* device-pointer launchers (``ntt_fwd``, ``ntt_inv``, ``dyadic_multiply``,
  ``KsPlan.keyswitch``) taking ``torch`` CUDA tensors of dtype int64/uint64
  (only ``data_ptr()`` and the current stream are used: torch is plumbing);
* the reference's host API under its own names (``acquire_FPGA_resources``,
  ``set_worksize_NTT`` / ``NTT`` / ``NTTCompleted`` ...; reference
  host/inc/hexl-fpga.h:15-161) taking contiguous ``numpy.uint64`` arrays.

There is deliberately no CPU implementation here: if the shared library is
missing or no CUDA device is present every call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhexl_b200.so")
CXX_LIB_PATH = os.path.join(_HERE, "lib", "libhexl-fpga.so")

u64 = C.c_uint64
p64 = C.POINTER(C.c_uint64)
vp = C.c_void_p


class HexlB200Error(RuntimeError):
    pass


# every exported symbol of include/hexl_b200.h with its signature
_SIGNATURES = {
    "hexl_b200_version": ([], C.c_int),
    "hexl_b200_last_error": ([], C.c_char_p),
    "hexl_b200_device_count": ([], C.c_int),
    "hexl_b200_ntt_fwd": ([vp, vp, vp, u64, u64, u64, vp], C.c_int),
    "hexl_b200_ntt_inv": ([vp, vp, vp, u64, u64, u64, u64, u64, vp], C.c_int),
    "hexl_b200_ntt_fwd_ex": ([vp, vp, vp, u64, u64, u64, u64, u64, vp], C.c_int),
    "hexl_b200_ntt_inv_ex": ([vp, vp, vp, u64, u64, u64, u64, u64, u64, u64, vp], C.c_int),
    "hexl_b200_dyadic_multiply": ([vp, vp, vp, u64, vp, u64, u64, C.c_int, vp], C.c_int),
    "hexl_b200_poly_multiply": ([vp, vp, vp, vp, vp, vp, vp, u64, u64, u64, u64, u64, vp], C.c_int),
    "hexl_b200_ks_plan_create": ([C.POINTER(vp), u64, u64, u64, u64, u64, vp, vp, vp, vp], C.c_int),
    "hexl_b200_ks_plan_destroy": ([vp], C.c_int),
    "hexl_b200_keyswitch": ([vp, vp, vp, u64, vp], C.c_int),
    "hexl_b200_set_option": ([C.c_char_p, C.c_int64], C.c_int),
    "hexl_b200_compute_twiddles": ([u64, u64, vp, p64, p64], C.c_int),
    "hexl_b200_host_acquire": ([], C.c_int),
    "hexl_b200_host_release": ([], C.c_int),
    "hexl_b200_host_set_worksize_dyadic_multiply": ([u64], C.c_int),
    "hexl_b200_host_dyadic_multiply": ([vp, vp, vp, u64, vp, u64], C.c_int),
    "hexl_b200_host_dyadic_multiply_completed": ([], C.c_int),
    "hexl_b200_host_set_worksize_keyswitch": ([u64], C.c_int),
    "hexl_b200_host_keyswitch": ([vp, vp, u64, u64, u64, u64, u64, vp, vp, vp, vp], C.c_int),
    "hexl_b200_host_keyswitch_completed": ([], C.c_int),
    "hexl_b200_host_set_worksize_ntt": ([u64], C.c_int),
    "hexl_b200_host_ntt": ([vp, vp, vp, u64, u64], C.c_int),
    "hexl_b200_host_ntt_completed": ([], C.c_int),
    "hexl_b200_host_set_worksize_intt": ([u64], C.c_int),
    "hexl_b200_host_intt": ([vp, vp, vp, u64, u64, u64, u64], C.c_int),
    "hexl_b200_host_intt_completed": ([], C.c_int),
    "hexl_b200_host_ntt_many": ([vp, u64, u64, vp, vp, u64, u64], C.c_int),
    "hexl_b200_host_intt_many": ([vp, u64, u64, vp, vp, u64, u64, u64, u64], C.c_int),
    "hexl_b200_host_dyadic_multiply_many": ([vp, vp, vp, u64, u64, vp, u64], C.c_int),
    "hexl_b200_host_keyswitch_many": ([vp, vp, u64, u64, u64, u64, u64, u64, vp, vp, vp, vp], C.c_int),
    "hexl_b200_get_stats": ([vp], C.c_int),
    "hexl_b200_host_device_stats": ([C.c_int, vp], C.c_int),
    "hexl_b200_kernel_times": ([vp, u64, vp], C.c_int),
    "hexl_b200_host_pin_buffer": ([vp, u64], C.c_int),
    "hexl_b200_host_unpin_buffer": ([vp], C.c_int),
    "hexl_b200_reset_stats": ([], C.c_int),
}

_lib = None


def lib():
    """Load libhexl_b200.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HexlB200Error(
                f"{LIB_PATH} not found: build it with `make -C hexl-fpga_b200` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (args, res) in _SIGNATURES.items():
            f = getattr(l, name)
            f.argtypes = args
            f.restype = res
        _lib = l
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def last_error():
    return lib().hexl_b200_last_error().decode()


def _check(rc, what):
    if rc != 0:
        raise HexlB200Error(f"{what} failed ({rc}): {last_error()}")


def set_option(name, value):
    _check(lib().hexl_b200_set_option(name.encode(), int(value)), f"set_option({name})")


def compute_twiddles(n, q):
    """(roots, precon, inv_roots, precon_inv, inv_n, inv_n_w) for x^n+1 mod q."""
    out = np.zeros(4 * n, dtype=np.uint64)
    a, b = u64(), u64()
    _check(lib().hexl_b200_compute_twiddles(n, q, out.ctypes.data, C.byref(a), C.byref(b)), "compute_twiddles")
    return out[:n], out[n:2 * n], out[2 * n:3 * n], out[3 * n:], a.value, b.value


class Stats(C.Structure):
    _fields_ = [("kernel_launches", u64), ("h2d_bytes", u64), ("d2h_bytes", u64)]


def get_stats():
    s = Stats()
    _check(lib().hexl_b200_get_stats(C.byref(s)), "get_stats")
    return {"kernel_launches": s.kernel_launches, "h2d_bytes": s.h2d_bytes, "d2h_bytes": s.d2h_bytes}


class DeviceStats(C.Structure):
    _fields_ = [("device", C.c_int32), ("batches", C.c_uint64), ("items", C.c_uint64)]


def device_stats(worker):
    """work done by worker `worker` (0 .. NUM_DEV-1) of the host-pointer runtime since acquire"""
    st = DeviceStats()
    _check(lib().hexl_b200_host_device_stats(worker, C.addressof(st)), "device_stats")
    return {"device": st.device, "batches": st.batches, "items": st.items}


def kernel_times(cap=4096):
    """durations (ms) of the kernel launches timed since the last call (option "time_kernels"), in launch order"""
    ms = np.zeros(cap, dtype=np.float32)
    cnt = u64()
    _check(lib().hexl_b200_kernel_times(ms.ctypes.data, cap, C.byref(cnt)), "kernel_times")
    return ms[:min(cap, cnt.value)].copy()


def pin_buffer(arr):
    """register a numpy array's memory with CUDA in place (hexl_b200_host_pin_buffer); the host API then
    DMAs straight out of / into it instead of staging it through its pinned ring"""
    _check(lib().hexl_b200_host_pin_buffer(arr.ctypes.data, arr.nbytes), "pin_buffer")


def unpin_buffer(arr):
    _check(lib().hexl_b200_host_unpin_buffer(arr.ctypes.data), "unpin_buffer")


def reset_stats():
    _check(lib().hexl_b200_reset_stats(), "reset_stats")


# --------------------------------------------------------------------------
# device-pointer launchers (torch CUDA tensors)
# --------------------------------------------------------------------------
def _dptr(t):
    import torch

    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HexlB200Error("expected a CUDA tensor")
    if t.element_size() != 8 or not t.is_contiguous():
        raise HexlB200Error("expected a contiguous 64-bit integer tensor")
    return t.data_ptr()


def _stream():
    import torch

    return torch.cuda.current_stream().cuda_stream


def ntt_fwd(operand, roots, precon, q, n, input_mod_factor=1, output_mod_factor=1):
    """In-place batched forward NTT of `operand` ([batch, n] on the GPU); mod factors as in the reference's
    NTT::ComputeForward (tests/test_utils/ntt.cpp:442-455)."""
    batch = operand.numel() // n
    _check(lib().hexl_b200_ntt_fwd_ex(_dptr(operand), _dptr(roots), _dptr(precon), q, n, batch, input_mod_factor,
                                      output_mod_factor, _stream()), "ntt_fwd")


def ntt_inv(operand, inv_roots, precon_inv, q, inv_n, inv_n_w, n, input_mod_factor=1, output_mod_factor=1):
    batch = operand.numel() // n
    _check(lib().hexl_b200_ntt_inv_ex(_dptr(operand), _dptr(inv_roots), _dptr(precon_inv), q, inv_n, inv_n_w, n, batch,
                                      input_mod_factor, output_mod_factor, _stream()), "ntt_inv")


def poly_multiply(result, a, b, roots, precon, inv_roots, precon_inv, q, inv_n, inv_n_w, n):
    """result = a * b mod (x^n + 1, q) for every polynomial of the [batch, n] GPU tensors."""
    batch = a.numel() // n
    _check(lib().hexl_b200_poly_multiply(_dptr(result), _dptr(a), _dptr(b), _dptr(roots), _dptr(precon),
                                         _dptr(inv_roots), _dptr(precon_inv), q, inv_n, inv_n_w, n, batch,
                                         _stream()), "poly_multiply")


def dyadic_multiply(results, op1, op2, n, moduli, n_moduli, batch, moduli_per_item=False):
    _check(lib().hexl_b200_dyadic_multiply(_dptr(results), _dptr(op1), _dptr(op2), n, _dptr(moduli),
                                           n_moduli, batch, int(bool(moduli_per_item)), _stream()),
           "dyadic_multiply")


def _np64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


class KsPlan:
    """Device-resident constants of one keyswitch key set (hexl_b200_ks_plan)."""

    def __init__(self, n, decomp, key_mod, rns, key_comp, moduli, keys, modswitch, twiddles=None):
        self.n, self.decomp = n, decomp
        moduli = _np64(moduli)
        modswitch = _np64(modswitch)
        self._keys = [_np64(k) for k in keys]
        arr = (vp * len(self._keys))(*[k.ctypes.data for k in self._keys])
        tw = _np64(twiddles) if twiddles is not None else None
        h = vp()
        _check(lib().hexl_b200_ks_plan_create(C.byref(h), n, decomp, key_mod, rns, key_comp,
                                              moduli.ctypes.data, C.cast(arr, vp), modswitch.ctypes.data,
                                              tw.ctypes.data if tw is not None else None), "ks_plan_create")
        self._h = h

    def keyswitch(self, result, t_target, batch):
        _check(lib().hexl_b200_keyswitch(self._h, _dptr(result), _dptr(t_target), batch, _stream()),
               "keyswitch")

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().hexl_b200_ks_plan_destroy(self._h)
            except Exception:      # interpreter shutdown
                pass
            self._h = None

    def __del__(self):
        self.close()


# --------------------------------------------------------------------------
# the reference's host API, same names (numpy uint64 arrays = host pointers)
# --------------------------------------------------------------------------
def _hptr(a, writable=False):
    if not isinstance(a, np.ndarray) or a.dtype != np.uint64 or not a.flags.c_contiguous:
        raise HexlB200Error("expected a C-contiguous numpy uint64 array")
    if writable and not a.flags.writeable:
        raise HexlB200Error("output array is read-only")
    return a.ctypes.data


def acquire_FPGA_resources():
    _check(lib().hexl_b200_host_acquire(), "acquire_FPGA_resources")


def release_FPGA_resources():
    _check(lib().hexl_b200_host_release(), "release_FPGA_resources")


def set_worksize_DyadicMultiply(ws):
    _check(lib().hexl_b200_host_set_worksize_dyadic_multiply(ws), "set_worksize_DyadicMultiply")


def DyadicMultiply(results, operand1, operand2, n, moduli, n_moduli):
    _check(lib().hexl_b200_host_dyadic_multiply(_hptr(results, True), _hptr(operand1), _hptr(operand2), n,
                                                _hptr(moduli), n_moduli), "DyadicMultiply")


def DyadicMultiplyCompleted():
    _check(lib().hexl_b200_host_dyadic_multiply_completed(), "DyadicMultiplyCompleted")
    return True


def set_worksize_KeySwitch(ws):
    _check(lib().hexl_b200_host_set_worksize_keyswitch(ws), "set_worksize_KeySwitch")


class KeyArray:
    """A `const uint64_t**` for KeySwitch; keep it alive (and reuse it) across
    calls -- like the reference, the library caches keys by this pointer."""

    def __init__(self, keys):
        self.keys = [_np64(k) for k in keys]
        self.arr = (vp * len(self.keys))(*[k.ctypes.data for k in self.keys])

    @property
    def ptr(self):
        return C.cast(self.arr, vp)


def KeySwitch(result, t_target_iter_ptr, n, decomp_modulus_size, key_modulus_size, rns_modulus_size,
              key_component_count, moduli, k_switch_keys, modswitch_factors, twiddle_factors=None):
    if not isinstance(k_switch_keys, KeyArray):
        raise HexlB200Error("k_switch_keys must be a KeyArray (pointer identity is the cache key)")
    _check(lib().hexl_b200_host_keyswitch(
        _hptr(result, True), _hptr(t_target_iter_ptr), n, decomp_modulus_size, key_modulus_size,
        rns_modulus_size, key_component_count, _hptr(moduli), k_switch_keys.ptr, _hptr(modswitch_factors),
        _hptr(twiddle_factors) if twiddle_factors is not None else None), "KeySwitch")


def KeySwitchCompleted():
    _check(lib().hexl_b200_host_keyswitch_completed(), "KeySwitchCompleted")
    return True


def set_worksize_NTT(ws):
    _check(lib().hexl_b200_host_set_worksize_ntt(ws), "_set_worksize_NTT")


def NTT(operand, root_of_unity_powers, precon_root_of_unity_powers, coeff_modulus, n):
    _check(lib().hexl_b200_host_ntt(_hptr(operand, True), _hptr(root_of_unity_powers),
                                    _hptr(precon_root_of_unity_powers), coeff_modulus, n), "_NTT")


def NTTCompleted():
    _check(lib().hexl_b200_host_ntt_completed(), "_NTTCompleted")
    return True


def set_worksize_INTT(ws):
    _check(lib().hexl_b200_host_set_worksize_intt(ws), "_set_worksize_INTT")


def INTT(operand, inv_root_of_unity_powers, precon_inv_root_of_unity_powers, coeff_modulus, inv_n, inv_n_w, n):
    _check(lib().hexl_b200_host_intt(_hptr(operand, True), _hptr(inv_root_of_unity_powers),
                                     _hptr(precon_inv_root_of_unity_powers), coeff_modulus, inv_n, inv_n_w,
                                     n), "_INTT")


def INTTCompleted():
    _check(lib().hexl_b200_host_intt_completed(), "_INTTCompleted")
    return True


# bulk submission: `count` calls in one FFI crossing; `ptr` arguments are raw
# host addresses (numpy .ctypes.data or a pinned torch tensor's data_ptr()).
def NTT_many(base_ptr, stride_words, count, roots, precon, q, n):
    _check(lib().hexl_b200_host_ntt_many(base_ptr, stride_words, count, _hptr(roots), _hptr(precon), q, n),
           "_NTT x count")


def INTT_many(base_ptr, stride_words, count, inv_roots, precon_inv, q, inv_n, inv_n_w, n):
    _check(lib().hexl_b200_host_intt_many(base_ptr, stride_words, count, _hptr(inv_roots), _hptr(precon_inv), q,
                                          inv_n, inv_n_w, n), "_INTT x count")


def DyadicMultiply_many(res_ptr, op1_ptr, op2_ptr, count, n, moduli, n_moduli):
    _check(lib().hexl_b200_host_dyadic_multiply_many(res_ptr, op1_ptr, op2_ptr, count, n, _hptr(moduli),
                                                     n_moduli), "DyadicMultiply x count")


def KeySwitch_many(res_ptr, t_ptr, count, n, D, K, R, Cc, moduli, keys, msf, twiddles=None):
    _check(lib().hexl_b200_host_keyswitch_many(res_ptr, t_ptr, count, n, D, K, R, Cc, _hptr(moduli), keys.ptr,
                                               _hptr(msf), _hptr(twiddles) if twiddles is not None else None),
           "KeySwitch x count")


# the reference spells the deprecated entry points with a leading underscore
_set_worksize_NTT, _NTT, _NTTCompleted = set_worksize_NTT, NTT, NTTCompleted
_set_worksize_INTT, _INTT, _INTTCompleted = set_worksize_INTT, INTT, INTTCompleted
