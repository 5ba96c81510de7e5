"""Runs the reference's OWN keyswitch device code on the CPU (oracle/_ref/ks_ref_emul: the
reference's device/keyswitch.cpp compiled unmodified against oracle/sycl_shim, see
oracle/ref_ks_emul.cpp).  Test infrastructure only."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ks_ref_emul")


def available():
    return os.path.exists(BIN)


def keyswitch(result, t_target, n, D, K, moduli, keys, msf, batch=1, timeout=1800):
    """result after the reference pipeline's accumulate, `batch` contiguous items (K must be 7, D <= 6)."""
    parts = [np.array([n, D, K, batch], dtype=np.uint64), np.asarray(moduli, dtype=np.uint64),
             np.asarray(msf, dtype=np.uint64)]
    parts += [np.ascontiguousarray(k, dtype=np.uint64).reshape(-1) for k in keys]
    parts += [np.ascontiguousarray(t_target, dtype=np.uint64).reshape(-1),
              np.ascontiguousarray(result, dtype=np.uint64).reshape(-1)]
    with tempfile.TemporaryDirectory() as d:
        pin, pout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        np.concatenate(parts).tofile(pin)
        subprocess.run([BIN, pin, pout], check=True, timeout=timeout)
        return np.fromfile(pout, dtype=np.uint64)


def _run(binary, header, arrays, out_words, timeout=900):
    path = os.path.join(ROOT, "oracle", "_ref", binary)
    parts = [np.array(header, dtype=np.uint64)] + [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1) for a in arrays]
    with tempfile.TemporaryDirectory() as d:
        pin, pout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        np.concatenate(parts).tofile(pin)
        subprocess.run([path, pin, pout], check=True, timeout=timeout)
        out = np.fromfile(pout, dtype=np.uint64)
    assert out.size == out_words
    return out


def device_available(kind):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "dev_ref_emul_" + kind))


def fwd_ntt(data, q, roots, precon):
    """reference device/fwd_ntt.cpp (N = 16384) on `data` [batch][16384]"""
    data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 16384)
    return _run("dev_ref_emul_ntt", [data.shape[0], q], [roots, precon, data], data.size).reshape(data.shape)


def inv_ntt(data, q, inv_n, inv_n_w, inv_roots, precon_inv):
    """reference device/inv_ntt.cpp (N = 16384)"""
    data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 16384)
    return _run("dev_ref_emul_intt", [data.shape[0], q, inv_n, inv_n_w], [inv_roots, precon_inv, data],
                data.size).reshape(data.shape)


def dyadic(op1, op2, n, moduli, batch):
    """reference device/dyadic_multiply.cpp; moduli [batch][M] (per item, as its tests pass them)"""
    moduli = np.ascontiguousarray(moduli, dtype=np.uint64).reshape(batch, -1)
    M = moduli.shape[1]
    return _run("dev_ref_emul_dyadic", [batch, n, M], [moduli, op1, op2], batch * 3 * M * n)
