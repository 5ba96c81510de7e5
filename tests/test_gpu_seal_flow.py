"""SEAL-layout caller flow: what the reference's SEAL
bridge hands to intel::hexl::KeySwitch for CKKS relinearisation
(experimental/bridge-seal/patches/hexl-fpga-BRIDGE-seal-4.0.0.patch:205-263):

    result            = (c0, c1) of the size-3 ciphertext, NTT form, [c][i][coeff]
    t_target          = c2, NTT form, limbs of the ciphertext's moduli
    k_switch_keys[j]  = relinearisation key data: for digit j, 2 x K polynomials
                        (-(a_j s + e_j) + q_k s^2 [i == j], a_j) in NTT form
    moduli            = key_parms.coeff_modulus (special prime last)
    modswitch_factors = q_k^-1 mod q_i

After the call result = (c0 + ks0, c1 + ks1), and decrypting it with s must give
the same plaintext as decrypting (c0, c1, c2) with (1, s, s^2), up to the
key-switching noise.  Keys are REAL RLWE keys built here with numpy; the GPU
path is driven through the host API exactly as the patched SEAL drives it."""
import numpy as np
import pytest

import oracle_binding as ob

pytestmark = pytest.mark.gpu


def pmul(a, b, q):
    """pointwise product mod q of uint64 vectors (q < 2^61) via Python ints"""
    return np.array((a.astype(object) * b.astype(object)) % q, dtype=np.uint64)


# the last two: the reference's largest shape 6/7/7/2 and BASELINE's decomp 7 / key 8 at N = 16384 -- an
# algebraic check at the headline size that does not route through the oracle's keyswitch formula
@pytest.mark.parametrize("n,D,K,bits", [(2048, 3, 4, 40), (4096, 2, 4, 50), (16384, 6, 7, 51), (16384, 7, 8, 51)])
def test_relinearize_like_seal(acquired, n, D, K, bits):
    hb = acquired
    rng = np.random.default_rng(7 * n + D)
    moduli = ob.primes(K, bits, n)
    qk = moduli[K - 1]
    tabs = [ob.Tables(n, q) for q in moduli]

    def ntt(poly, i):   # small signed coefficients -> NTT form mod q_i
        return ob.fwd_ntt(np.array([int(x) % moduli[i] for x in poly], dtype=np.uint64), tabs[i])

    s = rng.integers(-1, 2, n)
    s_ntt = [ntt(s, i) for i in range(K)]
    s2_ntt = [pmul(s_ntt[i], s_ntt[i], moduli[i]) for i in range(K)]
    # relinearisation keys, one per digit j
    keys = []
    for j in range(D):
        e = rng.integers(-3, 4, n)
        k = np.zeros((2, K, n), dtype=np.uint64)
        for i in range(K):
            q = moduli[i]
            a = ob.splitmix(n, 1000 * j + i + 5, q)
            c0 = (q - (pmul(a, s_ntt[i], q).astype(object) + ntt(e, i).astype(object)) % q) % q
            if i == j:
                c0 = (c0 + (qk % q) * s2_ntt[i].astype(object)) % q
            k[0, i] = np.array([int(x) for x in c0], dtype=np.uint64)
            k[1, i] = a
        keys.append(k.reshape(-1))
    o = ob.oracle()
    msf = np.array([o.ho_inv_mod(qk % q, q) for q in moduli], dtype=np.uint64)
    msf[K - 1] = 0
    mod_arr = np.array(moduli, dtype=np.uint64)
    # a batch of size-3 "ciphertexts" (uniform limbs; the identity holds for any c)
    B = 3 if n < 16384 else 2
    key_arr = hb.KeyArray(keys)
    cts = []
    hb.set_worksize_KeySwitch(B)
    for b in range(B):
        c = np.stack([[ob.splitmix(n, 50 * b + 10 * comp + i, moduli[i]) for i in range(D)] for comp in range(3)])
        result = c[:2].reshape(-1).copy()                       # (c0, c1), [c][i][coeff]
        t_target = c[2].reshape(-1).copy()                      # c2
        cts.append((c, result, t_target))
        hb.KeySwitch(result, t_target, n, D, K, D + 1, 2, mod_arr, key_arr, msf)
    assert hb.KeySwitchCompleted()
    for c, result, _ in cts:
        r = result.reshape(2, D, n)
        first = None
        for i in range(D):
            q = moduli[i]
            lhs = (r[0, i].astype(object) + pmul(r[1, i], s_ntt[i], q).astype(object)) % q
            rhs = (c[0, i].astype(object) + pmul(c[1, i], s_ntt[i], q).astype(object)
                   + pmul(c[2, i], s2_ntt[i], q).astype(object)) % q
            diff = np.array([int(x) for x in (lhs - rhs) % q], dtype=np.uint64)
            noise = [int(x) for x in ob.inv_ntt(diff, tabs[i])]
            noise = [x - q if x > q // 2 else x for x in noise]
            assert max(abs(x) for x in noise) < 8 * D * n, max(abs(x) for x in noise)
            if first is None:
                first = noise
            else:
                assert noise == first      # the same integer polynomial in every limb
