/*
 * hexl_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see hexl_oracle.h).
 *
 * Plain-C restatement of the reference algorithms.  Every function cites the
 * reference file:line (relative to /root/reference) it follows.  Nothing in
 * hexl-fpga_b200/ may link this file.
 */
#include "hexl_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------ */
/* number theory                                                            */
/* ------------------------------------------------------------------------ */

/* tests/test_utils/ntt.cpp:43-52 (MultiplyUIntMod: 128-bit product, exact). */
uint64_t ho_mul_mod(uint64_t a, uint64_t b, uint64_t q) {
    return (uint64_t)(((u128)a * b) % q);
}

/* tests/test_utils/ntt.cpp:64-72 (AddUIntMod, inputs < q). */
uint64_t ho_add_mod(uint64_t a, uint64_t b, uint64_t q) {
    uint64_t s = a + b;
    return (s >= q || s < a) ? s - q : s;
}

/* tests/test_utils/ntt.cpp:74-82 (SubUIntMod, inputs < q). */
uint64_t ho_sub_mod(uint64_t a, uint64_t b, uint64_t q) {
    return (a >= b) ? a - b : a + q - b;
}

/* tests/test_utils/ntt.cpp:84-96 (PowMod, square and multiply). */
uint64_t ho_pow_mod(uint64_t base, uint64_t exp, uint64_t q) {
    uint64_t r = 1 % q;
    base %= q;
    while (exp) {
        if (exp & 1) r = ho_mul_mod(r, base, q);
        base = ho_mul_mod(base, base, q);
        exp >>= 1;
    }
    return r;
}

/* tests/test_utils/ntt.cpp:16-41 (InverseUIntMod, extended Euclid). */
uint64_t ho_inv_mod(uint64_t a, uint64_t q) {
    __int128 t = 0, newt = 1;
    __int128 r = q, newr = a % q;
    while (newr != 0) {
        __int128 quo = r / newr;
        __int128 tmp = t - quo * newt;
        t = newt;
        newt = tmp;
        tmp = r - quo * newr;
        r = newr;
        newr = tmp;
    }
    if (r != 1) return 0; /* not invertible */
    if (t < 0) t += q;
    return (uint64_t)t;
}

/* tests/test_utils/ntt.cpp:160-171 (ReverseBitsUInt). */
uint64_t ho_reverse_bits(uint64_t x, uint64_t bit_width) {
    uint64_t rev = 0;
    for (uint64_t i = 0; i < bit_width; ++i) {
        rev = (rev << 1) | (x & 1);
        x >>= 1;
    }
    return rev;
}

/* tests/test_utils/ntt.cpp:173-222 (Miller-Rabin with the 12 fixed bases that
 * are sufficient below 2^64). */
int ho_is_prime(uint64_t n) {
    static const uint64_t bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (int i = 0; i < 12; ++i) {
        if (n == bases[i]) return 1;
        if (n % bases[i] == 0) return 0;
    }
    uint64_t d = n - 1;
    int r = 0;
    while ((d & 1) == 0) {
        d >>= 1;
        ++r;
    }
    for (int i = 0; i < 12; ++i) {
        uint64_t x = ho_pow_mod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int composite = 1;
        for (int k = 1; k < r; ++k) {
            x = ho_mul_mod(x, x, n);
            if (x == n - 1) {
                composite = 0;
                break;
            }
        }
        if (composite) return 0;
    }
    return 1;
}

/* tests/test_utils/ntt.cpp:224-247 (GeneratePrimes: scan 2^b+1, step 2N). */
size_t ho_generate_primes(uint64_t* out, size_t num_primes, size_t bit_size,
                          size_t ntt_size) {
    uint64_t v = ((uint64_t)1 << bit_size) + 1;
    uint64_t hi = (uint64_t)1 << (bit_size + 1);
    size_t found = 0;
    while (v < hi && found < num_primes) {
        if (ho_is_prime(v)) out[found++] = v;
        v += 2 * (uint64_t)ntt_size;
    }
    return found;
}

/* tests/test_utils/ntt.cpp:98-107 (IsPrimitiveRoot). */
int ho_is_primitive_root(uint64_t root, uint64_t degree, uint64_t q) {
    if (root == 0) return 0;
    return ho_pow_mod(root, degree / 2, q) == q - 1;
}

/* tests/test_utils/ntt.cpp:109-158 (GeneratePrimitiveRoot + MinimalPrimitive-
 * Root).  The reference draws random candidates; the minimum over the orbit
 * {g^(odd)} is independent of which primitive root g is found, so a
 * deterministic candidate scan gives the identical value. */
uint64_t ho_min_primitive_root(uint64_t degree, uint64_t q) {
    uint64_t quot = (q - 1) / degree;
    uint64_t g = 0;
    for (uint64_t cand = 2; cand < q; ++cand) {
        uint64_t r = ho_pow_mod(cand, quot, q);
        if (ho_is_primitive_root(r, degree, q)) {
            g = r;
            break;
        }
    }
    if (!g) return 0;
    uint64_t gsq = ho_mul_mod(g, g, q);
    uint64_t cur = g, best = g;
    for (uint64_t i = 0; i < degree / 2; ++i) {
        if (cur < best) best = cur;
        cur = ho_mul_mod(cur, gsq, q);
    }
    return best;
}

/* tests/test_utils/ntt.hpp:17-37 (MultiplyFactor, bit_shift 64). */
uint64_t ho_mult_factor64(uint64_t operand, uint64_t q) {
    return (uint64_t)((((u128)operand) << 64) / q);
}

/* ------------------------------------------------------------------------ */
/* twiddle tables                                                           */
/* ------------------------------------------------------------------------ */

static unsigned ilog2(uint64_t n) {
    unsigned l = 0;
    while (((uint64_t)1 << l) < n) ++l;
    return l;
}

/* tests/test_utils/ntt.cpp:290-384 (ComputeRootOfUnityPowers). */
void ho_compute_roots(uint64_t n, uint64_t q, uint64_t w, uint64_t* roots,
                      uint64_t* precon, uint64_t* inv_roots,
                      uint64_t* precon_inv) {
    unsigned bits = ilog2(n);
    uint64_t* inv_br = (uint64_t*)malloc(n * sizeof(uint64_t));
    roots[0] = 1;
    inv_br[0] = 1;
    uint64_t prev = 0;
    for (uint64_t i = 1; i < n; ++i) {
        uint64_t idx = ho_reverse_bits(i, bits);
        roots[idx] = ho_mul_mod(roots[prev], w, q);
        prev = idx;
    }
    /* inverse of w^k is w^(2n-k): avoid n extended-Euclid calls */
    {
        uint64_t winv = ho_inv_mod(w, q);
        uint64_t cur = 1;
        prev = 0;
        for (uint64_t i = 1; i < n; ++i) {
            uint64_t idx = ho_reverse_bits(i, bits);
            cur = ho_mul_mod(cur, winv, q);
            inv_br[idx] = cur;
        }
    }
    if (inv_roots) {
        uint64_t idx = 1;
        inv_roots[0] = inv_br[0];
        for (uint64_t m = n >> 1; m > 0; m >>= 1)
            for (uint64_t i = 0; i < m; ++i) inv_roots[idx++] = inv_br[m + i];
    }
    if (precon)
        for (uint64_t i = 0; i < n; ++i) precon[i] = ho_mult_factor64(roots[i], q);
    if (precon_inv && inv_roots)
        for (uint64_t i = 0; i < n; ++i)
            precon_inv[i] = ho_mult_factor64(inv_roots[i], q);
    free(inv_br);
}

/* host/src/twiddle-factors.cpp:16-62 + host/src/fpga.cpp:1102-1109. */
void ho_compute_roots_keyswitch(uint64_t n, uint64_t q, uint64_t w,
                                uint64_t* t) {
    uint64_t* inv_roots = t;
    uint64_t* precon_inv = t + n;
    uint64_t* roots = t + 2 * n;
    uint64_t* precon = t + 3 * n;
    uint64_t* tmp_inv = (uint64_t*)malloc(n * sizeof(uint64_t));
    ho_compute_roots(n, q, w, roots, precon, tmp_inv, NULL);
    precon[0] = 0;
    /* 0-based: drop the leading 1, terminate with 0 */
    for (uint64_t i = 0; i + 1 < n; ++i) inv_roots[i] = tmp_inv[i + 1];
    inv_roots[n - 1] = 0;
    for (uint64_t i = 0; i < n; ++i)
        precon_inv[i] = ho_mult_factor64(inv_roots[i], q);
    free(tmp_inv);
}

/* ------------------------------------------------------------------------ */
/* NTT                                                                       */
/* ------------------------------------------------------------------------ */

/* tests/test_utils/ntt.hpp:87-101 (MultiplyUIntModLazy<64>): result in
 * [0, 2q) for any x when y < q. */
static inline uint64_t mul_lazy(uint64_t x, uint64_t y, uint64_t y_precon,
                                uint64_t q) {
    uint64_t Q = (uint64_t)(((u128)x * y_precon) >> 64);
    return y * x - Q * q;
}

/* tests/test_utils/ntt.cpp:474-548 (ForwardTransformToBitReverse64 with
 * output_mod_factor == 1); same op order as device/fwd_ntt.cpp:282-386. */
/* tests/test_utils/ntt.cpp:474-548 with output_mod_factor = 4: the final correction loop (:535-546)
 * is skipped and the lazy words in [0, 4q) are the result. */
void ho_fwd_ntt_lazy(uint64_t* a, uint64_t n, uint64_t q, const uint64_t* roots,
                     const uint64_t* precon) {
    uint64_t twice_q = q << 1;
    uint64_t t = n >> 1;
    for (uint64_t m = 1; m < n; m <<= 1) {
        uint64_t j1 = 0;
        for (uint64_t i = 0; i < m; ++i) {
            uint64_t W = roots[m + i], Wp = precon[m + i];
            uint64_t* X = a + j1;
            uint64_t* Y = X + t;
            for (uint64_t j = 0; j < t; ++j) {
                uint64_t x = X[j];
                uint64_t tx = (x >= twice_q) ? x - twice_q : x;
                uint64_t T = mul_lazy(Y[j], W, Wp, q);
                X[j] = tx + T;
                Y[j] = tx + twice_q - T;
            }
            j1 += t << 1;
        }
        t >>= 1;
    }
}

void ho_fwd_ntt(uint64_t* a, uint64_t n, uint64_t q, const uint64_t* roots,
                const uint64_t* precon) {
    uint64_t twice_q = q << 1;
    uint64_t t = n >> 1;
    for (uint64_t m = 1; m < n; m <<= 1) {
        uint64_t j1 = 0;
        for (uint64_t i = 0; i < m; ++i) {
            uint64_t W = roots[m + i], Wp = precon[m + i];
            uint64_t* X = a + j1;
            uint64_t* Y = X + t;
            for (uint64_t j = 0; j < t; ++j) {
                uint64_t x = X[j];
                uint64_t tx = (x >= twice_q) ? x - twice_q : x;
                uint64_t T = mul_lazy(Y[j], W, Wp, q);
                X[j] = tx + T;
                Y[j] = tx + twice_q - T;
            }
            j1 += t << 1;
        }
        t >>= 1;
    }
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t v = a[i];
        if (v >= twice_q) v -= twice_q;
        if (v >= q) v -= q;
        a[i] = v;
    }
}

/* tests/test_utils/ntt.cpp:550-578 (ReferenceForwardTransformToBitReverse). */
void ho_fwd_ntt_reference(uint64_t* a, uint64_t n, uint64_t q,
                          const uint64_t* roots) {
    uint64_t t = n >> 1;
    for (uint64_t m = 1; m < n; m <<= 1) {
        uint64_t j1 = 0;
        for (uint64_t i = 0; i < m; ++i) {
            uint64_t W = roots[m + i];
            for (uint64_t j = j1; j < j1 + t; ++j) {
                uint64_t x = a[j];
                uint64_t wy = ho_mul_mod(a[j + t], W, q);
                a[j] = ho_add_mod(x, wy, q);
                a[j + t] = ho_sub_mod(x, wy, q);
            }
            j1 += t << 1;
        }
        t >>= 1;
    }
}

/* tests/test_utils/ntt.cpp:580-659 (InverseTransformFromBitReverse64 with
 * output_mod_factor == 1); same op order as device/inv_ntt.cpp:149-437.  The
 * FPGA API receives inv_n / inv_n_w from the caller (host/inc/hexl-fpga.h:
 * 139-156); their Shoup factors are derived here exactly as
 * MultiplyUIntModLazy<64>(x, y, modulus) does (ntt.hpp:105-126). */
void ho_inv_ntt(uint64_t* a, uint64_t n, uint64_t q, const uint64_t* inv_roots,
                const uint64_t* precon_inv, uint64_t inv_n, uint64_t inv_n_w) {
    uint64_t twice_q = q << 1;
    uint64_t t = 1;
    uint64_t r = 1;
    for (uint64_t m = n >> 1; m > 1; m >>= 1) {
        uint64_t j1 = 0;
        for (uint64_t i = 0; i < m; ++i, ++r) {
            uint64_t W = inv_roots[r], Wp = precon_inv[r];
            uint64_t* X = a + j1;
            uint64_t* Y = X + t;
            for (uint64_t j = 0; j < t; ++j) {
                uint64_t tx = X[j] + Y[j];
                uint64_t ty = X[j] + twice_q - Y[j];
                X[j] = (tx >= twice_q) ? tx - twice_q : tx;
                Y[j] = mul_lazy(ty, W, Wp, q);
            }
            j1 += t << 1;
        }
        t <<= 1;
    }
    uint64_t inv_n_p = ho_mult_factor64(inv_n, q);
    uint64_t inv_n_w_p = ho_mult_factor64(inv_n_w, q);
    uint64_t* X = a;
    uint64_t* Y = a + (n >> 1);
    for (uint64_t j = 0; j < (n >> 1); ++j) {
        uint64_t tx = X[j] + Y[j];
        if (tx >= twice_q) tx -= twice_q;
        uint64_t ty = X[j] + twice_q - Y[j];
        X[j] = mul_lazy(tx, inv_n, inv_n_p, q);
        Y[j] = mul_lazy(ty, inv_n_w, inv_n_w_p, q);
    }
    for (uint64_t i = 0; i < n; ++i)
        if (a[i] >= q) a[i] -= q;
}

/* tests/test_utils/ntt.cpp:580-659 with output_mod_factor = 2: the final loop (:648-657) is skipped and
 * the words stay in [0, 2q). */
void ho_inv_ntt_lazy(uint64_t* a, uint64_t n, uint64_t q, const uint64_t* inv_roots,
                     const uint64_t* precon_inv, uint64_t inv_n, uint64_t inv_n_w) {
    uint64_t twice_q = q << 1;
    uint64_t t = 1;
    uint64_t r = 1;
    for (uint64_t m = n >> 1; m > 1; m >>= 1) {
        uint64_t j1 = 0;
        for (uint64_t i = 0; i < m; ++i, ++r) {
            uint64_t W = inv_roots[r], Wp = precon_inv[r];
            uint64_t* X = a + j1;
            uint64_t* Y = X + t;
            for (uint64_t j = 0; j < t; ++j) {
                uint64_t tx = X[j] + Y[j];
                uint64_t ty = X[j] + twice_q - Y[j];
                X[j] = (tx >= twice_q) ? tx - twice_q : tx;
                Y[j] = mul_lazy(ty, W, Wp, q);
            }
            j1 += t << 1;
        }
        t <<= 1;
    }
    uint64_t inv_n_p = ho_mult_factor64(inv_n, q);
    uint64_t inv_n_w_p = ho_mult_factor64(inv_n_w, q);
    uint64_t* X = a;
    uint64_t* Y = a + (n >> 1);
    for (uint64_t j = 0; j < (n >> 1); ++j) {
        uint64_t tx = X[j] + Y[j];
        if (tx >= twice_q) tx -= twice_q;
        uint64_t ty = X[j] + twice_q - Y[j];
        X[j] = mul_lazy(tx, inv_n, inv_n_p, q);
        Y[j] = mul_lazy(ty, inv_n_w, inv_n_w_p, q);
    }
}

void ho_fwd_ntt_batch(uint64_t* a, uint64_t batch, uint64_t n, uint64_t q,
                      const uint64_t* roots, const uint64_t* precon,
                      int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t b = 0; b < (int64_t)batch; ++b)
        ho_fwd_ntt(a + (uint64_t)b * n, n, q, roots, precon);
}

void ho_inv_ntt_batch(uint64_t* a, uint64_t batch, uint64_t n, uint64_t q,
                      const uint64_t* inv_roots, const uint64_t* precon_inv,
                      uint64_t inv_n, uint64_t inv_n_w, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t b = 0; b < (int64_t)batch; ++b)
        ho_inv_ntt(a + (uint64_t)b * n, n, q, inv_roots, precon_inv, inv_n,
                   inv_n_w);
}

/* ------------------------------------------------------------------------ */
/* dyadic multiply                                                           */
/* ------------------------------------------------------------------------ */

/* device/dyadic_multiply.cpp:204-226; expected values
 * tests/test_dyadic_multiply.cpp:54-84 (exact products, any modulus >= 1,
 * inputs not necessarily reduced). */
void ho_dyadic_multiply(uint64_t* res, const uint64_t* op1,
                        const uint64_t* op2, uint64_t n,
                        const uint64_t* moduli, uint64_t n_moduli) {
    uint64_t M = n_moduli;
    for (uint64_t m = 0; m < M; ++m) {
        uint64_t q = moduli[m];
        const uint64_t* x0 = op1 + m * n;
        const uint64_t* x1 = op1 + (M + m) * n;
        const uint64_t* y0 = op2 + m * n;
        const uint64_t* y1 = op2 + (M + m) * n;
        uint64_t* r0 = res + m * n;
        uint64_t* r1 = res + (M + m) * n;
        uint64_t* r2 = res + (2 * M + m) * n;
        for (uint64_t i = 0; i < n; ++i) {
            u128 a = (u128)x0[i] * y1[i];
            u128 b = (u128)x1[i] * y0[i];
            r0[i] = (uint64_t)(((u128)x0[i] * y0[i]) % q);
            r1[i] = (uint64_t)(((a % q) + (b % q)) % q);
            r2[i] = (uint64_t)(((u128)x1[i] * y1[i]) % q);
        }
    }
}

void ho_dyadic_multiply_batch(uint64_t* res, const uint64_t* op1,
                              const uint64_t* op2, uint64_t n,
                              const uint64_t* moduli, uint64_t n_moduli,
                              uint64_t batch, int moduli_per_item,
                              int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t b = 0; b < (int64_t)batch; ++b)
        ho_dyadic_multiply(res + (uint64_t)b * 3 * n_moduli * n,
                           op1 + (uint64_t)b * 2 * n_moduli * n,
                           op2 + (uint64_t)b * 2 * n_moduli * n, n,
                           moduli + (moduli_per_item ? (uint64_t)b * n_moduli : 0),
                           n_moduli);
}

/* ------------------------------------------------------------------------ */
/* keyswitch                                                                 */
/* ------------------------------------------------------------------------ */

typedef struct {
    uint64_t q, w, inv_n, inv_n_w;
    uint64_t *roots, *precon, *inv_roots, *precon_inv;
} ks_tab;

static void ks_tab_init(ks_tab* t, uint64_t n, uint64_t q) {
    t->q = q;
    t->w = ho_min_primitive_root(2 * n, q); /* host/src/fpga.cpp:1098-1101 */
    t->roots = (uint64_t*)malloc(4 * n * sizeof(uint64_t));
    t->precon = t->roots + n;
    t->inv_roots = t->roots + 2 * n;
    t->precon_inv = t->roots + 3 * n;
    ho_compute_roots(n, q, t->w, t->roots, t->precon, t->inv_roots,
                     t->precon_inv);
    t->inv_n = ho_inv_mod(n % q, q);
    t->inv_n_w = ho_mul_mod(t->inv_n, t->inv_roots[n - 1], q);
}

static void ks_tab_free(ks_tab* t) { free(t->roots); }

/* Primary restatement: the FPGA pipeline stage by stage, every stage fully
 * reduced (device/keyswitch/load.hpp:48-128 -> intt_core.hpp:72-93,332-348 ->
 * intt1_redu.hpp:36-38 -> ntt_core.hpp:285-293 -> dyadmult.hpp:128-158 ->
 * intt2_redu.hpp:24-51 -> ntt2.hpp -> ms.hpp:68-83 -> host accumulate
 * host/src/fpga.cpp:441-475). */
static int ks_impl(uint64_t* result, const uint64_t* t_target, uint64_t n,
                   uint64_t D, uint64_t K, uint64_t R, uint64_t C,
                   const uint64_t* moduli, const uint64_t* const* keys,
                   const uint64_t* msf, const ks_tab* tab) {
    if (C != 2 || R != D + 1 || D + 1 > K) return -1;
    uint64_t qk = moduli[K - 1];
    uint64_t* u = (uint64_t*)malloc(D * n * sizeof(uint64_t));   /* coeff-form digits */
    uint64_t* acc = (uint64_t*)calloc(2 * R * n, sizeof(uint64_t)); /* [c][r][n] */
    uint64_t* tmp = (uint64_t*)malloc(n * sizeof(uint64_t));

    /* INTT1: u_j = INTT_{q_j}(t_j), canonical */
    for (uint64_t j = 0; j < D; ++j) {
        memcpy(u + j * n, t_target + j * n, n * sizeof(uint64_t));
        ho_inv_ntt(u + j * n, n, tab[j].q, tab[j].inv_roots, tab[j].precon_inv,
                   tab[j].inv_n, tab[j].inv_n_w);
    }
    /* base conversion + NTT1 + key multiply-accumulate */
    for (uint64_t r = 0; r < R; ++r) {
        uint64_t idx = (r == D) ? K - 1 : r;
        uint64_t qi = moduli[idx];
        for (uint64_t j = 0; j < D; ++j) {
            for (uint64_t l = 0; l < n; ++l) tmp[l] = u[j * n + l] % qi;
            ho_fwd_ntt(tmp, n, qi, tab[idx].roots, tab[idx].precon);
            for (uint64_t c = 0; c < 2; ++c) {
                const uint64_t* key = keys[j] + (c * K + idx) * n;
                uint64_t* a = acc + (c * R + r) * n;
                for (uint64_t l = 0; l < n; ++l)
                    a[l] = ho_add_mod(a[l], ho_mul_mod(tmp[l], key[l] % qi, qi), qi);
            }
        }
    }
    /* special-prime branch: INTT2, round, base-convert, NTT2, modswitch */
    uint64_t qk_half = qk >> 1;
    for (uint64_t c = 0; c < 2; ++c) {
        uint64_t* v = acc + (c * R + D) * n;
        ho_inv_ntt(v, n, qk, tab[K - 1].inv_roots, tab[K - 1].precon_inv,
                   tab[K - 1].inv_n, tab[K - 1].inv_n_w);
        for (uint64_t l = 0; l < n; ++l) v[l] = ho_add_mod(v[l], qk_half, qk);
        for (uint64_t i = 0; i < D; ++i) {
            uint64_t qi = moduli[i];
            uint64_t fix = qi - (qk_half % qi);
            uint64_t f = msf[i] % qi; /* host/src/fpga.cpp:1057-1061 */
            for (uint64_t l = 0; l < n; ++l)
                tmp[l] = (uint64_t)(((u128)v[l] + fix) % qi);
            ho_fwd_ntt(tmp, n, qi, tab[i].roots, tab[i].precon);
            const uint64_t* a = acc + (c * R + i) * n;
            uint64_t* res = result + (c * D + i) * n;
            for (uint64_t l = 0; l < n; ++l) {
                uint64_t d = ho_sub_mod(a[l], tmp[l], qi);
                uint64_t out = ho_mul_mod(d, f, qi);
                res[l] = ho_add_mod(res[l], out, qi);
            }
        }
    }
    free(tmp);
    free(acc);
    free(u);
    return 0;
}

/* Alternative restatement in "intel-hexl order" (the CPU branch the reference
 * dispatches to, host/src/fpga_int.cpp:473-477; intel-hexl v1.2.4 is not
 * vendored, so this follows its published algorithm): the digit whose modulus
 * equals the output modulus is taken straight from the NTT-form input, no
 * reduction when q_j <= q_i, products summed in 128 bits and reduced once,
 * lazy 4q difference before the modswitch multiply. */
static int ks_impl_alt(uint64_t* result, const uint64_t* t_target, uint64_t n,
                       uint64_t D, uint64_t K, uint64_t R, uint64_t C,
                       const uint64_t* moduli, const uint64_t* const* keys,
                       const uint64_t* msf, const ks_tab* tab) {
    if (C != 2 || R != D + 1 || D + 1 > K) return -1;
    uint64_t qk = moduli[K - 1];
    uint64_t* u = (uint64_t*)malloc(D * n * sizeof(uint64_t));
    uint64_t* prod = (uint64_t*)malloc(2 * R * n * sizeof(uint64_t));
    u128* lazy = (u128*)malloc(2 * n * sizeof(u128));
    uint64_t* tntt = (uint64_t*)malloc(n * sizeof(uint64_t));
    memcpy(u, t_target, D * n * sizeof(uint64_t));
    for (uint64_t j = 0; j < D; ++j)
        ho_inv_ntt(u + j * n, n, tab[j].q, tab[j].inv_roots, tab[j].precon_inv,
                   tab[j].inv_n, tab[j].inv_n_w);
    for (uint64_t r = 0; r < R; ++r) {
        uint64_t idx = (r == D) ? K - 1 : r;
        uint64_t qi = moduli[idx];
        memset(lazy, 0, 2 * n * sizeof(u128));
        for (uint64_t j = 0; j < D; ++j) {
            const uint64_t* operand;
            if (r == j) {
                operand = t_target + j * n;
            } else {
                if (moduli[j] <= qi) {
                    memcpy(tntt, u + j * n, n * sizeof(uint64_t));
                } else {
                    for (uint64_t l = 0; l < n; ++l) tntt[l] = u[j * n + l] % qi;
                }
                ho_fwd_ntt_reference(tntt, n, qi, tab[idx].roots);
                operand = tntt;
            }
            for (uint64_t c = 0; c < 2; ++c) {
                const uint64_t* key = keys[j] + (c * K + idx) * n;
                for (uint64_t l = 0; l < n; ++l) {
                    /* keep the 128-bit sum from overflowing for 62-bit q */
                    lazy[c * n + l] = (lazy[c * n + l] + (u128)operand[l] * key[l]) % qi;
                }
            }
        }
        for (uint64_t c = 0; c < 2; ++c)
            for (uint64_t l = 0; l < n; ++l)
                prod[(c * R + r) * n + l] = (uint64_t)lazy[c * n + l];
    }
    uint64_t qk_half = qk >> 1;
    for (uint64_t c = 0; c < 2; ++c) {
        uint64_t* last = prod + (c * R + D) * n;
        ho_inv_ntt(last, n, qk, tab[K - 1].inv_roots, tab[K - 1].precon_inv,
                   tab[K - 1].inv_n, tab[K - 1].inv_n_w);
        for (uint64_t l = 0; l < n; ++l) last[l] = (last[l] + qk_half) % qk;
        for (uint64_t i = 0; i < D; ++i) {
            uint64_t qi = moduli[i];
            uint64_t fix = qi - (qk_half % qi);
            for (uint64_t l = 0; l < n; ++l)
                tntt[l] = (qk > qi ? last[l] % qi : last[l]) + fix; /* [0,2qi) */
            for (uint64_t l = 0; l < n; ++l) tntt[l] %= qi;
            ho_fwd_ntt_reference(tntt, n, qi, tab[i].roots);
            const uint64_t* a = prod + (c * R + i) * n;
            uint64_t* res = result + (c * D + i) * n;
            uint64_t f = msf[i];
            for (uint64_t l = 0; l < n; ++l) {
                u128 diff = (u128)a[l] + ((u128)qi << 2) - tntt[l];
                uint64_t out = (uint64_t)(((diff % qi) * (f % qi)) % qi);
                res[l] = (uint64_t)(((u128)res[l] + out) % qi);
            }
        }
    }
    free(tntt);
    free(lazy);
    free(prod);
    free(u);
    return 0;
}

typedef int (*ks_fn)(uint64_t*, const uint64_t*, uint64_t, uint64_t, uint64_t,
                     uint64_t, uint64_t, const uint64_t*,
                     const uint64_t* const*, const uint64_t*, const ks_tab*);

static int ks_run(ks_fn fn, uint64_t* result, const uint64_t* t_target,
                  uint64_t batch, uint64_t n, uint64_t D, uint64_t K,
                  uint64_t R, uint64_t C, const uint64_t* moduli,
                  const uint64_t* const* keys, const uint64_t* msf,
                  int threads) {
    if (K == 0 || K > 64) return -1;
    ks_tab tab[64];
    for (uint64_t i = 0; i < K; ++i) ks_tab_init(&tab[i], n, moduli[i]);
    int rc = 0;
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t b = 0; b < (int64_t)batch; ++b) {
        int r = fn(result + (uint64_t)b * 2 * D * n, t_target + (uint64_t)b * D * n,
                   n, D, K, R, C, moduli, keys, msf, tab);
        if (r) {
#pragma omp atomic write
            rc = r;
        }
    }
    for (uint64_t i = 0; i < K; ++i) ks_tab_free(&tab[i]);
    return rc;
}

int ho_keyswitch(uint64_t* result, const uint64_t* t_target, uint64_t n,
                 uint64_t D, uint64_t K, uint64_t R, uint64_t C,
                 const uint64_t* moduli, const uint64_t* const* keys,
                 const uint64_t* msf) {
    return ks_run(ks_impl, result, t_target, 1, n, D, K, R, C, moduli, keys, msf, 1);
}

int ho_keyswitch_alt(uint64_t* result, const uint64_t* t_target, uint64_t n,
                     uint64_t D, uint64_t K, uint64_t R, uint64_t C,
                     const uint64_t* moduli, const uint64_t* const* keys,
                     const uint64_t* msf) {
    return ks_run(ks_impl_alt, result, t_target, 1, n, D, K, R, C, moduli, keys, msf, 1);
}

int ho_keyswitch_batch(uint64_t* result, const uint64_t* t_target,
                       uint64_t batch, uint64_t n, uint64_t D, uint64_t K,
                       uint64_t R, uint64_t C, const uint64_t* moduli,
                       const uint64_t* const* keys, const uint64_t* msf,
                       int threads) {
    return ks_run(ks_impl, result, t_target, batch, n, D, K, R, C, moduli, keys,
                  msf, threads);
}

int ho_keyswitch_alt_batch(uint64_t* result, const uint64_t* t_target,
                           uint64_t batch, uint64_t n, uint64_t D, uint64_t K,
                           uint64_t R, uint64_t C, const uint64_t* moduli,
                           const uint64_t* const* keys, const uint64_t* msf,
                           int threads) {
    return ks_run(ks_impl_alt, result, t_target, batch, n, D, K, R, C, moduli,
                  keys, msf, threads);
}

/* ------------------------------------------------------------------------ */
/* helpers                                                                   */
/* ------------------------------------------------------------------------ */

uint64_t ho_fnv1a(const uint64_t* v, size_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; ++i) {
        uint64_t x = v[i];
        for (int b = 0; b < 8; ++b) {
            h ^= (x >> (8 * b)) & 0xff;
            h *= 1099511628211ULL;
        }
    }
    return h;
}

uint64_t ho_splitmix_fill(uint64_t* out, size_t n, uint64_t s, uint64_t q) {
    for (size_t i = 0; i < n; ++i) {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        out[i] = q ? z % q : z;
    }
    return s;
}

int ho_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
