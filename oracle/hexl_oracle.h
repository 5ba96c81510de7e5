/*
 * hexl_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A from-scratch plain-C restatement of the arithmetic that intel/hexl-fpga's
 * four primitives (forward NTT, inverse NTT, dyadic multiply, keyswitch)
 * compute.  It exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs can check or time the algorithm on the
 * host.  NOTHING in the product path (hexl-fpga_b200/) may include, link or
 * call it: the product path is CUDA only and fails loudly without a GPU.
 *
 * Parity status:
 *   - fwd/inv NTT, number theory, twiddle tables: PINNED against the
 *     known-answer vectors of SURVEY.md Appendix B and against the
 *     reference's own scalar NTT (tests/test_utils/ntt.cpp compiled
 *     unmodified into oracle/_ref/libhexl_ref.so; fixtures committed under
 *     tests/golden/).
 *   - dyadic multiply: PINNED against the closed-form expectation of the
 *     reference's tests/test_dyadic_multiply.cpp:35-84.
 *   - keyswitch: PARITY UNPINNED by any artefact inside /root/reference (its
 *     only golden vectors live in an external testdata.zip that is not
 *     available offline).  Anchored instead on the reference's device
 *     dataflow (device/keyswitch/ *.hpp), two independently written
 *     restatements, and an RLWE decryption-noise self test.
 *
 * Reference citations are relative to /root/reference.
 */
#ifndef HEXL_ORACLE_H_
#define HEXL_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- number theory (tests/test_utils/ntt.cpp:16-247,
 *                     host/src/number_theory_util.cpp:12-154) ---- */
uint64_t ho_mul_mod(uint64_t a, uint64_t b, uint64_t q);
uint64_t ho_add_mod(uint64_t a, uint64_t b, uint64_t q);
uint64_t ho_sub_mod(uint64_t a, uint64_t b, uint64_t q);
uint64_t ho_pow_mod(uint64_t base, uint64_t exp, uint64_t q);
uint64_t ho_inv_mod(uint64_t a, uint64_t q);
uint64_t ho_reverse_bits(uint64_t x, uint64_t bit_width);
int ho_is_prime(uint64_t n);
/* primes in (2^bit_size, 2^(bit_size+1)) with p = 1 mod 2*ntt_size; returns
 * the number found (<= num_primes). */
size_t ho_generate_primes(uint64_t* out, size_t num_primes, size_t bit_size,
                          size_t ntt_size);
int ho_is_primitive_root(uint64_t root, uint64_t degree, uint64_t q);
/* smallest primitive degree-th root of unity mod q (degree a power of 2). */
uint64_t ho_min_primitive_root(uint64_t degree, uint64_t q);
/* floor(operand * 2^64 / q): the 64-bit Shoup/"Barrett" factor
 * (tests/test_utils/ntt.hpp:17-37). */
uint64_t ho_mult_factor64(uint64_t operand, uint64_t q);

/* ---- twiddle tables ---- */
/* hexl layout (tests/test_utils/ntt.cpp:290-384): roots in bit-reversed
 * order, inv_roots in 1-based stage order; each array has n entries. */
void ho_compute_roots(uint64_t n, uint64_t q, uint64_t w, uint64_t* roots,
                      uint64_t* precon, uint64_t* inv_roots,
                      uint64_t* precon_inv);
/* keyswitch 4-table layout per modulus, host/src/twiddle-factors.cpp:16-62 and
 * host/src/fpga.cpp:1102-1109: [inv_roots | precon_inv | roots | precon], the
 * inverse table 0-based with [n-1] = 0 and precon_roots[0] = 0. */
void ho_compute_roots_keyswitch(uint64_t n, uint64_t q, uint64_t w,
                                uint64_t* table4n);

/* ---- forward / inverse negacyclic NTT, exact op sequence ----
 * tests/test_utils/ntt.cpp:474-548 and :580-659 (== device/fwd_ntt.cpp:282-386
 * and device/inv_ntt.cpp:149-437).  In place, wrap-around uint64 arithmetic,
 * so out-of-range inputs give the same (garbage) words as the reference. */
void ho_fwd_ntt(uint64_t* a, uint64_t n, uint64_t q, const uint64_t* roots,
                const uint64_t* precon);
void ho_inv_ntt(uint64_t* a, uint64_t n, uint64_t q, const uint64_t* inv_roots,
                const uint64_t* precon_inv, uint64_t inv_n, uint64_t inv_n_w);
/* textbook fully reduced forward transform (ntt.cpp:550-578) used as an
 * independent cross-check of ho_fwd_ntt for in-range inputs. */
void ho_fwd_ntt_reference(uint64_t* a, uint64_t n, uint64_t q,
                          const uint64_t* roots);
/* batch versions, OpenMP over items (for the CPU baseline timing). */
void ho_fwd_ntt_batch(uint64_t* a, uint64_t batch, uint64_t n, uint64_t q,
                      const uint64_t* roots, const uint64_t* precon,
                      int threads);
void ho_inv_ntt_batch(uint64_t* a, uint64_t batch, uint64_t n, uint64_t q,
                      const uint64_t* inv_roots, const uint64_t* precon_inv,
                      uint64_t inv_n, uint64_t inv_n_w, int threads);

/* ---- dyadic multiply (device/dyadic_multiply.cpp:204-226, expected values
 * tests/test_dyadic_multiply.cpp:54-84) ---- one item:
 * op1/op2 = [2][n_moduli][n], res = [3][n_moduli][n]. */
void ho_dyadic_multiply(uint64_t* res, const uint64_t* op1,
                        const uint64_t* op2, uint64_t n,
                        const uint64_t* moduli, uint64_t n_moduli);
void ho_dyadic_multiply_batch(uint64_t* res, const uint64_t* op1,
                              const uint64_t* op2, uint64_t n,
                              const uint64_t* moduli, uint64_t n_moduli,
                              uint64_t batch, int moduli_per_item,
                              int threads);

/* ---- keyswitch (device/keyswitch/ *.hpp dataflow,
 * host accumulate host/src/fpga.cpp:441-475) ---- one item, accumulates into
 * result[2][decomp][n].  Twiddles are derived from moduli with the minimal
 * primitive root (host/src/fpga.cpp:1098-1109).  Returns 0 on success. */
int ho_keyswitch(uint64_t* result, const uint64_t* t_target, uint64_t n,
                 uint64_t decomp_modulus_size, uint64_t key_modulus_size,
                 uint64_t rns_modulus_size, uint64_t key_component_count,
                 const uint64_t* moduli, const uint64_t* const* k_switch_keys,
                 const uint64_t* modswitch_factors);
/* second, independently structured restatement ("hexl order": digit reuse when
 * i == j, 128-bit lazy accumulation, lazy 4q transforms); must agree bit for
 * bit with ho_keyswitch. */
int ho_keyswitch_alt(uint64_t* result, const uint64_t* t_target, uint64_t n,
                     uint64_t decomp_modulus_size, uint64_t key_modulus_size,
                     uint64_t rns_modulus_size, uint64_t key_component_count,
                     const uint64_t* moduli,
                     const uint64_t* const* k_switch_keys,
                     const uint64_t* modswitch_factors);
/* batch: items contiguous (result stride 2*decomp*n, t_target stride
 * decomp*n), one shared key set. */
int ho_keyswitch_batch(uint64_t* result, const uint64_t* t_target,
                       uint64_t batch, uint64_t n,
                       uint64_t decomp_modulus_size, uint64_t key_modulus_size,
                       uint64_t rns_modulus_size,
                       uint64_t key_component_count, const uint64_t* moduli,
                       const uint64_t* const* k_switch_keys,
                       const uint64_t* modswitch_factors, int threads);

/* the same over the second restatement (the cheaper of the two: digit reuse, lazy
 * transforms) -- bench.py's CPU keyswitch baseline */
int ho_keyswitch_alt_batch(uint64_t* result, const uint64_t* t_target,
                           uint64_t batch, uint64_t n,
                           uint64_t decomp_modulus_size, uint64_t key_modulus_size,
                           uint64_t rns_modulus_size,
                           uint64_t key_component_count, const uint64_t* moduli,
                           const uint64_t* const* k_switch_keys,
                           const uint64_t* modswitch_factors, int threads);

/* forward / inverse transform with the reference's output_mod_factor = 4 / 2: no final correction,
 * the lazy words of the Harvey butterflies are the result (tests/test_utils/ntt.cpp:535-546, 648-657) */
void ho_fwd_ntt_lazy(uint64_t* a, uint64_t n, uint64_t q, const uint64_t* roots, const uint64_t* precon);
void ho_inv_ntt_lazy(uint64_t* a, uint64_t n, uint64_t q, const uint64_t* inv_roots, const uint64_t* precon_inv,
                     uint64_t inv_n, uint64_t inv_n_w);

/* ---- helpers shared by tests ---- */
/* 64-bit FNV-1a over the little-endian bytes of v[0..n). */
uint64_t ho_fnv1a(const uint64_t* v, size_t n);
/* splitmix64 stream: fills out[0..n) with successive outputs, optionally
 * reduced mod q (q == 0: raw).  Returns the updated state. */
uint64_t ho_splitmix_fill(uint64_t* out, size_t n, uint64_t state, uint64_t q);
int ho_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* HEXL_ORACLE_H_ */
