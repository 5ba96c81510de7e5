// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" window onto the reference's OWN scalar NTT implementation
// (/root/reference/tests/test_utils/ntt.{hpp,cpp}, namespace hetest::utils),
// which is compiled unmodified from where it lies by oracle/Makefile into
// oracle/_ref/libhexl_ref.so.  No reference source is copied into this repo:
// this file only calls the reference's public functions.
//
// Used (a) to validate oracle/hexl_oracle.c, (b) to generate tests/golden/,
// (c) as the "reference" CPU baseline of bench.py (cpu_baseline.kind).
#include <cstdint>
#include <cstring>
#include <vector>

#include "test_utils/ntt.hpp"

using namespace hetest::utils;

extern "C" {

// GeneratePrimes (tests/test_utils/ntt.cpp:224-247)
uint64_t ref_generate_primes(uint64_t* out, uint64_t num, uint64_t bits,
                             uint64_t ntt_size) {
    std::vector<uint64_t> p = GeneratePrimes(num, bits, ntt_size);
    for (size_t i = 0; i < p.size(); ++i) out[i] = p[i];
    return p.size();
}

// MinimalPrimitiveRoot (ntt.cpp:137-158)
uint64_t ref_min_primitive_root(uint64_t degree, uint64_t q) {
    return MinimalPrimitiveRoot(degree, q);
}

uint64_t ref_inverse_mod(uint64_t a, uint64_t q) { return InverseUIntMod(a, q); }
uint64_t ref_multiply_mod(uint64_t a, uint64_t b, uint64_t q) {
    return MultiplyUIntMod(a, b, q);
}

// NTTImpl tables (ntt.cpp:290-384): four arrays of n entries.
void ref_tables(uint64_t n, uint64_t q, uint64_t* roots, uint64_t* precon,
                uint64_t* inv_roots, uint64_t* precon_inv) {
    NTT::NTTImpl ntt(n, q);
    std::memcpy(roots, ntt.GetRootOfUnityPowersPtr(), n * 8);
    std::memcpy(precon, ntt.GetPrecon64RootOfUnityPowersPtr(), n * 8);
    std::memcpy(inv_roots, ntt.GetInvRootOfUnityPowersPtr(), n * 8);
    std::memcpy(precon_inv, ntt.GetPrecon64InvRootOfUnityPowersPtr(), n * 8);
}

// NTTImpl::ComputeForward / ComputeInverse with mod factors (1,1), exactly the
// calls tests/test_fwd_ntt.cpp:103-108 and tests/test_inv_ntt.cpp:113-116 make.
void ref_fwd_ntt(uint64_t* a, uint64_t n, uint64_t q) {
    NTT::NTTImpl ntt(n, q);
    ntt.ComputeForward(a, a, 1, 1);
}
void ref_inv_ntt(uint64_t* a, uint64_t n, uint64_t q) {
    NTT::NTTImpl ntt(n, q);
    ntt.ComputeInverse(a, a, 1, 1);
}

// the reference's mod factors (ntt.cpp:442-470): output_mod_factor 4 (forward) / 2 (inverse) leaves the
// lazy words of the butterflies as the result
void ref_fwd_ntt_factors(uint64_t* a, uint64_t n, uint64_t q, uint64_t in_f, uint64_t out_f) {
    NTT::NTTImpl ntt(n, q);
    ntt.ComputeForward(a, a, in_f, out_f);
}
void ref_inv_ntt_factors(uint64_t* a, uint64_t n, uint64_t q, uint64_t in_f, uint64_t out_f) {
    NTT::NTTImpl ntt(n, q);
    ntt.ComputeInverse(a, a, in_f, out_f);
}

// Batch over caller-supplied tables (free functions ntt.cpp:474-548, 580-659);
// OpenMP over items -- the timing loop of the CPU baseline.
void ref_fwd_ntt_batch(uint64_t* a, uint64_t batch, uint64_t n, uint64_t q,
                       const uint64_t* roots, const uint64_t* precon,
                       int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t b = 0; b < (int64_t)batch; ++b)
        ForwardTransformToBitReverse64(a + (uint64_t)b * n, n, q, roots, precon,
                                       1, 1);
}
void ref_inv_ntt_batch(uint64_t* a, uint64_t batch, uint64_t n, uint64_t q,
                       const uint64_t* inv_roots, const uint64_t* precon_inv,
                       int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t b = 0; b < (int64_t)batch; ++b)
        InverseTransformFromBitReverse64(a + (uint64_t)b * n, n, q, inv_roots,
                                         precon_inv, 1, 1);
}

}  // extern "C"
