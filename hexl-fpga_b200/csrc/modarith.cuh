// modarith.cuh -- 64-bit modular arithmetic building blocks (host+device).
//
// Everything here is plain uint64 wrap-around arithmetic so that the exact
// Harvey butterflies reproduce the reference's words even for out-of-range
// inputs (reference: tests/test_utils/ntt.cpp:494-547, 618-657 and
// device/fwd_ntt.cpp:282-386, device/inv_ntt.cpp:149-437).
//
// The functions are __host__ __device__ so that tests/cpu_emul can run the very
// same index/arith code thread-by-thread on the CPU (test infrastructure, never
// a product fallback: the product entry points in capi.cu only launch kernels).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#define HB_D __device__ __forceinline__
#else
#define HB_HD inline
#define HB_D inline
#endif

namespace hb {

HB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// Shoup/Harvey lazy product: w*x - floor(x*wp/2^64)*q  (mod 2^64), in [0,2q)
// when w < q and wp = floor(w*2^64/q).   tests/test_utils/ntt.hpp:87-101
HB_HD uint64_t mul_lazy(uint64_t x, uint64_t w, uint64_t wp, uint64_t q) {
    return w * x - mulhi64(x, wp) * q;
}

// Forward (Cooley-Tukey) Harvey butterfly, exact op order of
// tests/test_utils/ntt.cpp:519-546 / device/fwd_ntt.cpp:282-386.
HB_HD void fwd_bfly(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wp,
                    uint64_t q, uint64_t twoq) {
    uint64_t tx = X - ((X >= twoq) ? twoq : 0);
    uint64_t T = mul_lazy(Y, w, wp, q);
    X = tx + T;
    Y = tx + twoq - T;
}

// Inverse (Gentleman-Sande) Harvey butterfly, tests/test_utils/ntt.cpp:618-634
// / device/inv_ntt.cpp:149-398.
HB_HD void inv_bfly(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wp,
                    uint64_t q, uint64_t twoq) {
    uint64_t tx = X + Y;
    uint64_t ty = X + twoq - Y;
    X = tx - ((tx >= twoq) ? twoq : 0);
    Y = mul_lazy(ty, w, wp, q);
}

// Last inverse stage fused with the n^-1 scaling, ntt.cpp:640-657 /
// device/inv_ntt.cpp:400-437.  Outputs are fully reduced to [0,q).
HB_HD void inv_last_bfly(uint64_t& X, uint64_t& Y, uint64_t inv_n,
                         uint64_t inv_n_p, uint64_t inv_n_w, uint64_t inv_n_w_p,
                         uint64_t q, uint64_t twoq) {
    uint64_t tx = X + Y;
    tx -= (tx >= twoq) ? twoq : 0;
    uint64_t ty = X + twoq - Y;
    uint64_t x = mul_lazy(tx, inv_n, inv_n_p, q);
    uint64_t y = mul_lazy(ty, inv_n_w, inv_n_w_p, q);
    X = x - ((x >= q) ? q : 0);
    Y = y - ((y >= q) ? q : 0);
}

// x mod q for any x < 2^64 with mu = floor(2^64/q)   (q < 2^63).
// device/keyswitch/intt1_redu.hpp:36-38 computes the same canonical value.
HB_HD uint64_t barrett_reduce64(uint64_t x, uint64_t q, uint64_t mu) {
    uint64_t r = x - mulhi64(x, mu) * q;  // in [0, 2q)
    return r - ((r >= q) ? q : 0);
}

HB_HD uint64_t add_mod(uint64_t a, uint64_t b, uint64_t q) {  // a,b < q < 2^63
    uint64_t s = a + b;
    return s - ((s >= q) ? q : 0);
}
HB_HD uint64_t sub_mod(uint64_t a, uint64_t b, uint64_t q) {  // a,b < q
    return (a >= b) ? a - b : a + q - b;
}

// ---- generic 128-by-64 remainder (any modulus >= 1), Moeller-Granlund
// "division by invariant integers" 2-by-1 step with a precomputed reciprocal.
struct Divisor {
    uint64_t d;   // modulus shifted left until its top bit is set
    uint64_t v;   // floor((2^128-1)/d) - 2^64
    uint32_t s;   // shift
    uint32_t pad;
    uint64_t q;   // the modulus itself
};

HB_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}

HB_HD Divisor make_divisor(uint64_t q) {
    Divisor dv;
    dv.q = q;
    dv.pad = 0;
    if (q == 0) {  // undefined in the reference; keep the arithmetic total
        dv.d = (uint64_t)1 << 63;
        dv.s = 63;
        dv.v = ~(uint64_t)0;
        return dv;
    }
    dv.s = (uint32_t)clz64(q);
    dv.d = q << dv.s;
    unsigned __int128 all = ~(unsigned __int128)0;
    dv.v = (uint64_t)(all / dv.d);  // low 64 bits == floor((2^128-1)/d) - 2^64
    return dv;
}

// remainder of (u1:u0) by d, d normalised, u1 < d.
HB_HD uint64_t rem_2by1(uint64_t u1, uint64_t u0, uint64_t d, uint64_t v) {
    // (q1:q0) = v*u1 + (u1:u0)
    uint64_t q0 = v * u1;
    uint64_t q1 = mulhi64(v, u1);
    uint64_t t0 = q0 + u0;
    q1 += u1 + ((t0 < q0) ? 1 : 0);
    q0 = t0;
    q1 += 1;
    uint64_t r = u0 - q1 * d;
    if (r > q0) r += d;
    if (r >= d) r -= d;
    return r;
}

// x mod q for one 64-bit word.
HB_HD uint64_t mod64(uint64_t x, const Divisor& dv) {
    if (x < dv.q) return x;
    uint64_t u1 = dv.s ? (x >> (64 - dv.s)) : 0;
    uint64_t u0 = x << dv.s;
    return rem_2by1(u1, u0, dv.d, dv.v) >> dv.s;
}

// (a*b) mod q with a,b < q (product's high word is then < q).
HB_HD uint64_t mulmod_reduced(uint64_t a, uint64_t b, const Divisor& dv) {
    uint64_t lo = a * b;
    uint64_t hi = mulhi64(a, b);
    uint64_t u1 = dv.s ? ((hi << dv.s) | (lo >> (64 - dv.s))) : hi;
    uint64_t u0 = lo << dv.s;
    return rem_2by1(u1, u0, dv.d, dv.v) >> dv.s;
}

}  // namespace hb
