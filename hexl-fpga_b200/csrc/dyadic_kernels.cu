// dyadic_kernels.cu -- ciphertext x ciphertext dyadic multiply (sm_100a).
// Replaces device/dyadic_multiply.cpp:195-228 (+ MultMod, device/mod_ops.hpp:
// 31-84).  Pure streaming kernel: 4 input words and 3 output words per
// coefficient, 16-byte accesses, one CTA per slab of one (item, modulus).
//
//   res[m]      = x0*y0 mod q_m
//   res[M+m]    = (x0*y1 mod q_m + x1*y0 mod q_m) mod q_m
//   res[2M+m]   = x1*y1 mod q_m
//
// Any modulus >= 1 and unreduced operands are accepted, as the reference's
// tests require (tests/test_dyadic_multiply.cpp:35-84).
#include "launch.h"

namespace hb {

constexpr int kDyThreads = 256;
constexpr int kDyCoeffPerThread = 2;   // one 16-byte access per array
constexpr int kDySlab = 2048;          // coefficients per CTA

// reciprocal / shift of every modulus of the call, once (the 128-by-64 division
// inside make_divisor is a few hundred instructions: far too slow to repeat in
// every CTA of the streaming kernel behind a block barrier)
__global__ void k_dyadic_prep(const uint64_t* __restrict__ moduli, Divisor* __restrict__ divs, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) divs[i] = make_divisor(moduli[i]);
}

__global__ void __launch_bounds__(kDyThreads)
k_dyadic(uint64_t* __restrict__ res, const uint64_t* __restrict__ op1,
         const uint64_t* __restrict__ op2, uint32_t n, const Divisor* __restrict__ divs,
         uint32_t M, int moduli_per_item, uint32_t slabs_per_poly) {
    // blockIdx.x = ((item * M) + m) * slabs_per_poly + slab
    const uint64_t bid = blockIdx.x;
    const uint32_t slab = (uint32_t)(bid % slabs_per_poly);
    const uint64_t im = bid / slabs_per_poly;
    const uint32_t m = (uint32_t)(im % M);
    const uint64_t item = im / M;
    Divisor dv;
    {
        const ulonglong2* dp = reinterpret_cast<const ulonglong2*>(divs + (moduli_per_item ? item * M : 0) + m);
        const ulonglong2 a = __ldg(dp), b = __ldg(dp + 1);
        dv.d = a.x;
        dv.v = a.y;
        dv.s = (uint32_t)b.x;
        dv.pad = 0;
        dv.q = b.y;
    }
    const bool lazy = (dv.q >> 63) == 0;   // CTA-uniform
    const uint32_t qh = (uint32_t)(dv.q >> 32);

    const uint64_t in_item = item * 2ull * M * n;
    const uint64_t out_item = item * 3ull * M * n;
    const uint64_t* x0p = op1 + in_item + (uint64_t)m * n;
    const uint64_t* x1p = op1 + in_item + (uint64_t)(M + m) * n;
    const uint64_t* y0p = op2 + in_item + (uint64_t)m * n;
    const uint64_t* y1p = op2 + in_item + (uint64_t)(M + m) * n;
    uint64_t* r0p = res + out_item + (uint64_t)m * n;
    uint64_t* r1p = res + out_item + (uint64_t)(M + m) * n;
    uint64_t* r2p = res + out_item + (uint64_t)(2 * M + m) * n;

    const uint32_t begin = slab * kDySlab;
    const uint32_t end = min(begin + kDySlab, n);
#pragma unroll 2
    for (uint32_t i = begin + threadIdx.x * kDyCoeffPerThread; i < end;
         i += kDyThreads * kDyCoeffPerThread) {
        uint64_t x0[2], x1[2], y0[2], y1[2], r0[2], r1[2], r2[2];
        if (i + 1 < end) {
            ld2(x0p + i, x0[0], x0[1]);
            ld2(x1p + i, x1[0], x1[1]);
            ld2(y0p + i, y0[0], y0[1]);
            ld2(y1p + i, y1[0], y1[1]);
        } else {  // odd tail
            x0[0] = x0p[i]; x1[0] = x1p[i]; y0[0] = y0p[i]; y1[0] = y1p[i];
            x0[1] = x1[1] = y0[1] = y1[1] = 0;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            uint64_t a0 = x0[k], a1 = x1[k], b0 = y0[k], b1 = y1[k];
            // operands are normally reduced already: one test on the high words
            // (sufficient, not necessary) skips the four 64-bit compares
            const uint32_t mh = max(max((uint32_t)(a0 >> 32), (uint32_t)(a1 >> 32)),
                                    max((uint32_t)(b0 >> 32), (uint32_t)(b1 >> 32)));
            if (mh >= qh) {   // rare (always for q < 2^32): exact per-operand reduction
                a0 = mod64(a0, dv); a1 = mod64(a1, dv); b0 = mod64(b0, dv); b1 = mod64(b1, dv);
            }
            const uint64_t as0 = a0 << dv.s, as1 = a1 << dv.s;
            r0[k] = mulmod_preshifted(as0, b0, dv);
            r2[k] = mulmod_preshifted(as1, b1, dv);
            if (lazy) {
                // q < 2^63: the cross term is reduced once, from the 128-bit sum of its two products
                r1[k] = mul2add_mod_preshifted(as0, b1, as1, b0, dv);
            } else {
                const uint64_t c = mulmod_preshifted(as0, b1, dv);
                const uint64_t d = mulmod_preshifted(as1, b0, dv);
                uint64_t s = c + d;                      // c,d < q <= 2^64-1: detect wrap
                if (s < c || s >= dv.q) s -= dv.q;
                r1[k] = s;
            }
        }
        if (i + 1 < end) {
            st2(r0p + i, r0[0], r0[1]);
            st2(r1p + i, r1[0], r1[1]);
            st2(r2p + i, r2[0], r2[1]);
        } else {
            r0p[i] = r0[0]; r1p[i] = r1[0]; r2p[i] = r2[0];
        }
    }
}

size_t dyadic_scratch_bytes(uint64_t n_moduli, uint64_t batch, int moduli_per_item) {
    return (size_t)(moduli_per_item ? batch * n_moduli : n_moduli) * sizeof(Divisor);
}

cudaError_t launch_dyadic(uint64_t* res, const uint64_t* op1, const uint64_t* op2, uint64_t n,
                          const uint64_t* moduli, uint64_t n_moduli, uint64_t batch,
                          int moduli_per_item, void* scratch, cudaStream_t st) {
    if (batch == 0 || n == 0 || n_moduli == 0) return cudaSuccess;
    static_assert(sizeof(Divisor) == 32 && alignof(Divisor) == 8, "k_dyadic reads a Divisor as two 16-byte words");
    Divisor* divs = static_cast<Divisor*>(scratch);
    const uint64_t n_div = moduli_per_item ? batch * n_moduli : n_moduli;
    if (n_div >> 32) return cudaErrorInvalidValue;
    k_dyadic_prep<<<(unsigned)((n_div + 127) / 128), 128, 0, st>>>(moduli, divs, (uint32_t)n_div);
    // 16-byte accesses need even n for every slab base to stay aligned.
    if (n & 1) return cudaErrorInvalidValue;
    const uint32_t slabs = (uint32_t)((n + kDySlab - 1) / kDySlab);
    const uint64_t per_item = n_moduli * slabs;
    const uint64_t kMaxGrid = 1u << 30;
    const uint64_t items_per_launch = kMaxGrid / per_item ? kMaxGrid / per_item : 1;
    for (uint64_t off = 0; off < batch; off += items_per_launch) {
        const uint64_t cnt = batch - off < items_per_launch ? batch - off : items_per_launch;
        k_dyadic<<<(unsigned)(cnt * per_item), kDyThreads, 0, st>>>(
            res + off * 3 * n_moduli * n, op1 + off * 2 * n_moduli * n,
            op2 + off * 2 * n_moduli * n, (uint32_t)n,
            divs + (moduli_per_item ? off * n_moduli : 0), (uint32_t)n_moduli,
            moduli_per_item, slabs);
    }
    return cudaGetLastError();
}

}  // namespace hb
