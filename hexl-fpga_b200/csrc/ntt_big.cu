// ntt_big.cu -- N = 32768 transforms ("other N"): a polynomial of 256 KiB no longer fits
// one SM's shared memory, so the transform is cut along its first (forward) / last (inverse) radix-2 stage:
//
//   forward   stage 0 pairs a[j], a[j + N/2] with the single twiddle roots[1]: one streaming kernel; after it
//             the two halves are independent N/2-point transforms whose stage-s twiddles are the big
//             table's entries roots[2m' + h m' + i'] (m' = 2^s blocks, half h) -- they run on the
//             shared-memory kernels of ntt_block.cuh, two items per polynomial;
//   inverse   the mirror image: two N/2-point inverse transforms (scaling 1), then one streaming kernel for
//             the last stage with n^-1 and n^-1 w (tests/test_utils/ntt.cpp:636-657).
// Traffic is twice the one-pass minimum (each polynomial crosses HBM twice); the reference's own NTT entry
// points stop at 16384 (host/src/ntt.cpp:24), so this size exists for callers of the dyadic shapes at 32768
// (SURVEY Appendix D).  In-contract inputs only: the reference op sequence on out-of-range words is not
// reproduced across the cut.
#include "ntt_launch.cuh"

namespace hb {

// sub-tables of half h = blockIdx.y in hexl layout (index m' + i'), N' = N / 2 entries each
__global__ void k_big_split_fwd(const uint64_t* __restrict__ roots, const uint64_t* __restrict__ precon,
                                uint64_t* __restrict__ sub, uint32_t n_half) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y;
    if (e >= n_half) return;
    uint64_t* r = sub + (size_t)h * 2 * n_half;
    uint64_t* p = r + n_half;
    if (e == 0) {
        r[0] = 1;
        p[0] = 0;
        return;
    }
    const uint32_t mp = 1u << (31 - __clz(e)), ip = e - mp;     // e = m' + i'
    r[e] = roots[2 * mp + h * mp + ip];
    p[e] = precon[2 * mp + h * mp + ip];
}
// inverse tables, 1-based stage order: sub[1 + N' - 2m' + i'] = big[1 + N - 4m' + h m' + i'], m' = N'/2 .. 2
__global__ void k_big_split_inv(const uint64_t* __restrict__ inv_roots, const uint64_t* __restrict__ precon_inv,
                                uint64_t* __restrict__ sub, uint32_t n_half) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y;
    if (e >= n_half) return;
    uint64_t* r = sub + (size_t)h * 2 * n_half;
    uint64_t* p = r + n_half;
    if (e == 0 || e == n_half - 1) {       // [0] = 1 by convention; [N'-1] (stage m' = 1) is applied by the last-stage kernel
        r[e] = 1;
        p[e] = 0;
        return;
    }
    // e - 1 = (N' - 2m') + i'  with 0 <= i' < m'  <=>  N' - e + 1 in (m', 2m']
    const uint32_t d = n_half - e;                 // = 2m' - 1 - i', in [m', 2m' - 1]
    const uint32_t mp = 1u << (31 - __clz(d));     // the stage's block count m'
    const uint32_t ip = e - 1 - (n_half - 2 * mp);
    const uint32_t n = 2 * n_half;
    r[e] = inv_roots[1 + n - 4 * mp + h * mp + ip];
    p[e] = precon_inv[1 + n - 4 * mp + h * mp + ip];
}

// forward stage 0 (ntt.cpp:494-533 with m = 1), outputs fully reduced so that both halves enter their
// sub-transforms inside the fast paths' contracts
__global__ void __launch_bounds__(256) k_big_stage0_fwd(uint64_t* __restrict__ data, const uint64_t* __restrict__ roots,
                                                        const uint64_t* __restrict__ precon, uint64_t q, uint32_t n_half,
                                                        uint64_t batch) {
    const uint64_t w = roots[1], wp = precon[1], twoq = q << 1;
    const size_t total = (size_t)batch * n_half / 2;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t b = e / (n_half / 2), j = (e % (n_half / 2)) * 2;
        uint64_t* X = data + b * 2 * n_half + j;
        uint64_t* Y = X + n_half;
        uint64_t x[2], y[2];
        ld2(X, x[0], x[1]);
        ld2(Y, y[0], y[1]);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            fwd_bfly(x[k], y[k], w, wp, q, twoq);
            x[k] -= (x[k] >= twoq) ? twoq : 0;
            x[k] -= (x[k] >= q) ? q : 0;
            y[k] -= (y[k] >= twoq) ? twoq : 0;
            y[k] -= (y[k] >= q) ? q : 0;
        }
        st2(X, x[0], x[1]);
        st2(Y, y[0], y[1]);
    }
}

// inverse last stage: X = lower[j], Y = upper[j] are the sub-transform outputs WITHOUT their own last twiddle
// (stage m = 2 of the big transform, inv_roots[N - 3 + h], applies to the upper half of each sub-transform);
// then ntt.cpp:636-657:  lower = (X + Y) n^-1,  upper = (X - Y) n^-1 w,  fully reduced
__global__ void __launch_bounds__(256) k_big_last_inv(uint64_t* __restrict__ data, const uint64_t* __restrict__ inv_roots,
                                                      uint64_t q, uint64_t inv_n, uint64_t inv_n_w, uint32_t n_half,
                                                      uint64_t batch, const Divisor dv) {
    const uint64_t w0 = inv_roots[2 * n_half - 3], w1 = inv_roots[2 * n_half - 2];
    const size_t total = (size_t)batch * n_half;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t b = e / n_half, j = e % n_half;
        uint64_t* X = data + b * 2 * n_half + j;
        uint64_t* Y = X + n_half;
        uint64_t x = mod64(*X, dv), y = mod64(*Y, dv);
        if (j >= n_half / 2) {
            x = mulmod_reduced(x, w0, dv);
            y = mulmod_reduced(y, w1, dv);
        }
        *X = mulmod_reduced(add_mod(x, y, q), inv_n, dv);
        *Y = mulmod_reduced(sub_mod(x, y, q), inv_n_w, dv);
    }
}

cudaError_t launch_big_split(bool fwd, const uint64_t* tab0, const uint64_t* tab1, uint64_t* sub, uint32_t n_half,
                             cudaStream_t st) {
    dim3 g((n_half + 255) / 256, 2);
    if (fwd) k_big_split_fwd<<<g, 256, 0, st>>>(tab0, tab1, sub, n_half);
    else k_big_split_inv<<<g, 256, 0, st>>>(tab0, tab1, sub, n_half);
    return cudaGetLastError();
}
cudaError_t launch_big_stage0_fwd(uint64_t* data, const uint64_t* roots, const uint64_t* precon, uint64_t q,
                                  uint32_t n_half, uint64_t batch, cudaStream_t st) {
    k_big_stage0_fwd<<<148 * 8, 256, 0, st>>>(data, roots, precon, q, n_half, batch);
    return cudaGetLastError();
}
cudaError_t launch_big_last_inv(uint64_t* data, const uint64_t* inv_roots, uint64_t q, uint64_t inv_n, uint64_t inv_n_w,
                                uint32_t n_half, uint64_t batch, cudaStream_t st) {
    k_big_last_inv<<<148 * 8, 256, 0, st>>>(data, inv_roots, q, inv_n, inv_n_w, n_half, batch, make_divisor(q));
    return cudaGetLastError();
}

// the N/2-point sub-transforms of half `half` of every polynomial (items at stride 2)
cudaError_t launch_ntt_half(bool fwd, uint64_t* data, const ModTab& tab, uint32_t logn_half, uint64_t batch, int variant,
                            uint32_t* list, cudaStream_t st, int* launches, uint32_t half) {
    if (batch == 0) return cudaSuccess;
    if (logn_half != 14 || (variant & 1) != 1) return cudaErrorInvalidValue;   // only N = 32768 is cut this way
    using C = NttCfg<14, 5>;
    if (fwd) return launch_one<C, true>(data, tab, batch, false, list, st, launches, nullptr, 2, half);
    return launch_one<C, false>(data, tab, batch, false, list, st, launches, nullptr, 2, half);
}

}  // namespace hb
