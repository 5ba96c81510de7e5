// ubench5.cu -- can the FP64 pipe and the integer multiplier pipe carry butterflies AT THE SAME TIME?
// Registers only, 16 warps per SM (four per scheduler) as in the transform kernels.  A warp runs either the
// FP64-pipe butterfly (modarith.cuh fwd_bfly_fp64) or the integer fast butterfly (fwd_bfly_fast, no extra
// corrections); `nint` of the four warps of every scheduler are integer warps.  Reported: cycles until the
// LAST warp is done when every warp runs the same number of butterflies (the transform kernels' situation:
// the polynomial is dealt out evenly), and the butterfly counts per warp kind that would finish together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../hexl-fpga_b200/csrc -o ubench5 ubench5.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "modarith.cuh"
using namespace hb;

#define NB 16

__global__ void __launch_bounds__(512, 1) k(uint64_t* out, const uint64_t* tw, uint64_t q, unsigned long long* cyc,
                                            int nint, int iters_fp, int iters_int) {
    const Fp64Mod m = make_fp64mod(q, 1, 1);
    const FastMod fm = make_fastmod(q);
    const int warp = threadIdx.x >> 5;
    const bool is_int = (warp >> 2) < nint;          // warps w, w+4, w+8, w+12 share a scheduler
    uint64_t X[NB], Y[NB];
    uint64_t w[4], wi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[i] = tw[2 * (i + (threadIdx.x & 3))];
        wi[i] = tw[2 * (i + (threadIdx.x & 3)) + 1];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const uint64_t a = (threadIdx.x * 977u + i * 131u + blockIdx.x) % 1000003u, b = (a * 7919u + 13u) % 1000003u;
        X[i] = is_int ? a : d2u((double)a);
        Y[i] = is_int ? b : d2u((double)b);
    }
    __syncthreads();
    const unsigned long long t0 = clock64();
    if (is_int) {
        for (int it = 0; it < iters_int; ++it) {
#pragma unroll
            for (int i = 0; i < NB; ++i) fwd_bfly_fast(X[i], Y[i], w[i & 3] & 0xfffffffffffffull, wi[i & 3], fm);
        }
    } else {
        for (int it = 0; it < iters_fp; ++it) {
#pragma unroll
            for (int i = 0; i < NB; ++i) fwd_bfly_fp64(X[i], Y[i], w[i & 3], wi[i & 3], m);
        }
    }
    const unsigned long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) s += X[i] ^ Y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
}

int main() {
    const uint64_t q = 2251799814045697ULL;
    uint64_t h_tw[64];
    for (int i = 0; i < 32; ++i) {
        const uint64_t r = (q / 3 + 1234567ULL * i) % q;
        const double ws = fp_centred(r, q);
        h_tw[2 * i] = d2u(ws);
        h_tw[2 * i + 1] = d2u(ws / (double)q);
    }
    uint64_t *out, *tw;
    unsigned long long* cyc;
    cudaMalloc(&out, 148 * 512 * 8);
    cudaMalloc(&tw, sizeof h_tw);
    cudaMalloc(&cyc, 148 * 16 * 8);
    cudaMemcpy(tw, h_tw, sizeof h_tw, cudaMemcpyHostToDevice);
    const int base = 512;
    struct Cfg { int nint, ifp, iint; } cfgs[] = {
        {0, base, base}, {4, base, base}, {2, base, base}, {1, base, base},
        {2, base, base / 2}, {2, base, base * 5 / 8}, {2, base, base * 3 / 4}, {1, base, base * 3 / 2}, {1, base, base * 2},
        {1, base, base * 5 / 4}, {3, base, base / 4}};
    for (const Cfg& c : cfgs) {
        unsigned long long h[16];
        for (int rep = 0; rep < 2; ++rep) {
            k<<<148, 512>>>(out, tw, q, cyc, c.nint, c.ifp, c.iint);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        unsigned long long tmax = 0, tfp = 0, tint = 0;
        for (int w = 0; w < 16; ++w) {
            tmax = h[w] > tmax ? h[w] : tmax;
            if ((w >> 2) < c.nint) tint = h[w] > tint ? h[w] : tint; else tfp = h[w] > tfp ? h[w] : tfp;
        }
        // butterflies per scheduler (warp-wide) and scheduler cycles per butterfly overall
        const double bf = (double)NB * ((4 - c.nint) * c.ifp + c.nint * c.iint);
        printf("{\"ubench5\": \"int warps per scheduler %d, fp iters %d, int iters %d\", \"cycles_fp\": %llu, \"cycles_int\": %llu, "
               "\"smsp_cycles_per_warp_butterfly_overall\": %.2f}\n", c.nint, c.ifp, c.iint, tfp, tint, (double)tmax / bf);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
