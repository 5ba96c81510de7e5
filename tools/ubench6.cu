// ubench6.cu -- FP64 pipe details behind the forward butterflies (development tool):
//   * issue cost of DADD / DMUL / DFMA alone (is the 2.15 cycles per instruction of ubench4 the
//     same for two- and three-operand instructions?)
//   * the every-other-stage butterfly pair of the forward kernels (fwd_bfly_fp64_a + _b)
//   * the same with the quotient rounded by FRND.F64 (cvt.rni.f64.f64) instead of the
//     magic-constant add/subtract: 5 FP64 + 1 conversion per product instead of 6 FP64
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../hexl-fpga_b200/csrc -o ubench6 ubench6.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "modarith.cuh"
using namespace hb;

#define ITERS 512
#define NB 16

__device__ __forceinline__ double rni(double x) {
    double c;
    asm("cvt.rni.f64.f64 %0, %1;" : "=d"(c) : "d"(x));
    return c;
}
__device__ __forceinline__ double mulmod_frnd(double y, double w, double wi, const Fp64Mod& m) {
    const double c = rni(fp_mul(y, wi));
    const double h = fp_mul(y, w);
    const double l = fp_fma(y, w, -h);
    const double d = fp_fma(c, m.nq, h);
    return fp_add(d, l);
}
__device__ __forceinline__ double cred_full_frnd(double x, const Fp64Mod& m) {
    const double c = rni(fp_mul(x, m.inv_q));
    return fp_fma(c, m.nq, x);
}

template <int KIND>
__global__ void __launch_bounds__(512, 1) k(uint64_t* out, const uint64_t* tw, uint64_t q, unsigned long long* cyc) {
    const Fp64Mod m = make_fp64mod(q, 1, 1);
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
        X[i] = d2u((double)a);
        Y[i] = d2u((double)b);
    }
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (KIND == 0) {        // stage A then stage B on the same pair (two butterflies per iteration)
                fwd_bfly_fp64_a(X[i], Y[i], w[i & 3], wi[i & 3], m);
                fwd_bfly_fp64_b(X[i], Y[i], w[(i + 1) & 3], wi[(i + 1) & 3], m);
            }
            if (KIND == 1) {        // the same with FRND quotients
                double x = cred_full_frnd(u2d(X[i]), m);
                double r = mulmod_frnd(u2d(Y[i]), u2d(w[i & 3]), u2d(wi[i & 3]), m);
                double x2 = fp_add(x, r), y2 = fp_add(x, -r);
                r = mulmod_frnd(y2, u2d(w[(i + 1) & 3]), u2d(wi[(i + 1) & 3]), m);
                X[i] = d2u(fp_add(x2, r));
                Y[i] = d2u(fp_add(x2, -r));
            }
            if (KIND == 2) {        // 8 dependent-free DFMA per slot
                double x = u2d(X[i]), y = u2d(Y[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    x = fp_fma(x, u2d(w[j]), y);
                    y = fp_fma(y, u2d(wi[j]), x);
                }
                X[i] = d2u(x);
                Y[i] = d2u(y);
            }
            if (KIND == 3) {        // 8 DADD
                double x = u2d(X[i]), y = u2d(Y[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    x = fp_add(x, y);
                    y = fp_add(y, u2d(w[j]));
                }
                X[i] = d2u(x);
                Y[i] = d2u(y);
            }
            if (KIND == 4) {        // 8 DMUL
                double x = u2d(X[i]), y = u2d(Y[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    x = fp_mul(x, u2d(wi[j]));
                    y = fp_mul(y, u2d(wi[j]));
                }
                X[i] = d2u(x);
                Y[i] = d2u(y);
            }
            if (KIND == 5) {        // product only, FRND
                Y[i] = d2u(mulmod_frnd(u2d(Y[i]), u2d(w[i & 3]), u2d(wi[i & 3]), m));
            }
            if (KIND == 6) {        // 8 FRND.F64 only
                double x = u2d(X[i]), y = u2d(Y[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    x = rni(x);
                    y = rni(y);
                    asm volatile("" : "+d"(x), "+d"(y));
                }
                X[i] = d2u(x);
                Y[i] = d2u(y);
            }
        }
    }
    const unsigned long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) s += X[i] ^ Y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, double units, uint64_t* out, const uint64_t* tw, uint64_t q, unsigned long long* cyc) {
    unsigned long long h[148];
    for (int rep = 0; rep < 2; ++rep) {
        k<KIND><<<148, 512>>>(out, tw, q, cyc);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    // four warps per scheduler: scheduler cycles per warp-wide unit (butterfly or instruction)
    const double per = (double)h[0] / ((double)ITERS * NB * 4.0 * units);
    printf("{\"ubench6\": \"%s\", \"cycles\": %llu, \"smsp_cycles_per_warp_unit\": %.3f}\n", name, h[0], per);
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
    cudaMalloc(&cyc, 148 * 8);
    cudaMemcpy(tw, h_tw, sizeof h_tw, cudaMemcpyHostToDevice);
    run<0>("butterfly, stages A+B, magic-constant quotients (unit = butterfly)", 2, out, tw, q, cyc);
    run<1>("butterfly, stages A+B, FRND quotients (unit = butterfly)", 2, out, tw, q, cyc);
    run<2>("DFMA only (unit = instruction)", 8, out, tw, q, cyc);
    run<3>("DADD only (unit = instruction)", 8, out, tw, q, cyc);
    run<4>("DMUL only (unit = instruction)", 8, out, tw, q, cyc);
    run<5>("product only, FRND quotient (unit = product)", 1, out, tw, q, cyc);
    run<6>("FRND.F64 only (unit = instruction)", 8, out, tw, q, cyc);
    return cudaDeviceSynchronize() != cudaSuccess;
}
