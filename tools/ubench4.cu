// ubench4.cu -- steady-state rate of the FP64-pipe butterfly (development tool).
// The kernels run the very butterflies of csrc/modarith.cuh on registers only (no shared
// memory, no barriers, twiddles in registers), 16 independent butterflies per thread and four
// warps per scheduler like the transform kernels, so the cycles per butterfly measured here are
// what the arithmetic itself costs on the SM; the transform kernels' extra is memory and sync.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../hexl-fpga_b200/csrc -o ubench4 ubench4.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "modarith.cuh"
using namespace hb;

#define ITERS 512
#define NB 16

// KIND 0: forward FP64 butterfly   1: inverse FP64 butterfly   2: product only (6 FP64)
//      3: forward without the conditional correction (8 FP64, 0 ALU)   4: integer fast butterfly
//      5: forward FP64, twiddle re-read from a small global table every butterfly (L1 hits)
template <int KIND>
__global__ void __launch_bounds__(512, 1) k(uint64_t* out, const uint64_t* tw, uint64_t q, unsigned long long* cyc) {
    const Fp64Mod m = make_fp64mod(q, 1, 1);
    const FastMod fm = make_fastmod(q);
    __shared__ ulonglong2 stw[32];
    if (threadIdx.x < 32) stw[threadIdx.x] = reinterpret_cast<const ulonglong2*>(tw)[threadIdx.x];
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
        X[i] = KIND == 4 ? a : d2u((double)a);
        Y[i] = KIND == 4 ? b : d2u((double)b);
    }
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (KIND == 0) fwd_bfly_fp64(X[i], Y[i], w[i & 3], wi[i & 3], m);
            if (KIND == 1) inv_bfly_fp64(X[i], Y[i], w[i & 3], wi[i & 3], m);
            if (KIND == 2) {
                Y[i] = d2u(fp_mulmod(u2d(Y[i]), u2d(w[i & 3]), u2d(wi[i & 3]), m));
            }
            if (KIND == 3) {
                const double x = u2d(X[i]);
                const double r = fp_mulmod(u2d(Y[i]), u2d(w[i & 3]), u2d(wi[i & 3]), m);
                X[i] = d2u(fp_add(x, r) * 0.5);      // keep the magnitudes bounded without the correction
                Y[i] = d2u(fp_add(x, -r));
            }
            if (KIND == 4) {
                fwd_bfly_fast(X[i], Y[i], w[i & 3], wi[i & 3], fm);
                X[i] = csub(X[i], fm.q4);
                Y[i] = csub(Y[i], fm.q4);
            }
            if (KIND == 6 || KIND == 7) {   // alternative conditional corrections
                double x = u2d(X[i]);
                if (KIND == 6) {            // two predicated DADDs, one signed and one unsigned compare of the high word
                    asm("{\n\t.reg .pred p1, p2;\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\t"
                        "setp.gt.s32 p1, hi, %2;\n\tsetp.gt.u32 p2, hi, %3;\n\t"
                        "@p1 sub.rn.f64 %0, %0, %1;\n\t@p2 add.rn.f64 %0, %0, %1;\n\t}"
                        : "+d"(x)
                        : "d"(m.q), "r"(m.half_hi), "r"(m.half_hi | 0x80000000u));
                } else {                    // one predicated DADD of copysign(q, x)
                    asm("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi, a, s;\n\t.reg .f64 c;\n\tmov.b64 {lo, hi}, %0;\n\t"
                        "and.b32 a, hi, 0x7fffffff;\n\tsetp.gt.u32 p, a, %2;\n\t"
                        "lop3.b32 s, hi, 0x80000000, %3, 0xf8;\n\tmov.b64 c, {%4, s};\n\t"
                        "@p sub.rn.f64 %0, %0, c;\n\t}"
                        : "+d"(x)
                        : "d"(m.q), "r"(m.half_hi), "r"(m.q_hi), "r"(m.q_lo));
                }
                const double r = fp_mulmod(u2d(Y[i]), u2d(w[i & 3]), u2d(wi[i & 3]), m);
                X[i] = d2u(fp_add(x, r));
                Y[i] = d2u(fp_add(x, -r));
            }
            if (KIND == 8) {                // twiddle from shared memory (broadcast LDS.128) per butterfly
                const ulonglong2 t = stw[(it + i) & 31];
                fwd_bfly_fp64(X[i], Y[i], t.x, t.y, m);
            }
            if (KIND == 9) {                // one L1 twiddle load per 4 butterflies (the kernels' average is 0.41 per butterfly)
                if ((i & 3) == 0) {
                    const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(tw) + ((it + i) & 31));
                    w[0] = t.x;
                    wi[0] = t.y;
                }
                fwd_bfly_fp64(X[i], Y[i], w[0], wi[0], m);
            }
            if (KIND == 5) {
                const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(tw) + ((it + i) & 31));
                fwd_bfly_fp64(X[i], Y[i], t.x, t.y, m);
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
void run(const char* name, uint64_t* out, const uint64_t* tw, uint64_t q, unsigned long long* cyc) {
    unsigned long long h[148];
    for (int rep = 0; rep < 2; ++rep) {
        k<KIND><<<148, 512>>>(out, tw, q, cyc);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    // four warps per scheduler: cycles of the scheduler per warp-wide butterfly
    const double per = (double)h[0] / ((double)ITERS * NB * 4.0);
    printf("{\"ubench4\": \"%s\", \"cycles\": %llu, \"smsp_cycles_per_warp_butterfly\": %.2f}\n", name, h[0], per);
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
    run<0>("forward FP64 butterfly (9 FP64 + 5 ALU)", out, tw, q, cyc);
    run<1>("inverse FP64 butterfly (9 FP64 + 5 ALU)", out, tw, q, cyc);
    run<2>("FP64 modular product only (6 FP64)", out, tw, q, cyc);
    run<3>("forward FP64 butterfly without the conditional correction (8 FP64 + 1 DMUL)", out, tw, q, cyc);
    run<4>("integer fast butterfly + 2 corrections (5 IMAD.WIDE + 4 IMAD)", out, tw, q, cyc);
    run<5>("forward FP64 butterfly, twiddle from L1 per butterfly", out, tw, q, cyc);
    run<8>("forward FP64 butterfly, twiddle from shared memory per butterfly", out, tw, q, cyc);
    run<9>("forward FP64 butterfly, one L1 twiddle load per 4 butterflies", out, tw, q, cyc);
    run<6>("forward FP64 butterfly, correction = 2 ISETP + 2 predicated DADD", out, tw, q, cyc);
    run<7>("forward FP64 butterfly, correction = LOP3 + ISETP + LOP3 + MOV + predicated DADD", out, tw, q, cyc);
    return cudaDeviceSynchronize() != cudaSuccess;
}
