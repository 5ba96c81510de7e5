// ubench2.cu -- second round of pipe microbenchmarks (development tool):
// what does an IMAD.WIDE really cost next to other instructions, and what can
// the FP64 pipe do concurrently?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 1024
#define ILP 8

template <int KIND>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, uint32_t seed, unsigned long long* cyc, int threads_active) {
    uint32_t a[ILP], a2[ILP], b = seed | 1, c = seed * 3 + 7;
    uint64_t w[ILP], w2[ILP];
    double d[ILP], e = 1.0000001 + seed * 1e-9, f = 0.5 + seed * 1e-9;
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = threadIdx.x + i * seed; a2[i] = a[i] * 7 + 1;
        w[i] = ((uint64_t)a[i] << 32) | (i + seed); w2[i] = w[i] * 3;
        d[i] = 1.0 + a[i] * 1e-6;
    }
    if ((int)threadIdx.x >= threads_active) return;
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
            if (KIND == 1) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(c));
            if (KIND == 2) {  // wide product feeding a 32-bit add of its high half (dependent ALU op)
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"((uint32_t)(w[i] >> 32)));
            }
            if (KIND == 3) {  // 5 wide + 4 lo, no adds
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(c));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2[i]) : "r"(a2[i]), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a2[i]), "r"(c));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2[i]) : "r"(a[i]), "r"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a2[i]) : "r"(b), "r"(c));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(c), "r"(c));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a2[i]) : "r"(b), "r"(b));
            }
            if (KIND == 4) {  // 1 wide(acc) + 1 three-input add on other registers
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
                asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b), "r"(c));
            }
            if (KIND == 5) {  // 1 wide(acc) + 2 independent 3-input adds
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
                asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a2[i]) : "r"(b), "r"(c));
            }
            if (KIND == 6) {  // 1 lo + 1 add
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a2[i]) : "r"(b), "r"(c));
            }
            if (KIND == 7) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e), "d"(f));
            if (KIND == 8) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(f));
            if (KIND == 9) {  // DFMA + IMAD.WIDE: separate pipes?
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e), "d"(f));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
            }
            if (KIND == 10) {  // DFMA + 2 IMAD.WIDE + 2 adds
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e), "d"(f));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2[i]) : "r"(b), "r"(c));
                asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a2[i]) : "r"(b), "r"(c));
            }
            if (KIND == 11) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 12) {  // wide with both halves consumed by 32-bit carry adds (as in the butterfly)
                uint32_t lo, hi;
                asm volatile("{.reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t;}" : "=r"(lo), "=r"(hi) : "r"(a[i]), "r"(c));
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(a2[i]) : "r"(lo), "r"(hi));
            }
            if (KIND == 13) {  // 64-bit IADD pair only
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(a2[i]) : "r"(b), "r"(c));
            }
            if (KIND == 14) {  // mad.lo x2 vs wide: low64 pieces
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a2[i]) : "r"(b), "r"(c));
            }
            if (KIND == 15) {  // F2I-free trick: DFMA + DADD (magic) + LOP on low word
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e), "d"(f));
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + a2[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32) + (uint32_t)w2[i] + (uint32_t)(w2[i] >> 32) + (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, uint32_t* out, unsigned long long* cyc, int threads_active = 1024) {
    unsigned long long h[148];
    for (int rep = 0; rep < 2; ++rep) { k<KIND><<<148, 1024>>>(out, 12345, cyc, threads_active); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = (double)h[0];
    // cycles per (warp, group-iteration) per SM sub-partition: warps per SMSP = threads_active/128
    double per = c / ((double)ITERS * ILP * (threads_active / 128.0));
    printf("{\"ubench2\": \"%s\", \"threads\": %d, \"cycles\": %.0f, \"smsp_cycles_per_warp_group\": %.2f}\n", name, threads_active, c, per);
}

int main() {
    uint32_t* out; unsigned long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8);
    cudaMalloc(&cyc, 148 * 8);
    run<0>("mad.wide acc (same operands)", out, cyc);
    run<1>("mul.wide (no acc, varying a)", out, cyc);
    run<2>("mul.wide + dependent add of hi", out, cyc);
    run<3>("5 mad.wide + 4 mad.lo", out, cyc);
    run<4>("mad.wide + IADD3", out, cyc);
    run<5>("mad.wide + 2 IADD3", out, cyc);
    run<6>("mad.lo + IADD3", out, cyc);
    run<7>("DFMA", out, cyc);
    run<8>("DADD", out, cyc);
    run<9>("DFMA + mad.wide", out, cyc);
    run<10>("DFMA + 2 mad.wide + 2 IADD3", out, cyc);
    run<11>("mad.hi", out, cyc);
    run<12>("mul.wide + add.cc/addc of both halves", out, cyc);
    run<13>("add.cc/addc pair", out, cyc);
    run<14>("2 mad.lo", out, cyc);
    run<15>("DFMA + DADD + LOP3", out, cyc);
    run<0>("mad.wide acc (same operands)", out, cyc, 512);
    run<3>("5 mad.wide + 4 mad.lo", out, cyc, 512);
    run<7>("DFMA", out, cyc, 512);
    return 0;
}
