// ubench.cu -- integer-pipe microbenchmarks for sm_100a (development tool).
// Measures sustained per-SM throughput of the instructions the Harvey
// butterfly is made of, and of the butterfly itself, with all 148 SMs busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../hexl-fpga_b200/csrc/modarith.cuh"

#define ITERS 2048
#define ILP 8

template <int KIND>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, uint32_t seed, unsigned long long* cyc) {
    uint32_t a[ILP], b = seed | 1, c = seed * 3 + 7;
    uint64_t w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x + i * seed; w[i] = ((uint64_t)a[i] << 32) | (i + seed); }
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
            if (KIND == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 3) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 5) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)b << 32 | c));
            if (KIND == 6) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)b << 32 | c));
            if (KIND == 7) {  // IMAD + IADD3-ish pair: can they dual-issue?
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(uint32_t*)&w[i]) : "r"(b), "r"(c));
            }
            if (KIND == 8) asm volatile("add.cc.u32 %0, %0, %1; addc.u32 %2, %2, %3;" : "+r"(a[i]), "+r"(b) : "r"(c), "r"(seed));
            if (KIND == 9) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (KIND == 10) {  // IMAD.WIDE + independent 3-input add
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
                asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b), "r"(c));
            }
            if (KIND == 11) {  // 1 IMAD.WIDE : 2 adds
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
                asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                asm volatile("addc.u32 %0, %0, %1;" : "+r"(a[(i + 1) % ILP]) : "r"(c));
            }
            if (KIND == 12) {  // 2 IMAD : 1 LOP3
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[(i + 3) % ILP]) : "r"(b), "r"(c));
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// register-resident Harvey butterflies: 8 independent (X,Y) pairs per thread
template <int VAR>
__global__ void __launch_bounds__(1024, 1) kb(uint64_t* out, uint64_t q, uint64_t w, uint64_t wp,
                                               unsigned long long* cyc) {
    uint64_t X[ILP], Y[ILP];
    const uint64_t twoq = 2 * q;
    const hb::FastMod fm = hb::make_fastmod(q);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { X[i] = (threadIdx.x * 77 + i) % q; Y[i] = (threadIdx.x * 131 + 5 * i) % q; }
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (VAR == 0) hb::fwd_bfly(X[i], Y[i], w, wp, q, twoq);
            if (VAR == 1) hb::inv_bfly(X[i], Y[i], w, wp, q, twoq);
            if (VAR == 3) hb::fwd_bfly_fast(X[i], Y[i], w, wp, fm);
            if (VAR == 4) hb::inv_bfly_fast(X[i], Y[i], w, wp, fm);
            if (VAR == 2) {  // lazy: no per-stage correction
                uint64_t T = hb::mul_lazy(Y[i], w, wp, q);
                uint64_t x = X[i];
                X[i] = x + T;
                Y[i] = x + twoq - T;
            }
        }
        w += it; wp ^= it;
    }
    unsigned long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += X[i] ^ Y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    uint32_t* out; unsigned long long* cyc;
    const int blocks = 148, threads = 1024;
    cudaMalloc(&out, blocks * threads * 8);
    cudaMalloc(&cyc, blocks * 8);
    unsigned long long h[148];
    const char* names[] = {"IMAD.lo32", "IMAD.WIDE.U32", "IMAD.HI.U32", "IADD x2", "LOP3", "mul.hi.u64", "mul.lo.u64",
                           "IMAD+LOP3 pair", "add.cc+addc", "SHF", "IMAD.WIDE+IADD3", "IMAD.WIDE+add.cc+addc",
                           "IMAD+IMAD.WIDE+LOP3"};
    const double ops_per_iter[] = {1, 1, 1, 2, 1, 1, 1, 2, 2, 1, 2, 3, 3};
    for (int kind = 0; kind < 13; ++kind) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (kind) {
                case 0: k<0><<<blocks, threads>>>(out, 12345, cyc); break;
                case 1: k<1><<<blocks, threads>>>(out, 12345, cyc); break;
                case 2: k<2><<<blocks, threads>>>(out, 12345, cyc); break;
                case 3: k<3><<<blocks, threads>>>(out, 12345, cyc); break;
                case 4: k<4><<<blocks, threads>>>(out, 12345, cyc); break;
                case 5: k<5><<<blocks, threads>>>(out, 12345, cyc); break;
                case 6: k<6><<<blocks, threads>>>(out, 12345, cyc); break;
                case 7: k<7><<<blocks, threads>>>(out, 12345, cyc); break;
                case 8: k<8><<<blocks, threads>>>(out, 12345, cyc); break;
                case 9: k<9><<<blocks, threads>>>(out, 12345, cyc); break;
                case 10: k<10><<<blocks, threads>>>(out, 12345, cyc); break;
                case 11: k<11><<<blocks, threads>>>(out, 12345, cyc); break;
                case 12: k<12><<<blocks, threads>>>(out, 12345, cyc); break;
            }
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double c = (double)h[0];
        double per_clk = (double)threads * ITERS * ILP * ops_per_iter[kind] / c;
        printf("{\"ubench\": \"%s\", \"cycles\": %.0f, \"thread_ops_per_clk_per_sm\": %.1f}\n", names[kind], c, per_clk);
    }
    const uint64_t q = 2251799814045697ull, w = 1111640190223217ull;
    const uint64_t wp = (uint64_t)((((unsigned __int128)w) << 64) / q);
    const char* bn[] = {"fwd_bfly exact", "inv_bfly exact", "fwd_bfly lazy", "fwd_bfly_fast", "inv_bfly_fast"};
    for (int v = 0; v < 5; ++v) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (v == 0) kb<0><<<blocks, threads>>>((uint64_t*)out, q, w, wp, cyc);
            if (v == 1) kb<1><<<blocks, threads>>>((uint64_t*)out, q, w, wp, cyc);
            if (v == 2) kb<2><<<blocks, threads>>>((uint64_t*)out, q, w, wp, cyc);
            if (v == 3) kb<3><<<blocks, threads>>>((uint64_t*)out, q, w, wp, cyc);
            if (v == 4) kb<4><<<blocks, threads>>>((uint64_t*)out, q, w, wp, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double c = (double)h[0];
        double bf = (double)threads * (ITERS / 4) * ILP;
        printf("{\"ubench\": \"%s\", \"cycles\": %.0f, \"butterflies_per_clk_per_sm\": %.3f, \"clk_per_warp_bfly\": %.2f, "
               "\"ms\": %.4f, \"sm_mhz_eff\": %.0f, \"ntt16384_per_s_ceiling\": %.3e}\n",
               bn[v], c, bf / c, 32.0 * c / bf, ms, c / ms / 1e3, 148.0 * bf / (ms * 1e-3) / 114688.0);
    }
    return 0;
}
