// ubench3.cu -- per-SM global store / load rate with few CTAs (L2-resident footprint), so
// that HBM is not the limit: how fast can one SM push results out?  (development tool)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench3 ubench3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(512, 1) k(uint64_t* buf, unsigned long long* cyc, int iters) {
    // each CTA owns 1 MiB (131072 words), walked repeatedly
    uint64_t* p = buf + (size_t)blockIdx.x * 131072;
    const uint32_t tid = threadIdx.x;
    uint64_t acc = tid;
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t off = ((it & 63) * 2048);   // 16 KiB per iteration per CTA
        if (KIND == 0) {  // 32-byte stores, 1 KiB contiguous per warp instruction
            asm volatile("st.global.v4.b64 [%0], {%1, %1, %1, %1};" ::"l"(p + off + tid * 4), "l"(acc) : "memory");
        } else if (KIND == 1) {  // 8-byte stores, 256 B contiguous per warp instruction, 4 per iteration
#pragma unroll
            for (int j = 0; j < 4; ++j) p[off + j * 512 + tid] = acc;
        } else {  // 16-byte loads (for comparison)
            uint64_t a0, a1, b0, b1;
            asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "l"(p + off + tid * 2));
            asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "l"(p + off + 1024 + tid * 2));
            acc += a0 + b1;
        }
    }
    unsigned long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x1234567) buf[0] = acc;
}

int main() {
    uint64_t* buf; unsigned long long* cyc;
    cudaMalloc(&buf, (size_t)148 << 20);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    const char* names[] = {"st.v4.b64 (32 B/thread)", "st.b64 x4 (coalesced 8 B)", "ld.v2.b64 x2"};
    for (int ctas : {1, 8, 148}) {
        for (int kind = 0; kind < 3; ++kind) {
            unsigned long long h[148];
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) k<0><<<ctas, 512>>>(buf, cyc, iters);
                if (kind == 1) k<1><<<ctas, 512>>>(buf, cyc, iters);
                if (kind == 2) k<2><<<ctas, 512>>>(buf, cyc, iters);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            double bytes = (double)iters * 16384;
            printf("{\"ubench3\": \"%s\", \"ctas\": %d, \"bytes_per_clk_per_sm\": %.1f}\n", names[kind], ctas, bytes / (double)h[0]);
        }
    }
    return 0;
}
