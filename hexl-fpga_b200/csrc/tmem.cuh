// tmem.cuh -- the SM's 256 KiB of tensor memory (TMEM: 512 columns x 128 lanes x 32 bit) used as a plain
// scratchpad.  The kernels here have no tensor-core work (integer modular arithmetic is not a contraction),
// which leaves the whole TMEM free: a CTA allocates all of it and parks per-thread state there through
// tcgen05.st / tcgen05.ld (.32x32b: thread t of a warp reaches lane 32 * (warp % 4) + t only, so state
// stays with the thread that owns it) -- the keyswitch sums of keyswitch_fused.cu, NTT(a) in the
// single-launch polynomial multiply of polymul_fused.cu.
#pragma once
#include "ntt_block.cuh"

namespace hb {

HB_D void tmem_alloc_all(uint32_t* smem_slot) {   // one warp; 512 columns = the whole TMEM of the SM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
HB_D void tmem_dealloc_all(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
// 16 consecutive 32-bit columns of this thread's lane = 8 accumulator words
HB_D void tmem_ld16(uint32_t taddr, uint64_t* a) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ((uint64_t)r[2 * i + 1] << 32) | r[2 * i];
}
HB_D void tmem_st16(uint32_t taddr, const uint64_t* a) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"((uint32_t)a[0]), "r"((uint32_t)(a[0] >> 32)), "r"((uint32_t)a[1]), "r"((uint32_t)(a[1] >> 32)),
          "r"((uint32_t)a[2]), "r"((uint32_t)(a[2] >> 32)), "r"((uint32_t)a[3]), "r"((uint32_t)(a[3] >> 32)),
          "r"((uint32_t)a[4]), "r"((uint32_t)(a[4] >> 32)), "r"((uint32_t)a[5]), "r"((uint32_t)(a[5] >> 32)),
          "r"((uint32_t)a[6]), "r"((uint32_t)(a[6] >> 32)), "r"((uint32_t)a[7]), "r"((uint32_t)(a[7] >> 32))
        : "memory");
}
HB_D void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


// tensor-memory address of column `col` of the calling thread's lane
HB_D uint32_t tmem_thread_addr(uint32_t tmem_base, uint32_t col) {
    return tmem_base + ((((threadIdx.x >> 5) & 3u) * 32u) << 16) + col;
}

}  // namespace hb
