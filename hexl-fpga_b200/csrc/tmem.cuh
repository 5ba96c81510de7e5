// tmem.cuh -- the SM's 256 KiB of tensor memory (TMEM: 512 columns x 128 lanes x 32 bit) used as a plain
// scratchpad.  The kernels here have no tensor-core work (integer modular arithmetic is not a contraction),
// which leaves the whole TMEM free: a CTA allocates all of it and parks per-thread state there through
// tcgen05.st / tcgen05.ld (.32x32b: thread t of a warp reaches lane 32 * (warp % 4) + t only, so state
// stays with the thread that owns it) -- the keyswitch sums of keyswitch_fused.cu, NTT(a) in the
// single-launch polynomial multiply of polymul_fused.cu.
#pragma once
#include "ntt_block.cuh"   // the primitives (tmem_alloc_all, tmem_ld16, tmem_st16, ...) live next to the mbarrier / TMA ones
