// ntt_block.cuh -- CTA-level forward / inverse NTT: persistent CTAs, one
// polynomial at a time in shared memory, the NEXT polynomial prefetched by TMA
// into the same buffer while the last pass of the current one runs out of
// registers.  Device only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "ntt_core.cuh"

namespace hb {

// Everything a kernel needs to transform under one modulus.
struct ModTab {
    uint64_t q, twoq;
    uint64_t mu;            // floor(2^64 / q), for barrett_reduce64
    InvScale sc;            // inv_n, inv_n_w and their Shoup factors
    FastMod fm;
    const TwPair* ftw;      // packed forward twiddles  (NttCfg::FWD_ENTRIES)
    const TwPair* itw;      // packed inverse twiddles  (NttCfg::INV_ENTRIES)
    uint32_t fwd_fast_ok;   // modulus small enough for the lazy forward path
    uint32_t inv_fast_ok;
};

// ---- load transforms (applied to each word as it enters the transform) ----
struct XfIdent {
    HB_D uint64_t operator()(uint64_t x) const { return x; }
};
// base conversion of a coefficient-form word to this modulus
// (device/keyswitch/intt1_redu.hpp:36-38)
struct XfReduce {
    uint64_t q, mu;
    HB_D uint64_t operator()(uint64_t x) const { return barrett_reduce64(x, q, mu); }
};
// keyswitch rounding + base conversion (device/keyswitch/intt2_redu.hpp:24-51):
// v = (x + floor(qk/2)) mod qk ; out = (v mod qi + fix_i) mod qi
struct XfKsRound {
    uint64_t qk, qk_half, q, mu, fix;
    HB_D uint64_t operator()(uint64_t x) const {
        uint64_t v = x + qk_half;
        v -= (v >= qk) ? qk : 0;
        uint64_t r = barrett_reduce64(v, q, mu) + fix;
        return r - ((r >= q) ? q : 0);
    }
};

// ---- output functors ----
struct OfRows {  // forward: 16 contiguous words of row `row` -> 8 x 16-byte stores
    uint64_t* dst;
    HB_D void row(uint32_t r, const uint64_t* v) const {
#pragma unroll
        for (int c = 0; c < 8; ++c) st2(dst + r * 16 + 2 * c, v[2 * c], v[2 * c + 1]);
    }
};
struct OfWords {  // inverse: one word at its natural index (coalesced along lo)
    uint64_t* dst;
    HB_D void word(uint32_t idx, uint64_t x) const { dst[idx] = x; }
};

// ---------------------------------------------------------------------------
// mbarrier / TMA primitives (PTX; SASS: SYNCS.*, UTMALDG)
// ---------------------------------------------------------------------------
HB_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
HB_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
HB_D void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
HB_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
HB_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
HB_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
HB_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 2-D tensor TMA load: box (16 words x box_rows) at row coordinate `row`
HB_D void tma_load_rows(void* smem_dst, const CUtensorMap* map, uint64_t* bar, uint32_t row) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(0), "r"(row)
        : "memory");
}

// One polynomial = N/16 rows of 128 bytes; a TMA box holds at most 256 rows.
template <class C>
struct TmaGeom {
    static constexpr uint32_t ROWS = C::N / 16;
    static constexpr uint32_t BOX_ROWS = ROWS < 256 ? ROWS : 256;
    static constexpr uint32_t BOXES = ROWS / BOX_ROWS;
    static constexpr uint32_t BYTES = C::N * 8;
};

// issued by ONE thread: arm the barrier and start the copy of the polynomial
// whose first row is `row0` into the (1024-byte aligned) buffer W
template <class C>
HB_D void issue_poly_load(uint64_t* W, const CUtensorMap* map, uint64_t* bar, uint32_t row0) {
    using G = TmaGeom<C>;
    mbar_expect_tx(bar, G::BYTES);
#pragma unroll
    for (uint32_t b = 0; b < G::BOXES; ++b)
        tma_load_rows(W + (size_t)b * G::BOX_ROWS * 16, map, bar, row0 + b * G::BOX_ROWS);
}

// The TMA prefetch of the next polynomial: one live register (the tensor-map
// row, or kNoPrefetch); buffer and barrier addresses are recomputed from the
// shared-memory base so they cost no registers in the butterfly code.
constexpr uint32_t kNoPrefetch = 0xffffffffu;

template <class C>
HB_D uint64_t* smem_poly() {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    return reinterpret_cast<uint64_t*>(smem_raw);
}

struct Prefetch {
    const CUtensorMap* map;
    uint32_t row;     // first tensor-map row of the next polynomial, or kNoPrefetch
    template <class C>
    HB_D void issue() const {
        if (row != kNoPrefetch) {
            uint64_t* W = smem_poly<C>();
            fence_proxy_async();
            issue_poly_load<C>(W, map, W + C::N, row);
        }
    }
};

// ---------------------------------------------------------------------------
// forward transform of the polynomial in W
// ---------------------------------------------------------------------------
template <class C, int P, class A>
HB_D void fwd_mid_passes(uint32_t tid, uint64_t* W, const TwPair* tw, const A& a) {
    if constexpr (P < C::NP) {
        fwd_head_pass<C, P>(tid, W, tw, a);
        __syncthreads();
        fwd_mid_passes<C, P + 1>(tid, W, tw, a);
    }
}

template <class C, class A, class Of>
HB_D void fwd_rest(uint32_t tid, uint64_t* W, uint64_t* v, const TwPair* tw, const A& a, const Of& of,
                   const Prefetch& pf) {
    using P0 = FwdPass<C, 0>;
    fwd_head_compute<C, 0>(tid, v, tw, a);
    head_store<C, P0::R, P0::LS>(tid, W, v);
    __syncthreads();
    fwd_mid_passes<C, 1>(tid, W, tw, a);
    tail_load<C>(tid, W, v, XfIdent());
    __syncthreads();        // every word of W is in registers now
    pf.template issue<C>(); // ... so the buffer can take the next polynomial
    fwd_tail_compute<C>(tid, v, tw, a);
#pragma unroll
    for (int ri = 0; ri < C::E / 16; ++ri) of.row(tid + ri * C::NT, v + ri * 16);
}

// reference op sequence for polynomials with out-of-contract words (rare):
// its own function so its register needs do not leak into the fast path.
// Re-reads the input from W (nothing has been stored yet).
template <class C, class Xf, class Of>
__device__ __noinline__ void fwd_exact_cta(uint64_t* W, const ModTab* t, Xf xf, Of of, Prefetch pf) {
    using P0 = FwdPass<C, 0>;
    const uint32_t tid = threadIdx.x;
    uint64_t v[C::E];
    head_load<C, P0::R, P0::LS>(tid, W, v, xf);
    const ExactArith a = {t->q, t->twoq};
    fwd_rest<C>(tid, W, v, t->ftw, a, of, pf);
}

// W holds the input (natural order, swizzled rows); output bit-reversed, [0,q)
template <class C, bool ASSUME_OK, class Xf, class Of>
HB_D void ntt_fwd_cta(uint64_t* W, const ModTab& t, const Xf& xf, const Of& of, const Prefetch& pf) {
    using P0 = FwdPass<C, 0>;
    const uint32_t tid = threadIdx.x;
    uint64_t v[C::E];
    head_load<C, P0::R, P0::LS>(tid, W, v, xf);
    if constexpr (ASSUME_OK) {   // perf-exploration knob: caller guarantees the contract
        const FastArith a = {t.fm};
        fwd_rest<C>(tid, W, v, t.ftw, a, of, pf);
        return;
    }
    // forward contract: every word < 4q (tests/test_utils/ntt.cpp:483-486)
    int bad = 0;
#pragma unroll
    for (int e = 0; e < C::E; ++e) bad |= (v[e] >= t.fm.q4);
    if (__syncthreads_or(bad) != 0 || !t.fwd_fast_ok) {
        fwd_exact_cta<C>(W, &t, xf, of, pf);
    } else {
        const FastArith a = {t.fm};
        fwd_rest<C>(tid, W, v, t.ftw, a, of, pf);
    }
}

// ---------------------------------------------------------------------------
// inverse transform of the polynomial in W
// ---------------------------------------------------------------------------
template <class C, int P, class A>
HB_D void inv_mid_passes(uint32_t tid, uint64_t* W, const TwPair* tw, const A& a) {
    if constexpr (P < C::NP - 1) {
        inv_head_pass<C, P>(tid, W, tw, a);
        __syncthreads();
        inv_mid_passes<C, P + 1>(tid, W, tw, a);
    }
}

template <class C, class A, class Of>
HB_D void inv_rest(uint32_t tid, uint64_t* W, uint64_t* v, const ModTab& t, const A& a, const Of& of,
                   const Prefetch& pf) {
    using PL = InvPass<C, C::NP - 1>;
    inv_tail_compute<C>(tid, v, t.itw, a);
    tail_store<C>(tid, W, v);
    __syncthreads();
    inv_mid_passes<C, 0>(tid, W, t.itw, a);
    head_load<C, PL::R, PL::LS>(tid, W, v, XfIdent());
    __syncthreads();
    pf.template issue<C>();
    inv_head_compute<C, C::NP - 1>(tid, v, t.itw, a, t.sc);
#pragma unroll
    for (int gi = 0; gi < (C::E >> PL::R); ++gi)
#pragma unroll
        for (int k = 0; k < (1 << PL::R); ++k) of.word(inv_last_index<C>(tid, gi, k), v[gi * (1 << PL::R) + k]);
}

template <class C, class Xf, class Of>
__device__ __noinline__ void inv_exact_cta(uint64_t* W, const ModTab* t, Xf xf, Of of, Prefetch pf) {
    const uint32_t tid = threadIdx.x;
    uint64_t v[C::E];
    tail_load<C>(tid, W, v, xf);
    const ExactArith a = {t->q, t->twoq};
    inv_rest<C>(tid, W, v, *t, a, of, pf);
}

// W holds the input (bit-reversed order); output natural order, [0,q)
template <class C, bool ASSUME_OK, class Xf, class Of>
HB_D void ntt_inv_cta(uint64_t* W, const ModTab& t, const Xf& xf, const Of& of, const Prefetch& pf) {
    const uint32_t tid = threadIdx.x;
    uint64_t v[C::E];
    tail_load<C>(tid, W, v, xf);
    if constexpr (ASSUME_OK) {
        const FastArith a = {t.fm};
        inv_rest<C>(tid, W, v, t, a, of, pf);
        return;
    }
    // inverse contract: every word < 2q (ntt.cpp:600-606)
    int bad = 0;
#pragma unroll
    for (int e = 0; e < C::E; ++e) bad |= (v[e] >= t.twoq);
    if (__syncthreads_or(bad) != 0 || !t.inv_fast_ok) {
        inv_exact_cta<C>(W, &t, xf, of, pf);
    } else {
        const FastArith a = {t.fm};
        inv_rest<C>(tid, W, v, t, a, of, pf);
    }
}

// ---------------------------------------------------------------------------
// the persistent kernel skeleton
// ---------------------------------------------------------------------------
// Job concept:
//   uint32_t src_row(item)            first tensor-map row of the item's input
//   const ModTab& mod(item)
//   Xf xf(item), Of of(item)
template <class C, bool FWD, class Job, bool ASSUME_OK = false>
HB_D void ntt_persistent(const CUtensorMap* tmap, const Job& job, uint32_t n_items) {
    // The 128-byte TMA swizzle needs the buffer 1024-byte aligned; the dynamic
    // shared window of a kernel without static shared memory starts aligned.
    uint64_t* W = smem_poly<C>();
    uint64_t* bar = W + C::N;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        if (smem_u32(W) & 1023u) __trap();
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t item = blockIdx.x;
    if (tid == 0 && item < n_items) issue_poly_load<C>(W, tmap, bar, job.src_row(item));
    uint32_t parity = 0;
    for (; item < n_items; item += gridDim.x) {
        const uint32_t next = item + gridDim.x;
        Prefetch pf;
        pf.map = tmap;
        pf.row = (tid == 0 && next < n_items) ? job.src_row(next) : kNoPrefetch;
        mbar_wait(bar, parity);
        parity ^= 1;
        if constexpr (FWD)
            ntt_fwd_cta<C, ASSUME_OK>(W, job.mod(item), job.xf(item), job.of(item), pf);
        else
            ntt_inv_cta<C, ASSUME_OK>(W, job.mod(item), job.xf(item), job.of(item), pf);
    }
}

template <class C>
constexpr size_t ntt_smem_bytes() {
    return (size_t)C::N * 8 + 16 /* mbarrier */;
}

}  // namespace hb
