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
    // small-modulus (q < 2^30) 32-bit path, when available for the shape
    Small32 sm32;
    const Tw32* ftw32;
    const Tw32* itw32;
    uint32_t small_ok;
    uint32_t inv_lazy_ok;   // q < 2^52: the correction-free inverse butterflies may be used
    // FP64-pipe path (modarith.cuh): constants and the {centred root, root / q} tables
    Fp64Mod fd;
    const TwPair* ftwd;
    const TwPair* itwd;
    uint32_t fp64_ok;       // 2^36 <= q <= 2^53 / 3 and the tables above are there
    uint32_t fp64_alt_ok;   // ... and q <= 2^51 (1 + 1/32): forward butterflies that correct every other stage
    uint32_t lazy_out;      // the caller wants the lazy words of the reference's output_mod_factor 4 / 2: exact kernels only
    uint32_t l2_prefetch;   // pull the polynomial after the current one towards L2 while the current one is transformed (option "l2_prefetch")
};

// ---- load transforms (applied to each word as it enters the transform) ----
struct XfIdent {
    static constexpr bool kPost = false;
    static constexpr bool kFp64Out = false;   // the words come out as integers (XfKsConvertFp: as doubles)
    HB_D uint64_t operator()(uint64_t x) const { return x; }
};
// the words in the buffer are doubles already, inside the forward contract (keyswitch S2 behind an S1 that hands
// over doubles): no entry conversion
struct XfIdentFp {
    static constexpr bool kPost = false;
    static constexpr bool kFp64Out = true;
    HB_D uint64_t operator()(uint64_t x) const { return x; }
};
// base conversion of a coefficient-form word to this modulus
// (device/keyswitch/intt1_redu.hpp:36-38)
struct XfReduce {
    static constexpr bool kPost = false;
    static constexpr bool kFp64Out = false;   // the words come out as integers (XfKsConvertFp: as doubles)
    uint64_t q, mu;
    uint32_t skip;   // the words are already inside the transform's input contract (see KsDev::s2_no_reduce)
    HB_D uint64_t operator()(uint64_t x) const { return skip ? x : barrett_reduce64(x, q, mu); }
};
// keyswitch base conversion after the rounding (device/keyswitch/intt2_redu.hpp:
// 24-51): the stage-S4 output is already v = (x + floor(qk/2)) mod qk; here
// out = (v mod qi + fix_i) mod qi.  When qk < 2*qi (same-size primes, the common
// case) "v mod qi" is one conditional subtraction instead of a Barrett product.
struct XfKsConvert {
    static constexpr bool kPost = false;
    static constexpr bool kFp64Out = false;   // the words come out as integers (XfKsConvertFp: as doubles)
    uint64_t q, mu, fix;
    uint32_t small;   // qk < 2*q
    HB_D uint64_t operator()(uint64_t v) const {
        uint64_t r = small ? (v - ((v >= q) ? q : 0)) : barrett_reduce64(v, q, mu);
        r += fix;
        return r - ((r >= q) ? q : 0);
    }
};

// The same conversion for the FP64-pipe forward kernels when qk <= 1.25 q (same-size primes) and
// q <= 2^51 (1 + 1/32): the transform only needs SOME representative of (v + fix) mod q with magnitude below
// 2^52, so the word is converted and the centred fix added -- |x| <= qk + q/2 <= 1.75 q -- and the two
// conditional subtractions, the integer add and the separate entry conversion fall away (kFp64Out).
struct XfKsConvertFp {
    static constexpr bool kPost = false;
    static constexpr bool kFp64Out = true;
    double fix_c;    // centred representative of fix mod q
    HB_D uint64_t operator()(uint64_t v) const { return d2u(fp_add(fp_from_int(v), fix_c)); }
};

// Fused polynomial multiply: the polynomial in shared memory is NTT(a); as the
// first inverse pass pulls its rows into registers every word is multiplied by
// the matching word of NTT(b), read straight from global memory (a row is 128
// contiguous bytes), so the dyadic product never makes an HBM round trip.
// Both factors are variable, so this is the generic 2-by-1 division of the
// dyadic kernel (any modulus, unreduced words tolerated); the products are
// canonical, inside the inverse transform's contract.
struct XfMulGlobal {
    static constexpr bool kPost = true;
    static constexpr bool kFp64Out = false;   // the words come out as integers (XfKsConvertFp: as doubles)
    const uint64_t* other;   // NTT(b) of this item, same (bit-reversed) order
    const uint64_t* next;    // NTT(b) of the item this CTA transforms next (L2 prefetch), or nullptr
    Divisor dv;
    HB_D uint64_t operator()(uint64_t x) const { return x; }
    HB_D uint64_t mul(uint64_t x, uint64_t y) const {
        if (((x | y) >> 32) >= (dv.q >> 32)) {   // rarely: operands not obviously below q
            x = mod64(x, dv);
            y = mod64(y, dv);
        }
        return mulmod_preshifted(x << dv.s, y, dv);
    }
    template <class C>
    HB_D void post(uint32_t tid, uint64_t* v) const;
};

// ---- output functors ----
// Shared-memory plan of a transform CTA:
//   [ W: N words, the polynomial ][ staging: one 4 KiB slice per warp ][ mbarrier ]
// The staging slices exist when they fit (kStagedStore): forward results then
// leave through per-warp TMA tensor stores (coalesced by the copy engine)
// instead of 16-byte stores at a 128-byte lane stride.
template <class C>
struct SmemPlan {
    static constexpr bool kStagedStore = ((size_t)C::N * 8 + (size_t)(C::NT / 32) * 4096 + 64) <= 227u * 1024u;
    static constexpr uint32_t STAGE_WORDS = kStagedStore ? (C::NT / 32) * 512 : 0;
    static constexpr uint32_t BAR_WORD = C::N + STAGE_WORDS;
    static constexpr uint32_t FLAG_WORD = BAR_WORD + 1;   // range-vote flag (fast-vote kernels)
    static constexpr uint32_t KQ_WORD = BAR_WORD + 2;     // two tables of k*q, k < 64 (iteration parity)
    static constexpr uint32_t CNT_WORD = BAR_WORD + 2 + 128;   // warps done with the buffer (C::WARPTAIL)
    static constexpr uint32_t TMEM_WORD = CNT_WORD + 1;        // tensor-memory base address (kernels that allocate TMEM)
    static constexpr uint32_t WBAR_WORD = CNT_WORD + 2;        // one mbarrier per warp (epilogues that TMA-load into the warp's slice)
    static constexpr uint32_t END_WORD = WBAR_WORD + C::NT / 32;
    static constexpr size_t BYTES = (size_t)END_WORD * 8;
    // head-pass twiddles of the FP64 plain kernels (Fp64ArithS), 16-byte entries behind everything else
    static constexpr uint32_t HEAD_TW_FWD = C::fwd_off(C::NP), HEAD_TW_INV = C::INV_ENTRIES - C::N;
    static constexpr uint32_t HEAD_TW = HEAD_TW_FWD > HEAD_TW_INV ? HEAD_TW_FWD : HEAD_TW_INV;
    static constexpr uint32_t TW_WORD = (END_WORD + 1) & ~1u;
    static_assert(TW_WORD >= END_WORD, "shared-memory plan: the twiddle area overlaps the control words");
    static constexpr size_t BYTES_TW = (size_t)TW_WORD * 8 + (size_t)HEAD_TW * 16;
    static constexpr bool kHeadTwFits = BYTES_TW <= 227u * 1024u;
};

// forward output: the 16 contiguous words of one row per thread
struct OfRows {
    uint64_t* dst;               // direct path: polynomial base in global memory
    const CUtensorMap* smap;     // staged path: store map (box 16 words x 32 rows, 128B swizzle)
    uint32_t row0;               //   first tensor-map row of the output polynomial
    // PREPARED: prepare() has run since the warp's previous store (the wait for the staging slice then sits in front
    // of the row's final conversion, and the converted words go straight into the 16-byte stores: with the wait
    // between them the compiler gathered every store's four registers with moves, 64 per transform)
    template <class C, bool PREPARED = false>
    HB_D void store(uint32_t row, const uint64_t* v) const;
    template <class C>
    HB_D void prepare() const;
    template <class C>
    HB_D void prefetch(uint32_t) const {}
};
// output functors that want a call in front of a row's final conversion
template <class Of, class C, class = void>
struct HasPrepare : std::false_type {};
template <class Of, class C>
struct HasPrepare<Of, C, std::void_t<decltype(std::declval<const Of&>().template prepare<C>())>> : std::true_type {};
struct OfWords {  // inverse: one word at its natural index (coalesced along lo)
    uint64_t* dst;
    HB_D void word(uint32_t idx, uint64_t x) const { dst[idx] = x; }
};
struct OfWordsD {  // the same for a kernel that produces canonical integers where the consumer expects doubles
    uint64_t* dst;
    HB_D void word(uint32_t idx, uint64_t x) const { dst[idx] = d2u(fp_from_int(x)); }
};
// inverse output with the keyswitch rounding v = (x + floor(qk/2)) mod qk
// (device/keyswitch/intt2_redu.hpp:24-25,43) applied once per special-prime word
struct OfWordsRound {
    uint64_t* dst;
    uint64_t qk, half;
    HB_D void word(uint32_t idx, uint64_t x) const {
        const uint64_t v = x + half;
        dst[idx] = v - ((v >= qk) ? qk : 0);
    }
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

// the same box towards L2 only (no shared-memory destination, no barrier)
HB_D void tma_prefetch_rows(const CUtensorMap* map, uint32_t row) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"((uint64_t)map), "r"(0), "r"(row)
                 : "memory");
}

// 2-D tensor TMA store of one staged box back to global memory (bulk async group)
HB_D void tma_store_rows(const CUtensorMap* map, const void* smem_src, uint32_t row) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map),
                 "r"(smem_u32(smem_src)), "r"(0), "r"(row)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// 3-D variant for the small-modulus forward epilogue: the map views the output as
// [row of 32 words][half j][16 words]; one box = half j of 32 consecutive rows
HB_D void tma_store_rows3(const CUtensorMap* map, const void* smem_src, uint32_t half, uint32_t row) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)map),
                 "r"(smem_u32(smem_src)), "r"(0), "r"(half), "r"(row)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all earlier bulk stores of this thread have finished READING shared memory
HB_D void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------------------
// tensor memory (TMEM: 512 columns x 128 lanes x 32 bit per SM) as a plain per-thread scratchpad, see tmem.cuh
// ---------------------------------------------------------------------------
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


// Split load: issue now, use after tmem_ld_wait16 on the same registers.  The wait instruction covers every
// tcgen05.ld the thread has issued so far; passing the destination registers through it ("+r") is what keeps
// the compiler from scheduling a consumer between the load and the wait.
HB_D void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
HB_D void tmem_ld_wait16(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// packed twiddle k (0..3) of a 16-register block: {w, wp} = columns 4k .. 4k+3
HB_D TwPair tmem_pair(const uint32_t* r, int k) {
    TwPair t;
    t.w = ((uint64_t)r[4 * k + 1] << 32) | r[4 * k];
    t.wp = ((uint64_t)r[4 * k + 3] << 32) | r[4 * k + 2];
    return t;
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
        if (row != kNoPrefetch && (!C::WARPTAIL || threadIdx.x == 0)) {
            uint64_t* W = smem_poly<C>();
            fence_proxy_async();
            issue_poly_load<C>(W, map, W + SmemPlan<C>::BAR_WORD, row);
        }
    }
    // C::WARPTAIL: no block barrier before the prefetch.  Every warp reports that its words have
    // left the buffer; the warp that completes the count (it never resets: NT/32 is a power of two)
    // starts the copy.  `row` is valid in lane 0 of every warp.
    template <class C>
    HB_D void issue_when_all_warps_done() const {
        // the copy engine will overwrite words this warp wrote in the passes before: order every lane's
        // generic-proxy stores in front of the async proxy before the warp reports (the issuing lane's own
        // fence below covers only what that thread has observed)
        fence_proxy_async();
        __syncwarp();
        if ((threadIdx.x & 31u) == 0) {
            uint64_t* W = smem_poly<C>();
            uint32_t old;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;"
                         : "=r"(old)
                         : "r"(smem_u32(W + SmemPlan<C>::CNT_WORD))
                         : "memory");
            if ((old & (C::NT / 32 - 1)) == C::NT / 32 - 1 && row != kNoPrefetch) {
                fence_proxy_async();
                issue_poly_load<C>(W, map, W + SmemPlan<C>::BAR_WORD, row);
            }
        }
    }
};

// A thread's row of NTT(b) is 128 contiguous bytes, so reading it directly makes every 16-byte
// load of a warp touch 32 different lines, each of which is fetched eight times (the lines of 16
// warps do not survive in the L1 left next to the buffer).  The warp's 32 rows are 4 KiB contiguous:
// they are loaded with coalesced 512-byte accesses and transposed through the warp's staging slice
// (which the inverse kernels do not use otherwise), same chunk swizzle as the polynomial buffer.
template <class C>
HB_D void XfMulGlobal::post(uint32_t tid, uint64_t* v) const {
    static_assert(C::ROW == 16, "64-bit rows");
    if constexpr (SmemPlan<C>::kStagedStore) {
        const uint32_t lane = tid & 31u;
        uint64_t* slice = smem_poly<C>() + C::N + (tid >> 5) * 512;
#pragma unroll
        for (int ri = 0; ri < C::E / 16; ++ri) {
            const uint32_t row0 = tail_row<C>(tid, ri) - lane;
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(other + (size_t)row0 * 16);
            ulonglong2 t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = __ldg(src + i * 32 + lane);
            // the same rows of the next item towards L2, a whole transform ahead (one line per lane)
            if (next) asm volatile("prefetch.global.L2 [%0];" ::"l"(next + (size_t)row0 * 16 + lane * 16));
            __syncwarp();                      // the slice's previous contents have been read
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t j = (uint32_t)i * 32 + lane, r = j >> 3, c = j & 7u;
                st2(slice + r * 16 + ((c ^ (r & 7u)) << 1), t[i].x, t[i].y);
            }
            __syncwarp();
            // one test for the whole row: with the (rare) unreduced-operand branch inside every product
            // the 16 products end up in separate basic blocks and run one dependent chain at a time
            uint64_t y[16];
            uint32_t top = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                ld2(slice + lane * 16 + (((uint32_t)c ^ (lane & 7u)) << 1), y[2 * c], y[2 * c + 1]);
                top |= (uint32_t)((v[ri * 16 + 2 * c] | y[2 * c] | v[ri * 16 + 2 * c + 1] | y[2 * c + 1]) >> 32);
            }
            if (top < (uint32_t)(dv.q >> 32)) {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[ri * 16 + e] = mulmod_preshifted(v[ri * 16 + e] << dv.s, y[e], dv);
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[ri * 16 + e] = mul(v[ri * 16 + e], y[e]);
            }
        }
    } else {
#pragma unroll
        for (int ri = 0; ri < C::E / 16; ++ri) {
            const uint64_t* p = other + (size_t)tail_row<C>(tid, ri) * 16;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const ulonglong2 y = __ldg(reinterpret_cast<const ulonglong2*>(p + 2 * c));
                v[ri * 16 + 2 * c] = mul(v[ri * 16 + 2 * c], y.x);
                v[ri * 16 + 2 * c + 1] = mul(v[ri * 16 + 2 * c + 1], y.y);
            }
        }
    }
}

template <class C>
HB_D void OfRows::prepare() const {
    if constexpr (SmemPlan<C>::kStagedStore) {
        if ((threadIdx.x & 31u) == 0) tma_store_wait_read();   // previous box of this slice has left
        __syncwarp();
    }
}
template <class C, bool PREPARED>
HB_D void OfRows::store(uint32_t row, const uint64_t* v) const {
    if constexpr (SmemPlan<C>::kStagedStore) {
        // rows of a warp are consecutive: stage them in the warp's slice (same
        // chunk swizzle as the polynomial buffer) and let one lane hand the
        // 4 KiB box to the TMA store engine
        const uint32_t lane = threadIdx.x & 31u;
        uint64_t* slice = smem_poly<C>() + C::N + (threadIdx.x >> 5) * 512;
        if constexpr (!PREPARED) prepare<C>();
#pragma unroll
        for (int c = 0; c < 8; ++c)
            st2(slice + lane * 16 + (((uint32_t)c ^ (lane & 7u)) << 1), v[2 * c], v[2 * c + 1]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) tma_store_rows(smap, slice, row0 + row);
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) st2(dst + row * 16 + 2 * c, v[2 * c], v[2 * c + 1]);
    }
}

// ---------------------------------------------------------------------------
// forward transform of the polynomial in W
// ---------------------------------------------------------------------------
template <class C, int P, class A>
HB_D void fwd_mid_passes(uint32_t tid, uint64_t* W, const TwPair* tw, const A& a) {
    if constexpr (P < C::NP) {
        fwd_head_pass<C, P>(tid, W, tw, a);
        if constexpr (C::WARPTAIL && P == C::NP - 1) __syncwarp();   // the tail reads what this warp wrote
        else __syncthreads();
        fwd_mid_passes<C, P + 1>(tid, W, tw, a);
    }
}

// How a kernel treats the input contract (modarith.cuh):
//   kFastVote   fast arithmetic; the range check rides on the first barrier and
//               an out-of-contract polynomial is left untouched and appended to
//               the deferred list for the exact kernel
//   kFastTrust  fast arithmetic, inputs are in range by construction (internal
//               keyswitch stages, whose load transform reduces every word)
//   kExactAll   reference op sequence for every item
//   kExactList  reference op sequence for the items of the deferred list
enum NttMode { kFastVote = 0, kFastTrust = 1, kExactAll = 2, kExactList = 3 };

// deferred list: word 0 = count, word 1 = reserved (zero), words 2.. = item indices; word 0 is zero when a voting
// kernel starts
constexpr uint32_t kListHead = 2;
HB_D void defer_item(uint32_t* list, uint32_t item) {
    const uint32_t slot = atomicAdd(list, 1u);
    list[kListHead + slot] = item;
}
// Range vote on the freshly loaded words: nonzero when some word may be >=
// bound.  For bounds above 2^32 only the high words are compared (one max per
// word): conservative -- words in [hi32(bound)*2^32, bound) are flagged too and
// merely take the exact kernel, which is always right -- and canonical inputs
// (< q, and bound >= 2q) never trip it.
template <int E>
HB_D int out_of_range(const uint64_t* v, uint64_t bound) {
    const uint32_t bh = (uint32_t)(bound >> 32);
    if (bh != 0) {
        uint32_t m = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) m = max(m, (uint32_t)(v[e] >> 32));
        return m >= bh;
    }
    int bad = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) bad |= (v[e] >= bound);
    return bad;
}

// Block-wide "some word out of range?" without a reduction barrier: warps that
// saw a bad word raise a shared flag before the pass barrier that is there
// anyway; everybody reads the flag after it.  Returns true (after clearing the
// flag for the next polynomial) when the polynomial must be deferred.
template <class C>
HB_D void vote_raise(int bad) {
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31u) == 0) smem_poly<C>()[SmemPlan<C>::FLAG_WORD] = 1;
}
template <class C>
HB_D bool vote_read() {
    uint64_t* flag = smem_poly<C>() + SmemPlan<C>::FLAG_WORD;
    if (*flag == 0) return false;
    __syncthreads();                 // everyone has seen it
    if (threadIdx.x == 0) *flag = 0;
    return true;
}

// ---------------------------------------------------------------------------
// tail-pass twiddles in tensor memory (arithmetic policies with kTmemTail, ntt_core.cuh)
// ---------------------------------------------------------------------------
// Layout in a thread's lane: row ri, packed slot s (1..15, slot 0 is padding) at columns (ri * 16 + s) * 4 ..+3
// = {w lo, w hi, w' lo, w' hi}.  Compile-time switch for A/B measurements.
#ifndef HB_TMEM_TAIL
#define HB_TMEM_TAIL 1
#endif
template <class C>
struct TmemTail {
    static constexpr uint32_t ROWS = C::E / C::ROW;
    static constexpr uint32_t COLS = ROWS * C::ROW * 4;                  // per thread
    static constexpr uint32_t TOTAL = ((C::NT / 32 + 3) / 4) * COLS;     // four warps share a lane quarter
    static constexpr bool kFits = C::LOGROW == 4 && TOTAL <= 512;
};
// this thread's first column
template <class C>
HB_D uint32_t tmem_tail_addr(uint32_t tmem_base) {
    return tmem_thread_addr(tmem_base, (threadIdx.x >> 7) * TmemTail<C>::COLS);
}
// once per launch: every thread parks the twiddles of its own tail rows; `tail` = first tail entry of the
// packed table (forward: ftwd + fwd_off(NP), inverse: itwd)
template <class C>
HB_D void tail_tw_to_tmem(uint32_t tid, const TwPair* tail, uint32_t ttail) {
#pragma unroll
    for (int ri = 0; ri < (int)TmemTail<C>::ROWS; ++ri) {
        const TwPair* src = tail + tail_tw_base<C>(tail_row<C>(tid, ri));
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
            uint64_t w[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int slot = blk * 4 + k;
                TwPair p = {0, 0};
                if (slot != 0) p = ldpair(src + 32 * slot);
                w[2 * k] = p.w;
                w[2 * k + 1] = p.wp;
            }
            tmem_st16(ttail + (uint32_t)(ri * 64 + blk * 16), w);
        }
    }
    tmem_wait_st();
}
// no instruction: the registers of a second load become "defined here" for the compiler (placed right
// behind the tmem_ld_wait16 that covers both loads)
HB_D void tmem_touch16(uint32_t* r) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}

// fwd_tail_compute (ntt_core.cuh) with the twiddles read from tensor memory, four packed slots per load, each
// load issued one butterfly stage ahead of its use.  Every load has its own destination registers (the arrays
// are fully scalarised; at most two or three are alive at a time): re-using two buffers in turns made the
// compiler copy twiddles out of the way of the next load, ~100 moves per transform.
template <class C, class A, class F, class G>
HB_D void fwd_tail_compute_tmem(uint64_t* v, const A& a, const F& after_row, const G& before_final) {
    constexpr int ROWS = (int)TmemTail<C>::ROWS, S0 = C::HEAD;
    uint32_t r[ROWS][4][16];              // [row][slots 4g .. 4g+3]
    tmem_ld16_issue(a.ttail, r[0][0]);
    tmem_ld16_issue(a.ttail + 16, r[0][1]);
    static_for<0, ROWS>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        uint64_t* x = v + ri * 16;
        const uint32_t ta = a.ttail + (uint32_t)ri * 64u;
        tmem_ld_wait16(r[ri][0]);
        tmem_touch16(r[ri][1]);
        {   // stage 0: slot 1, pairs (j, j + 8)
            const TwPair t = tmem_pair(r[ri][0], 1);
            static_for<0, 8>([&](auto jc) { a.template fwd_at<S0>(x[decltype(jc)::value], x[decltype(jc)::value + 8], t); });
        }
        static_for<0, 2>([&](auto bc) {   // stage 1: slots 2, 3
            constexpr int blk = decltype(bc)::value;
            const TwPair t = tmem_pair(r[ri][0], 2 + blk);
            static_for<0, 4>([&](auto jc) {
                a.template fwd_at<S0 + 1>(x[blk * 8 + decltype(jc)::value], x[blk * 8 + decltype(jc)::value + 4], t);
            });
        });
        tmem_ld16_issue(ta + 32, r[ri][2]);     // slots 8..11
        static_for<0, 4>([&](auto bc) {   // stage 2: slots 4..7
            constexpr int blk = decltype(bc)::value;
            const TwPair t = tmem_pair(r[ri][1], blk);
            static_for<0, 2>([&](auto jc) {
                a.template fwd_at<S0 + 2>(x[blk * 4 + decltype(jc)::value], x[blk * 4 + decltype(jc)::value + 2], t);
            });
        });
        tmem_ld_wait16(r[ri][2]);
        tmem_ld16_issue(ta + 48, r[ri][3]);     // slots 12..15
        static_for<0, 4>([&](auto bc) {   // stage 3, first half: slots 8..11
            constexpr int blk = decltype(bc)::value;
            a.template fwd_at<S0 + 3>(x[blk * 2], x[blk * 2 + 1], tmem_pair(r[ri][2], blk));
        });
        tmem_ld_wait16(r[ri][3]);
        if constexpr (ri + 1 < ROWS) tmem_ld16_issue(ta + 64, r[ri + 1 < ROWS ? ri + 1 : ri][0]);        // next row, slots 0..3
        static_for<0, 4>([&](auto bc) {   // stage 3, second half: slots 12..15
            constexpr int blk = decltype(bc)::value;
            a.template fwd_at<S0 + 3>(x[8 + blk * 2], x[8 + blk * 2 + 1], tmem_pair(r[ri][3], blk));
        });
        if constexpr (ri + 1 < ROWS) tmem_ld16_issue(ta + 64 + 16, r[ri + 1 < ROWS ? ri + 1 : ri][1]);   // next row, slots 4..7
        before_final(ri);
        static_for<0, C::ROW>([&](auto kc) { x[decltype(kc)::value] = a.fwd_final(x[decltype(kc)::value]); });
        after_row(ri);
    });
}

// inv_tail_compute with the twiddles from tensor memory (stage d uses slots 2^(3-d) + blk)
template <class C, class A>
HB_D void inv_tail_compute_tmem(uint64_t* v, const A& a) {
    constexpr int ROWS = (int)TmemTail<C>::ROWS;
    uint32_t r[ROWS][4][16];              // [row][slots 4g .. 4g+3]
    tmem_ld16_issue(a.ttail + 32, r[0][2]);
    tmem_ld16_issue(a.ttail + 48, r[0][3]);
    static_for<0, ROWS>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        uint64_t* x = v + ri * 16;
        const uint32_t ta = a.ttail + (uint32_t)ri * 64u;
        tmem_ld_wait16(r[ri][2]);
        tmem_touch16(r[ri][3]);
        static_for<0, 4>([&](auto bc) {   // stage 0, first half: slots 8..11
            constexpr int blk = decltype(bc)::value;
            a.template inv_at<0>(x[blk * 2], x[blk * 2 + 1], tmem_pair(r[ri][2], blk));
        });
        tmem_ld16_issue(ta + 16, r[ri][1]);     // slots 4..7
        static_for<0, 4>([&](auto bc) {   // stage 0, second half: slots 12..15
            constexpr int blk = decltype(bc)::value;
            a.template inv_at<0>(x[8 + blk * 2], x[8 + blk * 2 + 1], tmem_pair(r[ri][3], blk));
        });
        tmem_ld_wait16(r[ri][1]);
        tmem_ld16_issue(ta, r[ri][0]);          // slots 0..3
        static_for<0, 4>([&](auto bc) {   // stage 1: slots 4..7
            constexpr int blk = decltype(bc)::value;
            const TwPair t = tmem_pair(r[ri][1], blk);
            static_for<0, 2>([&](auto jc) {
                a.template inv_at<1>(x[blk * 4 + decltype(jc)::value], x[blk * 4 + decltype(jc)::value + 2], t);
            });
        });
        tmem_ld_wait16(r[ri][0]);
        if constexpr (ri + 1 < ROWS) tmem_ld16_issue(ta + 64 + 32, r[ri + 1 < ROWS ? ri + 1 : ri][2]);   // next row, slots 8..11
        static_for<0, 2>([&](auto bc) {   // stage 2: slots 2, 3
            constexpr int blk = decltype(bc)::value;
            const TwPair t = tmem_pair(r[ri][0], 2 + blk);
            static_for<0, 4>([&](auto jc) {
                a.template inv_at<2>(x[blk * 8 + decltype(jc)::value], x[blk * 8 + decltype(jc)::value + 4], t);
            });
        });
        {   // stage 3: slot 1
            const TwPair t = tmem_pair(r[ri][0], 1);
            static_for<0, 8>([&](auto jc) { a.template inv_at<3>(x[decltype(jc)::value], x[decltype(jc)::value + 8], t); });
        }
        if constexpr (ri + 1 < ROWS) tmem_ld16_issue(ta + 64 + 48, r[ri + 1 < ROWS ? ri + 1 : ri][3]);   // next row, slots 12..15
    });
}

// returns false when the polynomial was deferred (fast-vote mode only); PF_ON_FAIL = false: the buffer is then
// left free (nothing prefetched into it) because the caller re-loads the same polynomial for the exact pass
template <class C, int MODE, bool PF_ON_FAIL = true, class A, class Xf, class Of>
HB_D bool ntt_fwd_cta(uint64_t* W, const ModTab& t, const A& a, const Xf& xf, const Of& of, const Prefetch& pf) {
    using P0 = FwdPass<C, 0>;
    const uint32_t tid = threadIdx.x;
    uint64_t v[C::E];
    head_load<C, P0::R, P0::LS>(tid, W, v, xf);
    of.template prefetch<C>(tid);     // epilogue operands (keyswitch) towards L2 early
    int bad = 0;
    if constexpr (MODE == kFastVote) {
        // forward contract: every word < 4q (tests/test_utils/ntt.cpp:483-486); the FP64
        // arithmetic takes words below 1.25 q, the rest goes to the exact kernel
        bad = out_of_range<C::E>(v, A::kFp64 ? t.fd.vote : t.fm.q4);
    }
    if constexpr (MODE == kFastVote) vote_raise<C>(bad);
    const TwPair* ftw = A::kFp64 ? t.ftwd : t.ftw;
    const TwPair* ftw_head = ftw;            // twiddles of the head passes
    if constexpr (A::kSmemHead) ftw_head = a.fwd_base();
    if constexpr (A::kFp64 && !Xf::kFp64Out) {
#pragma unroll
        for (int e = 0; e < C::E; ++e) v[e] = a.enter_fwd(v[e]);
    }
    fwd_head_compute<C, 0>(tid, v, ftw_head, a, [&](int gi, int k0, int k1) {
        head_store_word<C, P0::R, P0::LS>(tid, W, gi, k0, v[gi * (1 << P0::R) + k0]);
        head_store_word<C, P0::R, P0::LS>(tid, W, gi, k1, v[gi * (1 << P0::R) + k1]);
    });
    __syncthreads();
    if constexpr (MODE == kFastVote) {
        if (vote_read<C>()) {          // global memory still holds the untouched input
            if constexpr (PF_ON_FAIL) pf.template issue<C>();
            return false;
        }
    }
    fwd_mid_passes<C, 1>(tid, W, ftw_head, a);
    tail_load<C>(tid, W, v, XfIdent());
    if constexpr (C::WARPTAIL) {
        pf.template issue_when_all_warps_done<C>();
    } else {
        __syncthreads();        // every word of W is in registers now
        pf.template issue<C>(); // ... so the buffer can take the next polynomial
    }
    // each row leaves as soon as it is final: its staged TMA store drains while the next row is computed
    if constexpr (A::kTmemTail)
    {
        if constexpr (HasPrepare<Of, C>::value)
            fwd_tail_compute_tmem<C>(
                v, a, [&](int ri) { of.template store<C, true>(tail_row<C>(tid, ri), v + ri * 16); },
                [&](int) { of.template prepare<C>(); });
        else
            fwd_tail_compute_tmem<C>(
                v, a, [&](int ri) { of.template store<C>(tail_row<C>(tid, ri), v + ri * 16); }, [](int) {});
    }
    else
        fwd_tail_compute<C>(tid, v, ftw, a, [&](int ri) { of.template store<C>(tail_row<C>(tid, ri), v + ri * 16); });
    return true;
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

template <class C, int MODE, bool PF_ON_FAIL = true, class A, class Xf, class Of>
HB_D bool ntt_inv_cta(uint64_t* W, const ModTab& t, const A& a, const Xf& xf, const Of& of, const Prefetch& pf) {
    using PL = InvPass<C, C::NP - 1>;
    const uint32_t tid = threadIdx.x;
    uint64_t v[C::E];
    tail_load<C>(tid, W, v, xf);
    if constexpr (Xf::kPost) xf.template post<C>(tid, v);
    int bad = 0;
    if constexpr (MODE == kFastVote) {
        // inverse contract: every word < 2q (ntt.cpp:600-606); FP64 arithmetic: below 1.25 q
        bad = out_of_range<C::E>(v, A::kFp64 ? t.fd.vote : t.twoq);
    }
    const TwPair* itw = A::kFp64 ? t.itwd : t.itw;
    const TwPair* itw_head = itw;            // twiddles of the head passes
    if constexpr (A::kSmemHead) itw_head = a.template inv_base<C>();
    if constexpr (A::kFp64) {
#pragma unroll
        for (int e = 0; e < C::E; ++e) v[e] = a.enter_inv(v[e]);
    }
    if constexpr (A::kTmemTail) inv_tail_compute_tmem<C>(v, a);
    else inv_tail_compute<C>(tid, v, itw, a);
    tail_store<C>(tid, W, v);
    if constexpr (C::WARPTAIL) {
        // the first head pass stays inside the words this warp has just written: no block barrier
        // before it, and the range vote rides on the one after it
        static_assert(C::NP >= 2, "WARPTAIL inverse needs a warp-local head pass");
        __syncwarp();
        inv_head_pass<C, 0>(tid, W, itw_head, a);
        if constexpr (MODE == kFastVote) vote_raise<C>(bad);
        __syncthreads();
        if constexpr (MODE == kFastVote) {
            if (vote_read<C>()) {
                if constexpr (PF_ON_FAIL) pf.template issue<C>();
                return false;
            }
        }
        inv_mid_passes<C, 1>(tid, W, itw_head, a);
        head_load<C, PL::R, PL::LS>(tid, W, v, XfIdent());
        pf.template issue_when_all_warps_done<C>();
    } else {
        if constexpr (MODE == kFastVote) vote_raise<C>(bad);   // right before the barrier it rides on
        __syncthreads();
        if constexpr (MODE == kFastVote) {
            if (vote_read<C>()) {
                if constexpr (PF_ON_FAIL) pf.template issue<C>();
                return false;
            }
        }
        inv_mid_passes<C, 0>(tid, W, itw_head, a);
        head_load<C, PL::R, PL::LS>(tid, W, v, XfIdent());
        __syncthreads();
        pf.template issue<C>();
    }
    // every pair is stored as soon as its last butterfly has made it final: the stores (32 B/clk
    // per SM at most) drain under the remaining butterflies instead of as one burst at the end
    inv_head_compute<C, C::NP - 1>(tid, v, itw_head, a, [&](int gi, int k0, int k1) {
        of.word(inv_last_index<C>(tid, gi, k0), v[gi * (1 << PL::R) + k0]);
        of.word(inv_last_index<C>(tid, gi, k1), v[gi * (1 << PL::R) + k1]);
    });
    return true;
}

// ---------------------------------------------------------------------------
// the persistent kernel skeleton
// ---------------------------------------------------------------------------
// Job concept:
//   uint32_t src_row(item)            first tensor-map row of the item's input
//   const ModTab& mod(item)
//   Xf xf(item), Of of(item, store_map)
// `list`: the deferred list (written in kFastVote, read in kExactList mode).
// FP64: 0 integer butterflies, 1 FP64-pipe butterflies, 2 (forward only) FP64 with a full correction every other stage,
//       3 (forward only) as 2 with the raw doubles as output (|v| <= 1.92 q, any representative of the residue)
// Plain forward FP64 voting kernels (kFoldExact): a polynomial whose vote fails is transformed with the
// reference op sequence by the same CTA at once (re-loaded, the first pass has overwritten the buffer), so
// these kernels need neither the deferred list nor the separate pass over it.  That pass costs 6 - 8 us per
// call (2 % of a 4096-polynomial call) although the list is empty in normal use, because the next call's
// kernels queue behind it.  Measured per 4096-polynomial forward call: 346.1 us with the separate pass, 342.5
// like this, 337.8 voting without any exact path (the extra code costs the main loop ~1.4 %), 332.5 without a
// vote.  A first version kept the list and ran the pass at the end of the kernel behind a grid barrier: same
// time, plus a residency assumption.  In the inverse kernels the extra code costs the main loop as much as
// the launch saves (HB_FOLD_INV = 1: 390 us per call either way, 384 voting without any exact path), so they
// keep the list.
// ... and alternating with other kernels, as every real caller does (bench.py: forward / inverse in turns), the
// forward call with the exact path inside takes 358.5 us instead of 344.8 with the separate pass, although
// the same call repeated back to back takes 343.4 instead of 346.4 (tools/time_pattern.py): the kernel is
// 40 KB larger and comes back into a cold instruction cache every time.  Both are therefore OFF; the code
// stays for the measurement (make CFLAGS_EXTRA="-DHB_FOLD_FWD=1").
#ifndef HB_FOLD_FWD
#define HB_FOLD_FWD 0
#endif
#ifndef HB_FOLD_INV
#define HB_FOLD_INV 0
#endif
// jobs whose load transform reads a second operand (JobInvMul::kPostXf) keep the list
template <class J, class = void>
struct JobPostXf : std::false_type {};
template <class J>
struct JobPostXf<J, std::void_t<decltype(J::kPostXf)>> : std::integral_constant<bool, J::kPostXf> {};
template <class C, bool FWD, int MODE, int FP64, class Job>
constexpr bool kFoldExact = (FWD ? HB_FOLD_FWD != 0 : HB_FOLD_INV != 0) && MODE == kFastVote && FP64 != 0 && Job::kOneModulus && !JobPostXf<Job>::value;
template <class C, bool FWD, int MODE, class Job, bool LAZY = false, int FP64 = 0>
HB_D void ntt_persistent(const CUtensorMap* tmap, const CUtensorMap* smap, const Job& job, uint32_t n_items,
                         uint32_t* list) {
    static_assert(!LAZY || (!FWD && (MODE == kFastVote || MODE == kFastTrust)), "LAZY is an inverse fast-path option");
    static_assert(FP64 == 0 || (!LAZY && (MODE == kFastVote || MODE == kFastTrust)), "FP64 is a fast-path option");
    // The 128-byte TMA swizzle needs the buffer 1024-byte aligned; the dynamic
    // shared window of a kernel without static shared memory starts aligned.
    // launched with programmatic stream serialization (ntt_launch.cuh): wait for the kernel in front
    // of us (a no-op otherwise), then let the one behind us be scheduled
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    constexpr bool FOLD = kFoldExact<C, FWD, MODE, FP64, Job>;
    uint64_t* W = smem_poly<C>();
    uint64_t* bar = W + SmemPlan<C>::BAR_WORD;
    const uint32_t tid = threadIdx.x;
    if constexpr (MODE == kExactList) n_items = list[0];
    auto item_of = [&](uint32_t i) -> uint32_t {
        if constexpr (MODE == kExactList) return list[kListHead + i];
        else return job.order(i);      // the order in which the grid walks over the items (identity, or modulus-major)
    };
    if (tid == 0) {
        if (smem_u32(W) & 1023u) __trap();
        mbar_init(bar, 1);
        W[SmemPlan<C>::FLAG_WORD] = 0;
        W[SmemPlan<C>::CNT_WORD] = 0;
        for (uint32_t w = 0; w < C::NT / 32; ++w) mbar_init(W + SmemPlan<C>::WBAR_WORD + w, 1);
        fence_barrier_init();
    }
    // FP64 kernels whose CTAs keep a modulus for long runs of items (the plain batched calls: one modulus per
    // launch; the keyswitch stages: items dealt out modulus-major, Job::order), launched with BYTES_TW of shared
    // memory: the twiddles of the head passes sit next to the buffer, those of the tail pass in tensor memory
    // (one CTA per SM at this shape: the whole TMEM), both (re)loaded when the modulus changes
    constexpr bool SMEM_HEAD = FP64 != 0 && (Job::kOneModulus || Job::kModulusRuns) && SmemPlan<C>::kHeadTwFits;
    constexpr bool TMEM_TAIL = SMEM_HEAD && HB_TMEM_TAIL != 0 && C::LOGN == 14 && TmemTail<C>::kFits;
    if constexpr (TMEM_TAIL) {
        if (tid < 32) tmem_alloc_all(reinterpret_cast<uint32_t*>(W + SmemPlan<C>::TMEM_WORD));
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    uint32_t tmem_base = 0, ttail = 0;
    if constexpr (TMEM_TAIL) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_base = *reinterpret_cast<volatile uint32_t*>(W + SmemPlan<C>::TMEM_WORD);
        ttail = tmem_tail_addr<C>(tmem_base);
    }
    auto load_twiddles = [&](const ModTab& tm) {
        if constexpr (SMEM_HEAD) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(FWD ? tm.ftwd : tm.itwd + C::inv_off(0));
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(W + SmemPlan<C>::TW_WORD);
            constexpr uint32_t COUNT = FWD ? SmemPlan<C>::HEAD_TW_FWD : SmemPlan<C>::HEAD_TW_INV;
            for (uint32_t e = tid; e < COUNT; e += C::NT) dst[e] = src[e];
        }
        if constexpr (TMEM_TAIL) tail_tw_to_tmem<C>(tid, FWD ? tm.ftwd + C::fwd_off(C::NP) : tm.itwd, ttail);
    };
    // How the grid walks over the items.  Strided (CTA c takes c, c + grid, ...) by default.  Jobs that walk
    // modulus-major (Job::kModulusRuns: the keyswitch stages) give every CTA ONE contiguous stretch of the order
    // instead: it then crosses a modulus boundary once or twice per launch, not once per modulus, and the 262 KiB
    // of twiddles it keeps in shared / tensor memory are reloaded that rarely (same makespan: ceil(items / grid)).
    bool blocked = false;
    if constexpr (Job::kModulusRuns && MODE != kExactList) blocked = job.ks.walk_blocked != 0;
    uint32_t i_first = blockIdx.x, i_end = n_items, i_step = gridDim.x;
    if (blocked) {
        const uint32_t per = (n_items + gridDim.x - 1) / gridDim.x;
        i_first = blockIdx.x * per;
        i_end = i_first + per < n_items ? i_first + per : n_items;
        if (i_first > n_items) i_first = n_items;
        i_step = 1;
    }
    // the first polynomial is on its way while the twiddles are brought in
    if (tid == 0 && i_first < i_end) issue_poly_load<C>(W, tmap, bar, job.src_row(item_of(i_first)));
    const TwPair* cur_tw = nullptr;      // whose twiddles are resident
    if constexpr (SMEM_HEAD) {
        const ModTab& t0 = job.mod(i_first < n_items ? item_of(i_first) : 0);
        load_twiddles(t0);
        cur_tw = FWD ? t0.ftwd : t0.itwd;
        __syncthreads();
    }
    uint32_t head_s = 0;
    if constexpr (SMEM_HEAD) {
        // volatile: the address (and with it every load of these twiddles) stays behind the barrier above
        asm volatile("mov.u32 %0, %1;" : "=r"(head_s) : "r"(smem_u32(W + SmemPlan<C>::TW_WORD)));
    }
    uint32_t i = i_first;
    uint32_t parity = 0;
    for (; i < i_end; i += i_step) {
        const uint32_t item = item_of(i);
        const uint32_t next = i + i_step;
        Prefetch pf;
        pf.map = tmap;
        pf.row = ((C::WARPTAIL ? (tid & 31u) == 0 : tid == 0) && next < i_end) ? job.src_row(item_of(next))
                                                                                        : kNoPrefetch;
        const ModTab& t = job.mod(item);
        if constexpr (SMEM_HEAD && Job::kModulusRuns) {
            if ((FWD ? t.ftwd : t.itwd) != cur_tw) {     // uniform: the next run of items, under another modulus
                __syncthreads();                          // slower warps may still be reading the old head twiddles
                load_twiddles(t);
                cur_tw = FWD ? t.ftwd : t.itwd;
                __syncthreads();
            }
        }
        // The copy of the next polynomial into the buffer can only start when this one has left it (the tail of this
        // transform); its lines are asked for now, a whole transform earlier, so that the copy finds them in L2.
        if (tid == 0 && next < i_end && t.l2_prefetch) {
            const uint32_t row0 = job.src_row(item_of(next));
#pragma unroll
            for (uint32_t b = 0; b < TmaGeom<C>::BOXES; ++b) tma_prefetch_rows(tmap, row0 + b * TmaGeom<C>::BOX_ROWS);
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        bool done;
        if constexpr (TMEM_TAIL && FWD && FP64 == 3) {
            Fp64ArithRawST a;
            a.m = t.fd;
            a.head_s = head_s;
            a.ttail = ttail;
            done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (TMEM_TAIL && FWD && FP64 == 2) {
            Fp64AltArithST a;
            a.m = t.fd;
            a.head_s = head_s;
            a.ttail = ttail;
            done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (TMEM_TAIL && !FWD && FP64 == 4) {
            Fp64ArithSTD a;      // doubles out
            a.m = t.fd;
            a.head_s = head_s;
            a.ttail = ttail;
            done = ntt_inv_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (TMEM_TAIL) {
            Fp64ArithST a;
            a.m = t.fd;
            a.head_s = head_s;
            a.ttail = ttail;
            if constexpr (FWD) done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
            else done = ntt_inv_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (SMEM_HEAD && FWD && FP64 == 2) {
            Fp64AltArithS a;
            a.m = t.fd;
            a.head_s = head_s;
            done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (FWD && FP64 == 3) {
            Fp64ArithRaw a;      // as FP64 == 2, raw doubles out
            a.m = t.fd;
            done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (FWD && FP64 == 2) {
            Fp64AltArith a;
            a.m = t.fd;
            done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (SMEM_HEAD) {
            Fp64ArithS a;
            a.m = t.fd;
            a.head_s = head_s;
            if constexpr (FWD) done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
            else done = ntt_inv_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (FP64 != 0) {
            const Fp64Arith a = {t.fd};
            if constexpr (FWD) done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
            else done = ntt_inv_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (LAZY) {
            const LazyInvArith a = {t.fm, t.sc};
            done = ntt_inv_cta<C, MODE>(W, t, a, job.xf(item), job.of(item, smap), pf);
        } else if constexpr (MODE == kFastVote || MODE == kFastTrust) {
            if constexpr (FWD) {
                // multiples of this item's modulus for the final reduction; the table of the
                // previous item may still be read by slower warps, hence two of them.  Written
                // before the first barrier of the transform, read after its last one.
                uint64_t* kq = W + SmemPlan<C>::KQ_WORD + (parity ? 0 : 64);
                if (tid < 64) kq[tid] = (uint64_t)tid * t.fm.q;
                FastArithTab a;
                a.m = t.fm;
                a.sc = t.sc;
                a.kq = kq;
                done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
            } else {
                const FastArith a = {t.fm, t.sc};
                done = ntt_inv_cta<C, MODE>(W, t, a, job.xf(item), job.of(item, smap), pf);
            }
        } else {
            const ExactArith a = {t.q, t.twoq, t.sc, t.lazy_out};
            if constexpr (FWD) done = ntt_fwd_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
            else done = ntt_inv_cta<C, MODE, !FOLD>(W, t, a, job.xf(item), job.of(item, smap), pf);
        }
        if constexpr (FOLD) {
            if (!done) {     // out of contract (not in normal use): the reference op sequence, right away
                if (tid == 0) {
                    fence_proxy_async();
                    issue_poly_load<C>(W, tmap, bar, job.src_row(item));
                }
                mbar_wait(bar, parity);
                parity ^= 1;
                const ExactArith ax = {t.q, t.twoq, t.sc, t.lazy_out};
                if constexpr (FWD) ntt_fwd_cta<C, kExactAll>(W, t, ax, job.xf(item), job.of(item, smap), pf);
                else ntt_inv_cta<C, kExactAll>(W, t, ax, job.xf(item), job.of(item, smap), pf);
            }
        } else {
            if (MODE == kFastVote && !done && tid == 0) defer_item(list, item);
        }
    }
    // staged TMA stores read shared memory asynchronously: drain before exit
    if (FWD && SmemPlan<C>::kStagedStore && (tid & 31u) == 0) tma_store_wait_read();
    if constexpr (TMEM_TAIL) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32) tmem_dealloc_all(tmem_base);
    }
}

// ---------------------------------------------------------------------------
// small-modulus (q < 2^30) path: uint32 registers and a uint32 working buffer
// ---------------------------------------------------------------------------
// Shared memory: [ W: N uint64 (TMA landing) ][ S: N uint32 (working copy) ][ mbarrier, flag ]
// The landing buffer is free again as soon as the first pass has narrowed the
// words into S, so the next polynomial is prefetched right after the first
// barrier, a whole transform ahead of its use.
template <class C32>
struct SmallPlan {
    static constexpr uint32_t S_WORD = C32::N;                 // in uint64 words from the base
    static constexpr uint32_t BAR_WORD = C32::N + C32::N / 2;
    static constexpr uint32_t FLAG_WORD = BAR_WORD + 1;
    // head-pass twiddles (8-byte entries), copied once per launch (SmallArithS)
    static constexpr uint32_t HEAD_TW_FWD = C32::fwd_off(C32::NP), HEAD_TW_INV = C32::INV_ENTRIES - C32::N;
    static constexpr uint32_t HEAD_TW = HEAD_TW_FWD > HEAD_TW_INV ? HEAD_TW_FWD : HEAD_TW_INV;
    static constexpr uint32_t TW_WORD = BAR_WORD + 2;
    static constexpr size_t BYTES = (size_t)(TW_WORD + HEAD_TW) * 8;
    static_assert(BYTES <= 227u * 1024u, "small-modulus shared-memory plan does not fit");
};

// narrow a landed word to 32 bits while collecting the range vote
struct XfNarrowVote {
    uint32_t* hi_or;
    uint32_t* lo_max;
    HB_D uint32_t operator()(uint64_t x) const {
        *hi_or |= (uint32_t)(x >> 32);
        *lo_max = max(*lo_max, (uint32_t)x);
        return (uint32_t)x;
    }
};
struct XfSame32 {
    HB_D uint32_t operator()(uint32_t x) const { return x; }
};

template <class C32>
HB_D void small_vote_raise(int bad, uint64_t* base) {
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31u) == 0) base[SmallPlan<C32>::FLAG_WORD] = 1;
}
template <class C32>
HB_D bool small_vote_read(uint64_t* base) {
    uint64_t* flag = base + SmallPlan<C32>::FLAG_WORD;
    if (*flag == 0) return false;
    __syncthreads();
    if (threadIdx.x == 0) *flag = 0;
    return true;
}

struct PrefetchSmall {
    const CUtensorMap* map;
    uint32_t row;
    template <class C64, class C32>
    HB_D void issue(uint64_t* base) const {
        if (row != kNoPrefetch) {
            fence_proxy_async();
            issue_poly_load<C64>(base, map, base + SmallPlan<C32>::BAR_WORD, row);
        }
    }
};

template <class C32, int P, class A>
HB_D void fwd_mid_passes32(uint32_t tid, uint32_t* S, const Tw32* tw, const A& a) {
    if constexpr (P < C32::NP) {
        fwd_head_pass<C32, P>(tid, S, tw, a);
        __syncthreads();
        fwd_mid_passes32<C32, P + 1>(tid, S, tw, a);
    }
}
template <class C32, int P, class A>
HB_D void inv_mid_passes32(uint32_t tid, uint32_t* S, const Tw32* tw, const A& a) {
    if constexpr (P < C32::NP - 1) {
        inv_head_pass<C32, P>(tid, S, tw, a);
        __syncthreads();
        inv_mid_passes32<C32, P + 1>(tid, S, tw, a);
    }
}

// Forward results of a warp (32 rows x 32 uint32 words = 1024 consecutive
// outputs) leave through the warp's 4 KiB slice of the dead working buffer:
// rows go in with the tail pattern, come out with the lanes along consecutive
// words, are widened to uint64 and written as 32-byte stores -- 1 KiB contiguous
// per warp instruction instead of 32-byte pieces at a 256-byte lane stride.
template <class C32>
HB_D void store_rows32_coalesced(uint32_t* S, uint64_t* dst, uint32_t row, const uint32_t* v) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t* slice = S + (threadIdx.x >> 5) * 1024;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) st_chunk(slice + lane * 32 + (((uint32_t)c ^ (lane & 7u)) << 2), v + 4 * c);
    __syncwarp();
    uint64_t* out = dst + (size_t)(row - lane) * C32::ROW;     // first output word of the warp's 32 rows
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t r = i * 4 + (lane >> 3), ch = lane & 7u;
        uint32_t w[4];
        ld_chunk(slice + r * 32 + ((ch ^ (r & 7u)) << 2), w);
        asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(out + i * 128 + lane * 4), "l"((uint64_t)w[0]),
                     "l"((uint64_t)w[1]), "l"((uint64_t)w[2]), "l"((uint64_t)w[3])
                     : "memory");
    }
}

// Asynchronous version: the rows are widened to uint64 into the warp's 4 KiB slice
// of the dead working buffer (128-byte tensor rows, TMA swizzle) and handed to the
// TMA store engine, 16 words of 32 rows per box, so the 128 KiB of results drain
// in the background instead of as one burst of register stores that every warp
// waits on.  `row` = first of the warp's 32 consecutive rows in the store map.
template <class C32>
HB_D void store_rows32_tma(uint64_t* slice, const CUtensorMap* smap32, uint32_t row, const uint32_t* v) {
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (lane == 0) tma_store_wait_read();     // the previous box of this slice has left
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c)
            st2(slice + lane * 16 + (((uint32_t)c ^ (lane & 7u)) << 1), (uint64_t)v[j * 16 + 2 * c],
                (uint64_t)v[j * 16 + 2 * c + 1]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) tma_store_rows3(smap32, slice, (uint32_t)j, row);
    }
}

// forward, small modulus: base -> dst (bit-reversed order, [0,q)); false = deferred
template <class C64, class C32, int MODE>
HB_D bool ntt_fwd_small_cta(uint64_t* base, const ModTab& t, uint64_t* dst, const PrefetchSmall& pf,
                            const CUtensorMap* smap32, uint32_t item, uint32_t head_s) {
    using P0 = FwdPass<C32, 0>;
    const uint32_t tid = threadIdx.x;
    uint32_t* S = reinterpret_cast<uint32_t*>(base + SmallPlan<C32>::S_WORD);
    SmallArithS a;
    a.m = t.sm32;
    a.head_s = head_s;
    const Tw32* tw_head = a.fwd_base();      // head-pass twiddles: shared memory
    uint32_t v[C32::E];
    uint32_t hi_or = 0, lo_max = 0;
    head_load<C32, P0::R, P0::LS>(tid, base, v, XfNarrowVote{&hi_or, &lo_max});
    if constexpr (MODE == kFastVote)   // forward contract: every word < 4q
        small_vote_raise<C32>((hi_or != 0) | (lo_max >= 2u * t.sm32.twoq), base);
    fwd_head_compute<C32, 0>(tid, v, tw_head, a, [&](int gi, int k0, int k1) {
        head_store_word<C32, P0::R, P0::LS>(tid, S, gi, k0, v[gi * (1 << P0::R) + k0]);
        head_store_word<C32, P0::R, P0::LS>(tid, S, gi, k1, v[gi * (1 << P0::R) + k1]);
    });
    __syncthreads();
    pf.template issue<C64, C32>(base);       // the landing buffer is free from here on
    if constexpr (MODE == kFastVote) {
        if (small_vote_read<C32>(base)) return false;
    }
    fwd_mid_passes32<C32, 1>(tid, S, tw_head, a);
    tail_load<C32>(tid, S, v, XfSame32());
    __syncthreads();                          // S may be overwritten by the next polynomial's first pass
    // each row leaves as soon as it is final, so its stores drain while the next row is computed
    if (smap32) {
        uint64_t* slice = base + SmallPlan<C32>::S_WORD + (tid >> 5) * 512;
        fwd_tail_compute<C32>(tid, v, t.ftw32, a, [&](int ri) {
            store_rows32_tma<C32>(slice, smap32, item * (C32::N / C32::ROW) + (tid & ~31u) + ri * C32::NT,
                                  v + ri * C32::ROW);
        });
        if ((tid & 31u) == 0) tma_store_wait_read();
    } else {
        fwd_tail_compute<C32>(tid, v, t.ftw32, a, [&](int ri) {
            store_rows32_coalesced<C32>(S, dst, tid + ri * C32::NT, v + ri * C32::ROW);
        });
    }
    // the next transform's first pass writes S: all staged rows must have been read back
    __syncthreads();
    return true;
}

// inverse, small modulus: base (bit-reversed order) -> dst (natural order, [0,q))
template <class C64, class C32, int MODE>
HB_D bool ntt_inv_small_cta(uint64_t* base, const ModTab& t, uint64_t* dst, const PrefetchSmall& pf,
                            uint32_t head_s) {
    using PL = InvPass<C32, C32::NP - 1>;
    const uint32_t tid = threadIdx.x;
    uint32_t* S = reinterpret_cast<uint32_t*>(base + SmallPlan<C32>::S_WORD);
    SmallArithS a;
    a.m = t.sm32;
    a.head_s = head_s;
    const Tw32* tw_head = a.template inv_base<C32>();   // head-pass twiddles: shared memory
    uint32_t v[C32::E];
    uint32_t hi_or = 0, lo_max = 0;
    const XfNarrowVote xf = {&hi_or, &lo_max};
    // a row of 32 words = two 16-word rows of the uint64 landing buffer
#pragma unroll
    for (int ri = 0; ri < C32::E / C32::ROW; ++ri) {
        const uint32_t row64 = (tid + ri * C32::NT) * 2;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint64_t x, y;
                ld2(base + (row64 + h) * 16 + (((uint32_t)c ^ ((row64 + h) & 7u)) << 1), x, y);
                v[ri * C32::ROW + h * 16 + 2 * c] = xf(x);
                v[ri * C32::ROW + h * 16 + 2 * c + 1] = xf(y);
            }
    }
    if constexpr (MODE == kFastVote)   // inverse contract: every word < 2q
        small_vote_raise<C32>((hi_or != 0) | (lo_max >= t.sm32.twoq), base);
    inv_tail_compute<C32>(tid, v, t.itw32, a);
    tail_store<C32>(tid, S, v);
    __syncthreads();
    pf.template issue<C64, C32>(base);
    if constexpr (MODE == kFastVote) {
        if (small_vote_read<C32>(base)) return false;
    }
    inv_mid_passes32<C32, 0>(tid, S, tw_head, a);
    head_load<C32, PL::R, PL::LS>(tid, S, v, XfSame32());
    __syncthreads();
    inv_head_compute<C32, C32::NP - 1>(tid, v, tw_head, a, [&](int gi, int k0, int k1) {
        dst[inv_last_index<C32>(tid, gi, k0)] = v[gi * (1 << PL::R) + k0];
        dst[inv_last_index<C32>(tid, gi, k1)] = v[gi * (1 << PL::R) + k1];
    });
    return true;
}

// persistent skeleton of the small-modulus kernels (plain in-place batches)
template <class C64, class C32, bool FWD, int MODE>
HB_D void ntt_persistent_small(const CUtensorMap* tmap, uint64_t* data, const ModTab& t, uint32_t n_items,
                               uint32_t* list, const CUtensorMap* smap32) {
    asm volatile("griddepcontrol.wait;" ::: "memory");      // see ntt_persistent
    asm volatile("griddepcontrol.launch_dependents;");
    uint64_t* base = smem_poly<C64>();
    uint64_t* bar = base + SmallPlan<C32>::BAR_WORD;
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t ROWS = C64::N / 16;
    if (tid == 0) {
        if (smem_u32(base) & 1023u) __trap();
        mbar_init(bar, 1);
        base[SmallPlan<C32>::FLAG_WORD] = 0;
        fence_barrier_init();
    }
    {   // the twiddles of the head passes move next to the buffers once per launch
        const uint2* src = reinterpret_cast<const uint2*>(FWD ? t.ftw32 : t.itw32 + C32::inv_off(0));
        uint2* dstw = reinterpret_cast<uint2*>(base + SmallPlan<C32>::TW_WORD);
        constexpr uint32_t COUNT = FWD ? SmallPlan<C32>::HEAD_TW_FWD : SmallPlan<C32>::HEAD_TW_INV;
        for (uint32_t e = tid; e < COUNT; e += C32::NT) dstw[e] = src[e];
    }
    __syncthreads();
    uint32_t head_s;
    // volatile: the address (and every load through it) stays behind the barrier above
    asm volatile("mov.u32 %0, %1;" : "=r"(head_s) : "r"(smem_u32(base + SmallPlan<C32>::TW_WORD)));
    uint32_t item = blockIdx.x;
    if (tid == 0 && item < n_items) issue_poly_load<C64>(base, tmap, bar, item * ROWS);
    uint32_t parity = 0;
    for (; item < n_items; item += gridDim.x) {
        const uint32_t next = item + gridDim.x;
        PrefetchSmall pf;
        pf.map = tmap;
        pf.row = (tid == 0 && next < n_items) ? next * ROWS : kNoPrefetch;
        mbar_wait(bar, parity);
        parity ^= 1;
        uint64_t* dst = data + (size_t)item * C64::N;
        bool done;
        if constexpr (FWD) done = ntt_fwd_small_cta<C64, C32, MODE>(base, t, dst, pf, smap32, item, head_s);
        else done = ntt_inv_small_cta<C64, C32, MODE>(base, t, dst, pf, head_s);
        if (MODE == kFastVote && !done && tid == 0) defer_item(list, item);
        // the next transform's first pass writes S: everyone must have left this one
        // (its last reads of S are followed by a block barrier inside the cta functions)
    }
}

#ifdef HB_EXPERIMENTAL_VARIANTS
// ---------------------------------------------------------------------------
// small-modulus path, second generation ("small2"): no landing buffer
// ---------------------------------------------------------------------------
// The first pass reads the uint64 polynomial straight from global memory into
// registers (forward: columns, coalesced along the lanes; inverse: the 1024
// consecutive words of a warp's 32 rows, dropped into the warp's own rows of
// the working buffer), so a CTA needs only the 64 KiB uint32 working buffer
// and 64 registers per thread: TWO CTAs per SM, whose barrier / load / store
// phases overlap each other's butterflies.  The polynomial of the next
// iteration is pulled towards L2 while the current one is transformed.
template <class C32>
struct Small2Plan {
    static constexpr uint32_t FLAG_U32 = C32::N;     // two range-vote flags (iteration parity)
    static constexpr size_t BYTES = (size_t)C32::N * 4 + 16;
};

HB_D uint64_t ldg_stream(const uint64_t* p) {
    uint64_t x;
    asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(x) : "l"(p));
    return x;
}
HB_D void ldg_stream2(const uint64_t* p, uint64_t& a, uint64_t& b) {
    asm volatile("ld.global.cs.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
HB_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Range vote of the small2 kernels.  Flag (iter & 1) belongs to this iteration;
// the other one (raised at most in the previous iteration, read by everybody
// before this iteration's barrier) is cleared here, a full iteration ahead of
// its next use.
template <class C32>
HB_D bool small2_vote(uint32_t* S, int bad, uint32_t iter) {
    volatile uint32_t* flag = S + Small2Plan<C32>::FLAG_U32;
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31u) == 0) flag[iter & 1u] = 1;
    __syncthreads();
    const bool deferred = flag[iter & 1u] != 0;
    if (threadIdx.x == 0) flag[(iter & 1u) ^ 1u] = 0;
    return deferred;
}

// forward first pass: word k of group g is poly[base(g) + (k << LS)]; the lanes
// of a warp read consecutive words (256 contiguous bytes per instruction)
template <class C, int R, int LS>
HB_D void head_load_global_narrow(uint32_t tid, const uint64_t* poly, uint32_t* v, uint32_t& hi_or,
                                  uint32_t& lo_max) {
    using Gm = HeadGeom<C, R, LS>;
    static_for<0, Gm::G>([&](auto gc) {
        constexpr int gi = decltype(gc)::value;
        const uint64_t* p = poly + Gm::base(tid + gi * C::NT);
        static_for<0, (1 << R) / 8>([&](auto cc) {
            constexpr int c0 = decltype(cc)::value * 8;
            uint64_t t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = ldg_stream(p + ((uint32_t)(c0 + j) << LS));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                hi_or |= (uint32_t)(t[j] >> 32);
                lo_max = max(lo_max, (uint32_t)t[j]);
                v[gi * (1 << R) + c0 + j] = (uint32_t)t[j];
            }
        });
    });
}

template <class C32, int MODE>
HB_D bool ntt_fwd_small2_cta(uint32_t* S, const ModTab& t, uint64_t* poly, uint32_t iter) {
    using P0 = FwdPass<C32, 0>;
    const uint32_t tid = threadIdx.x;
    const SmallArith a = {t.sm32};
    uint32_t v[C32::E];
    uint32_t hi_or = 0, lo_max = 0;
    head_load_global_narrow<C32, P0::R, P0::LS>(tid, poly, v, hi_or, lo_max);
    fwd_head_compute<C32, 0>(tid, v, t.ftw32, a, [&](int gi, int k0, int k1) {
        head_store_word<C32, P0::R, P0::LS>(tid, S, gi, k0, v[gi * (1 << P0::R) + k0]);
        head_store_word<C32, P0::R, P0::LS>(tid, S, gi, k1, v[gi * (1 << P0::R) + k1]);
    });
    if constexpr (MODE == kFastVote) {
        // forward contract: every word < 4q; global memory still holds the untouched input
        if (small2_vote<C32>(S, (hi_or != 0) | (lo_max >= 2u * t.sm32.twoq), iter)) return false;
    } else {
        __syncthreads();
    }
    fwd_mid_passes32<C32, 1>(tid, S, t.ftw32, a);
    tail_load<C32>(tid, S, v, XfSame32());
    __syncthreads();                          // the staging slices below overwrite S
    fwd_tail_compute<C32>(tid, v, t.ftw32, a);
#pragma unroll
    for (int ri = 0; ri < C32::E / C32::ROW; ++ri)
        store_rows32_coalesced<C32>(S, poly, tid + ri * C32::NT, v + ri * C32::ROW);
    __syncthreads();                          // the next transform's first pass writes S
    return true;
}

template <class C32, int MODE>
HB_D bool ntt_inv_small2_cta(uint32_t* S, const ModTab& t, uint64_t* poly, uint32_t iter) {
    using PL = InvPass<C32, C32::NP - 1>;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const SmallArith a = {t.sm32};
    uint32_t v[C32::E];
    uint32_t hi_or = 0, lo_max = 0;
    // rows tid + ri*NT: the 32 rows of a warp are 1024 consecutive words of the
    // polynomial -> coalesced 16-byte reads, narrowed into the warp's rows of S
#pragma unroll
    for (int ri = 0; ri < C32::E / C32::ROW; ++ri) {
        const uint32_t row0 = (tid - lane) + ri * C32::NT;
        const uint64_t* src = poly + (size_t)row0 * C32::ROW;
#pragma unroll
        for (int i0 = 0; i0 < 16; i0 += 4) {
            uint64_t x[4], y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) ldg_stream2(src + (i0 + j) * 64 + lane * 2, x[j], y[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hi_or |= (uint32_t)(x[j] >> 32) | (uint32_t)(y[j] >> 32);
                lo_max = max(lo_max, max((uint32_t)x[j], (uint32_t)y[j]));
                const uint32_t e = (i0 + j) * 64 + lane * 2;
                const uint32_t r = row0 + (e >> 5), col = e & 31u;
                uint32_t* d = S + r * C32::ROW + ((((col >> 2) ^ (r & 7u)) << 2) | (col & 3u));
                *reinterpret_cast<uint2*>(d) = make_uint2((uint32_t)x[j], (uint32_t)y[j]);
            }
        }
    }
    __syncwarp();
    tail_load<C32>(tid, S, v, XfSame32());
    __syncwarp();                             // tail_store below rewrites the warp's rows
    inv_tail_compute<C32>(tid, v, t.itw32, a);
    tail_store<C32>(tid, S, v);
    if constexpr (MODE == kFastVote) {
        // inverse contract: every word < 2q
        if (small2_vote<C32>(S, (hi_or != 0) | (lo_max >= t.sm32.twoq), iter)) return false;
    } else {
        __syncthreads();
    }
    inv_mid_passes32<C32, 0>(tid, S, t.itw32, a);
    head_load<C32, PL::R, PL::LS>(tid, S, v, XfSame32());
    __syncthreads();                          // S is free for the next transform's rows
    inv_head_compute<C32, C32::NP - 1>(tid, v, t.itw32, a);
#pragma unroll
    for (int gi = 0; gi < (C32::E >> PL::R); ++gi)
#pragma unroll
        for (int k = 0; k < (1 << PL::R); ++k) poly[inv_last_index<C32>(tid, gi, k)] = v[gi * (1 << PL::R) + k];
    return true;
}

template <class C32, bool FWD, int MODE>
HB_D void ntt_persistent_small2(uint64_t* data, const ModTab& t, uint32_t n_items, uint32_t* list) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t* S = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    if (tid < 2) S[Small2Plan<C32>::FLAG_U32 + tid] = 0;
    __syncthreads();
    uint32_t iter = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
        const uint32_t next = item + gridDim.x;
        if (next < n_items) {
            // 128 KiB = 1024 lines of 128 bytes: two per thread
            const uint8_t* nx = reinterpret_cast<const uint8_t*>(data + (size_t)next * C32::N);
#pragma unroll
            for (uint32_t l = tid; l < C32::N * 8 / 128; l += C32::NT) prefetch_l2(nx + (size_t)l * 128);
        }
        uint64_t* poly = data + (size_t)item * C32::N;
        bool done;
        if constexpr (FWD) done = ntt_fwd_small2_cta<C32, MODE>(S, t, poly, iter);
        else done = ntt_inv_small2_cta<C32, MODE>(S, t, poly, iter);
        if (MODE == kFastVote && !done && tid == 0) defer_item(list, item);
    }
}

// ---------------------------------------------------------------------------
// small-modulus path, third generation ("small3"): two transforms per SM
// ---------------------------------------------------------------------------
// One CTA of 1024 threads = two groups of 512, each running its own stream of polynomials
// with its own named barrier, so that the store burst, the barrier bubbles and the shared-
// memory phases of one transform hide under the butterflies of the other.  Shared memory:
// three 64 KiB regions.  Group A lands a polynomial (uint64, 128 KiB) in R0 (lower half) +
// R1 (upper half), narrows it IN PLACE into R0 during the first pass and works there;
// group B uses R2 + R1 the same way.  R1 is needed only between a group's TMA issue and the
// end of its first-pass loads, so the groups take turns on it: releases are arrivals on one
// mbarrier (phase k = k-th release), A acquires on even turns, B on odd ones.
template <class C32>
struct Small3Plan {
    static constexpr uint32_t REGION = C32::N / 2;          // uint64 words per 64 KiB region
    static constexpr uint32_t BAR_WORD = 3 * REGION;        // full[A], full[B], R1 token
    static constexpr uint32_t FLAG_WORD = BAR_WORD + 3;     // four uint32 flags: [group][iteration parity]
    static constexpr size_t BYTES = (size_t)(FLAG_WORD + 2) * 8;
};

HB_D void group_sync(uint32_t g) { asm volatile("bar.sync %0, 512;" ::"r"(g + 1u) : "memory"); }
HB_D void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// range vote of a group: flag (iter & 1) belongs to this iteration, the other one is cleared
// a full iteration ahead of its next use (cf. small2_vote); includes the group barrier
HB_D bool small3_vote(volatile uint32_t* flags, int bad, uint32_t iter, uint32_t g, uint32_t gtid, bool vote) {
    if (vote && __any_sync(0xffffffffu, bad) && (gtid & 31u) == 0) flags[iter & 1u] = 1;
    group_sync(g);
    if (!vote) return false;
    const bool deferred = flags[iter & 1u] != 0;
    if (gtid == 0) flags[(iter & 1u) ^ 1u] = 0;
    return deferred;
}

// 32 consecutive result words of one row, widened, straight from registers (8 x 32 bytes)
HB_D void store_row32_direct(uint64_t* dst, const uint32_t* v) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * c), "l"((uint64_t)v[4 * c]),
                     "l"((uint64_t)v[4 * c + 1]), "l"((uint64_t)v[4 * c + 2]), "l"((uint64_t)v[4 * c + 3])
                     : "memory");
}

template <class C32, int P>
HB_D void fwd_mid_passes3(uint32_t gtid, uint32_t g, uint32_t* S, const Tw32* tw, const SmallArith& a) {
    if constexpr (P < C32::NP) {
        fwd_head_pass<C32, P>(gtid, S, tw, a);
        group_sync(g);
        fwd_mid_passes3<C32, P + 1>(gtid, g, S, tw, a);
    }
}
template <class C32, int P>
HB_D void inv_mid_passes3(uint32_t gtid, uint32_t g, uint32_t* S, const Tw32* tw, const SmallArith& a) {
    if constexpr (P < C32::NP - 1) {
        inv_head_pass<C32, P>(gtid, S, tw, a);
        group_sync(g);
        inv_mid_passes3<C32, P + 1>(gtid, g, S, tw, a);
    }
}

// forward: landing (low | up) -> poly (bit-reversed order, [0,q)); false = deferred
template <class C32, int MODE>
HB_D bool ntt_fwd_small3(const uint64_t* low, const uint64_t* up, uint32_t* S, uint64_t* token,
                         volatile uint32_t* flags, const ModTab& t, uint64_t* poly, uint32_t iter, uint32_t g,
                         uint32_t gtid) {
    using P0 = FwdPass<C32, 0>;
    static_assert(P0::LS == C32::LOGN - P0::R && (C32::E >> P0::R) == 1, "one column per thread in the first pass");
    const SmallArith a = {t.sm32};
    uint32_t v[C32::E];
    uint32_t hi_or = 0, lo_max = 0;
    // column gtid: word k sits at k * 2^LS + gtid; k below E/2 in the lower landing half
#pragma unroll
    for (int c0 = 0; c0 < C32::E; c0 += 8) {
        uint64_t tt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = c0 + j;
            const uint64_t* src = (k < C32::E / 2) ? low : up;
            tt[j] = src[swz(((uint32_t)(k & (C32::E / 2 - 1)) << P0::LS) + gtid)];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            hi_or |= (uint32_t)(tt[j] >> 32);
            lo_max = max(lo_max, (uint32_t)tt[j]);
            v[c0 + j] = (uint32_t)tt[j];
        }
    }
    // everybody has read the landing halves: R1 goes to the other group, the lower half becomes S
    const bool deferred =
        small3_vote(flags, (hi_or != 0) | (lo_max >= 2u * t.sm32.twoq), iter, g, gtid, MODE == kFastVote);
    if (gtid == 0) mbar_arrive(token);
    if (deferred) return false;
    fwd_head_compute<C32, 0>(gtid, v, t.ftw32, a, [&](int gi, int k0, int k1) {
        head_store_word<C32, P0::R, P0::LS>(gtid, S, gi, k0, v[gi * (1 << P0::R) + k0]);
        head_store_word<C32, P0::R, P0::LS>(gtid, S, gi, k1, v[gi * (1 << P0::R) + k1]);
    });
    group_sync(g);
    fwd_mid_passes3<C32, 1>(gtid, g, S, t.ftw32, a);
    tail_load<C32>(gtid, S, v, XfSame32());
    group_sync(g);                            // S (the lower region) is free for the next landing
    fwd_tail_compute<C32>(gtid, v, t.ftw32, a, [&](int ri) {
        store_row32_direct(poly + (size_t)(gtid + ri * C32::NT) * C32::ROW, v + ri * C32::ROW);
    });
    return true;
}

// inverse: landing (bit-reversed order) -> poly (natural order, [0,q))
template <class C32, int MODE>
HB_D bool ntt_inv_small3(const uint64_t* low, const uint64_t* up, uint32_t* S, uint64_t* token,
                         volatile uint32_t* flags, const ModTab& t, uint64_t* poly, uint32_t iter, uint32_t g,
                         uint32_t gtid) {
    using PL = InvPass<C32, C32::NP - 1>;
    static_assert(C32::E == C32::ROW, "one row per thread");
    const SmallArith a = {t.sm32};
    uint32_t v[C32::E];
    uint32_t hi_or = 0, lo_max = 0;
    // row gtid = uint64 landing rows 2*gtid, 2*gtid + 1; rows below N/64 live in the lower half
    const uint32_t row64 = gtid * 2;
    const uint64_t* src = (gtid < C32::NT / 2) ? low : up - Small3Plan<C32>::REGION;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint64_t x, y;
            ld2(src + (row64 + h) * 16 + (((uint32_t)c ^ ((row64 + h) & 7u)) << 1), x, y);
            hi_or |= (uint32_t)(x >> 32) | (uint32_t)(y >> 32);
            lo_max = max(lo_max, max((uint32_t)x, (uint32_t)y));
            v[h * 16 + 2 * c] = (uint32_t)x;
            v[h * 16 + 2 * c + 1] = (uint32_t)y;
        }
    const bool deferred = small3_vote(flags, (hi_or != 0) | (lo_max >= t.sm32.twoq), iter, g, gtid, MODE == kFastVote);
    if (gtid == 0) mbar_arrive(token);
    if (deferred) return false;
    inv_tail_compute<C32>(gtid, v, t.itw32, a);
    tail_store<C32>(gtid, S, v);
    group_sync(g);
    inv_mid_passes3<C32, 0>(gtid, g, S, t.itw32, a);
    head_load<C32, PL::R, PL::LS>(gtid, S, v, XfSame32());
    group_sync(g);                            // S is free for the next landing
    inv_head_compute<C32, C32::NP - 1>(gtid, v, t.itw32, a, [&](int gi, int k0, int k1) {
        poly[inv_last_index<C32>(gtid, gi, k0)] = v[gi * (1 << PL::R) + k0];
        poly[inv_last_index<C32>(gtid, gi, k1)] = v[gi * (1 << PL::R) + k1];
    });
    return true;
}

template <class C32, bool FWD, int MODE>
HB_D void ntt_persistent_small3(const CUtensorMap* tmap, uint64_t* data, const ModTab& t, uint32_t n_items,
                                uint32_t* list) {
    using P3 = Small3Plan<C32>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint64_t* base = reinterpret_cast<uint64_t*>(smem_raw);
    const uint32_t g = threadIdx.x >> 9, gtid = threadIdx.x & 511u;
    uint64_t* low = base + (g ? 2 * P3::REGION : 0);           // R0 for group A, R2 for group B
    uint64_t* up = base + P3::REGION;                          // R1, shared in turns
    uint32_t* S = reinterpret_cast<uint32_t*>(low);
    uint64_t* full = base + P3::BAR_WORD + g;
    uint64_t* token = base + P3::BAR_WORD + 2;
    volatile uint32_t* flags = reinterpret_cast<volatile uint32_t*>(base + P3::FLAG_WORD) + 2 * g;
    constexpr uint32_t ROWS = C32::N / 16;                     // uint64 tensor-map rows per polynomial
    if (threadIdx.x == 0) {
        if (smem_u32(base) & 1023u) __trap();
        mbar_init(base + P3::BAR_WORD, 1);
        mbar_init(base + P3::BAR_WORD + 1, 1);
        mbar_init(token, 1);
        uint32_t* f = reinterpret_cast<uint32_t*>(base + P3::FLAG_WORD);
        f[0] = f[1] = f[2] = f[3] = 0;
        fence_barrier_init();
    }
    __syncthreads();
    // the CTA's items are blockIdx.x + s * gridDim.x; group g takes the positions s = g, g + 2, ...
    auto item_at = [&](uint32_t j) { return blockIdx.x + (2 * j + g) * gridDim.x; };
    // leader: take R1 (turn 2j + g waits for release 2j + g - 1) and start the landing of item j
    auto land = [&](uint32_t j) {
        if (2 * j + g > 0) mbar_wait(token, (g + 1u) & 1u);    // A waits odd phases, B even ones
        fence_proxy_async();
        const uint32_t row0 = item_at(j) * ROWS;
        mbar_expect_tx(full, C32::N * 8);
        tma_load_rows(low, tmap, full, row0);
        tma_load_rows(low + 4096, tmap, full, row0 + 256);
        tma_load_rows(up, tmap, full, row0 + 512);
        tma_load_rows(up + 4096, tmap, full, row0 + 768);
    };
    if (item_at(0) >= n_items) return;                         // nothing for this group
    if (gtid == 0) land(0);
    for (uint32_t j = 0; item_at(j) < n_items; ++j) {
        const uint32_t item = item_at(j);
        mbar_wait(full, j & 1u);
        uint64_t* poly = data + (size_t)item * C32::N;
        bool done;
        if constexpr (FWD) done = ntt_fwd_small3<C32, MODE>(low, up, S, token, flags, t, poly, j, g, gtid);
        else done = ntt_inv_small3<C32, MODE>(low, up, S, token, flags, t, poly, j, g, gtid);
        if (MODE == kFastVote && !done && gtid == 0) defer_item(list, item);
        if (!done) group_sync(g);                              // nobody reads the landing any more
        if (gtid == 0 && item_at(j + 1) < n_items) land(j + 1);
    }
}

#endif  // HB_EXPERIMENTAL_VARIANTS

// shared memory of the FP64 kernels of the plain batched calls (head-pass twiddles included when they fit)
template <class C>
constexpr size_t ntt_smem_bytes_fp64_plain() {
    return SmemPlan<C>::kHeadTwFits ? SmemPlan<C>::BYTES_TW : SmemPlan<C>::BYTES;
}
template <class C>
constexpr size_t ntt_smem_bytes() {
    return SmemPlan<C>::BYTES;
}

}  // namespace hb
