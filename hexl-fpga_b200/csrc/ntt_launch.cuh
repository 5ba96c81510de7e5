// ntt_launch.cuh -- kernels and launch templates of the plain batched transforms, shared by
// ntt_kernels.cu (forward, twiddle packing, host helpers) and ntt_kernels_inv.cu (inverse), which
// are separate translation units only to halve the build time.
#pragma once
#include <type_traits>

#include "launch.h"

namespace hb {

// store map of the small-modulus forward epilogue (ntt_kernels.cu)
cudaError_t make_rows32_store_tmap(CUtensorMap* out, const void* base, uint64_t polys, uint32_t logn);

// ---- plain batched transform, in place ------------------------------------
template <class C>
struct JobPlain {
    static constexpr bool kOneModulus = true;   // every item of a launch is transformed under tab
    static constexpr bool kModulusRuns = false;
    static constexpr bool kPostXf = false;       // JobInvMul: the load transform reads a second operand
    HB_D uint32_t order(uint32_t i) const { return i; }
    uint64_t* data;
    ModTab tab;
    // item i is polynomial offset + i * stride of the array (1, 0: a plain batch; 2, h: the half-size
    // sub-transforms of an N = 2 * C::N transform, ntt_big.cu)
    uint32_t stride = 1, offset = 0;
    HB_D uint32_t poly(uint32_t item) const { return offset + item * stride; }
    HB_D uint32_t src_row(uint32_t item) const { return poly(item) * (C::N / 16); }
    HB_D const ModTab& mod(uint32_t) const { return tab; }
    HB_D XfIdent xf(uint32_t) const { return XfIdent(); }
};
template <class C>
struct JobFwd : JobPlain<C> {
    HB_D OfRows of(uint32_t item, const CUtensorMap* smap) const {
        return OfRows{this->data + (size_t)this->poly(item) * C::N, smap, this->poly(item) * (C::N / 16)};
    }
};
template <class C>
struct JobInv : JobPlain<C> {
    HB_D OfWords of(uint32_t item, const CUtensorMap*) const {
        return OfWords{this->data + (size_t)this->poly(item) * C::N};
    }
};

template <class C, int MODE, int FP64 = 0>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ntt_fwd(const __grid_constant__ CUtensorMap tmap,
                                                   const __grid_constant__ CUtensorMap smap, const JobFwd<C> job,
                                                   uint32_t n_items, uint32_t* list) {
    ntt_persistent<C, true, MODE, JobFwd<C>, false, FP64>(&tmap, &smap, job, n_items, list);
}
template <class C, int MODE, bool LAZY = false, int FP64 = 0>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ntt_inv(const __grid_constant__ CUtensorMap tmap, const JobInv<C> job,
                                                   uint32_t n_items, uint32_t* list) {
    ntt_persistent<C, false, MODE, JobInv<C>, LAZY, FP64>(&tmap, nullptr, job, n_items, list);
}

// inverse transform of NTT(a) (.) NTT(b): the fused tail of a polynomial multiply
template <class C>
struct JobInvMul : JobPlain<C> {
    static constexpr bool kPostXf = true;
    const uint64_t* other;
    Divisor dv;
    uint32_t n_items;
    HB_D XfMulGlobal xf(uint32_t item) const {
        const uint32_t nxt = item + gridDim.x;      // the persistent loop's stride
        return XfMulGlobal{other + (size_t)item * C::N, nxt < n_items ? other + (size_t)nxt * C::N : nullptr, dv};
    }
    HB_D OfWords of(uint32_t item, const CUtensorMap*) const { return OfWords{this->data + (size_t)item * C::N}; }
};
template <class C, int MODE, int FP64 = 0>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ntt_inv_mul(const __grid_constant__ CUtensorMap tmap,
                                                                   const JobInvMul<C> job, uint32_t n_items) {
    ntt_persistent<C, false, MODE, JobInvMul<C>, false, FP64>(&tmap, nullptr, job, n_items, nullptr);
}

// the same over the items of a deferred list, reference op sequence (polymul_fused.cu)
template <class C>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ntt_inv_mul_list(const __grid_constant__ CUtensorMap tmap,
                                                                        const JobInvMul<C> job, uint32_t n_items,
                                                                        uint32_t* list) {
    ntt_persistent<C, false, kExactList, JobInvMul<C>>(&tmap, nullptr, job, n_items, list);
}

// small-modulus kernels (q < 2^30): uint32 arithmetic, see ntt_block.cuh
// smap32: 3-D store map of the forward epilogue (use_tma_store = 0: plain coalesced stores)
template <class C64, class C32, bool FWD, int MODE>
__global__ void __launch_bounds__(C32::NT, 1) k_ntt_small(const __grid_constant__ CUtensorMap tmap,
                                                         const __grid_constant__ CUtensorMap smap32, uint64_t* data,
                                                         const ModTab tab, uint32_t n_items, uint32_t* list,
                                                         int use_tma_store) {
    ntt_persistent_small<C64, C32, FWD, MODE>(&tmap, data, tab, n_items, list,
                                              (FWD && use_tma_store) ? &smap32 : nullptr);
}

// Kernel variants that were built, verified and MEASURED SLOWER than the defaults are only
// compiled with -DHB_EXPERIMENTAL_VARIANTS (make EXPERIMENTAL=1): small_path 2 / 3, the lazy inverse,
// the 16-words-per-thread configuration at N >= 8192, ks_mac_items 1 / 2 / 8.
#ifdef HB_EXPERIMENTAL_VARIANTS
// second generation: no landing buffer, two CTAs per SM (ntt_block.cuh)
template <class C32, bool FWD, int MODE>
__global__ void __launch_bounds__(C32::NT, 2) k_ntt_small2(uint64_t* data, const ModTab tab, uint32_t n_items,
                                                          uint32_t* list) {
    ntt_persistent_small2<C32, FWD, MODE>(data, tab, n_items, list);
}

// third generation: two transforms per SM sharing three 64 KiB regions (ntt_block.cuh)
template <class C32, bool FWD, int MODE>
__global__ void __launch_bounds__(1024, 1) k_ntt_small3(const __grid_constant__ CUtensorMap tmap, uint64_t* data,
                                                       const ModTab tab, uint32_t n_items, uint32_t* list) {
    ntt_persistent_small3<C32, FWD, MODE>(&tmap, data, tab, n_items, list);
}
#endif  // HB_EXPERIMENTAL_VARIANTS

// the configuration with warp-dealt tail rows, where the shape allows it
template <class C>
struct WarpTailCfg {
    using type = C;
};
template <>
struct WarpTailCfg<NttCfg<14, 5, 4, 0>> {
    using type = NttCfg<14, 5, 4, 1>;
};

// Launch with programmatic stream serialization (option "pdl"): the grid may be scheduled while the
// kernel in front of it in the stream (this call's pack kernel, or the main kernel in front of the
// deferred-list pass) is still running; its CTAs execute griddepcontrol.wait before they touch
// anything (ntt_persistent*), so only launch latency and the drain between the kernels overlap.
template <class... KArgs, class... Args>
static cudaError_t launch_dep(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    if (g_time_kernels) {
        // bench.py's roofline leg: the kernel's own duration, CUDA events on its stream directly around the launch
        // (a stream marker between two kernels rules the programmatic overlap out, so the launch is a plain one)
        cudaEvent_t a, b;
        cudaError_t e;
        if ((e = cudaEventCreate(&a)) || (e = cudaEventCreate(&b))) return e;
        if ((e = cudaEventRecord(a, st))) return e;
        kern<<<grid, block, smem, st>>>(KArgs(args)...);
        if ((e = cudaGetLastError())) return e;
        if ((e = cudaEventRecord(b, st))) return e;
        note_kernel_events(a, b, grid);
        return cudaSuccess;
    }
    if (!g_pdl) {
        kern<<<grid, block, smem, st>>>(KArgs(args)...);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// *folded (voting mode): set when the launched kernel transforms out-of-contract polynomials itself
// (ntt_block.cuh, kFoldExact: the forward FP64 kernels), i.e. no pass over a deferred list has to follow
template <class C, bool FWD, int MODE>
static cudaError_t launch_mode(const CUtensorMap& tmap, const CUtensorMap& smap, uint64_t* base, const ModTab& tab,
                               uint64_t cnt, uint32_t* list, cudaStream_t st, uint32_t stride = 1, uint32_t offset = 0,
                               bool* folded = nullptr) {
    const size_t smem = ntt_smem_bytes<C>();
    cudaError_t e = cudaSuccess;
    auto fp64_kernel = [&]() {
        if ((FWD ? HB_FOLD_FWD != 0 : HB_FOLD_INV != 0) && MODE == kFastVote && folded) *folded = true;
    };
    if constexpr (FWD) {
        JobFwd<C> job;
        job.data = base;
        job.tab = tab;
        job.stride = stride;
        job.offset = offset;
        if constexpr (MODE == kFastVote || MODE == kFastTrust) {
            using CW = typename WarpTailCfg<C>::type;
            if (tab.fp64_ok && tab.fp64_alt_ok && g_warp_tail && !std::is_same<CW, C>::value) {
                fp64_kernel();
                // q <= 2^51 (1 + 1/32): full correction every other stage (modarith.cuh), ~15 % fewer scheduler cycles
                auto kern = k_ntt_fwd<CW, MODE, 2>;
                const size_t smemw = ntt_smem_bytes_fp64_plain<CW>();
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemw))) return e;
                JobFwd<CW> jobw;
                jobw.data = base;
                jobw.tab = tab;
                jobw.stride = stride;
                jobw.offset = offset;
                e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, CW::NT, smemw, cnt), CW::NT, smemw, st, tmap, smap, jobw, (uint32_t)cnt, list);
                return e != cudaSuccess ? e : cudaGetLastError();
            }
            if (tab.fp64_ok && g_warp_tail && !std::is_same<CW, C>::value) {
                fp64_kernel();
                auto kern = k_ntt_fwd<CW, MODE, true>;
                const size_t smemw = ntt_smem_bytes_fp64_plain<CW>();
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemw))) return e;
                JobFwd<CW> jobw;
                jobw.data = base;
                jobw.tab = tab;
                jobw.stride = stride;
                jobw.offset = offset;
                e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, CW::NT, smemw, cnt), CW::NT, smemw, st, tmap, smap, jobw, (uint32_t)cnt, list);
                return e != cudaSuccess ? e : cudaGetLastError();
            }
            if (tab.fp64_ok) {       // 36..51-bit modulus: butterflies on the FP64 pipe
                fp64_kernel();
                auto kern = k_ntt_fwd<C, MODE, true>;
                const size_t smemd = ntt_smem_bytes_fp64_plain<C>();
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemd))) return e;
                e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, C::NT, smemd, cnt), C::NT, smemd, st, tmap, smap, job, (uint32_t)cnt, list);
                return e != cudaSuccess ? e : cudaGetLastError();
            }
        }
        auto kern = k_ntt_fwd<C, MODE>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
        e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, C::NT, smem, cnt), C::NT, smem, st, tmap, smap, job, (uint32_t)cnt, list);
    } else {
        JobInv<C> job;
        job.data = base;
        job.tab = tab;
        job.stride = stride;
        job.offset = offset;
        if constexpr (MODE == kFastVote || MODE == kFastTrust) {
            using CW = typename WarpTailCfg<C>::type;
            if (tab.fp64_ok && g_warp_tail && !std::is_same<CW, C>::value) {
                fp64_kernel();
                auto kern = k_ntt_inv<CW, MODE, false, true>;
                const size_t smemw = ntt_smem_bytes_fp64_plain<CW>();
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemw))) return e;
                JobInv<CW> jobw;
                jobw.data = base;
                jobw.tab = tab;
                jobw.stride = stride;
                jobw.offset = offset;
                e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, CW::NT, smemw, cnt), CW::NT, smemw, st, tmap, jobw, (uint32_t)cnt, list);
                return e != cudaSuccess ? e : cudaGetLastError();
            }
            if (tab.fp64_ok) {
                fp64_kernel();
                auto kern = k_ntt_inv<C, MODE, false, true>;
                const size_t smemd = ntt_smem_bytes_fp64_plain<C>();
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemd))) return e;
                e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, C::NT, smemd, cnt), C::NT, smemd, st, tmap, job, (uint32_t)cnt, list);
                return e != cudaSuccess ? e : cudaGetLastError();
            }
#ifdef HB_EXPERIMENTAL_VARIANTS
            if (tab.inv_lazy_ok) {   // q < 2^52: butterflies without per-stage corrections
                auto kern = k_ntt_inv<C, MODE, true>;
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
                e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, C::NT, smem, cnt), C::NT, smem, st, tmap, job, (uint32_t)cnt, list);
                return e != cudaSuccess ? e : cudaGetLastError();
            }
#endif
        }
        auto kern = k_ntt_inv<C, MODE>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
        e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, C::NT, smem, cnt), C::NT, smem, st, tmap, job, (uint32_t)cnt, list);
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

// `list`: device scratch of 1 + batch words whose first word is zero on entry
// (launch_pack_twiddles resets it); `trust` skips the input-range vote.
// `src`: where the polynomials are read from (nullptr: in place, from `data`)
template <class C, bool FWD>
static cudaError_t launch_one(uint64_t* data, const ModTab& tab, uint64_t batch, bool trust, uint32_t* list,
                              cudaStream_t st, int* launches, const uint64_t* src = nullptr, uint32_t stride = 1,
                              uint32_t offset = 0) {
    CUtensorMap tmap, smap;
    cudaError_t e;
    // the tensor map's row coordinate is 32 bits
    const uint64_t kMaxPolys = ((1ull << 32) - 1) / (C::N / 16);
    if (batch * stride > kMaxPolys) return cudaErrorInvalidValue;
    if ((e = make_poly_tmap(&tmap, src ? src : data, batch * stride, C::LOGN)) != cudaSuccess) return e;
    if ((e = make_poly_tmap(&smap, data, batch * stride, C::LOGN, 32)) != cudaSuccess) return e;
    const bool fast = (FWD ? tab.fwd_fast_ok : tab.inv_fast_ok) && !tab.lazy_out;
    if (!fast) {
        *launches += 1;
        return launch_mode<C, FWD, kExactAll>(tmap, smap, data, tab, batch, list, st, stride, offset);
    }
    if constexpr (C::LOGN == 14 && C::LOGE == 5) {
        // q < 2^30: the 32-bit kernels (out-of-contract items still go to the
        // 64-bit exact kernel through the deferred list)
        if (tab.small_ok && stride == 1) {
            using C32 = NttCfg<14, 5, 5>;
#ifdef HB_EXPERIMENTAL_VARIANTS
            if (tab.small_ok == 3) {
                const size_t smem3 = Small3Plan<C32>::BYTES;
                int sms = 0, dev = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                const unsigned grid3 = (unsigned)(batch < (uint64_t)sms ? batch : (uint64_t)sms);
                if (trust) {
                    auto kern = k_ntt_small3<C32, FWD, kFastTrust>;
                    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)))
                        return e;
                    kern<<<grid3, 1024, smem3, st>>>(tmap, data, tab, (uint32_t)batch, list);
                    *launches += 1;
                    return cudaGetLastError();
                }
                auto kern = k_ntt_small3<C32, FWD, kFastVote>;
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3))) return e;
                kern<<<grid3, 1024, smem3, st>>>(tmap, data, tab, (uint32_t)batch, list);
                if ((e = cudaGetLastError())) return e;
                *launches += 2;
                return launch_mode<C, FWD, kExactList>(tmap, smap, data, tab, batch, list, st, stride, offset);
            }
            if (tab.small_ok == 2) {
                const size_t smem2 = Small2Plan<C32>::BYTES;
                if (trust) {
                    auto kern = k_ntt_small2<C32, FWD, kFastTrust>;
                    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)))
                        return e;
                    kern<<<persistent_grid((const void*)kern, C32::NT, smem2, batch), C32::NT, smem2, st>>>(
                        data, tab, (uint32_t)batch, list);
                    *launches += 1;
                    return cudaGetLastError();
                }
                auto kern = k_ntt_small2<C32, FWD, kFastVote>;
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2))) return e;
                kern<<<persistent_grid((const void*)kern, C32::NT, smem2, batch), C32::NT, smem2, st>>>(
                    data, tab, (uint32_t)batch, list);
                if ((e = cudaGetLastError())) return e;
                *launches += 2;
                return launch_mode<C, FWD, kExactList>(tmap, smap, data, tab, batch, list, st, stride, offset);
            }
#endif
            const size_t smem = SmallPlan<C32>::BYTES;
            CUtensorMap smap32;
            const int tma_store = FWD && g_small_tma_store;
            if ((e = make_rows32_store_tmap(&smap32, data, batch, C::LOGN)) != cudaSuccess) return e;
            if (trust) {
                auto kern = k_ntt_small<C, C32, FWD, kFastTrust>;
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
                e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, C32::NT, smem, batch), C32::NT, smem, st, tmap, smap32, data, tab, (uint32_t)batch, list, tma_store);
                *launches += 1;
                return e != cudaSuccess ? e : cudaGetLastError();
            }
            auto kern = k_ntt_small<C, C32, FWD, kFastVote>;
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
            e = launch_dep(kern, (unsigned)persistent_grid((const void*)kern, C32::NT, smem, batch), C32::NT, smem, st, tmap, smap32, data, tab, (uint32_t)batch, list, tma_store);
            if (e != cudaSuccess || (e = cudaGetLastError())) return e;
            *launches += 2;
            return launch_mode<C, FWD, kExactList>(tmap, smap, data, tab, batch, list, st, stride, offset);
        }
    }
    if (trust) {
        *launches += 1;
        return launch_mode<C, FWD, kFastTrust>(tmap, smap, data, tab, batch, list, st, stride, offset);
    }
    bool folded = false;
    if ((e = launch_mode<C, FWD, kFastVote>(tmap, smap, data, tab, batch, list, st, stride, offset, &folded))) return e;
    if (folded || g_debug_skip_list) {   // (debug_skip_list: measurement only, deferred items are dropped)
        *launches += 1;
        return cudaSuccess;
    }
    // polynomials with out-of-contract words (none in normal use): exact pass
    // over the deferred list; exits at once when the list is empty
    *launches += 2;
    return launch_mode<C, FWD, kExactList>(tmap, smap, data, tab, batch, list, st, stride, offset);
}

#ifdef HB_EXPERIMENTAL_VARIANTS
#define HB_DISPATCH_CFG(logn, variant, CALL)                                   \
    switch (logn) {                                                            \
        case 10: { using C = NttCfg<10, 4>; CALL; } break;                     \
        case 11: { using C = NttCfg<11, 4>; CALL; } break;                     \
        case 12: { using C = NttCfg<12, 4>; CALL; } break;                     \
        case 13:                                                               \
            if (((variant) & 1) == 1) { using C = NttCfg<13, 5>; CALL; }       \
            else { using C = NttCfg<13, 4>; CALL; }                            \
            break;                                                             \
        case 14:                                                               \
            if (((variant) & 1) == 1) { using C = NttCfg<14, 5>; CALL; }       \
            else { using C = NttCfg<14, 4>; CALL; }                            \
            break;                                                             \
        default: break;                                                        \
    }
#else
// default build: N = 16384 only with 32 words per thread (N = 8192 keeps both: the plain calls use 32
// words per thread, the keyswitch stages 16)
#define HB_DISPATCH_CFG(logn, variant, CALL)                                   \
    switch (logn) {                                                            \
        case 10: { using C = NttCfg<10, 4>; CALL; } break;                     \
        case 11: { using C = NttCfg<11, 4>; CALL; } break;                     \
        case 12: { using C = NttCfg<12, 4>; CALL; } break;                     \
        case 13:                                                               \
            if (((variant) & 1) == 1) { using C = NttCfg<13, 5>; CALL; }       \
            else { using C = NttCfg<13, 4>; CALL; }                            \
            break;                                                             \
        case 14: { using C = NttCfg<14, 5>; CALL; } break;                     \
        default: break;                                                        \
    }
#endif

}  // namespace hb
