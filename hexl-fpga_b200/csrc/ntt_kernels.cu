// ntt_kernels.cu -- batched forward / inverse negacyclic NTT kernels (sm_100a).
// Persistent CTAs, polynomial resident in shared memory, TMA prefetch of the
// next polynomial (see ntt_block.cuh).  Replaces device/fwd_ntt.cpp:81-646 and
// device/inv_ntt.cpp:82-607 of the reference.
#include <mutex>

#include "launch.h"

namespace hb {

int g_warp_tail = 1;         // FP64-pipe forward kernel with warp-dealt tail rows (one block barrier per transform instead of three), option "warp_tail"
int g_small_tma_store = 0;   // small-modulus forward epilogue through TMA stores (option "small_tma_store"): measured 5% slower than the coalesced register stores (slice reuse waits on the store engine), off by default

// ---- plain batched transform, in place ------------------------------------
template <class C>
struct JobPlain {
    uint64_t* data;
    ModTab tab;
    HB_D uint32_t src_row(uint32_t item) const { return item * (C::N / 16); }
    HB_D const ModTab& mod(uint32_t) const { return tab; }
    HB_D XfIdent xf(uint32_t) const { return XfIdent(); }
};
template <class C>
struct JobFwd : JobPlain<C> {
    HB_D OfRows of(uint32_t item, const CUtensorMap* smap) const {
        return OfRows{this->data + (size_t)item * C::N, smap, item * (C::N / 16)};
    }
};
template <class C>
struct JobInv : JobPlain<C> {
    HB_D OfWords of(uint32_t item, const CUtensorMap*) const { return OfWords{this->data + (size_t)item * C::N}; }
};

template <class C, int MODE, bool FP64 = false>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ntt_fwd(const __grid_constant__ CUtensorMap tmap,
                                                   const __grid_constant__ CUtensorMap smap, const JobFwd<C> job,
                                                   uint32_t n_items, uint32_t* list) {
    ntt_persistent<C, true, MODE, JobFwd<C>, false, FP64>(&tmap, &smap, job, n_items, list);
}
template <class C, int MODE, bool LAZY = false, bool FP64 = false>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ntt_inv(const __grid_constant__ CUtensorMap tmap, const JobInv<C> job,
                                                   uint32_t n_items, uint32_t* list) {
    ntt_persistent<C, false, MODE, JobInv<C>, LAZY, FP64>(&tmap, nullptr, job, n_items, list);
}

// inverse transform of NTT(a) (.) NTT(b): the fused tail of a polynomial multiply
template <class C>
struct JobInvMul : JobPlain<C> {
    const uint64_t* other;
    Divisor dv;
    HB_D XfMulGlobal xf(uint32_t item) const { return XfMulGlobal{other + (size_t)item * C::N, dv}; }
    HB_D OfWords of(uint32_t item, const CUtensorMap*) const { return OfWords{this->data + (size_t)item * C::N}; }
};
template <class C, int MODE, bool FP64 = false>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ntt_inv_mul(const __grid_constant__ CUtensorMap tmap,
                                                                   const JobInvMul<C> job, uint32_t n_items) {
    ntt_persistent<C, false, MODE, JobInvMul<C>, false, FP64>(&tmap, nullptr, job, n_items, nullptr);
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

// the configuration with warp-dealt tail rows, where the shape allows it
template <class C>
struct WarpTailCfg {
    using type = C;
};
template <>
struct WarpTailCfg<NttCfg<14, 5, 4, 0>> {
    using type = NttCfg<14, 5, 4, 1>;
};

// ---- packed twiddle builder -------------------------------------------------
template <class C>
__global__ void k_pack_twiddles(const uint64_t* __restrict__ roots, const uint64_t* __restrict__ precon,
                                TwPair* __restrict__ fwd_out, const uint64_t* __restrict__ inv_roots,
                                const uint64_t* __restrict__ precon_inv, TwPair* __restrict__ inv_out,
                                uint32_t* __restrict__ zero_count) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (zero_count && e == 0) *zero_count = 0;   // reset the deferred list of the call that follows
    if (fwd_out && e < (uint32_t)C::FWD_ENTRIES) {
        const int s = fwd_pack_src<C>(e);
        TwPair t = {0, 0};
        if (s >= 0) t = TwPair{roots[s], precon[s]};
        fwd_out[e] = t;
    }
    if (inv_out && e < (uint32_t)C::INV_ENTRIES) {
        const int s = inv_pack_src<C>(e);
        TwPair t = {0, 0};
        if (s >= 0) t = TwPair{inv_roots[s], precon_inv[s]};
        inv_out[e] = t;
    }
}

// FP64-pipe tables: {centred root as a double, its correctly rounded quotient by q}
template <class C>
__global__ void k_pack_twiddles_fp64(const uint64_t* __restrict__ roots, TwPair* __restrict__ fwd_out,
                                     const uint64_t* __restrict__ inv_roots, TwPair* __restrict__ inv_out,
                                     uint64_t q) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    auto entry = [&](uint64_t r) {
        const double ws = fp_centred(r < q ? r : r % q, q);
        return TwPair{d2u(ws), d2u(fp_quot(ws, q))};
    };
    if (fwd_out && e < (uint32_t)C::FWD_ENTRIES) {
        const int s = fwd_pack_src<C>(e);
        fwd_out[e] = s >= 0 ? entry(roots[s]) : TwPair{0, 0};
    }
    if (inv_out && e < (uint32_t)C::INV_ENTRIES) {
        const int s = inv_pack_src<C>(e);
        inv_out[e] = s >= 0 ? entry(inv_roots[s]) : TwPair{0, 0};
    }
}

template <class C32>
__global__ void k_pack_twiddles32(const uint64_t* __restrict__ roots, const uint64_t* __restrict__ precon,
                                  Tw32* __restrict__ fwd_out, const uint64_t* __restrict__ inv_roots,
                                  const uint64_t* __restrict__ precon_inv, Tw32* __restrict__ inv_out) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (fwd_out && e < (uint32_t)C32::FWD_ENTRIES) {
        const int s = fwd_pack_src<C32>(e);
        Tw32 t = {0, 0};
        if (s >= 0) t = Tw32{(uint32_t)roots[s], (uint32_t)(precon[s] >> 32)};
        fwd_out[e] = t;
    }
    if (inv_out && e < (uint32_t)C32::INV_ENTRIES) {
        const int s = inv_pack_src<C32>(e);
        Tw32 t = {0, 0};
        if (s >= 0) t = Tw32{(uint32_t)inv_roots[s], (uint32_t)(precon_inv[s] >> 32)};
        inv_out[e] = t;
    }
}

// ---- host side ---------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// store map of the small-modulus forward epilogue: [polys * N/32 rows][2 halves][16 words],
// box = 16 words x 1 half x 32 rows (4 KiB, 128-byte swizzle)
static cudaError_t make_rows32_store_tmap(CUtensorMap* out, const void* base, uint64_t polys, uint32_t logn) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    const uint64_t rows = polys * ((1ull << logn) / 32);
    if (rows == 0 || rows >> 32) return cudaErrorInvalidValue;
    const cuuint64_t gdim[3] = {16, 2, rows};
    const cuuint64_t gstride[2] = {128, 256};
    const cuuint32_t box[3] = {16, 1, 32};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t make_poly_tmap(CUtensorMap* out, const void* base, uint64_t polys, uint32_t logn, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    const uint64_t rows_per_poly = (1ull << logn) / 16;
    const uint64_t rows = polys * rows_per_poly;
    if (rows == 0 || rows >> 32) return cudaErrorInvalidValue;
    const cuuint64_t gdim[2] = {16, rows};
    const cuuint64_t gstride[1] = {128};
    if (box_rows == 0) box_rows = (uint32_t)(rows_per_poly < 256 ? rows_per_poly : 256);
    const cuuint32_t box[2] = {16, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

int persistent_grid(const void* kernel, int threads, size_t smem, uint64_t items) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    uint64_t g = (uint64_t)sms * per_sm;
    return (int)(items < g ? items : g);
}

template <class C, bool FWD, int MODE>
static cudaError_t launch_mode(const CUtensorMap& tmap, const CUtensorMap& smap, uint64_t* base, const ModTab& tab,
                               uint64_t cnt, uint32_t* list, cudaStream_t st) {
    const size_t smem = ntt_smem_bytes<C>();
    cudaError_t e;
    if constexpr (FWD) {
        JobFwd<C> job;
        job.data = base;
        job.tab = tab;
        if constexpr (MODE == kFastVote || MODE == kFastTrust) {
            using CW = typename WarpTailCfg<C>::type;
            if (tab.fp64_ok && g_warp_tail && !std::is_same<CW, C>::value) {
                auto kern = k_ntt_fwd<CW, MODE, true>;
                const size_t smemw = ntt_smem_bytes<CW>();
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemw))) return e;
                JobFwd<CW> jobw;
                jobw.data = base;
                jobw.tab = tab;
                kern<<<persistent_grid((const void*)kern, CW::NT, smemw, cnt), CW::NT, smemw, st>>>(tmap, smap, jobw,
                                                                                                   (uint32_t)cnt, list);
                return cudaGetLastError();
            }
            if (tab.fp64_ok) {       // 36..51-bit modulus: butterflies on the FP64 pipe
                auto kern = k_ntt_fwd<C, MODE, true>;
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
                kern<<<persistent_grid((const void*)kern, C::NT, smem, cnt), C::NT, smem, st>>>(tmap, smap, job,
                                                                                               (uint32_t)cnt, list);
                return cudaGetLastError();
            }
        }
        auto kern = k_ntt_fwd<C, MODE>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
        kern<<<persistent_grid((const void*)kern, C::NT, smem, cnt), C::NT, smem, st>>>(tmap, smap, job, (uint32_t)cnt,
                                                                                       list);
    } else {
        JobInv<C> job;
        job.data = base;
        job.tab = tab;
        if constexpr (MODE == kFastVote || MODE == kFastTrust) {
            using CW = typename WarpTailCfg<C>::type;
            if (tab.fp64_ok && g_warp_tail && !std::is_same<CW, C>::value) {
                auto kern = k_ntt_inv<CW, MODE, false, true>;
                const size_t smemw = ntt_smem_bytes<CW>();
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemw))) return e;
                JobInv<CW> jobw;
                jobw.data = base;
                jobw.tab = tab;
                kern<<<persistent_grid((const void*)kern, CW::NT, smemw, cnt), CW::NT, smemw, st>>>(tmap, jobw,
                                                                                                   (uint32_t)cnt, list);
                return cudaGetLastError();
            }
            if (tab.fp64_ok) {
                auto kern = k_ntt_inv<C, MODE, false, true>;
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
                kern<<<persistent_grid((const void*)kern, C::NT, smem, cnt), C::NT, smem, st>>>(tmap, job, (uint32_t)cnt,
                                                                                               list);
                return cudaGetLastError();
            }
            if (tab.inv_lazy_ok) {   // q < 2^52: butterflies without per-stage corrections
                auto kern = k_ntt_inv<C, MODE, true>;
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
                kern<<<persistent_grid((const void*)kern, C::NT, smem, cnt), C::NT, smem, st>>>(tmap, job, (uint32_t)cnt,
                                                                                               list);
                return cudaGetLastError();
            }
        }
        auto kern = k_ntt_inv<C, MODE>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
        kern<<<persistent_grid((const void*)kern, C::NT, smem, cnt), C::NT, smem, st>>>(tmap, job, (uint32_t)cnt, list);
    }
    return cudaGetLastError();
}

// `list`: device scratch of 1 + batch words whose first word is zero on entry
// (launch_pack_twiddles resets it); `trust` skips the input-range vote.
// `src`: where the polynomials are read from (nullptr: in place, from `data`)
template <class C, bool FWD>
static cudaError_t launch_one(uint64_t* data, const ModTab& tab, uint64_t batch, bool trust, uint32_t* list,
                              cudaStream_t st, int* launches, const uint64_t* src = nullptr) {
    CUtensorMap tmap, smap;
    cudaError_t e;
    // the tensor map's row coordinate is 32 bits
    const uint64_t kMaxPolys = ((1ull << 32) - 1) / (C::N / 16);
    if (batch > kMaxPolys) return cudaErrorInvalidValue;
    if ((e = make_poly_tmap(&tmap, src ? src : data, batch, C::LOGN)) != cudaSuccess) return e;
    if ((e = make_poly_tmap(&smap, data, batch, C::LOGN, 32)) != cudaSuccess) return e;
    const bool fast = FWD ? tab.fwd_fast_ok : tab.inv_fast_ok;
    if (!fast) {
        *launches += 1;
        return launch_mode<C, FWD, kExactAll>(tmap, smap, data, tab, batch, list, st);
    }
    if constexpr (C::LOGN == 14 && C::LOGE == 5) {
        // q < 2^30: the 32-bit kernels (out-of-contract items still go to the
        // 64-bit exact kernel through the deferred list)
        if (tab.small_ok) {
            using C32 = NttCfg<14, 5, 5>;
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
                return launch_mode<C, FWD, kExactList>(tmap, smap, data, tab, batch, list, st);
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
                return launch_mode<C, FWD, kExactList>(tmap, smap, data, tab, batch, list, st);
            }
            const size_t smem = SmallPlan<C32>::BYTES;
            CUtensorMap smap32;
            const int tma_store = FWD && g_small_tma_store;
            if ((e = make_rows32_store_tmap(&smap32, data, batch, C::LOGN)) != cudaSuccess) return e;
            if (trust) {
                auto kern = k_ntt_small<C, C32, FWD, kFastTrust>;
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
                kern<<<persistent_grid((const void*)kern, C32::NT, smem, batch), C32::NT, smem, st>>>(
                    tmap, smap32, data, tab, (uint32_t)batch, list, tma_store);
                *launches += 1;
                return cudaGetLastError();
            }
            auto kern = k_ntt_small<C, C32, FWD, kFastVote>;
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
            kern<<<persistent_grid((const void*)kern, C32::NT, smem, batch), C32::NT, smem, st>>>(
                tmap, smap32, data, tab, (uint32_t)batch, list, tma_store);
            if ((e = cudaGetLastError())) return e;
            *launches += 2;
            return launch_mode<C, FWD, kExactList>(tmap, smap, data, tab, batch, list, st);
        }
    }
    if (trust) {
        *launches += 1;
        return launch_mode<C, FWD, kFastTrust>(tmap, smap, data, tab, batch, list, st);
    }
    if ((e = launch_mode<C, FWD, kFastVote>(tmap, smap, data, tab, batch, list, st))) return e;
    // polynomials with out-of-contract words (none in normal use): exact pass
    // over the deferred list; exits at once when the list is empty
    *launches += 2;
    return launch_mode<C, FWD, kExactList>(tmap, smap, data, tab, batch, list, st);
}

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

bool ntt_shape_supported(uint32_t logn) { return logn >= 10 && logn <= 14; }

size_t packed_fwd_entries(uint32_t logn, int variant) {
    HB_DISPATCH_CFG(logn, variant, return C::FWD_ENTRIES);
    return 0;
}
size_t packed_inv_entries(uint32_t logn, int variant) {
    HB_DISPATCH_CFG(logn, variant, return C::INV_ENTRIES);
    return 0;
}

template <class C>
static cudaError_t pack_one(const uint64_t* roots, const uint64_t* precon, TwPair* fwd_out, const uint64_t* inv_roots,
                            const uint64_t* precon_inv, TwPair* inv_out, uint32_t* zero_count, cudaStream_t st) {
    const int total = C::FWD_ENTRIES > C::INV_ENTRIES ? C::FWD_ENTRIES : C::INV_ENTRIES;
    k_pack_twiddles<C><<<(total + 255) / 256, 256, 0, st>>>(roots, precon, fwd_out, inv_roots, precon_inv, inv_out,
                                                            zero_count);
    return cudaGetLastError();
}

cudaError_t launch_pack_twiddles(uint32_t logn, int variant, const uint64_t* roots, const uint64_t* precon,
                                 TwPair* fwd_out, const uint64_t* inv_roots, const uint64_t* precon_inv,
                                 TwPair* inv_out, uint32_t* zero_count, cudaStream_t st) {
    HB_DISPATCH_CFG(logn, variant,
                    return pack_one<C>(roots, precon, fwd_out, inv_roots, precon_inv, inv_out, zero_count, st));
    return cudaErrorInvalidValue;
}

template <class C>
static cudaError_t pack_fp64_one(const uint64_t* roots, TwPair* fwd_out, const uint64_t* inv_roots, TwPair* inv_out,
                                 uint64_t q, cudaStream_t st) {
    const int total = C::FWD_ENTRIES > C::INV_ENTRIES ? C::FWD_ENTRIES : C::INV_ENTRIES;
    k_pack_twiddles_fp64<C><<<(total + 255) / 256, 256, 0, st>>>(roots, fwd_out, inv_roots, inv_out, q);
    return cudaGetLastError();
}
cudaError_t launch_pack_twiddles_fp64(uint32_t logn, int variant, const uint64_t* roots, TwPair* fwd_out,
                                      const uint64_t* inv_roots, TwPair* inv_out, uint64_t q, cudaStream_t st) {
    HB_DISPATCH_CFG(logn, variant, return pack_fp64_one<C>(roots, fwd_out, inv_roots, inv_out, q, st));
    return cudaErrorInvalidValue;
}

bool small_path_available(uint32_t logn, int variant) { return logn == 14 && (variant & 1) == 1; }
size_t packed32_fwd_entries() { return NttCfg<14, 5, 5>::FWD_ENTRIES; }
size_t packed32_inv_entries() { return NttCfg<14, 5, 5>::INV_ENTRIES; }
cudaError_t launch_pack_twiddles32(const uint64_t* roots, const uint64_t* precon, Tw32* fwd_out,
                                   const uint64_t* inv_roots, const uint64_t* precon_inv, Tw32* inv_out,
                                   cudaStream_t st) {
    using C32 = NttCfg<14, 5, 5>;
    const int total = C32::FWD_ENTRIES > C32::INV_ENTRIES ? C32::FWD_ENTRIES : C32::INV_ENTRIES;
    k_pack_twiddles32<C32><<<(total + 255) / 256, 256, 0, st>>>(roots, precon, fwd_out, inv_roots, precon_inv, inv_out);
    return cudaGetLastError();
}

cudaError_t launch_ntt_fwd(uint64_t* data, const ModTab& tab, uint32_t logn, uint64_t batch, int variant,
                           uint32_t* list, cudaStream_t st, int* launches, const uint64_t* src) {
    if (batch == 0) return cudaSuccess;
    const bool trust = (variant & 2) != 0;
    HB_DISPATCH_CFG(logn, variant, return (launch_one<C, true>(data, tab, batch, trust, list, st, launches, src)));
    return cudaErrorInvalidValue;
}

template <class C>
static cudaError_t launch_inv_mul_one(uint64_t* data, const uint64_t* other, const ModTab& tab, uint64_t batch,
                                      cudaStream_t st) {
    static_assert(SmemPlan<C>::kStagedStore || true, "");
    CUtensorMap tmap;
    cudaError_t e;
    const uint64_t kMaxPolys = ((1ull << 32) - 1) / (C::N / 16);
    if (batch > kMaxPolys) return cudaErrorInvalidValue;
    if ((e = make_poly_tmap(&tmap, data, batch, C::LOGN)) != cudaSuccess) return e;
    const size_t smem = ntt_smem_bytes<C>();
    JobInvMul<C> job;
    job.data = data;
    job.tab = tab;
    job.other = other;
    job.dv = make_divisor(tab.q);
    if (tab.inv_fast_ok && tab.fp64_ok) {
        auto kern = k_ntt_inv_mul<C, kFastTrust, true>;   // canonical products into the FP64-pipe butterflies
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
        kern<<<persistent_grid((const void*)kern, C::NT, smem, batch), C::NT, smem, st>>>(tmap, job, (uint32_t)batch);
    } else if (tab.inv_fast_ok) {
        auto kern = k_ntt_inv_mul<C, kFastTrust>;   // the products are canonical: no range vote needed
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
        kern<<<persistent_grid((const void*)kern, C::NT, smem, batch), C::NT, smem, st>>>(tmap, job, (uint32_t)batch);
    } else {
        auto kern = k_ntt_inv_mul<C, kExactAll>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
        kern<<<persistent_grid((const void*)kern, C::NT, smem, batch), C::NT, smem, st>>>(tmap, job, (uint32_t)batch);
    }
    return cudaGetLastError();
}

cudaError_t launch_ntt_inv_mul(uint64_t* data, const uint64_t* other, const ModTab& tab, uint32_t logn,
                               uint64_t batch, int variant, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    HB_DISPATCH_CFG(logn, variant, return (launch_inv_mul_one<C>(data, other, tab, batch, st)));
    return cudaErrorInvalidValue;
}
cudaError_t launch_ntt_inv(uint64_t* data, const ModTab& tab, uint32_t logn, uint64_t batch, int variant,
                           uint32_t* list, cudaStream_t st, int* launches) {
    if (batch == 0) return cudaSuccess;
    const bool trust = (variant & 2) != 0;
    HB_DISPATCH_CFG(logn, variant, return (launch_one<C, false>(data, tab, batch, trust, list, st, launches)));
    return cudaErrorInvalidValue;
}

}  // namespace hb
