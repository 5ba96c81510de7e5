// ntt_kernels.cu -- batched forward / inverse negacyclic NTT kernels (sm_100a).
// Persistent CTAs, polynomial resident in shared memory, TMA prefetch of the
// next polynomial (see ntt_block.cuh).  Replaces device/fwd_ntt.cpp:81-646 and
// device/inv_ntt.cpp:82-607 of the reference.
#include <mutex>
#include <vector>

#include "ntt_launch.cuh"

namespace hb {

int g_pdl = 1;               // plain NTT kernels launched with programmatic stream serialization (option "pdl", ntt_launch.cuh)
int g_debug_skip_list = 0;   // measurement only, see launch.h
int g_time_kernels = 0;      // measurement only, see launch.h
namespace {
std::mutex g_kt_mu;
struct TimedLaunch {
    cudaEvent_t a, b;
    unsigned grid;
};
std::vector<TimedLaunch> g_kt;
}  // namespace
void note_kernel_events(cudaEvent_t a, cudaEvent_t b, unsigned grid) {
    std::lock_guard<std::mutex> lk(g_kt_mu);
    g_kt.push_back({a, b, grid});
}
// durations (ms) of the launches timed since the last call, in launch order; the events are released
cudaError_t take_kernel_times(float* ms, uint64_t cap, uint64_t* count) {
    std::lock_guard<std::mutex> lk(g_kt_mu);
    cudaError_t err = cudaSuccess;
    uint64_t n = 0;
    for (const TimedLaunch& t : g_kt) {
        cudaError_t e = cudaEventSynchronize(t.b);
        float v = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&v, t.a, t.b);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        if (n < cap && ms) ms[n] = v;
        ++n;
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    g_kt.clear();
    *count = n;
    return err;
}
int g_warp_tail = 1;         // FP64-pipe forward kernel with warp-dealt tail rows (one block barrier per transform instead of three), option "warp_tail"
int g_small_tma_store = 0;   // small-modulus forward epilogue through TMA stores (option "small_tma_store"): measured 5% slower than the coalesced register stores (slice reuse waits on the store engine), off by default

// ---- packed twiddle builder -------------------------------------------------
template <class C>
__global__ void k_pack_twiddles(const uint64_t* __restrict__ roots, const uint64_t* __restrict__ precon,
                                TwPair* __restrict__ fwd_out, const uint64_t* __restrict__ inv_roots,
                                const uint64_t* __restrict__ precon_inv, TwPair* __restrict__ inv_out,
                                uint32_t* __restrict__ zero_count, const PackExtra x) {
    // launched with programmatic stream serialization itself (pack_one): this grid may be scheduled while the
    // kernel in front of it (the previous call's deferred-list pass) still runs, and waits here until that one has
    // finished -- it overwrites tables the kernels in front read
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // the transform kernel behind us may be scheduled now; it waits for this grid before it reads anything
    asm volatile("griddepcontrol.launch_dependents;");
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (zero_count && e == 0) zero_count[0] = zero_count[1] = 0;   // reset the deferred list (count, barrier word) of the call that follows
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
    // the same call's tables in the other two formats (one launch instead of two or three)
    auto entry_d = [&](uint64_t r) {
        const double ws = fp_centred(r < x.q ? r : r % x.q, x.q);
        return TwPair{d2u(ws), d2u(fp_quot(ws, x.q))};
    };
    if (x.fwd_d && e < (uint32_t)C::FWD_ENTRIES) {
        const int s = fwd_pack_src<C>(e);
        x.fwd_d[e] = s >= 0 ? entry_d(roots[s]) : TwPair{0, 0};
    }
    if (x.inv_d && e < (uint32_t)C::INV_ENTRIES) {
        const int s = inv_pack_src<C>(e);
        x.inv_d[e] = s >= 0 ? entry_d(inv_roots[s]) : TwPair{0, 0};
    }
    if constexpr (C::LOGN == 14 && C::LOGE == 5) {
        using C32 = NttCfg<14, 5, 5>;
        if (x.fwd32 && e < (uint32_t)C32::FWD_ENTRIES) {
            const int s = fwd_pack_src<C32>(e);
            x.fwd32[e] = s >= 0 ? Tw32{(uint32_t)roots[s], (uint32_t)(precon[s] >> 32)} : Tw32{0, 0};
        }
        if (x.inv32 && e < (uint32_t)C32::INV_ENTRIES) {
            const int s = inv_pack_src<C32>(e);
            x.inv32[e] = s >= 0 ? Tw32{(uint32_t)inv_roots[s], (uint32_t)(precon_inv[s] >> 32)} : Tw32{0, 0};
        }
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
cudaError_t make_rows32_store_tmap(CUtensorMap* out, const void* base, uint64_t polys, uint32_t logn) {
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
                            const uint64_t* precon_inv, TwPair* inv_out, uint32_t* zero_count, cudaStream_t st,
                            const PackExtra& x) {
    int total = C::FWD_ENTRIES > C::INV_ENTRIES ? C::FWD_ENTRIES : C::INV_ENTRIES;
    if (x.fwd32 || x.inv32) {
        using C32 = NttCfg<14, 5, 5>;
        if (!(C::LOGN == 14 && C::LOGE == 5)) return cudaErrorInvalidValue;
        const int t32 = C32::FWD_ENTRIES > C32::INV_ENTRIES ? C32::FWD_ENTRIES : C32::INV_ENTRIES;
        total = t32 > total ? t32 : total;
    }
    return launch_dep(k_pack_twiddles<C>, (unsigned)((total + 255) / 256), 256u, (size_t)0, st, roots, precon, fwd_out,
                      inv_roots, precon_inv, inv_out, zero_count, x);
}

cudaError_t launch_pack_twiddles(uint32_t logn, int variant, const uint64_t* roots, const uint64_t* precon,
                                 TwPair* fwd_out, const uint64_t* inv_roots, const uint64_t* precon_inv,
                                 TwPair* inv_out, uint32_t* zero_count, cudaStream_t st, const PackExtra& x) {
    HB_DISPATCH_CFG(logn, variant,
                    return pack_one<C>(roots, precon, fwd_out, inv_roots, precon_inv, inv_out, zero_count, st, x));
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

}  // namespace hb
