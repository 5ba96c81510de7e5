// ntt_kernels.cu -- batched forward / inverse negacyclic NTT kernels (sm_100a).
// One CTA per polynomial, polynomial resident in shared memory (see
// ntt_core.cuh).  Replaces device/fwd_ntt.cpp:81-497 and
// device/inv_ntt.cpp:82-442 of the reference.
#include "launch.h"

namespace hb {

template <class C>
__global__ void __launch_bounds__(C::NT) k_ntt_fwd(uint64_t* data, const ModTab tab) {
    extern __shared__ __align__(1024) uint64_t sm[];
    uint64_t* poly = data + (size_t)blockIdx.x * C::N;
    ntt_fwd_block<C>(sm, poly, poly, XfIdent(), OfStore16(), tab);
}

template <class C>
__global__ void __launch_bounds__(C::NT) k_ntt_inv(uint64_t* data, const ModTab tab) {
    extern __shared__ __align__(1024) uint64_t sm[];
    uint64_t* poly = data + (size_t)blockIdx.x * C::N;
    ntt_inv_block<C>(sm, poly, poly, XfIdent(), OfStore1(), tab);
}

template <class C, bool FWD>
static cudaError_t launch_one(uint64_t* data, const ModTab& tab, uint64_t batch, cudaStream_t st) {
    auto kern = FWD ? k_ntt_fwd<C> : k_ntt_inv<C>;
    const size_t smem = (size_t)C::N * sizeof(uint64_t);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // grid.x is limited to 2^31-1 CTAs; batches beyond that are split.
    const uint64_t kMaxGrid = 1u << 30;
    for (uint64_t off = 0; off < batch; off += kMaxGrid) {
        uint64_t cnt = batch - off < kMaxGrid ? batch - off : kMaxGrid;
        kern<<<(unsigned)cnt, C::NT, smem, st>>>(data + off * C::N, tab);
    }
    return cudaGetLastError();
}

template <bool FWD>
static cudaError_t dispatch(uint64_t* data, const ModTab& tab, uint32_t logn, uint64_t batch,
                            int variant, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    switch (logn) {
        case 10: return launch_one<NttCfg<10, 4>, FWD>(data, tab, batch, st);
        case 11: return launch_one<NttCfg<11, 4>, FWD>(data, tab, batch, st);
        case 12: return launch_one<NttCfg<12, 4>, FWD>(data, tab, batch, st);
        case 13: return launch_one<NttCfg<13, 4>, FWD>(data, tab, batch, st);
        case 14:
            return variant == 1 ? launch_one<NttCfg<14, 5>, FWD>(data, tab, batch, st)
                                : launch_one<NttCfg<14, 4>, FWD>(data, tab, batch, st);
        default: return cudaErrorInvalidValue;
    }
}

bool ntt_shape_supported(uint32_t logn) { return logn >= 10 && logn <= 14; }

cudaError_t launch_ntt_fwd(uint64_t* data, const ModTab& tab, uint32_t logn, uint64_t batch,
                           int variant, cudaStream_t st) {
    return dispatch<true>(data, tab, logn, batch, variant, st);
}
cudaError_t launch_ntt_inv(uint64_t* data, const ModTab& tab, uint32_t logn, uint64_t batch,
                           int variant, cudaStream_t st) {
    return dispatch<false>(data, tab, logn, batch, variant, st);
}

}  // namespace hb
