// ntt_kernels_inv.cu -- inverse-transform launchers (kernels: ntt_launch.cuh, ntt_block.cuh).
// Replaces device/inv_ntt.cpp:82-607 of the reference.
#include "ntt_launch.cuh"

namespace hb {

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
    job.n_items = (uint32_t)batch;
    if (tab.inv_fast_ok && tab.fp64_ok) {
        auto kern = k_ntt_inv_mul<C, kFastTrust, true>;   // canonical products into the FP64-pipe butterflies
        const size_t smemd = ntt_smem_bytes_fp64_plain<C>();
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemd))) return e;
        kern<<<persistent_grid((const void*)kern, C::NT, smemd, batch), C::NT, smemd, st>>>(tmap, job, (uint32_t)batch);
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
