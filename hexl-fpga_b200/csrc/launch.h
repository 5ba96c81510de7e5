// launch.h -- host-callable launchers of the sm_100a kernels (internal; the
// public boundary is include/hexl_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ntt_block.cuh"

namespace hb {

// Shape of one keyswitch (reference host/inc/hexl-fpga.h:54-64 arguments) plus
// the device-resident constants a plan owns.
struct KsDev {
    uint32_t logn, D, K, R;
    const ModTab* tabs;     // [K]
    const Divisor* divs;    // [K]
    const uint64_t* keys;   // [D][2][K][N]  == k_switch_keys[j][(c*K+i)*N + l]
    const uint64_t* msf;    // [K] modswitch_factors[i] mod q_i
    const uint64_t* msf_p;  // [K] Shoup factors of msf
};

// variant: 0 = LOGE 4 (16 words / thread), 1 = LOGE 5 (32 words / thread)
cudaError_t launch_ntt_fwd(uint64_t* data, const ModTab& tab, uint32_t logn, uint64_t batch,
                           int variant, cudaStream_t st);
cudaError_t launch_ntt_inv(uint64_t* data, const ModTab& tab, uint32_t logn, uint64_t batch,
                           int variant, cudaStream_t st);

cudaError_t launch_dyadic(uint64_t* res, const uint64_t* op1, const uint64_t* op2, uint64_t n,
                          const uint64_t* moduli, uint64_t n_moduli, uint64_t batch,
                          int moduli_per_item, cudaStream_t st);

// keyswitch stages over a chunk of `items` ciphertexts (scratch layouts in
// keyswitch_kernels.cu)
cudaError_t launch_ks_chunk(const KsDev& ks, uint64_t* result, const uint64_t* t_target,
                            uint64_t items, uint64_t* U, uint64_t* V, uint64_t* ACC,
                            cudaStream_t st);

bool ntt_shape_supported(uint32_t logn);

}  // namespace hb
