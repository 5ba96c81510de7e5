// keyswitch_kernels.cu -- RNS keyswitch (sm_100a), staged version.
//
// Implements SURVEY.md Appendix A.4 == the reference's device pipeline
// (device/keyswitch/load.hpp -> intt_core.hpp -> intt1_redu.hpp -> ntt_core.hpp
// -> dyadmult.hpp -> intt2_redu.hpp -> ntt2.hpp -> ms.hpp) plus the host-side
// accumulate into `result` (host/src/fpga.cpp:441-475), for a chunk of items:
//
//   S1  U[b][j]      = INTT_{q_j}(t[b][j])                          grid items*D
//   S2  V[b][r][j]   = NTT_{q_idx(r)}(U[b][j] mod q_idx(r)), j != r grid items*R*D
//   S3  ACC[b][c][r] = sum_j V[b][r][j] (.) key[j][c][idx(r)]       elementwise
//                      (digit j == r is taken from t[b][j] directly:
//                       NTT(INTT(t_j)) == t_j)
//   S4  ACC[b][c][D] = INTT_{q_k}(ACC[b][c][D])                     grid items*2
//   S5  w = NTT_{q_i}(round/convert(ACC[b][c][D]));
//       result[b][c][i] += (ACC[b][c][i] - w) * msf_i  (mod q_i)    grid items*2*D
//
// idx(r) = r for r < D and K-1 (the special prime) for r == D.  Every stage
// output is canonical in [0,q), so the result is the unique value the
// reference pipeline produces.
#include "launch.h"

namespace hb {

// ---- S1 -------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NT) k_ks_intt1(KsDev ks, const uint64_t* t_target, uint64_t* U) {
    extern __shared__ __align__(1024) uint64_t sm[];
    const uint32_t j = blockIdx.x % ks.D;
    const ModTab tab = ks.tabs[j];
    const size_t off = (size_t)blockIdx.x * C::N;
    ntt_inv_block<C>(sm, t_target + off, U + off, XfIdent(), OfStore1(), tab);
}

// ---- S2 -------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NT) k_ks_ntt1(KsDev ks, const uint64_t* U, uint64_t* V) {
    extern __shared__ __align__(1024) uint64_t sm[];
    const uint32_t j = blockIdx.x % ks.D;
    const uint32_t r = (blockIdx.x / ks.D) % ks.R;
    const uint32_t b = blockIdx.x / (ks.D * ks.R);
    if (j == r) return;  // S3 reads t_target for this digit
    const uint32_t idx = (r == ks.D) ? ks.K - 1 : r;
    const ModTab tab = ks.tabs[idx];
    XfReduce xf = {tab.q, tab.mu};
    ntt_fwd_block<C>(sm, U + ((size_t)b * ks.D + j) * C::N, V + (size_t)blockIdx.x * C::N, xf,
                     OfStore16(), tab);
}

// ---- S3 -------------------------------------------------------------------
// grid: (N / (256*2), R, items); each thread owns two adjacent coefficients.
__global__ void __launch_bounds__(256)
k_ks_mac(KsDev ks, const uint64_t* __restrict__ t_target, const uint64_t* __restrict__ V,
         uint64_t* __restrict__ ACC) {
    const uint32_t N = 1u << ks.logn;
    const uint32_t l = (blockIdx.x * 256 + threadIdx.x) * 2;
    const uint32_t r = blockIdx.y, b = blockIdx.z;
    const uint32_t idx = (r == ks.D) ? ks.K - 1 : r;
    const Divisor dv = ks.divs[idx];
    uint64_t a0[2] = {0, 0}, a1[2] = {0, 0};
    for (uint32_t j = 0; j < ks.D; ++j) {
        const uint64_t* op = (j == r) ? t_target + ((size_t)b * ks.D + j) * N
                                      : V + (((size_t)b * ks.R + r) * ks.D + j) * N;
        const uint64_t* k0 = ks.keys + (((size_t)j * 2 + 0) * ks.K + idx) * N;
        const uint64_t* k1 = ks.keys + (((size_t)j * 2 + 1) * ks.K + idx) * N;
        uint64_t x[2], u[2], w[2];
        ld2(op + l, x[0], x[1]);
        ld2(k0 + l, u[0], u[1]);
        ld2(k1 + l, w[0], w[1]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const uint64_t xe = mod64(x[e], dv);
            a0[e] = add_mod(a0[e], mulmod_reduced(xe, mod64(u[e], dv), dv), dv.q);
            a1[e] = add_mod(a1[e], mulmod_reduced(xe, mod64(w[e], dv), dv), dv.q);
        }
    }
    st2(ACC + (((size_t)b * 2 + 0) * ks.R + r) * N + l, a0[0], a0[1]);
    st2(ACC + (((size_t)b * 2 + 1) * ks.R + r) * N + l, a1[0], a1[1]);
}

// ---- S4 -------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NT) k_ks_intt2(KsDev ks, uint64_t* ACC) {
    extern __shared__ __align__(1024) uint64_t sm[];
    const ModTab tab = ks.tabs[ks.K - 1];
    uint64_t* p = ACC + ((size_t)blockIdx.x * ks.R + ks.D) * C::N;  // [b][c][D]
    ntt_inv_block<C>(sm, p, p, XfIdent(), OfStore1(), tab);
}

// ---- S5 -------------------------------------------------------------------
// tail-pass output functor: modswitch + accumulate into result
// (device/keyswitch/ms.hpp:68-83, host/src/fpga.cpp:453-468)
struct OfKsFinal {
    const uint64_t* acc;   // ACC[b][c][i]
    uint64_t* result;      // result[b][c][i]
    uint64_t q, msf, msf_p;
    HB_D void operator()(uint64_t*, uint32_t off, const uint64_t (&v)[16]) const {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint64_t a[2], r[2];
            ld2(acc + off + 2 * c, a[0], a[1]);
            ld2(result + off + 2 * c, r[0], r[1]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint64_t d = sub_mod(a[e], v[2 * c + e], q);
                uint64_t o = mul_lazy(d, msf, msf_p, q);
                o -= (o >= q) ? q : 0;
                r[e] = add_mod(r[e], o, q);
            }
            st2(result + off + 2 * c, r[0], r[1]);
        }
    }
};

template <class C>
__global__ void __launch_bounds__(C::NT) k_ks_ntt2(KsDev ks, const uint64_t* ACC, uint64_t* result) {
    extern __shared__ __align__(1024) uint64_t sm[];
    const uint32_t i = blockIdx.x % ks.D;
    const uint32_t bc = blockIdx.x / ks.D;  // b*2 + c
    const ModTab tab = ks.tabs[i];
    const uint64_t qk = ks.tabs[ks.K - 1].q;
    const uint64_t qk_half = qk >> 1;
    XfKsRound xf = {qk, qk_half, tab.q, tab.mu, tab.q - barrett_reduce64(qk_half, tab.q, tab.mu)};
    OfKsFinal of = {ACC + ((size_t)bc * ks.R + i) * C::N,
                    result + ((size_t)bc * ks.D + i) * C::N, tab.q, ks.msf[i], ks.msf_p[i]};
    ntt_fwd_block<C>(sm, ACC + ((size_t)bc * ks.R + ks.D) * C::N, nullptr, xf, of, tab);
}

template <class C>
static cudaError_t ks_chunk(const KsDev& ks, uint64_t* result, const uint64_t* t_target,
                            uint64_t items, uint64_t* U, uint64_t* V, uint64_t* ACC,
                            cudaStream_t st) {
    const size_t smem = (size_t)C::N * sizeof(uint64_t);
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_ks_intt1<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
    if ((e = cudaFuncSetAttribute(k_ks_ntt1<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
    if ((e = cudaFuncSetAttribute(k_ks_intt2<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
    if ((e = cudaFuncSetAttribute(k_ks_ntt2<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
    k_ks_intt1<C><<<(unsigned)(items * ks.D), C::NT, smem, st>>>(ks, t_target, U);
    k_ks_ntt1<C><<<(unsigned)(items * ks.R * ks.D), C::NT, smem, st>>>(ks, U, V);
    dim3 g(C::N / 512, ks.R, (unsigned)items);
    k_ks_mac<<<g, 256, 0, st>>>(ks, t_target, V, ACC);
    k_ks_intt2<C><<<(unsigned)(items * 2), C::NT, smem, st>>>(ks, ACC);
    k_ks_ntt2<C><<<(unsigned)(items * 2 * ks.D), C::NT, smem, st>>>(ks, ACC, result);
    return cudaGetLastError();
}

cudaError_t launch_ks_chunk(const KsDev& ks, uint64_t* result, const uint64_t* t_target,
                            uint64_t items, uint64_t* U, uint64_t* V, uint64_t* ACC,
                            cudaStream_t st) {
    if (items == 0) return cudaSuccess;
    if (items > 65535) return cudaErrorInvalidValue;  // gridDim.z of the MAC stage
    switch (ks.logn) {
        case 10: return ks_chunk<NttCfg<10, 4>>(ks, result, t_target, items, U, V, ACC, st);
        case 11: return ks_chunk<NttCfg<11, 4>>(ks, result, t_target, items, U, V, ACC, st);
        case 12: return ks_chunk<NttCfg<12, 4>>(ks, result, t_target, items, U, V, ACC, st);
        case 13: return ks_chunk<NttCfg<13, 4>>(ks, result, t_target, items, U, V, ACC, st);
        case 14: return ks_chunk<NttCfg<14, 4>>(ks, result, t_target, items, U, V, ACC, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace hb
