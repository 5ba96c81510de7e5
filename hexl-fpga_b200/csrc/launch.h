// launch.h -- host-callable launchers of the sm_100a kernels (internal; the
// public boundary is include/hexl_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ntt_block.cuh"

namespace hb {

// Division of a 32-bit index by an invariant divisor: x / d = hi64(x * m) with m = floor(2^64 / d) + 1, exact for
// every 32-bit x (the error term x * (m - 2^64/d) / 2^64 stays below 1/d as long as x * d < 2^64).  The keyswitch
// jobs decode (item, modulus, digit) from a running index several times per transform in every thread; the
// compiler's division by a run-time value is ~25 instructions each (I2F / MUFU.RCP / F2I and two fix-up
// branches), which added up to 5 % of stage S2's instruction stream.
struct FastDiv {
    uint32_t d = 0;
    uint64_t m = 0;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    f.m = d >= 2 ? ~(uint64_t)0 / d + 1 : 0;
    return f;
}
HB_HD uint32_t fdiv(uint32_t x, const FastDiv& f) {
    if (f.d <= 1) return x;
#if defined(__CUDA_ARCH__)
    return (uint32_t)__umul64hi((uint64_t)x, f.m);
#else
    return (uint32_t)(((unsigned __int128)x * f.m) >> 64);
#endif
}
// x / d with the prepared divisor when it is the one asked for (jobs built for another chunk size fall back)
HB_HD uint32_t fdiv(uint32_t x, uint32_t d, const FastDiv& f) { return f.d == d ? fdiv(x, f) : x / d; }

// Shape of one keyswitch (reference host/inc/hexl-fpga.h:54-64 arguments) plus
// the device-resident constants a plan owns.
struct KsDev {
    uint32_t logn, D, K, R;
    uint32_t fast_ok;       // every modulus admits the fast arithmetic (q < 2^58)
    uint32_t fp64_ok;       // ... and the FP64-pipe butterflies (2^36 <= q <= 2^53/3, tables in tabs[])
    // FP64 path with all moduli within 25 % of each other (same-size primes, the usual case): a digit
    // reduced mod q_j is below 1.25 q_r for every target modulus, i.e. already inside the forward
    // transform's input contract, and NTT_r(x) = NTT_r(x mod q_r): stage S2 skips its base conversion
    uint32_t s2_no_reduce;
    uint32_t fp64_alt_ok;   // every modulus <= 2^51 (1 + 1/32): forward stages correct every other stage (modarith.cuh)
    const ModTab* tabs;     // [K]
    const Divisor* divs;    // [K]
    const uint64_t* keys;   // [D][2][K][N]  == k_switch_keys[j][(c*K+i)*N + l]
    const TwPair* keys_sh;  // same index space: {key mod q_i, Shoup factor}; null if !fast_ok
    const uint64_t* msf;    // [K] modswitch_factors[i] mod q_i
    const uint64_t* msf_p;  // [K] Shoup factors of msf
    const void* keys_fused; // key quads in the layout of the fused kernel (keyswitch_fused.cu), or null
    // same index space as keys_sh: {centred key mod q_i, its quotient by q_i} as doubles, for the multiply-
    // accumulate on the FP64 pipe (k_ks_mac_fp64); null unless fp64_alt_ok
    const TwPair* keys_fp;
    const double* msf_fp;   // [2K]: centred msf_i and its quotient by q_i (FP64 epilogue of stage S5), or null
    // prepared divisors of the index arithmetic: D, D - 1, D * D (plan), and of the chunk in flight: its item
    // count B, B * (D - 1) and 2 * B (set by ks_chunk)
    FastDiv fD, fDm1, fDD, fB, fBlo, fB2;
    uint32_t walk_blocked;  // modulus-major stages: every CTA takes one contiguous stretch of the order (option "ks_blocked")
};

bool ntt_shape_supported(uint32_t logn);
// the NttCfg variant the keyswitch kernels are instantiated with (their packed tables must match)
inline int ks_variant_for(uint32_t logn) { return logn == 14 ? 1 : 0; }
// number of 16-byte entries of the packed forward / inverse twiddle tables
size_t packed_fwd_entries(uint32_t logn, int variant);
size_t packed_inv_entries(uint32_t logn, int variant);
// interleave caller tables (roots/precon and/or inv_roots/precon_inv, n words
// each, device pointers) into the packed per-group layout of ntt_core.cuh
// optional extra outputs of the same launch: the small-modulus (uint32) tables (n = 16384, 32 words per
// thread only) and the FP64-pipe tables {centred root, root / q} of modulus q
struct PackExtra {
    Tw32* fwd32 = nullptr;
    Tw32* inv32 = nullptr;
    TwPair* fwd_d = nullptr;
    TwPair* inv_d = nullptr;
    uint64_t q = 0;
};
cudaError_t launch_pack_twiddles(uint32_t logn, int variant, const uint64_t* roots, const uint64_t* precon,
                                 TwPair* fwd_out, const uint64_t* inv_roots, const uint64_t* precon_inv,
                                 TwPair* inv_out, uint32_t* zero_count, cudaStream_t st,
                                 const PackExtra& extra = PackExtra());

// FP64-pipe tables (modarith.cuh): same packed geometry, entries {centred root, root / q} as doubles
cudaError_t launch_pack_twiddles_fp64(uint32_t logn, int variant, const uint64_t* roots, TwPair* fwd_out,
                                      const uint64_t* inv_roots, TwPair* inv_out, uint64_t q, cudaStream_t st);

// small-modulus (q < 2^30) 32-bit kernels: available for N = 16384 with the 32-word configuration
bool small_path_available(uint32_t logn, int variant);
size_t packed32_fwd_entries();
size_t packed32_inv_entries();
cudaError_t launch_pack_twiddles32(const uint64_t* roots, const uint64_t* precon, Tw32* fwd_out,
                                   const uint64_t* inv_roots, const uint64_t* precon_inv, Tw32* inv_out,
                                   cudaStream_t st);

// 2-D tensor map over `polys` polynomials of 2^logn words starting at `base`
// (rows of 16 words, 128-byte swizzle), for the kernels' TMA loads
// box_rows = 0: the load box (min(256, rows per polynomial)); 32: the per-warp store box
cudaError_t make_poly_tmap(CUtensorMap* out, const void* base, uint64_t polys, uint32_t logn,
                           uint32_t box_rows = 0);

// variant bit 0: 32 words / thread at N = 16384 (else 16); bit 1: trust the
// caller about the input range (no vote).  `list`: 1 + batch words of device
// scratch, word 0 zero on entry (the deferred list of out-of-contract items).
// `src` (forward only): read the polynomials from there and write the transforms to `data`
cudaError_t launch_ntt_fwd(uint64_t* data, const ModTab& tab, uint32_t logn, uint64_t batch, int variant,
                           uint32_t* list, cudaStream_t st, int* launches, const uint64_t* src = nullptr);
// data[b] <- INTT(data[b] (.) other[b]): inverse transform whose first pass multiplies
// by the second NTT-form operand on the fly (fused polynomial multiply)
cudaError_t launch_ntt_inv_mul(uint64_t* data, const uint64_t* other, const ModTab& tab, uint32_t logn,
                               uint64_t batch, int variant, cudaStream_t st);
cudaError_t launch_ntt_inv(uint64_t* data, const ModTab& tab, uint32_t logn, uint64_t batch, int variant,
                           uint32_t* list, cudaStream_t st, int* launches);

// single-launch polynomial multiply (polymul_fused.cu): N = 16384, FP64-contract moduli
extern int g_polymul_fused;
bool polymul_fused_available(const ModTab& tab, uint32_t logn);
cudaError_t launch_polymul_fused(uint64_t* res, const uint64_t* a, const uint64_t* b, const ModTab& tab, uint64_t batch,
                                 uint32_t* list, cudaStream_t st);
cudaError_t launch_polymul_deferred(uint64_t* res, uint64_t* tb, const uint64_t* a, const uint64_t* b, const ModTab& tab,
                                    uint64_t batch, uint32_t* list, cudaStream_t st);

// N = 32768 (ntt_big.cu): table split, the streaming first / last stage, the half-size sub-transforms
cudaError_t launch_big_split(bool fwd, const uint64_t* tab0, const uint64_t* tab1, uint64_t* sub, uint32_t n_half,
                             cudaStream_t st);
cudaError_t launch_big_stage0_fwd(uint64_t* data, const uint64_t* roots, const uint64_t* precon, uint64_t q,
                                  uint32_t n_half, uint64_t batch, cudaStream_t st);
cudaError_t launch_big_last_inv(uint64_t* data, const uint64_t* inv_roots, uint64_t q, uint64_t inv_n, uint64_t inv_n_w,
                                uint32_t n_half, uint64_t batch, cudaStream_t st);
cudaError_t launch_ntt_half(bool fwd, uint64_t* data, const ModTab& tab, uint32_t logn_half, uint64_t batch, int variant,
                            uint32_t* list, cudaStream_t st, int* launches, uint32_t half);

size_t dyadic_scratch_bytes(uint64_t n_moduli, uint64_t batch, int moduli_per_item);
// `scratch`: dyadic_scratch_bytes() of device memory, 16-byte aligned (per-modulus reciprocals)
cudaError_t launch_dyadic(uint64_t* res, const uint64_t* op1, const uint64_t* op2, uint64_t n,
                          const uint64_t* moduli, uint64_t n_moduli, uint64_t batch, int moduli_per_item,
                          void* scratch, cudaStream_t st);

// keyswitch stages over a chunk of `items` ciphertexts (scratch layouts in
// keyswitch_kernels.cu); returns the number of kernel launches in *launches
extern int g_ks_mac_items;
extern int g_small_tma_store;
extern int g_warp_tail;
extern int g_pdl;
extern int g_time_kernels;      // measurement only (option "time_kernels"): every plain-NTT kernel launch stands alone between two CUDA events
void note_kernel_events(cudaEvent_t a, cudaEvent_t b, unsigned grid);
cudaError_t take_kernel_times(float* ms, uint64_t cap, uint64_t* count);
extern int g_debug_skip_list;   // measurement only (option "debug_skip_list"): out-of-contract items are NOT transformed
size_t ks_scratch_words_per_item(const KsDev& ks);
cudaError_t launch_ks_prepare_keys(const KsDev& ks, TwPair* out, cudaStream_t st);
cudaError_t launch_ks_prepare_keys_fp64(const KsDev& ks, TwPair* out, cudaStream_t st);
extern int g_ks_mac_fp64;
extern int g_ks_s5_fp64;
extern int g_ks_u_fp64;
cudaError_t launch_ks_chunk(const KsDev& ks, uint64_t* result, const uint64_t* t_target, uint64_t items,
                            uint64_t* scratch, cudaStream_t st, int* launches);

// fused S2 + S3 + S4 (keyswitch_fused.cu)
extern int g_ks_fused;
extern int g_ks_sub_items;
bool ks_fused_available(const KsDev& ks, const void* keys_f);
size_t ks_fused_key_bytes(const KsDev& ks);
cudaError_t launch_ks_prepare_keys_fused(const KsDev& ks, void* out, cudaStream_t st);
cudaError_t launch_ks_fused(const KsDev& ks, const void* keys_f, const uint64_t* t_target, const uint64_t* U,
                            uint64_t* ACC, uint64_t items, cudaStream_t st);

int persistent_grid(const void* kernel, int threads, size_t smem, uint64_t items);

}  // namespace hb
