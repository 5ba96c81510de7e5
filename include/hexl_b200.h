/*
 * hexl_b200.h -- C ABI of libhexl_b200.so, the B200 (sm_100a) device path that
 * replaces intel/hexl-fpga's FPGA bitstream libraries and device runtime.
 *
 * Two layers, both plain C (pointers + sizes, no C++/torch/CUDA types):
 *
 *  (1) DEVICE-POINTER entry points: the kernel launchers.  They stand where the
 *      reference's dlsym'd bitstream symbols stood
 *      (host/inc/dl_kernel_interfaces.hpp:45-136: ntt_input/fwd_ntt/ntt_output,
 *      intt_input/inv_ntt/intt_output, input_fifo_usm/output_nb_fifo_usm,
 *      load/store/launchStoreSwitchKeys/launchConfigurableKernels) -- those take
 *      sycl::queue& / sycl::buffer& and cannot be a C ABI, so each group is
 *      replaced by one batched launcher taking device pointers and a
 *      cudaStream_t (passed as void*).
 *
 *  (2) HOST-POINTER entry points hexl_b200_host_*: the 14 functions of the
 *      reference's public API (host/inc/hexl-fpga.h:15-161), same argument
 *      meaning, same asynchronous worksize/Completed protocol, same
 *      accumulate-into-result behaviour for KeySwitch.  This is what a cgo /
 *      JNI / ctypes / C++ binding of the reference's API binds to; the C++
 *      drop-in libhexl-fpga.so (hexl-fpga_b200/host) forwards to them 1:1.
 *
 * All functions return 0 on success and a non-zero code on failure (negative:
 * argument/shape error detected by this library; positive: a cudaError_t);
 * hexl_b200_last_error() gives the message for the calling thread.  The
 * library never falls back to a CPU computation: without a CUDA device every
 * compute entry point fails.
 */
#ifndef HEXL_B200_H_
#define HEXL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEXL_B200_OK 0
#define HEXL_B200_EINVAL (-1)   /* bad argument / unsupported shape */
#define HEXL_B200_ENODEV (-2)   /* no CUDA device / library not acquired */
#define HEXL_B200_ESTATE (-3)   /* protocol misuse (e.g. Completed without work) */

int hexl_b200_version(void);                 /* 10000*major + 100*minor + patch */
const char* hexl_b200_last_error(void);      /* thread-local, never NULL */
int hexl_b200_device_count(void);            /* visible CUDA devices, <0 on error */

/* ------------------------------------------------------------------------- */
/* (1) device-pointer launchers.  All pointers are device pointers on the    */
/*     current CUDA device, 16-byte aligned; `stream` is a cudaStream_t.      */
/* ------------------------------------------------------------------------- */

/* Batched in-place forward negacyclic NTT, natural -> bit-reversed order,
 * output in [0,q).  Replaces ntt_input + fwd_ntt + ntt_output
 * (device/fwd_ntt.cpp:499-646; host call host/src/fpga.cpp:1014-1021).
 * d_operand: batch*n words; d_roots/d_precon: n words each, index m+i
 * (host/inc/hexl-fpga.h:110-120).  n = 2^10 .. 2^14 (the reference accepts
 * only 16384, host/src/ntt.cpp:24). */
int hexl_b200_ntt_fwd(uint64_t* d_operand, const uint64_t* d_root_of_unity_powers,
                      const uint64_t* d_precon_root_of_unity_powers, uint64_t coeff_modulus,
                      uint64_t n, uint64_t batch, void* stream);

/* Batched in-place inverse NTT, bit-reversed -> natural, output in [0,q).
 * Replaces intt_input + inv_ntt + intt_output (device/inv_ntt.cpp:444-607;
 * host/inc/hexl-fpga.h:139-156 for the argument meaning). */
int hexl_b200_ntt_inv(uint64_t* d_operand, const uint64_t* d_inv_root_of_unity_powers,
                      const uint64_t* d_precon_inv_root_of_unity_powers, uint64_t coeff_modulus,
                      uint64_t inv_n, uint64_t inv_n_w, uint64_t n, uint64_t batch, void* stream);

/* The same with the mod factors of the reference's own NTT class (tests/test_utils/ntt.cpp:442-470,
 * hetest::utils::NTT::ComputeForward / ComputeInverse): input_mod_factor in {1, 2, 4} (forward) / {1, 2}
 * (inverse) is the caller's promise about the input range (operand < input_mod_factor * q; checked by
 * the range vote like every input); output_mod_factor = 4 (forward) / 2 (inverse) skips the final
 * correction and returns the lazy words of the Harvey butterflies exactly as the reference leaves them
 * (ntt.cpp:535-546, 648-657); 1 = fully reduced, what hexl_b200_ntt_fwd / _inv do. */
int hexl_b200_ntt_fwd_ex(uint64_t* d_operand, const uint64_t* d_root_of_unity_powers,
                         const uint64_t* d_precon_root_of_unity_powers, uint64_t coeff_modulus, uint64_t n,
                         uint64_t batch, uint64_t input_mod_factor, uint64_t output_mod_factor, void* stream);
int hexl_b200_ntt_inv_ex(uint64_t* d_operand, const uint64_t* d_inv_root_of_unity_powers,
                         const uint64_t* d_precon_inv_root_of_unity_powers, uint64_t coeff_modulus, uint64_t inv_n,
                         uint64_t inv_n_w, uint64_t n, uint64_t batch, uint64_t input_mod_factor,
                         uint64_t output_mod_factor, void* stream);

/* Batched negacyclic polynomial multiply  result = a * b  mod (x^n + 1, q):
 * INTT(NTT(a) (.) NTT(b)) with the dyadic product fused into the first pass
 * of the inverse transform (no HBM round trip of the product).  The step on
 * either side of the standalone primitives that the reference motivates
 * (README.md:47) but does not ship as one call; SURVEY section 8(f) row 4.
 * a, b, result: batch*n words, coefficient form, natural order; result may
 * alias a (not b).  Tables and scalars as for hexl_b200_ntt_fwd / _ntt_inv. */
int hexl_b200_poly_multiply(uint64_t* d_result, const uint64_t* d_a, const uint64_t* d_b,
                            const uint64_t* d_root_of_unity_powers,
                            const uint64_t* d_precon_root_of_unity_powers,
                            const uint64_t* d_inv_root_of_unity_powers,
                            const uint64_t* d_precon_inv_root_of_unity_powers, uint64_t coeff_modulus,
                            uint64_t inv_n, uint64_t inv_n_w, uint64_t n, uint64_t batch, void* stream);

/* Batched dyadic ciphertext multiply.  Replaces input_fifo_usm + dyadic_multiply
 * + output_nb_fifo_usm (device/dyadic_multiply.cpp:61-376).  Per item:
 * op1/op2 [2][n_moduli][n], results [3][n_moduli][n] (host/inc/hexl-fpga.h:
 * 28-44).  d_moduli holds n_moduli words (moduli_per_item == 0) or
 * batch*n_moduli words (moduli_per_item != 0, one set per item as in
 * tests/test_dyadic_multiply.cpp:36-38).  Any modulus >= 1, operands need not
 * be reduced. */
int hexl_b200_dyadic_multiply(uint64_t* d_results, const uint64_t* d_operand1,
                              const uint64_t* d_operand2, uint64_t n, const uint64_t* d_moduli,
                              uint64_t n_moduli, uint64_t batch, int moduli_per_item,
                              void* stream);

/* KeySwitch plan: the device-resident constants of one (shape, moduli, key set)
 * -- twiddle tables, Barrett/Shoup metadata and the switch keys.  Replaces the
 * host-side metadata/key upload of host/src/fpga.cpp:1039-1123,1158-1248 and
 * the launchStoreSwitchKeys / launchConfigurableKernels bitstream calls.
 * Arguments are HOST pointers with the meaning of host/inc/hexl-fpga.h:54-64;
 * k_switch_keys[j] points to key_component_count*key_modulus_size*n words.
 * twiddle_factors may be NULL (tables are then derived from `moduli` with the
 * minimal primitive 2n-th root, as host/src/fpga.cpp:1097-1109 does). */
typedef struct hexl_b200_ks_plan hexl_b200_ks_plan;
int hexl_b200_ks_plan_create(hexl_b200_ks_plan** plan, uint64_t n, uint64_t decomp_modulus_size,
                             uint64_t key_modulus_size, uint64_t rns_modulus_size,
                             uint64_t key_component_count, const uint64_t* moduli,
                             const uint64_t* const* k_switch_keys,
                             const uint64_t* modswitch_factors, const uint64_t* twiddle_factors);
int hexl_b200_ks_plan_destroy(hexl_b200_ks_plan* plan);

/* Batched keyswitch on device-resident items.  Replaces load + the autorun
 * pipeline + store (device/keyswitch.cpp:15-65) AND the host accumulate of
 * host/src/fpga.cpp:441-475: d_result[b] (2*decomp*n words, [c][i][coeff]) is
 * read-modify-written, d_t_target[b] is decomp*n words. */
int hexl_b200_keyswitch(hexl_b200_ks_plan* plan, uint64_t* d_result, const uint64_t* d_t_target,
                        uint64_t batch, void* stream);

/* Host-side twiddle generation (no GPU needed): the tables a caller of _NTT /
 * _INTT must supply, from the minimal primitive 2n-th root of unity mod q.
 * Replaces host/src/twiddle-factors.cpp:16-62 + number_theory_util.cpp
 * (ComputeRootOfUnityPowers / MinimalPrimitiveRoot / InverseUIntMod).
 * out4n receives [roots | precon_roots | inv_roots | precon_inv_roots], n words
 * each, in the layouts of host/inc/hexl-fpga.h:110-156; *inv_n = n^-1 mod q,
 * *inv_n_w = inv_n * inv_roots[n-1] mod q. */
int hexl_b200_compute_twiddles(uint64_t n, uint64_t modulus, uint64_t* out4n, uint64_t* inv_n,
                               uint64_t* inv_n_w);

/* Kernel selection knobs for benchmarking.  Not part of the reference surface.
 *   "ntt_variant"      bit 0: 32 words per thread at n = 16384 (default 1);
 *                      bit 1: skip the input-range vote (caller guarantees the contract)
 *   "small_path"       q < 2^30 kernels: 0 off, 1 uint32 kernels behind a TMA landing
 *                      buffer (default), 2 uint32 kernels with direct loads, two CTAs / SM,
 *                      3 two transforms per SM sharing three 64 KiB regions (both measured slower)
 *   "small_tma_store"  1: small-modulus forward results leave through TMA stores (default 0)
 *   "inv_lazy"         1: correction-free inverse butterflies for q < 2^52 (default 0)
 *   "fp64_path"        1 (default): hexl_b200_ntt_fwd / _inv run their butterflies on the FP64
 *                      pipe when 2^36 <= q <= 2^53 / 3 (bit-identical results); 0: integer kernels
 *   "pdl"              1 (default): the plain NTT kernels are launched with programmatic stream
 *                      serialization (launch latency overlaps the kernel in front); 0: plain launches
 *   "warp_tail"        1 (default): the FP64-pipe kernels at n = 16384 deal the tail rows out by warp
 *                      (one block barrier per transform instead of three); 0: by thread index
 *   "l2_prefetch"      1 (default): a transform CTA pulls the polynomial after the current one towards L2 a whole
 *                      transform ahead of its TMA load (read per call / at keyswitch plan creation)
 *   "ks_blocked"       1 (default): keyswitch stages that walk their items modulus-major give every CTA one
 *                      contiguous stretch of the order (read at plan creation); 0: strided
 *   "time_kernels"     1: time every plain-NTT kernel launch with CUDA events (hexl_b200_kernel_times)
 *   "ks_workspace_mb"  keyswitch scratch bound in MiB (>= 16, default 10240; a batch is cut into equal chunks that fit)
 *   "ks_mac_items"     items sharing one key load in the keyswitch MAC (1, 4, 8) */
int hexl_b200_set_option(const char* name, int64_t value);

/* Measurement aid for bench.py's roofline leg.  With option "time_kernels" = 1 every kernel launched by
 * hexl_b200_ntt_fwd / _inv / _poly_multiply stands alone between two CUDA events recorded on its stream
 * (plain launches: no programmatic overlap with the kernel in front).  This call waits for the launches timed
 * since the previous call and returns their durations in milliseconds, in launch order (at most `cap` of them
 * are written; *count is how many there were). */
int hexl_b200_kernel_times(float* ms, uint64_t cap, uint64_t* count);

/* ------------------------------------------------------------------------- */
/* (2) host-pointer API: the reference's public functions, C linkage.        */
/*     (host/inc/hexl-fpga.h line numbers in brackets)                        */
/* ------------------------------------------------------------------------- */
int hexl_b200_host_acquire(void);                       /* acquire_FPGA_resources [19] */
int hexl_b200_host_release(void);                       /* release_FPGA_resources [23] */

int hexl_b200_host_set_worksize_dyadic_multiply(uint64_t ws);                 /* [34] */
int hexl_b200_host_dyadic_multiply(uint64_t* results, const uint64_t* operand1,
                                   const uint64_t* operand2, uint64_t n,
                                   const uint64_t* moduli, uint64_t n_moduli);  /* [46] */
int hexl_b200_host_dyadic_multiply_completed(void);                           /* [55] */

int hexl_b200_host_set_worksize_keyswitch(uint64_t ws);                       /* [63] */
int hexl_b200_host_keyswitch(uint64_t* result, const uint64_t* t_target_iter_ptr, uint64_t n,
                             uint64_t decomp_modulus_size, uint64_t key_modulus_size,
                             uint64_t rns_modulus_size, uint64_t key_component_count,
                             const uint64_t* moduli, const uint64_t** k_switch_keys,
                             const uint64_t* modswitch_factors,
                             const uint64_t* twiddle_factors);                /* [83] */
int hexl_b200_host_keyswitch_completed(void);                                 /* [95] */

int hexl_b200_host_set_worksize_ntt(uint64_t ws);                             /* [108] */
int hexl_b200_host_ntt(uint64_t* operand, const uint64_t* root_of_unity_powers,
                       const uint64_t* precon_root_of_unity_powers, uint64_t coeff_modulus,
                       uint64_t n);                                           /* [120] */
int hexl_b200_host_ntt_completed(void);                                       /* [130] */

int hexl_b200_host_set_worksize_intt(uint64_t ws);                            /* [138] */
int hexl_b200_host_intt(uint64_t* operand, const uint64_t* inv_root_of_unity_powers,
                        const uint64_t* precon_inv_root_of_unity_powers, uint64_t coeff_modulus,
                        uint64_t inv_n, uint64_t inv_n_w, uint64_t n);        /* [152] */
int hexl_b200_host_intt_completed(void);                                      /* [161] */

/* Bulk submission helpers (not in the reference): exactly `count` calls of the
 * function above on items `base + i*stride_words`, made in one FFI crossing so
 * that bindings with an expensive call path (ctypes, JNI, cgo) do not pay it
 * per polynomial.  They do not call set_worksize / Completed. */
int hexl_b200_host_ntt_many(uint64_t* operand_base, uint64_t stride_words, uint64_t count,
                            const uint64_t* root_of_unity_powers,
                            const uint64_t* precon_root_of_unity_powers, uint64_t coeff_modulus,
                            uint64_t n);
int hexl_b200_host_intt_many(uint64_t* operand_base, uint64_t stride_words, uint64_t count,
                             const uint64_t* inv_root_of_unity_powers,
                             const uint64_t* precon_inv_root_of_unity_powers,
                             uint64_t coeff_modulus, uint64_t inv_n, uint64_t inv_n_w, uint64_t n);
int hexl_b200_host_dyadic_multiply_many(uint64_t* results_base, const uint64_t* operand1_base,
                                        const uint64_t* operand2_base, uint64_t count, uint64_t n,
                                        const uint64_t* moduli, uint64_t n_moduli);
int hexl_b200_host_keyswitch_many(uint64_t* result_base, const uint64_t* t_target_base,
                                  uint64_t count, uint64_t n, uint64_t decomp_modulus_size,
                                  uint64_t key_modulus_size, uint64_t rns_modulus_size,
                                  uint64_t key_component_count, const uint64_t* moduli,
                                  const uint64_t** k_switch_keys, const uint64_t* modswitch_factors,
                                  const uint64_t* twiddle_factors);

/* Pin a caller buffer in place (not in the reference).  Every caller of the reference passes pageable memory
 * (std::vector), which this library stages through its own pinned ring with CPU copy threads -- bound by
 * host cores and host-memory bandwidth (about 0.6 of the pinned rate on a 16-core box).  An integration that
 * owns long-lived buffers (a ciphertext pool) registers them once; from then on they are the source / target
 * of the DMA directly, like memory from cudaHostAlloc.  Wraps cudaHostRegister (page-granular, portable
 * across the NUM_DEV devices).  The buffer must stay mapped until hexl_b200_host_unpin_buffer(p) -- same
 * pointer -- or hexl_b200_host_release(), which unpins everything.  A range already covered returns 0; a
 * partial overlap with an earlier range is EINVAL. */
int hexl_b200_host_pin_buffer(void* p, uint64_t bytes);
int hexl_b200_host_unpin_buffer(void* p);

/* Counters since acquire: kernel launches issued by this library and bytes
 * moved host<->device by the host-pointer API (for bench.py's gpu_launches /
 * h2d / d2h fields). */
typedef struct {
    uint64_t kernel_launches;
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
} hexl_b200_stats;
int hexl_b200_get_stats(hexl_b200_stats* out);
int hexl_b200_reset_stats(void);

/* Work done so far by worker `worker` (0 .. NUM_DEV-1) of the host-pointer runtime since acquire: which
 * CUDA device it drives, how many batches and how many requests it has executed.  The reference's
 * DevicePool (host/src/fpga.cpp:1646-1673) has no such counter; tests use it to check that a run is dealt
 * out over all NUM_DEV workers. */
typedef struct {
    int32_t device;
    uint64_t batches;
    uint64_t items;
} hexl_b200_device_stats;
int hexl_b200_host_device_stats(int worker, hexl_b200_device_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* HEXL_B200_H_ */
