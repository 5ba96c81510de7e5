// capi.cu -- the device-pointer half of the C ABI (include/hexl_b200.h, part 1)
// and the keyswitch plan.  Host-pointer API: ../host/src/runtime.cpp.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hexl_b200.h"
#include "../host/inc/number_theory.h"
#include "internal.h"
#include "launch.h"

namespace nt = hexl_b200::nt;

namespace hexl_b200 {

static thread_local std::string g_err = "";
std::atomic<uint64_t> g_launches{0}, g_h2d{0}, g_d2h{0};
static std::atomic<int> g_ntt_variant{1};   // 1: 32 words/thread at N=16384 (default), 0: 16
static std::atomic<int64_t> g_ks_workspace_mb{10240};  // scratch budget of one keyswitch plan: ~1000 items per chunk at D/K = 7/8 (measured optimum, profiles/r2_time_ks_ws.jsonl: long runs of one modulus per CTA, few launch tails)

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}

static int ilog2_exact(uint64_t n) {
    if (n == 0 || (n & (n - 1))) return -1;
    int l = 0;
    while ((1ull << l) < n) ++l;
    return l;
}
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static std::atomic<int> g_inv_lazy{0};     // correction-free inverse butterflies for q < 2^52 (option "inv_lazy")
// 32-bit kernels for q < 2^30 (option "small_path"): 0 off, 1 TMA landing buffer
// (one CTA per SM), 2 direct global loads (two CTAs per SM)
static std::atomic<int> g_small_path{1};
// butterflies on the FP64 pipe for 2^36 <= q <= 2^53/3 in the plain NTT entry points (option "fp64_path")
static std::atomic<int> g_fp64_path{1};
// forward FP64 butterflies with a full correction every other stage for q <= 2^51 (1 + 1/32) (option "fp64_alt")
static std::atomic<int> g_fp64_alt{1};

// keyswitch stages that walk modulus-major: contiguous stretch of the order per CTA (option "ks_blocked", read at plan creation)
static std::atomic<int> g_ks_blocked{1};
// the polynomial after the current one is pulled towards L2 a transform ahead of its TMA load (option "l2_prefetch")
static std::atomic<int> g_l2_prefetch{1};

hb::ModTab make_modtab(uint64_t q, uint64_t inv_n, uint64_t inv_n_w, const hb::TwPair* ftw,
                       const hb::TwPair* itw, int logn, const hb::Tw32* ftw32, const hb::Tw32* itw32) {
    hb::ModTab t;
    t.q = q;
    t.twoq = q << 1;
    t.mu = nt::barrett_mu(q);
    t.sc.inv_n = inv_n;
    t.sc.inv_n_p = inv_n < q ? nt::shoup(inv_n, q) : 0;
    t.sc.inv_n_w = inv_n_w;
    t.sc.inv_n_w_p = inv_n_w < q ? nt::shoup(inv_n_w, q) : 0;
    t.fm = hb::make_fastmod(q);
    t.ftw = ftw;
    t.itw = itw;
    t.fwd_fast_ok = hb::fwd_fast_modulus_ok(q, logn) ? 1u : 0u;
    t.inv_fast_ok = hb::inv_fast_modulus_ok(q) ? 1u : 0u;
    t.sm32 = hb::make_small32(q, t.sc);
    t.ftw32 = ftw32;
    t.itw32 = itw32;
    t.small_ok = (hb::small_modulus_ok(q) && (ftw32 || itw32)) ? (uint32_t)g_small_path.load() : 0u;
    t.inv_lazy_ok = (g_inv_lazy.load() && hb::inv_lazy_modulus_ok(q)) ? 1u : 0u;
    t.fd = hb::make_fp64mod(q, inv_n < q ? inv_n : 0, inv_n_w < q ? inv_n_w : 0);
    t.l2_prefetch = (uint32_t)g_l2_prefetch.load();
    t.ftwd = nullptr;
    t.itwd = nullptr;
    t.fp64_ok = 0;          // set by the callers that build the FP64 tables
    t.fp64_alt_ok = (g_fp64_alt.load() && hb::fp64_alt_modulus_ok(q)) ? 1u : 0u;
    t.lazy_out = 0;
    return t;
}

// Scratch for the packed twiddles of a plain NTT call: one small buffer per
// (device, stream), allocated once and reused -- work on one stream is ordered,
// so the pack kernel of call k+1 cannot overtake the transform of call k.
struct StreamScratch {
    std::mutex mu;
    std::map<std::pair<int, cudaStream_t>, std::pair<void*, size_t>> bufs;
    cudaError_t get(cudaStream_t st, size_t bytes, void** out) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lk(mu);
        auto& b = bufs[{dev, st}];
        if (b.second < bytes) {
            if (b.first) {
                if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
                cudaFree(b.first);
                b = {nullptr, 0};
            }
            if ((e = cudaMalloc(&b.first, bytes)) != cudaSuccess) return e;
            b.second = bytes;
        }
        *out = b.first;
        return cudaSuccess;
    }
};
static StreamScratch g_scratch;

// N = 32768 (csrc/ntt_big.cu): stage 0 / the last stage as a streaming kernel, the two halves of every
// polynomial on the 16384-point shared-memory kernels.  Scratch (per stream): raw sub-tables of both halves
// (2 x 2 x 16384 words), per half the packed tables (integer + FP64 format) and a deferred list.
static int ntt_big(bool fwd, uint64_t* d_operand, const uint64_t* d_tab, const uint64_t* d_tab_p, uint64_t q,
                   uint64_t inv_n, uint64_t inv_n_w, uint64_t batch, uint32_t lazy_out, cudaStream_t st) {
    const uint32_t logh = 14, nh = 1u << logh;
    const int variant = 1;
    if (lazy_out) return fail(HEXL_B200_EINVAL, "n = 32768: output_mod_factor must be 1");
    const size_t raw_bytes = (size_t)4 * nh * 8, half_bytes = (size_t)2 << 20;
    const size_t list_bytes = ((batch + 2) * 4 + 255) & ~(size_t)255;
    uint8_t* scratch = nullptr;
    cudaError_t e = g_scratch.get(st, raw_bytes + 2 * half_bytes + 2 * list_bytes, (void**)&scratch);
    if (e != cudaSuccess) return cuda_fail(e, "ntt (n = 32768): scratch");
    uint64_t* sub = reinterpret_cast<uint64_t*>(scratch);
    if ((e = hb::launch_big_split(fwd, d_tab, d_tab_p, sub, nh, st))) return cuda_fail(e, "ntt (n = 32768): table split");
    int launches = 1;
    if (fwd) {
        if ((e = hb::launch_big_stage0_fwd(d_operand, d_tab, d_tab_p, q, nh, batch, st)))
            return cuda_fail(e, "ntt (n = 32768): stage 0");
        ++launches;
    }
    for (uint32_t h = 0; h < 2; ++h) {
        uint8_t* blk = scratch + raw_bytes + h * half_bytes;
        hb::TwPair* packed = reinterpret_cast<hb::TwPair*>(blk);
        hb::TwPair* packed_d = reinterpret_cast<hb::TwPair*>(blk + (1u << 20));
        uint32_t* list = reinterpret_cast<uint32_t*>(scratch + raw_bytes + 2 * half_bytes + h * list_bytes);
        const uint64_t* r = sub + (size_t)h * 2 * nh;
        hb::PackExtra px;
        const bool fp64 = g_fp64_path.load() && hb::fp64_modulus_ok(q);
        if (fp64) {
            (fwd ? px.fwd_d : px.inv_d) = packed_d;
            px.q = q;
        }
        e = fwd ? hb::launch_pack_twiddles(logh, variant, r, r + nh, packed, nullptr, nullptr, nullptr, list, st, px)
                : hb::launch_pack_twiddles(logh, variant, nullptr, nullptr, nullptr, r, r + nh, packed, list, st, px);
        if (e != cudaSuccess) return cuda_fail(e, "ntt (n = 32768): pack twiddles");
        ++launches;
        // the sub-transforms' own last stage is scaled by 1: the big transform's twiddle of that stage and
        // n^-1 are applied by the last-stage kernel
        hb::ModTab t = fwd ? make_modtab(q, 0, 0, packed, nullptr, (int)logh) : make_modtab(q, 1 % q, 1 % q, nullptr, packed, (int)logh);
        if (fp64) {
            (fwd ? t.ftwd : t.itwd) = packed_d;
            t.fp64_ok = 1;
        }
        if ((e = hb::launch_ntt_half(fwd, d_operand, t, logh, batch, variant, list, st, &launches, h)))
            return cuda_fail(e, "ntt (n = 32768): sub-transforms");
    }
    if (!fwd) {
        if ((e = hb::launch_big_last_inv(d_operand, d_tab, q, inv_n, inv_n_w, nh, batch, st)))
            return cuda_fail(e, "ntt (n = 32768): last stage");
        ++launches;
    }
    g_launches += launches;
    return 0;
}

}  // namespace hexl_b200

using namespace hexl_b200;

// keyswitch plan ------------------------------------------------------------
struct hexl_b200_ks_plan {
    int device = 0;
    uint64_t n = 0, D = 0, K = 0, R = 0;
    hb::KsDev dev{};
    hb::TwPair* d_packed = nullptr; // K * (FWD_ENTRIES + INV_ENTRIES) packed twiddles
    uint64_t* d_keys = nullptr;     // D * 2 * K * n
    hb::TwPair* d_keys_sh = nullptr;  // same, with Shoup factors (fast path)
    void* d_keys_fused = nullptr;     // key quads in the fused kernel's layout (keyswitch_fused.cu)
    hb::TwPair* d_keys_fp = nullptr;  // {centred key, key / q} doubles (FP64-pipe multiply-accumulate)
    uint64_t* d_small = nullptr;    // msf, msf_p
    hb::ModTab* d_tabs = nullptr;
    hb::Divisor* d_divs = nullptr;
    // workspace (grown lazily, reused across calls; guarded by mu)
    std::mutex mu;
    uint64_t* ws = nullptr;
    size_t ws_words = 0;
};

extern "C" {

int hexl_b200_version(void) { return 100; }
const char* hexl_b200_last_error(void) { return g_err.c_str(); }

int hexl_b200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HEXL_B200_ENODEV, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return n;
}

#ifdef HB_EXPERIMENTAL_VARIANTS
static const bool kExperimental = true;
#else
static const bool kExperimental = false;
#endif
#define HB_NEEDS_EXPERIMENTAL(cond)                                                                              \
    if ((cond) && !kExperimental)                                                                                \
        return fail(HEXL_B200_EINVAL, "option '%s' = %lld selects a kernel variant that is only compiled with "   \
                                      "make EXPERIMENTAL=1 (measured slower than the default)", name, (long long)value)

int hexl_b200_set_option(const char* name, int64_t value) {
    if (!name) return fail(HEXL_B200_EINVAL, "option name is NULL");
    if (!strcmp(name, "ntt_variant")) {
        // bit 0: 32 words/thread at N=16384; bit 1 (perf exploration only): skip the
        // input-range vote, i.e. the caller guarantees in-contract inputs
        if (value < 0 || value > 3) return fail(HEXL_B200_EINVAL, "ntt_variant must be 0..3");
        // (N = 16384 with 16 words per thread needs the experimental build; other sizes ignore the difference)
        (void)kExperimental;
        g_ntt_variant = (int)value;
        return 0;
    }
    if (!strcmp(name, "ks_workspace_mb")) {
        if (value < 16) return fail(HEXL_B200_EINVAL, "ks_workspace_mb must be >= 16");
        g_ks_workspace_mb = value;
        return 0;
    }
    if (!strcmp(name, "small_path")) {
        if (value < 0 || value > 3) return fail(HEXL_B200_EINVAL, "small_path must be 0, 1, 2 or 3");
        HB_NEEDS_EXPERIMENTAL(value >= 2);
        g_small_path = (int)value;
        return 0;
    }
    if (!strcmp(name, "small_tma_store")) {
        hb::g_small_tma_store = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "pdl")) {
        hb::g_pdl = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "ks_u_fp64")) {
        hb::g_ks_u_fp64 = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "ks_blocked")) {
        g_ks_blocked = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "l2_prefetch")) {
        g_l2_prefetch = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "time_kernels")) {
        hb::g_time_kernels = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "debug_skip_list")) {
        hb::g_debug_skip_list = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "warp_tail")) {
        hb::g_warp_tail = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "fp64_alt")) {
        g_fp64_alt = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "fp64_path")) {
        g_fp64_path = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "inv_lazy")) {
        HB_NEEDS_EXPERIMENTAL(value != 0);
        g_inv_lazy = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "polymul_fused")) {
        hb::g_polymul_fused = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "ks_fused")) {
        hb::g_ks_fused = value ? 1 : 0;     // read when a plan is created and on every call
        return 0;
    }
    if (!strcmp(name, "ks_s5_fp64")) {
        hb::g_ks_s5_fp64 = value ? 1 : 0;    // read on every call
        return 0;
    }
    if (!strcmp(name, "ks_mac_fp64")) {
        hb::g_ks_mac_fp64 = value ? 1 : 0;   // read on every call
        return 0;
    }
    if (!strcmp(name, "ks_sub_items")) {
        if (value < 0 || value > 65535) return fail(HEXL_B200_EINVAL, "ks_sub_items must be 0..65535");
        hb::g_ks_sub_items = (int)value;
        return 0;
    }
    if (!strcmp(name, "ks_mac_items")) {
        if (value != 1 && value != 2 && value != 4 && value != 8)
            return fail(HEXL_B200_EINVAL, "ks_mac_items must be 1, 2, 4 or 8");
        HB_NEEDS_EXPERIMENTAL(value != 4);
        hb::g_ks_mac_items = (int)value;
        return 0;
    }
    return fail(HEXL_B200_EINVAL, "unknown option '%s'", name);
}

int hexl_b200_compute_twiddles(uint64_t n, uint64_t q, uint64_t* out4n, uint64_t* inv_n, uint64_t* inv_n_w) {
    if (!out4n) return fail(HEXL_B200_EINVAL, "compute_twiddles: NULL output");
    if (ilog2_exact(n) < 1 || n > (1u << 20)) return fail(HEXL_B200_EINVAL, "compute_twiddles: bad n");
    if (q < 2 || q >> 62 || (q - 1) % (2 * n))
        return fail(HEXL_B200_EINVAL, "compute_twiddles: modulus must be a prime = 1 mod 2n below 2^62");
    nt::Tables t = nt::make_tables(n, q);
    if (t.roots.size() != n || t.inv_n == 0)
        return fail(HEXL_B200_EINVAL, "compute_twiddles: no primitive 2n-th root of unity mod %llu",
                    (unsigned long long)q);
    memcpy(out4n, t.roots.data(), n * 8);
    memcpy(out4n + n, t.precon.data(), n * 8);
    memcpy(out4n + 2 * n, t.inv_roots.data(), n * 8);
    memcpy(out4n + 3 * n, t.precon_inv.data(), n * 8);
    if (inv_n) *inv_n = t.inv_n;
    if (inv_n_w) *inv_n_w = t.inv_n_w;
    return 0;
}

int hexl_b200_get_stats(hexl_b200_stats* out) {
    if (!out) return fail(HEXL_B200_EINVAL, "stats pointer is NULL");
    out->kernel_launches = g_launches.load();
    out->h2d_bytes = g_h2d.load();
    out->d2h_bytes = g_d2h.load();
    return 0;
}
int hexl_b200_reset_stats(void) {
    g_launches = 0;
    g_h2d = 0;
    g_d2h = 0;
    return 0;
}

int hexl_b200_kernel_times(float* ms, uint64_t cap, uint64_t* count) {
    if (!count) return fail(HEXL_B200_EINVAL, "kernel_times: NULL count");
    const cudaError_t e = hb::take_kernel_times(ms, cap, count);
    if (e != cudaSuccess) return cuda_fail(e, "kernel_times");
    return 0;
}

int hexl_b200_ntt_fwd(uint64_t* d_operand, const uint64_t* d_roots, const uint64_t* d_precon,
                      uint64_t q, uint64_t n, uint64_t batch, void* stream) {
    return hexl_b200_ntt_fwd_ex(d_operand, d_roots, d_precon, q, n, batch, 1, 1, stream);
}

int hexl_b200_ntt_fwd_ex(uint64_t* d_operand, const uint64_t* d_roots, const uint64_t* d_precon, uint64_t q, uint64_t n,
                         uint64_t batch, uint64_t input_mod_factor, uint64_t output_mod_factor, void* stream) {
    // reference: tests/test_utils/ntt.cpp:442-455 (ComputeForward argument checks)
    if (input_mod_factor != 1 && input_mod_factor != 2 && input_mod_factor != 4)
        return fail(HEXL_B200_EINVAL, "ntt_fwd: input_mod_factor must be 1, 2 or 4");
    if (output_mod_factor != 1 && output_mod_factor != 4)
        return fail(HEXL_B200_EINVAL, "ntt_fwd: output_mod_factor must be 1 or 4");
    const int logn = ilog2_exact(n);
    if (logn < 0 || !(hb::ntt_shape_supported((uint32_t)logn) || logn == 15))
        return fail(HEXL_B200_EINVAL, "ntt_fwd: n=%llu unsupported (power of two in [1024,32768])",
                    (unsigned long long)n);
    if (!d_operand || !d_roots || !d_precon) return fail(HEXL_B200_EINVAL, "ntt_fwd: NULL pointer");
    if (!aligned16(d_operand)) return fail(HEXL_B200_EINVAL, "ntt_fwd: operand not 16-byte aligned");
    if (q < 2 || q >> 62) return fail(HEXL_B200_EINVAL, "ntt_fwd: modulus must be in [2, 2^62)");
    if (batch == 0) return 0;
    if (logn == 15)
        return ntt_big(true, d_operand, d_roots, d_precon, q, 0, 0, batch, output_mod_factor == 4, (cudaStream_t)stream);
    const int variant = g_ntt_variant.load();
    // per-stream scratch: [0,512K) packed forward twiddles, [512K,1M) packed
    // inverse twiddles, [1M,1.5M) / [1.5M,2M) the same for the FP64 path, then
    // two deferred lists of 1 + batch words
    uint8_t* scratch = nullptr;
    const size_t list_bytes = ((batch + 2) * 4 + 255) & ~(size_t)255;
    cudaError_t e = g_scratch.get((cudaStream_t)stream, (size_t)(2u << 20) + 2 * list_bytes, (void**)&scratch);
    if (e != cudaSuccess) return cuda_fail(e, "ntt_fwd: scratch");
    hb::TwPair* packed = reinterpret_cast<hb::TwPair*>(scratch);
    uint32_t* list = reinterpret_cast<uint32_t*>(scratch + (2u << 20));
    // one launch packs the call's twiddles in every format its kernels may need
    hb::PackExtra px;
    hb::Tw32* packed32 = nullptr;
    hb::TwPair* packed_d = nullptr;
    if (g_small_path.load() && hb::small_modulus_ok(q) && hb::small_path_available((uint32_t)logn, variant))
        px.fwd32 = packed32 = reinterpret_cast<hb::Tw32*>(scratch + 300 * 1024);
    if (g_fp64_path.load() && hb::fp64_modulus_ok(q)) {
        px.fwd_d = packed_d = reinterpret_cast<hb::TwPair*>(scratch + (1u << 20));
        px.q = q;
    }
    e = hb::launch_pack_twiddles((uint32_t)logn, variant, d_roots, d_precon, packed, nullptr, nullptr, nullptr, list,
                                 (cudaStream_t)stream, px);
    if (e != cudaSuccess) return cuda_fail(e, "ntt_fwd: pack twiddles");
    int launches = 1;
    hb::ModTab t = make_modtab(q, 0, 0, packed, nullptr, logn, packed32, nullptr);
    if (packed_d) {
        t.ftwd = packed_d;
        t.fp64_ok = 1;
    }
    // output_mod_factor 4: the words of the Harvey butterflies as they stand, in [0, 4q) (ntt.cpp:535): only
    // the reference op sequence defines them, so the exact kernel runs for every item
    t.lazy_out = output_mod_factor == 4 ? 1u : 0u;
    e = hb::launch_ntt_fwd(d_operand, t, (uint32_t)logn, batch, variant, list, (cudaStream_t)stream, &launches);
    if (e != cudaSuccess) return cuda_fail(e, "ntt_fwd launch");
    g_launches += launches;
    return 0;
}

int hexl_b200_ntt_inv(uint64_t* d_operand, const uint64_t* d_inv_roots, const uint64_t* d_precon_inv,
                      uint64_t q, uint64_t inv_n, uint64_t inv_n_w, uint64_t n, uint64_t batch,
                      void* stream) {
    return hexl_b200_ntt_inv_ex(d_operand, d_inv_roots, d_precon_inv, q, inv_n, inv_n_w, n, batch, 1, 1, stream);
}

int hexl_b200_ntt_inv_ex(uint64_t* d_operand, const uint64_t* d_inv_roots, const uint64_t* d_precon_inv, uint64_t q,
                         uint64_t inv_n, uint64_t inv_n_w, uint64_t n, uint64_t batch, uint64_t input_mod_factor,
                         uint64_t output_mod_factor, void* stream) {
    // reference: tests/test_utils/ntt.cpp:457-470 (ComputeInverse argument checks)
    if (input_mod_factor != 1 && input_mod_factor != 2)
        return fail(HEXL_B200_EINVAL, "ntt_inv: input_mod_factor must be 1 or 2");
    if (output_mod_factor != 1 && output_mod_factor != 2)
        return fail(HEXL_B200_EINVAL, "ntt_inv: output_mod_factor must be 1 or 2");
    const int logn = ilog2_exact(n);
    if (logn < 0 || !(hb::ntt_shape_supported((uint32_t)logn) || logn == 15))
        return fail(HEXL_B200_EINVAL, "ntt_inv: n=%llu unsupported (power of two in [1024,32768])",
                    (unsigned long long)n);
    if (!d_operand || !d_inv_roots || !d_precon_inv)
        return fail(HEXL_B200_EINVAL, "ntt_inv: NULL pointer");
    if (!aligned16(d_operand)) return fail(HEXL_B200_EINVAL, "ntt_inv: operand not 16-byte aligned");
    if (q < 2 || q >> 62) return fail(HEXL_B200_EINVAL, "ntt_inv: modulus must be in [2, 2^62)");
    if (inv_n >= q || inv_n_w >= q)
        return fail(HEXL_B200_EINVAL, "ntt_inv: inv_n / inv_n_w must be reduced mod q");
    if (batch == 0) return 0;
    if (logn == 15)
        return ntt_big(false, d_operand, d_inv_roots, d_precon_inv, q, inv_n, inv_n_w, batch, output_mod_factor == 2,
                       (cudaStream_t)stream);
    const int variant = g_ntt_variant.load();
    uint8_t* scratch = nullptr;
    const size_t list_bytes = ((batch + 2) * 4 + 255) & ~(size_t)255;
    cudaError_t e = g_scratch.get((cudaStream_t)stream, (size_t)(2u << 20) + 2 * list_bytes, (void**)&scratch);
    if (e != cudaSuccess) return cuda_fail(e, "ntt_inv: scratch");
    // second halves, so a forward and an inverse call may be queued back to back
    hb::TwPair* packed = reinterpret_cast<hb::TwPair*>(scratch + (1u << 19));
    uint32_t* list = reinterpret_cast<uint32_t*>(scratch + (2u << 20) + list_bytes);
    hb::PackExtra px;
    hb::Tw32* packed32 = nullptr;
    hb::TwPair* packed_d = nullptr;
    if (g_small_path.load() && hb::small_modulus_ok(q) && hb::small_path_available((uint32_t)logn, variant))
        px.inv32 = packed32 = reinterpret_cast<hb::Tw32*>(scratch + (1u << 19) + 300 * 1024);
    if (g_fp64_path.load() && hb::fp64_modulus_ok(q)) {
        px.inv_d = packed_d = reinterpret_cast<hb::TwPair*>(scratch + (3u << 19));
        px.q = q;
    }
    e = hb::launch_pack_twiddles((uint32_t)logn, variant, nullptr, nullptr, nullptr, d_inv_roots, d_precon_inv, packed,
                                 list, (cudaStream_t)stream, px);
    if (e != cudaSuccess) return cuda_fail(e, "ntt_inv: pack twiddles");
    int launches = 1;
    hb::ModTab t = make_modtab(q, inv_n, inv_n_w, nullptr, packed, logn, nullptr, packed32);
    if (packed_d) {
        t.itwd = packed_d;
        t.fp64_ok = 1;
    }
    t.lazy_out = output_mod_factor == 2 ? 1u : 0u;   // words left in [0, 2q) (ntt.cpp:648-657): exact kernel
    e = hb::launch_ntt_inv(d_operand, t, (uint32_t)logn, batch, variant, list, (cudaStream_t)stream, &launches);
    if (e != cudaSuccess) return cuda_fail(e, "ntt_inv launch");
    g_launches += launches;
    return 0;
}

int hexl_b200_poly_multiply(uint64_t* d_result, const uint64_t* d_a, const uint64_t* d_b,
                            const uint64_t* d_roots, const uint64_t* d_precon, const uint64_t* d_inv_roots,
                            const uint64_t* d_precon_inv, uint64_t q, uint64_t inv_n, uint64_t inv_n_w,
                            uint64_t n, uint64_t batch, void* stream) {
    const int logn = ilog2_exact(n);
    if (logn < 0 || !hb::ntt_shape_supported((uint32_t)logn))
        return fail(HEXL_B200_EINVAL, "poly_multiply: n=%llu unsupported (power of two in [1024,16384])",
                    (unsigned long long)n);
    if (!d_result || !d_a || !d_b || !d_roots || !d_precon || !d_inv_roots || !d_precon_inv)
        return fail(HEXL_B200_EINVAL, "poly_multiply: NULL pointer");
    if (!aligned16(d_result) || !aligned16(d_a) || !aligned16(d_b))
        return fail(HEXL_B200_EINVAL, "poly_multiply: buffers must be 16-byte aligned");
    if (q < 2 || q >> 62) return fail(HEXL_B200_EINVAL, "poly_multiply: modulus must be in [2, 2^62)");
    if (inv_n >= q || inv_n_w >= q)
        return fail(HEXL_B200_EINVAL, "poly_multiply: inv_n / inv_n_w must be reduced mod q");
    if (batch == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int variant = g_ntt_variant.load() & 1;     // the range vote always runs on caller data
    // per-stream scratch: packed forward / inverse twiddles, one deferred list, and NTT(b) of a chunk
    const uint64_t chunk_max = 2048;
    const uint64_t chunk = batch < chunk_max ? batch : chunk_max;
    const size_t list_bytes = ((chunk + 2) * 4 + 255) & ~(size_t)255;
    // (same layout as the NTT entry points: [1M,2M) holds the FP64-pipe tables)
    const size_t tb_off = (size_t)(2u << 20) + list_bytes;
    uint8_t* scratch = nullptr;
    cudaError_t e = g_scratch.get(st, tb_off + (size_t)chunk * n * 8, (void**)&scratch);
    if (e != cudaSuccess) return cuda_fail(e, "poly_multiply: scratch");
    hb::TwPair* pf = reinterpret_cast<hb::TwPair*>(scratch);
    hb::TwPair* pi = reinterpret_cast<hb::TwPair*>(scratch + (1u << 19));
    uint32_t* list = reinterpret_cast<uint32_t*>(scratch + (2u << 20));
    uint64_t* tb = reinterpret_cast<uint64_t*>(scratch + tb_off);
    hb::PackExtra px;
    if (g_fp64_path.load() && hb::fp64_modulus_ok(q)) {
        px.fwd_d = reinterpret_cast<hb::TwPair*>(scratch + (1u << 20));
        px.inv_d = reinterpret_cast<hb::TwPair*>(scratch + (3u << 19));
        px.q = q;
    }
    e = hb::launch_pack_twiddles((uint32_t)logn, variant, d_roots, d_precon, pf, d_inv_roots, d_precon_inv, pi, list, st, px);
    if (e != cudaSuccess) return cuda_fail(e, "poly_multiply: pack twiddles");
    hb::ModTab t = make_modtab(q, inv_n, inv_n_w, pf, pi, logn, nullptr, nullptr);
    int launches = 1;
    if (px.fwd_d) {
        t.ftwd = px.fwd_d;
        t.itwd = px.inv_d;
        t.fp64_ok = 1;
    }
    const bool fused = (variant & 1) && hb::polymul_fused_available(t, (uint32_t)logn);
    for (uint64_t off = 0; off < batch; off += chunk) {
        const uint64_t cnt = batch - off < chunk ? batch - off : chunk;
        uint64_t* res = d_result + off * n;
        if (fused) {
            // one launch: NTT(a) parked in tensor memory, NTT(b), product, INTT (polymul_fused.cu); items with
            // out-of-contract words go through the exact kernels behind it (deferred list, normally empty)
            if (off && (e = cudaMemsetAsync(list, 0, 8, st)) != cudaSuccess) return cuda_fail(e, "poly_multiply: list");
            e = hb::launch_polymul_fused(res, d_a + off * n, d_b + off * n, t, cnt, list, st);
            if (e != cudaSuccess) return cuda_fail(e, "poly_multiply: fused kernel");
            e = hb::launch_polymul_deferred(res, tb, d_a + off * n, d_b + off * n, t, cnt, list, st);
            if (e != cudaSuccess) return cuda_fail(e, "poly_multiply: deferred items");
            launches += 4;
            continue;
        }
        // NTT(a) -> result, NTT(b) -> scratch; out-of-contract words take the exact kernel (deferred list)
        if (off && (e = cudaMemsetAsync(list, 0, 8, st)) != cudaSuccess) return cuda_fail(e, "poly_multiply: list");
        e = hb::launch_ntt_fwd(res, t, (uint32_t)logn, cnt, variant, list, st, &launches, d_a + off * n);
        if (e != cudaSuccess) return cuda_fail(e, "poly_multiply: forward a");
        if ((e = cudaMemsetAsync(list, 0, 8, st)) != cudaSuccess) return cuda_fail(e, "poly_multiply: list");
        e = hb::launch_ntt_fwd(tb, t, (uint32_t)logn, cnt, variant, list, st, &launches, d_b + off * n);
        if (e != cudaSuccess) return cuda_fail(e, "poly_multiply: forward b");
        e = hb::launch_ntt_inv_mul(res, tb, t, (uint32_t)logn, cnt, variant, st);
        if (e != cudaSuccess) return cuda_fail(e, "poly_multiply: inverse");
        ++launches;
    }
    g_launches += launches;
    return 0;
}

int hexl_b200_dyadic_multiply(uint64_t* d_results, const uint64_t* d_op1, const uint64_t* d_op2,
                              uint64_t n, const uint64_t* d_moduli, uint64_t n_moduli,
                              uint64_t batch, int moduli_per_item, void* stream) {
    if (!d_results || !d_op1 || !d_op2 || !d_moduli)
        return fail(HEXL_B200_EINVAL, "dyadic_multiply: NULL pointer");
    if (n == 0 || (n & 1) || n > (1u << 20))
        return fail(HEXL_B200_EINVAL, "dyadic_multiply: n=%llu must be even and <= 2^20",
                    (unsigned long long)n);
    if (n_moduli == 0 || n_moduli > 4096)
        return fail(HEXL_B200_EINVAL, "dyadic_multiply: n_moduli=%llu out of range",
                    (unsigned long long)n_moduli);
    if (!aligned16(d_results) || !aligned16(d_op1) || !aligned16(d_op2))
        return fail(HEXL_B200_EINVAL, "dyadic_multiply: buffers must be 16-byte aligned");
    if (batch == 0) return 0;
    // per-stream scratch (shared with the NTT calls of the stream, which are ordered before / after us)
    uint8_t* scratch = nullptr;
    cudaError_t e = g_scratch.get((cudaStream_t)stream, hb::dyadic_scratch_bytes(n_moduli, batch, moduli_per_item),
                                  (void**)&scratch);
    if (e != cudaSuccess) return cuda_fail(e, "dyadic_multiply: scratch");
    e = hb::launch_dyadic(d_results, d_op1, d_op2, n, d_moduli, n_moduli, batch, moduli_per_item, scratch,
                          (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "dyadic_multiply launch");
    g_launches += 2;
    return 0;
}

int hexl_b200_ks_plan_create(hexl_b200_ks_plan** out, uint64_t n, uint64_t D, uint64_t K, uint64_t R,
                             uint64_t C, const uint64_t* moduli, const uint64_t* const* keys,
                             const uint64_t* msf, const uint64_t* twiddles) {
    if (!out) return fail(HEXL_B200_EINVAL, "ks_plan_create: NULL plan pointer");
    *out = nullptr;
    const int logn = ilog2_exact(n);
    // reference checks: host/src/keyswitch.cpp:23-34 (n in {1024..16384},
    // key_component_count == 2, moduli in [2^16, 2^52]); the FPGA's
    // key_modulus_size <= 7 limit is lifted (BASELINE config uses 8).
    if (logn < 0 || !hb::ntt_shape_supported((uint32_t)logn))
        return fail(HEXL_B200_EINVAL, "keyswitch: n=%llu unsupported", (unsigned long long)n);
    if (C != 2) return fail(HEXL_B200_EINVAL, "keyswitch: key_component_count must be 2");
    if (D == 0 || K < 2 || D + 1 > K || K > 64)
        return fail(HEXL_B200_EINVAL, "keyswitch: need 1 <= decomp < key_modulus_size <= 64");
    if (R != D + 1) return fail(HEXL_B200_EINVAL, "keyswitch: rns_modulus_size must be decomp+1");
    if (!moduli || !keys || !msf) return fail(HEXL_B200_EINVAL, "keyswitch: NULL pointer");
    for (uint64_t i = 0; i < K; ++i) {
        if (moduli[i] < 2 || moduli[i] >> 61)
            return fail(HEXL_B200_EINVAL, "keyswitch: modulus %llu out of range [2, 2^61)",
                        (unsigned long long)moduli[i]);
        if ((moduli[i] - 1) % (2 * n))
            return fail(HEXL_B200_EINVAL, "keyswitch: modulus %llu is not 1 mod 2n",
                        (unsigned long long)moduli[i]);
    }
    for (uint64_t j = 0; j < D; ++j)
        if (!keys[j]) return fail(HEXL_B200_EINVAL, "keyswitch: k_switch_keys[%d] is NULL", (int)j);

    auto* p = new hexl_b200_ks_plan;
    cudaError_t e = cudaGetDevice(&p->device);
    if (e != cudaSuccess) {
        delete p;
        return cuda_fail(e, "cudaGetDevice");
    }
    p->n = n; p->D = D; p->K = K; p->R = R;

    std::vector<uint64_t> h_tables(K * 4 * n), h_small(4 * K);   // msf, its Shoup factors, {centred msf, msf / q} as doubles
    std::vector<hb::ModTab> h_tabs(K);
    std::vector<hb::Divisor> h_divs(K);
    auto cleanup = [&](int rc) {
        hexl_b200_ks_plan_destroy(p);
        return rc;
    };
    const int ks_variant = hb::ks_variant_for((uint32_t)logn);
    const size_t fe = hb::packed_fwd_entries((uint32_t)logn, ks_variant), ie = hb::packed_inv_entries((uint32_t)logn, ks_variant);
    uint64_t* d_raw = nullptr;      // raw tables, only needed while packing
    // second half: the same tables in the FP64-pipe format (modarith.cuh)
    if ((e = cudaMalloc(&p->d_packed, 2 * K * (fe + ie) * sizeof(hb::TwPair)))) return cleanup(cuda_fail(e, "cudaMalloc tables"));
    bool fp64_all = g_fp64_path.load() != 0;
    for (uint64_t i = 0; i < K; ++i) fp64_all = fp64_all && hb::fp64_modulus_ok(moduli[i]);
    if ((e = cudaMalloc(&d_raw, K * 4 * n * 8))) return cleanup(cuda_fail(e, "cudaMalloc raw tables"));
    if ((e = cudaMalloc(&p->d_keys, D * 2 * K * n * 8))) return cleanup(cuda_fail(e, "cudaMalloc keys"));
    if ((e = cudaMalloc(&p->d_small, 4 * K * 8))) return cleanup(cuda_fail(e, "cudaMalloc small"));
    if ((e = cudaMalloc(&p->d_tabs, K * sizeof(hb::ModTab)))) return cleanup(cuda_fail(e, "cudaMalloc tabs"));
    if ((e = cudaMalloc(&p->d_divs, K * sizeof(hb::Divisor)))) return cleanup(cuda_fail(e, "cudaMalloc divs"));

    for (uint64_t i = 0; i < K; ++i) {
        const uint64_t q = moduli[i];
        nt::Tables t = twiddles ? nt::tables_from_keyswitch_block(n, q, twiddles + i * 4 * n)
                                : nt::make_tables(n, q);
        if (t.roots.size() != n || t.inv_n == 0)
            return cleanup(fail(HEXL_B200_EINVAL, "keyswitch: no primitive 2n-th root mod %llu",
                                (unsigned long long)q));
        uint64_t* h = h_tables.data() + i * 4 * n;
        memcpy(h, t.roots.data(), n * 8);
        memcpy(h + n, t.precon.data(), n * 8);
        memcpy(h + 2 * n, t.inv_roots.data(), n * 8);
        memcpy(h + 3 * n, t.precon_inv.data(), n * 8);
        hb::TwPair* pk = p->d_packed + i * (fe + ie);
        h_tabs[i] = make_modtab(q, t.inv_n, t.inv_n_w, pk, pk + fe, logn);
        if (fp64_all) {
            h_tabs[i].ftwd = pk + K * (fe + ie);
            h_tabs[i].itwd = pk + K * (fe + ie) + fe;
            h_tabs[i].fp64_ok = 1;
        }
        h_divs[i] = hb::make_divisor(q);
        h_small[i] = msf[i] % q;                       // host/src/fpga.cpp:1057-1061
        h_small[K + i] = nt::shoup(h_small[i], q);
        const double msf_c = hb::fp_centred(h_small[i], q);
        h_small[2 * K + i] = hb::d2u(msf_c);
        h_small[3 * K + i] = hb::d2u(hb::fp_quot(msf_c, q));
    }
    e = cudaMemcpy(d_raw, h_tables.data(), K * 4 * n * 8, cudaMemcpyHostToDevice);
    for (uint64_t i = 0; i < K && e == cudaSuccess; ++i) {
        const uint64_t* d = d_raw + i * 4 * n;
        hb::TwPair* pk = p->d_packed + i * (fe + ie);
        e = hb::launch_pack_twiddles((uint32_t)logn, ks_variant, d, d + n, pk, d + 2 * n, d + 3 * n, pk + fe, nullptr, 0);
        if (e == cudaSuccess && fp64_all)
            e = hb::launch_pack_twiddles_fp64((uint32_t)logn, ks_variant, d, pk + K * (fe + ie), d + 2 * n,
                                              pk + K * (fe + ie) + fe, moduli[i], 0);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d_raw);
    if (e != cudaSuccess) return cleanup(cuda_fail(e, "upload / pack tables"));
    g_launches += fp64_all ? 2 * K : K;
    for (uint64_t j = 0; j < D; ++j)
        if ((e = cudaMemcpy(p->d_keys + j * 2 * K * n, keys[j], 2 * K * n * 8, cudaMemcpyHostToDevice)))
            return cleanup(cuda_fail(e, "upload keys"));
    if ((e = cudaMemcpy(p->d_small, h_small.data(), 4 * K * 8, cudaMemcpyHostToDevice)))
        return cleanup(cuda_fail(e, "upload msf"));
    if ((e = cudaMemcpy(p->d_tabs, h_tabs.data(), K * sizeof(hb::ModTab), cudaMemcpyHostToDevice)))
        return cleanup(cuda_fail(e, "upload tabs"));
    if ((e = cudaMemcpy(p->d_divs, h_divs.data(), K * sizeof(hb::Divisor), cudaMemcpyHostToDevice)))
        return cleanup(cuda_fail(e, "upload divisors"));
    g_h2d += (K * 4 * n + D * 2 * K * n + 2 * K) * 8;

    p->dev.fast_ok = 1;
    for (uint64_t i = 0; i < K; ++i)
        if (!h_tabs[i].fwd_fast_ok || !h_tabs[i].inv_fast_ok) p->dev.fast_ok = 0;
    p->dev.fp64_ok = (p->dev.fast_ok && fp64_all) ? 1u : 0u;
    uint64_t qmin = moduli[0], qmax = moduli[0];
    for (uint64_t i = 1; i < K; ++i) {
        qmin = moduli[i] < qmin ? moduli[i] : qmin;
        qmax = moduli[i] > qmax ? moduli[i] : qmax;
    }
    p->dev.s2_no_reduce = (p->dev.fp64_ok && qmax <= qmin + (qmin >> 2)) ? 1u : 0u;   // q_j - 1 < vote bound of every q_r
    p->dev.fp64_alt_ok = p->dev.fp64_ok;
    for (uint64_t i = 0; i < K; ++i)
        if (!h_tabs[i].fp64_alt_ok) p->dev.fp64_alt_ok = 0;
    p->dev.logn = (uint32_t)logn;
    p->dev.D = (uint32_t)D; p->dev.K = (uint32_t)K; p->dev.R = (uint32_t)R;
    p->dev.walk_blocked = (uint32_t)g_ks_blocked.load();
    p->dev.fD = hb::make_fastdiv((uint32_t)D);
    p->dev.fDm1 = hb::make_fastdiv((uint32_t)D - 1);
    p->dev.fDD = hb::make_fastdiv((uint32_t)(D * D));
    p->dev.tabs = p->d_tabs; p->dev.divs = p->d_divs; p->dev.keys = p->d_keys;
    p->dev.msf = p->d_small; p->dev.msf_p = p->d_small + K;
    p->dev.msf_fp = p->dev.fp64_ok ? reinterpret_cast<const double*>(p->d_small + 2 * K) : nullptr;
    p->dev.keys_sh = nullptr;
    if (p->dev.fast_ok) {
        if ((e = cudaMalloc(&p->d_keys_sh, D * 2 * K * n * sizeof(hb::TwPair))))
            return cleanup(cuda_fail(e, "cudaMalloc Shoup keys"));
        if ((e = hb::launch_ks_prepare_keys(p->dev, p->d_keys_sh, 0)) || (e = cudaDeviceSynchronize()))
            return cleanup(cuda_fail(e, "prepare keys"));
        p->dev.keys_sh = p->d_keys_sh;
        g_launches += 1;
    }
    p->dev.keys_fp = nullptr;
    if (p->dev.fast_ok && p->dev.fp64_alt_ok && logn == 14) {
        if ((e = cudaMalloc(&p->d_keys_fp, D * 2 * K * n * sizeof(hb::TwPair))))
            return cleanup(cuda_fail(e, "cudaMalloc FP64 keys"));
        if ((e = hb::launch_ks_prepare_keys_fp64(p->dev, p->d_keys_fp, 0)) || (e = cudaDeviceSynchronize()))
            return cleanup(cuda_fail(e, "prepare FP64 keys"));
        p->dev.keys_fp = p->d_keys_fp;
        g_launches += 1;
    }
    p->dev.keys_fused = nullptr;
    {
        hb::KsDev probe = p->dev;
        probe.keys_fused = (const void*)1;
        if (hb::g_ks_fused && hb::ks_fused_available(probe, probe.keys_fused)) {
            if ((e = cudaMalloc(&p->d_keys_fused, hb::ks_fused_key_bytes(p->dev))))
                return cleanup(cuda_fail(e, "cudaMalloc fused keys"));
            if ((e = hb::launch_ks_prepare_keys_fused(p->dev, p->d_keys_fused, 0)) || (e = cudaDeviceSynchronize()))
                return cleanup(cuda_fail(e, "prepare fused keys"));
            p->dev.keys_fused = p->d_keys_fused;
            g_launches += 1;
        }
    }
    *out = p;
    return 0;
}

int hexl_b200_ks_plan_destroy(hexl_b200_ks_plan* p) {
    if (!p) return 0;
    cudaFree(p->d_packed); cudaFree(p->d_keys); cudaFree(p->d_keys_sh); cudaFree(p->d_keys_fp); cudaFree(p->d_keys_fused); cudaFree(p->d_small);
    cudaFree(p->d_tabs); cudaFree(p->d_divs); cudaFree(p->ws);
    delete p;
    return 0;
}

int hexl_b200_keyswitch(hexl_b200_ks_plan* p, uint64_t* d_result, const uint64_t* d_t, uint64_t batch,
                        void* stream) {
    if (!p) return fail(HEXL_B200_EINVAL, "keyswitch: NULL plan");
    if (batch == 0) return 0;
    if (!d_result || !d_t) return fail(HEXL_B200_EINVAL, "keyswitch: NULL pointer");
    if (!aligned16(d_result) || !aligned16(d_t))
        return fail(HEXL_B200_EINVAL, "keyswitch: buffers must be 16-byte aligned");
    const uint64_t n = p->n, D = p->D;
    const uint64_t per_item = hb::ks_scratch_words_per_item(p->dev);
    uint64_t chunk = ((uint64_t)g_ks_workspace_mb.load() << 20) / 8 / per_item;
    if (chunk < 1) chunk = 1;
    if (chunk > 16384) chunk = 16384;
    // equal chunks: a short last chunk runs the same number of launches over a fraction of the work
    chunk = (batch + (batch + chunk - 1) / chunk - 1) / ((batch + chunk - 1) / chunk);
    std::lock_guard<std::mutex> lk(p->mu);
    if (p->ws_words < chunk * per_item) {
        // stream-ordered work may still use the old buffer
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(e, "keyswitch: sync before workspace growth");
        cudaFree(p->ws);
        p->ws = nullptr; p->ws_words = 0;
        if ((e = cudaMalloc(&p->ws, chunk * per_item * 8))) return cuda_fail(e, "keyswitch: workspace");
        p->ws_words = chunk * per_item;
    }
    for (uint64_t off = 0; off < batch; off += chunk) {
        const uint64_t items = batch - off < chunk ? batch - off : chunk;
        int launches = 0;
        cudaError_t e = hb::launch_ks_chunk(p->dev, d_result + off * 2 * D * n, d_t + off * D * n, items, p->ws,
                                            (cudaStream_t)stream, &launches);
        if (e != cudaSuccess) return cuda_fail(e, "keyswitch launch");
        g_launches += launches;
    }
    return 0;
}

}  // extern "C"
