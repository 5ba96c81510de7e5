// ntt_core.cuh -- thread-level structure of the shared-memory negacyclic NTT.
//
// One polynomial (N = 2^LOGN uint64 words, N*8 bytes) lives in shared memory;
// a CTA of NT = N/E threads (E = 2^LOGE words per thread) walks over the LOGN
// radix-2 stages in a few "passes".  In a pass every thread pulls E words into
// registers, runs up to LOGE consecutive butterfly stages on them
// (register-resident radix-2^R groups) and puts them back in place, so a
// polynomial makes ceil((LOGN-4)/LOGE)+1 shared-memory round trips, not LOGN.
//
//   forward (Cooley-Tukey, reference tests/test_utils/ntt.cpp:494-547):
//     head passes  stages 0 .. LOGN-5   (strides >= 16, lanes along `lo`)
//     tail pass    stages LOGN-4..LOGN-1 (16 contiguous words per thread row)
//   inverse (Gentleman-Sande, ntt.cpp:580-659): the mirror image.
//
// The first pass of a transform loads all E words, so the kernel can vote on
// the input range and pick the arithmetic (exact reference op sequence vs the
// fast path, see modarith.cuh); the last pass keeps its E words in registers
// after reading them, which frees the shared buffer for the TMA prefetch of
// the next polynomial while the last stages are still being computed.
//
// Shared-memory layout: 128-byte rows of 16 words; the 16-byte chunk c of row
// r is stored at chunk c ^ (r & 7) -- the TMA SWIZZLE_128B pattern, so a tensor
// TMA load lands the polynomial already in this layout.  Both access shapes
// are conflict-free: head passes touch one whole row per half warp (a
// permutation inside a row still covers all 32 banks), the tail pass reads one
// 16-byte chunk of 8 different rows per quarter warp.
//
// Twiddles are read from a PACKED table (built once per call / per plan by
// k_pack_twiddles): for every (pass, group) the 2^R - 1 (root, precon) pairs
// the group needs sit in one contiguous, 16-byte-interleaved block, so a
// thread issues 16-byte loads at consecutive addresses instead of 8-byte
// gathers from two separate arrays.
//
// All functions are host+device so tests/cpu_emul can replay them thread by
// thread on the CPU to validate the index math before GPU time is spent.
#pragma once
#include <type_traits>

#include "modarith.cuh"

namespace hb {

// LOGROW_ = log2(words per 128-byte shared-memory row): 4 for uint64 words, 5
// for the uint32 small-modulus path.  The "tail" pass owns the last LOGROW
// stages (one full row per thread), the head passes the first LOGN - LOGROW.
// WARPTAIL_ = 1: the tail rows are dealt out by warp (rows 64w .. 64w+63 to warp w) instead of by
// thread index, so that a warp's tail rows are exactly the words it handled in the last head pass
// and the block barrier between the two becomes a __syncwarp (ntt_block.cuh).
template <int LOGN_, int LOGE_, int LOGROW_ = 4, int WARPTAIL_ = 0>
struct NttCfg {
    static constexpr int LOGN = LOGN_, LOGE = LOGE_, LOGROW = LOGROW_;
    static constexpr bool WARPTAIL = WARPTAIL_ != 0;
    static constexpr int N = 1 << LOGN, E = 1 << LOGE, NT = N / E, ROW = 1 << LOGROW;
    using elem = typename std::conditional<LOGROW_ == 4, uint64_t, uint32_t>::type;
    // CTAs per SM the register allocation must allow (launch bounds)
    static constexpr int MIN_CTAS = (NT * 128 <= 32768) ? (65536 / (NT * 128) > 4 ? 4 : 65536 / (NT * 128)) : 1;
    static constexpr int HEAD = LOGN - LOGROW;            // stages outside the one-row tail
    static constexpr int NP = (HEAD + LOGE - 1) / LOGE;   // number of head passes
    static constexpr int BASE = HEAD / NP, REM = HEAD % NP;
    static_assert(LOGROW == 4 || LOGROW == 5, "rows hold 16 uint64 or 32 uint32 words");
    static_assert(LOGE >= LOGROW && LOGN >= LOGE + LOGROW, "unsupported NTT shape");
    // two rows per thread and one full-size group per thread in the last head pass
    static_assert(!WARPTAIL || (LOGE == LOGROW + 1 && HEAD % LOGE == 0), "WARPTAIL needs E = 2 rows = one last-pass group");
    static constexpr int pass_r(int p) { return BASE + (p < REM ? 1 : 0); }
    static constexpr int pass_s0(int p) {
        int s = 0;
        for (int i = 0; i < p; ++i) s += pass_r(i);
        return s;
    }
    // ---- packed twiddle table geometry (entries) ----
    // forward: head pass P owns 2^S0 groups of 2^R entries, then N/ROW tail rows of ROW
    static constexpr int fwd_off(int p) {  // p == NP -> tail
        int o = 0;
        for (int i = 0; i < p; ++i) o += 1 << (pass_s0(i) + pass_r(i));
        return o;
    }
    static constexpr int FWD_ENTRIES = fwd_off(NP) + N;
    // inverse: N/ROW tail rows of ROW first, then head pass P: (N >> (U0+R)) groups of 2^R
    static constexpr int inv_u0(int p) { return LOGROW + pass_s0(p); }
    static constexpr int inv_off(int p) {
        int o = N;
        for (int i = 0; i < p; ++i) o += N >> inv_u0(i);
        return o;
    }
    static constexpr int INV_ENTRIES = inv_off(NP);
};

// 16-byte chunk c of 128-byte row r lives at chunk c ^ (r & 7)
HB_HD uint32_t swz(uint32_t idx) { return idx ^ (((idx >> 4) & 7u) << 1); }          // uint64 words
HB_HD uint32_t swz32(uint32_t idx) { return idx ^ (((idx >> 5) & 7u) << 2); }        // uint32 words
template <class T>
HB_HD uint32_t swz_t(uint32_t idx) {
    return sizeof(T) == 8 ? swz(idx) : swz32(idx);
}

// compile-time loop: f(std::integral_constant<int, I>) for I in [0, COUNT).
// (#pragma unroll does not reliably flatten loops whose bounds depend on an
// outer unrolled index, and a rolled loop would push the register arrays to
// local memory.)
template <int I, int COUNT, class F>
HB_HD void static_for(const F& f) {
    if constexpr (I < COUNT) {
        f(std::integral_constant<int, I>());
        static_for<I + 1, COUNT>(f);
    }
}

// ---------------------------------------------------------------------------
// packed twiddles
// ---------------------------------------------------------------------------
struct alignas(16) TwPair {
    uint64_t w, wp;
};

HB_HD TwPair ldpair(const TwPair* p) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(p));
    TwPair r;
    r.w = t.x;
    r.wp = t.y;
    return r;
#else
    return *p;
#endif
}

// L1 prefetch of one packed entry per lane (no destination register): used one
// butterfly stage ahead in the tail passes, whose twiddles are unique per row
// and would otherwise expose the full L2 latency at first use.
HB_HD void prefetch_pair(const TwPair* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

HB_HD int ilog2_u32(uint32_t x) {
    int l = 0;
    while (x >>= 1) ++l;
    return l;
}

// source index (into roots[] / precon[]) of packed forward entry e, or -1 for padding
template <class C>
HB_HD int fwd_pack_src(uint32_t e) {
    int s0 = 0;
    for (int p = 0; p <= C::NP; ++p) {
        const int r = (p == C::NP) ? C::LOGROW : C::pass_r(p);
        const uint32_t count = 1u << (s0 + r);
        if (e < count) {
            uint32_t hi = e >> r, slot = e & ((1u << r) - 1);
            if (p == C::NP) {  // tail: [row/32][slot][row%32] so a warp's loads coalesce
                slot = (e >> 5) & (C::ROW - 1);
                hi = ((e >> (5 + C::LOGROW)) << 5) | (e & 31u);
            }
            if (slot == 0) return -1;
            const int d = ilog2_u32(slot);
            return (int)((((1u << s0) + hi) << d) + (slot - (1u << d)));
        }
        e -= count;
        s0 += r;
    }
    return -1;
}

// source index (into inv_roots[] / precon_inv[]) of packed inverse entry e
template <class C>
HB_HD int inv_pack_src(uint32_t e) {
    constexpr uint32_t N = C::N;
    int u0 = 0;
    for (int p = -1; p < C::NP; ++p) {   // p == -1: tail
        const int r = (p < 0) ? C::LOGROW : C::pass_r(p);
        const uint32_t count = N >> u0;  // groups (N >> (u0+r)) * 2^r
        if (e < count) {
            uint32_t hi = e >> r, slot = e & ((1u << r) - 1);
            if (p < 0) {  // tail: [row/32][slot][row%32]
                slot = (e >> 5) & (C::ROW - 1);
                hi = ((e >> (5 + C::LOGROW)) << 5) | (e & 31u);
            }
            if (slot == 0) return -1;
            const int ee = ilog2_u32(slot);          // = R-1-d
            const int d = r - 1 - ee;
            const uint32_t base_u = 1u + N - (N >> (u0 + d));
            return (int)(base_u + (hi << ee) + (slot - (1u << ee)));
        }
        e -= count;
        u0 += r;
    }
    return -1;
}

// ---------------------------------------------------------------------------
// arithmetic policies: element type, twiddle type, butterflies
// ---------------------------------------------------------------------------
struct InvScale {
    uint64_t inv_n, inv_n_p, inv_n_w, inv_n_w_p;
};

// reference op sequence, bit-exact even on out-of-range words
struct ExactArith {
    using elem = uint64_t;
    using Tw = TwPair;
    static constexpr bool kFp64 = false;
    static constexpr bool kSmemHead = false;   // head-pass twiddles in shared memory (Fp64ArithS)
    static constexpr bool kTmemTail = false;   // tail-pass twiddles in tensor memory (Fp64ArithST)
    HB_HD Tw ld_head(const Tw* p) const { return ldpair(p); }
    uint64_t q, twoq;
    InvScale sc;
    uint32_t lazy_out;     // output_mod_factor 4 (forward) / 2 (inverse) of the reference: no final correction
    HB_HD Tw ld(const Tw* p) const { return ldpair(p); }
    HB_HD void fwd(uint64_t& X, uint64_t& Y, const TwPair& t) const { fwd_bfly(X, Y, t.w, t.wp, q, twoq); }
    template <int S> HB_HD void fwd_at(uint64_t& X, uint64_t& Y, const TwPair& t) const { fwd(X, Y, t); }
    HB_HD uint64_t fwd_final(uint64_t x) const {      // ntt.cpp:535-546
        if (lazy_out) return x;
        x -= (x >= twoq) ? twoq : 0;
        x -= (x >= q) ? q : 0;
        return x;
    }
    HB_HD void inv(uint64_t& X, uint64_t& Y, const TwPair& t) const { inv_bfly(X, Y, t.w, t.wp, q, twoq); }
    HB_HD void inv_last(uint64_t& X, uint64_t& Y) const {
        inv_last_bfly(X, Y, sc.inv_n, sc.inv_n_p, sc.inv_n_w, sc.inv_n_w_p, q, twoq, lazy_out != 0);
    }
    static constexpr bool kLazyInv = false;
    template <int E> HB_HD void inv_at(uint64_t& X, uint64_t& Y, const TwPair& t) const { inv(X, Y, t); }
    template <int E> HB_HD void inv_last_at(uint64_t& X, uint64_t& Y) const { inv_last(X, Y); }
};

// any-correct-algorithm path for in-contract inputs (modarith.cuh)
struct FastArith {
    using elem = uint64_t;
    using Tw = TwPair;
    static constexpr bool kFp64 = false;
    static constexpr bool kSmemHead = false;   // head-pass twiddles in shared memory (Fp64ArithS)
    static constexpr bool kTmemTail = false;   // tail-pass twiddles in tensor memory (Fp64ArithST)
    HB_HD Tw ld_head(const Tw* p) const { return ldpair(p); }
    FastMod m;
    InvScale sc;
    HB_HD Tw ld(const Tw* p) const { return ldpair(p); }
    HB_HD void fwd(uint64_t& X, uint64_t& Y, const TwPair& t) const { fwd_bfly_fast(X, Y, t.w, t.wp, m); }
    template <int S> HB_HD void fwd_at(uint64_t& X, uint64_t& Y, const TwPair& t) const { fwd(X, Y, t); }
    HB_HD uint64_t fwd_final(uint64_t x) const { return reduce_small_multiple(x, m); }
    HB_HD void inv(uint64_t& X, uint64_t& Y, const TwPair& t) const { inv_bfly_fast(X, Y, t.w, t.wp, m); }
    HB_HD void inv_last(uint64_t& X, uint64_t& Y) const {
        inv_last_bfly_fast(X, Y, sc.inv_n, sc.inv_n_p, sc.inv_n_w, sc.inv_n_w_p, m);
    }
    static constexpr bool kLazyInv = false;
    template <int E> HB_HD void inv_at(uint64_t& X, uint64_t& Y, const TwPair& t) const { inv(X, Y, t); }
    template <int E> HB_HD void inv_last_at(uint64_t& X, uint64_t& Y) const { inv_last(X, Y); }
};

// FastArith whose final forward reduction looks the multiple of q up in a 64-entry
// shared-memory table (kernels only; the table is built per item by the CTA)
struct FastArithTab : FastArith {
    const uint64_t* kq;
    HB_HD uint64_t fwd_final(uint64_t x) const { return reduce_by_table(x, m, kq); }
};

// FP64-pipe arithmetic (modarith.cuh): the registers and the shared buffer hold the bit
// patterns of integer-valued doubles, the packed twiddles are {centred root, root / q} as
// doubles (k_pack_twiddles_fp64); words are converted right after the range vote and come
// back as canonical integers from fwd_final / inv_last.
struct Fp64Arith {
    using elem = uint64_t;
    using Tw = TwPair;
    static constexpr bool kFp64 = true;
    static constexpr bool kSmemHead = false;   // head-pass twiddles in shared memory (Fp64ArithS)
    static constexpr bool kTmemTail = false;   // tail-pass twiddles in tensor memory (Fp64ArithST)
    HB_HD Tw ld_head(const Tw* p) const { return ldpair(p); }
    static constexpr bool kLazyInv = false;
    Fp64Mod m;
    HB_HD Tw ld(const Tw* p) const { return ldpair(p); }
    HB_HD uint64_t enter_fwd(uint64_t x) const { return d2u(fp_from_int(x)); }               // [0, 1.25q)
    HB_HD uint64_t enter_inv(uint64_t x) const { return d2u(fp_from_int(x)); }               // [0, 1.25q): stage E = 0 copes
    HB_HD void fwd(uint64_t& X, uint64_t& Y, const TwPair& t) const { fwd_bfly_fp64(X, Y, t.w, t.wp, m); }
    template <int S> HB_HD void fwd_at(uint64_t& X, uint64_t& Y, const TwPair& t) const { fwd(X, Y, t); }
    HB_HD uint64_t fwd_final(uint64_t x) const { return fp_to_canonical(u2d(x), m); }
    // E = 0 marks the very first stage of a transform (inv_tail_compute): words as they entered
    template <int E> HB_HD void inv_at(uint64_t& X, uint64_t& Y, const TwPair& t) const {
        if constexpr (E == 0) inv_bfly_fp64_first(X, Y, t.w, t.wp, m);
        else inv_bfly_fp64(X, Y, t.w, t.wp, m);
    }
    template <int E> HB_HD void inv_last_at(uint64_t& X, uint64_t& Y) const { inv_last_bfly_fp64(X, Y, m); }
};

// Fp64Arith with the twiddles of the head passes in shared memory (plain batched calls: a persistent
// CTA keeps one modulus, so its ~17 KiB of head-pass twiddles are copied next to the polynomial
// buffer once per launch).  A broadcast LDS.128 costs 2.9 cycles of the scheduler where an L1 hit
// costs 4.6-6.5 (tools/ubench4.cu).  The head-pass functions receive a "pointer" whose numeric value
// is the 32-bit shared-space address of the entry (fwd_base / inv_base below), so the index math
// of ntt_core.cuh applies unchanged; the tail twiddles stay in global memory.
struct Fp64ArithS : Fp64Arith {
    static constexpr bool kSmemHead = true;
    uint32_t head_s;   // shared-space byte address of head entry 0
    HB_HD const TwPair* fwd_base() const { return reinterpret_cast<const TwPair*>((uintptr_t)head_s); }
    template <class C>
    HB_HD const TwPair* inv_base() const {   // inverse tables start with the N tail entries
        return reinterpret_cast<const TwPair*>((uintptr_t)head_s - (uintptr_t)C::inv_off(0) * sizeof(TwPair));
    }
    HB_HD Tw ld_head(const Tw* p) const {
#if defined(__CUDA_ARCH__)
        TwPair t;
        // volatile: a pure load of a loop-invariant address would be hoisted out of the persistent
        // loop over the polynomials (all 62 of them, 248 registers)
        asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(t.w), "=l"(t.wp) : "r"((uint32_t)(uintptr_t)p));
        return t;
#else
        return *p;
#endif
    }
};

// Forward-only FP64 arithmetic with a full correction every other stage (modarith.cuh, fwd_bfly_fp64_a / _b):
// S = the global stage index, even stages correct.  Moduli up to 2^51 (1 + 1/32).
struct Fp64AltArith : Fp64Arith {
    template <int S> HB_HD void fwd_at(uint64_t& X, uint64_t& Y, const TwPair& t) const {
        if constexpr ((S & 1) == 0) fwd_bfly_fp64_a(X, Y, t.w, t.wp, m);
        else fwd_bfly_fp64_b(X, Y, t.w, t.wp, m);
    }
    HB_HD uint64_t fwd_final(uint64_t x) const { return fp_to_canonical_full(u2d(x), m); }
};
struct Fp64AltArithS : Fp64ArithS {
    template <int S> HB_HD void fwd_at(uint64_t& X, uint64_t& Y, const TwPair& t) const {
        if constexpr ((S & 1) == 0) fwd_bfly_fp64_a(X, Y, t.w, t.wp, m);
        else fwd_bfly_fp64_b(X, Y, t.w, t.wp, m);
    }
    HB_HD uint64_t fwd_final(uint64_t x) const { return fp_to_canonical_full(u2d(x), m); }
};

// Forward transform (full correction every other stage) whose tail hands over the raw doubles, |v| <= 1.92 q
// < 2^52, instead of canonical words: for consumers that multiply on the FP64 pipe anyway (the polynomial
// multiply of polymul_fused.cu, the keyswitch multiply-accumulate k_ks_mac_fp64)
struct Fp64ArithRaw : Fp64AltArith {
    HB_HD uint64_t fwd_final(uint64_t x) const { return x; }
};

// The S policies with the twiddles of the TAIL pass in tensor memory (ntt_block.cuh, tail_tw_to_tmem): a
// persistent CTA of the plain batched calls keeps one modulus, and the 15 twiddles of each of a thread's tail
// rows are the same for every polynomial it transforms, so each thread parks them once per launch in its own
// lane of the SM's otherwise idle TMEM (2 rows x 16 slots x 16 bytes = 128 columns) and reads them back with
// tcgen05.ld (12 cycles) instead of pulling 245 KiB per transform through L2.
struct Fp64ArithST : Fp64ArithS {
    static constexpr bool kTmemTail = true;
    uint32_t ttail;    // tensor-memory address of this thread's slot 0 (row 0)
};
struct Fp64AltArithST : Fp64AltArithS {
    static constexpr bool kTmemTail = true;
    uint32_t ttail;
};
// inverse transform whose last stage hands over non-negative doubles in [0, q) (modarith.cuh, inv_last_bfly_fp64_d)
struct Fp64ArithSTD : Fp64ArithST {
    template <int E> HB_HD void inv_last_at(uint64_t& X, uint64_t& Y) const { inv_last_bfly_fp64_d(X, Y, m); }
};
struct Fp64ArithRawST : Fp64AltArithST {     // raw doubles out (cf. Fp64ArithRaw)
    HB_HD uint64_t fwd_final(uint64_t x) const { return x; }
};

// inverse transform for q < 2^52 without per-stage corrections (modarith.cuh);
// E = log2 of the bound (in units of q) of the words entering the stage
struct LazyInvArith {
    using elem = uint64_t;
    using Tw = TwPair;
    static constexpr bool kFp64 = false;
    static constexpr bool kSmemHead = false;   // head-pass twiddles in shared memory (Fp64ArithS)
    static constexpr bool kTmemTail = false;   // tail-pass twiddles in tensor memory (Fp64ArithST)
    HB_HD Tw ld_head(const Tw* p) const { return ldpair(p); }
    FastMod m;
    InvScale sc;
    static constexpr bool kLazyInv = true;
    HB_HD Tw ld(const Tw* p) const { return ldpair(p); }
    template <int E> HB_HD void inv_at(uint64_t& X, uint64_t& Y, const TwPair& t) const {
        inv_bfly_lazy<E>(X, Y, t.w, t.wp, m);
    }
    template <int E> HB_HD void inv_last_at(uint64_t& X, uint64_t& Y) const {
        inv_last_bfly_lazy<E>(X, Y, sc.inv_n, sc.inv_n_p, sc.inv_n_w, sc.inv_n_w_p, m);
    }
    HB_HD uint64_t reduce(uint64_t x) const { return reduce_mid(x, m); }
};

// Bound schedule of the lazy inverse: words enter the transform below 2q (E = 1),
// every stage adds one to E, and a pass that ends with E >= 8 (and is not the
// last) reduces its words to [0,q) before storing them.
template <class C>
struct InvLazy {
    static constexpr int pass_stages(int p) { return p < 0 ? C::LOGROW : C::pass_r(p); }   // p = -1: tail
    static constexpr bool reduce_after_e(int p, int e_out) { return p < C::NP - 1 && e_out >= 8; }
    static constexpr int e_in(int p) {
        int e = 1;
        for (int i = -1; i < p; ++i) {
            e += pass_stages(i);
            if (reduce_after_e(i, e)) e = 1;
        }
        return e;
    }
    static constexpr int e_out(int p) { return e_in(p) + pass_stages(p); }
    static constexpr bool reduce_after(int p) { return reduce_after_e(p, e_out(p)); }
    static constexpr bool valid() {
        for (int p = -1; p < C::NP; ++p) {
            if (e_in(p) + pass_stages(p) - 1 > 11) return false;     // 2^(E+1) q < 2^64 at every stage
            if (reduce_after(p) && e_out(p) > 10) return false;      // reduce_mid handles < 1024q
        }
        return true;
    }
};

// Small-modulus path (q < 2^30): every word of a lazy Harvey transform stays
// below 4q < 2^32, so the whole butterfly runs in 32-bit arithmetic -- one
// IMAD.WIDE (high half = Shoup quotient), two IMAD and four ALU instructions
// instead of ~18 -- on uint32 registers and a uint32 shared buffer.  Twiddles
// are {w, floor(w*2^32/q)} = {w, hi32 of the caller's 64-bit factor}.
struct Tw32 {
    uint32_t w, wp;
};
struct Small32 {
    uint32_t q, nq, twoq;              // nq = 2^32 - q
    uint32_t inv_n, inv_n_p, inv_n_w, inv_n_w_p;
};
HB_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    // IMAD.WIDE (full rate) and keep the high half; IMAD.HI runs at half rate on sm_100
    uint32_t lo, hi;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
    (void)lo;
    return hi;
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
HB_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
// w*y - floor(y*wp/2^32)*q in [0,2q) for ANY y < 2^32 (w < q < 2^31)
HB_HD uint32_t mul_lazy32(uint32_t y, uint32_t w, uint32_t wp, uint32_t nq) {
    return w * y + mulhi32(y, wp) * nq;
}
struct SmallArith {
    using elem = uint32_t;
    using Tw = Tw32;
    static constexpr bool kSmemHead = false;
    static constexpr bool kTmemTail = false;
    HB_HD Tw ld_head(const Tw* p) const { return ld(p); }
    Small32 m;
    HB_HD Tw ld(const Tw* p) const {
#if defined(__CUDA_ARCH__)
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        return Tw32{t.x, t.y};
#else
        return *p;
#endif
    }
    // x - 2q if x >= 2q (x < 4q): the wrapped difference is huge when x < 2q
    HB_HD uint32_t csub32(uint32_t x, uint32_t b) const { return umin32(x, x - b); }
    HB_HD void fwd(uint32_t& X, uint32_t& Y, const Tw32& t) const {
        const uint32_t tx = csub32(X, m.twoq);
        const uint32_t T = mul_lazy32(Y, t.w, t.wp, m.nq);
        X = tx + T;
        Y = tx + m.twoq - T;
    }
    template <int S> HB_HD void fwd_at(uint32_t& X, uint32_t& Y, const Tw32& t) const { fwd(X, Y, t); }
    HB_HD uint32_t fwd_final(uint32_t x) const { return csub32(csub32(x, m.twoq), m.q); }
    HB_HD void inv(uint32_t& X, uint32_t& Y, const Tw32& t) const {   // values in [0,2q)
        const uint32_t tx = X + Y;
        const uint32_t ty = X + m.twoq - Y;
        X = csub32(tx, m.twoq);
        Y = mul_lazy32(ty, t.w, t.wp, m.nq);
    }
    HB_HD void inv_last(uint32_t& X, uint32_t& Y) const {
        const uint32_t tx = csub32(X + Y, m.twoq);
        const uint32_t ty = X + m.twoq - Y;
        X = csub32(mul_lazy32(tx, m.inv_n, m.inv_n_p, m.nq), m.q);
        Y = csub32(mul_lazy32(ty, m.inv_n_w, m.inv_n_w_p, m.nq), m.q);
    }
    static constexpr bool kLazyInv = false;
    template <int E> HB_HD void inv_at(uint32_t& X, uint32_t& Y, const Tw32& t) const { inv(X, Y, t); }
    template <int E> HB_HD void inv_last_at(uint32_t& X, uint32_t& Y) const { inv_last(X, Y); }
};
// SmallArith with the head-pass twiddles in shared memory (cf. Fp64ArithS: the "pointers" handed to
// the head-pass functions carry 32-bit shared-space addresses)
struct SmallArithS : SmallArith {
    static constexpr bool kSmemHead = true;
    uint32_t head_s;
    HB_HD const Tw32* fwd_base() const { return reinterpret_cast<const Tw32*>((uintptr_t)head_s); }
    template <class C>
    HB_HD const Tw32* inv_base() const {
        return reinterpret_cast<const Tw32*>((uintptr_t)head_s - (uintptr_t)C::inv_off(0) * sizeof(Tw32));
    }
    HB_HD Tw32 ld_head(const Tw32* p) const {
#if defined(__CUDA_ARCH__)
        Tw32 t;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t.w), "=r"(t.wp) : "r"((uint32_t)(uintptr_t)p));
        return t;
#else
        return *p;
#endif
    }
};
HB_HD bool small_modulus_ok(uint64_t q) { return q < ((uint64_t)1 << 30); }
HB_HD Small32 make_small32(uint64_t q, const InvScale& sc) {
    Small32 m;
    m.q = (uint32_t)q;
    m.nq = 0u - (uint32_t)q;
    m.twoq = (uint32_t)(q << 1);
    m.inv_n = (uint32_t)sc.inv_n;
    m.inv_n_p = (uint32_t)(sc.inv_n_p >> 32);      // floor(x*2^32/q) = hi32(floor(x*2^64/q))
    m.inv_n_w = (uint32_t)sc.inv_n_w;
    m.inv_n_w_p = (uint32_t)(sc.inv_n_w_p >> 32);
    return m;
}

// forward fast path is valid when every input word < 4q and 4q*(LOGN+1) < 2^64;
// inverse when every input word < 2q and 8q < 2^64.
HB_HD bool fwd_fast_modulus_ok(uint64_t q, int logn) {
    return q < (~(uint64_t)0) / (uint64_t)(4 * (logn + 2));
}
HB_HD bool inv_fast_modulus_ok(uint64_t q) { return q < ((uint64_t)1 << 60); }

// tail twiddles are stored [row/32][slot][row%32]: entry of (row, slot) is
// tail_tw_base(row) + 32*slot, so the 32 lanes of a warp (32 consecutive rows)
// read one contiguous run per slot
template <class C>
HB_HD uint32_t tail_tw_base(uint32_t row) {
    return ((row >> 5) << (5 + C::LOGROW)) | (row & 31u);
}

// ---------------------------------------------------------------------------
// register-resident groups
// ---------------------------------------------------------------------------

// R forward stages on 2^R registers.  Stage d (global stage s = S0+d) pairs
// k with k + 2^(R-1-d) inside blocks of 2^(R-d); its twiddle
// roots[2^s + (idx >> (LOGN-s))] (ntt.cpp:494-500) is packed at slot 2^d + blk.
// fin(k0, k1): called right after the group's LAST stage has made registers k0, k1 final,
// so that a pass can put the pair back (or send it out) while the next pairs are computed
// instead of in one burst of stores at the end
struct NoFin {
    HB_HD void operator()(int, int) const {}
    HB_HD void operator()(int, int, int) const {}
};
// S0: global index of the group's first stage (arithmetic policies whose butterfly depends on the stage)
template <int R, int TS, int S0, class A, class Fin = NoFin>
HB_HD void fwd_group(typename A::elem* v, const typename A::Tw* g, const A& a, const Fin& fin = Fin()) {
    static_for<0, R>([&](auto dc) {
        constexpr int d = decltype(dc)::value;
        constexpr int half = 1 << (R - 1 - d);
        static_for<0, (1 << d)>([&](auto bc) {
            constexpr int blk = decltype(bc)::value;
            const typename A::Tw t = (TS == 1) ? a.ld_head(g + ((1 << d) + blk) * TS) : a.ld(g + ((1 << d) + blk) * TS);
            static_for<0, half>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                a.template fwd_at<S0 + d>(v[blk * 2 * half + j], v[blk * 2 * half + j + half], t);
                if constexpr (d == R - 1) fin(blk * 2 * half + j, blk * 2 * half + j + half);
            });
        });
    });
}

// R inverse stages on 2^R registers.  Stage d (global stage u = U0+d, t = 2^u)
// pairs k with k + 2^d inside blocks of 2^(d+1); its twiddle
// inv_roots[1 + N - (N >> u) + (idx >> (u+1))] (ntt.cpp:600-636) is packed at
// slot 2^(R-1-d) + blk.  When LAST, the final stage is the inv_n-fused one.
template <int R, bool LAST, int TS, int E0, class A, class Fin = NoFin>
HB_HD void inv_group(typename A::elem* v, const typename A::Tw* g, const A& a, const Fin& fin = Fin()) {
    static_for<0, R>([&](auto dc) {
        constexpr int d = decltype(dc)::value;
        constexpr int half = 1 << d;
        static_for<0, (1 << (R - d - 1))>([&](auto bc) {
            constexpr int blk = decltype(bc)::value;
            if constexpr (LAST && d == R - 1) {
                static_for<0, half>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    a.template inv_last_at<E0 + d>(v[blk * 2 * half + j], v[blk * 2 * half + j + half]);
                    fin(blk * 2 * half + j, blk * 2 * half + j + half);
                });
            } else {
                const typename A::Tw t = (TS == 1) ? a.ld_head(g + ((1 << (R - 1 - d)) + blk) * TS)
                                                   : a.ld(g + ((1 << (R - 1 - d)) + blk) * TS);
                static_for<0, half>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    a.template inv_at<E0 + d>(v[blk * 2 * half + j], v[blk * 2 * half + j + half], t);
                    if constexpr (d == R - 1) fin(blk * 2 * half + j, blk * 2 * half + j + half);
                });
            }
        });
    });
}

// ---------------------------------------------------------------------------
// 16-byte accessors
// ---------------------------------------------------------------------------
HB_HD void ld2(const uint64_t* p, uint64_t& a, uint64_t& b) {
#if defined(__CUDA_ARCH__)
    ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p);
    a = t.x;
    b = t.y;
#else
    a = p[0];
    b = p[1];
#endif
}
HB_HD void st2(uint64_t* p, uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b);
#else
    p[0] = a;
    p[1] = b;
#endif
}
// one 16-byte chunk = 2 uint64 or 4 uint32 words, to / from registers
HB_HD void ld_chunk(const uint64_t* p, uint64_t* v) { ld2(p, v[0], v[1]); }
HB_HD void st_chunk(uint64_t* p, const uint64_t* v) { st2(p, v[0], v[1]); }
HB_HD void ld_chunk(const uint32_t* p, uint32_t* v) {
#if defined(__CUDA_ARCH__)
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
#else
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
#endif
}
HB_HD void st_chunk(uint32_t* p, const uint32_t* v) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]);
#else
    p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; p[3] = v[3];
#endif
}

// ---------------------------------------------------------------------------
// head-pass geometry shared by forward and inverse
// ---------------------------------------------------------------------------
// A head pass with R stages and stride 2^LS: group g (0 <= g < N >> R) holds
// idx = (hi << (LS+R)) + (k << LS) + lo,  lo = g mod 2^LS, hi = g >> LS.
// Thread tid handles groups tid + gi*NT, gi < E >> R.
template <class C, int R, int LS>
struct HeadGeom {
    static constexpr int G = C::E >> R;
    HB_HD static uint32_t hi(uint32_t g) { return (LS + R == C::LOGN) ? 0u : (g >> LS); }
    HB_HD static uint32_t base(uint32_t g) {
        return (hi(g) << (LS + R)) + (g & ((1u << LS) - 1));
    }
};

// Swizzled position of word k of a head-pass group with base index b and stride 2^LS.  The swizzle only looks
// at the row bits just above the 128-byte row (bits 4..6 of a uint64 index, 5..7 of a uint32 one): a stride that
// is a multiple of eight rows leaves them alone, so the group's words sit at swz(b) + k * 2^LS -- one address
// register and immediate offsets instead of an XOR and a shift per word (the compiler does not see through
// the XOR / add mix by itself and kept 32 live address registers through the first forward pass).
template <class S, int LS>
HB_HD uint32_t head_pos(uint32_t b, uint32_t k) {
    if constexpr (LS >= (sizeof(S) == 8 ? 7 : 8)) return swz_t<S>(b) + (k << LS);
    else return swz_t<S>(b + (k << LS));
}

// load all E words of a head pass from a (swizzled) shared buffer of S-typed
// words into T-typed registers through xf
template <class C, int R, int LS, class S, class T, class Xf>
HB_HD void head_load(uint32_t tid, const S* sm, T* v, const Xf& xf) {
    using Gm = HeadGeom<C, R, LS>;
    static_for<0, Gm::G>([&](auto gc) {
        constexpr int gi = decltype(gc)::value;
        const uint32_t b = Gm::base(tid + gi * C::NT);
        static_for<0, (1 << R)>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            v[gi * (1 << R) + k] = xf(sm[head_pos<S, LS>(b, (uint32_t)k)]);
        });
    });
}
template <class C, int R, int LS, class T>
HB_HD void head_store(uint32_t tid, T* sm, const T* v) {
    using Gm = HeadGeom<C, R, LS>;
    static_for<0, Gm::G>([&](auto gc) {
        constexpr int gi = decltype(gc)::value;
        const uint32_t b = Gm::base(tid + gi * C::NT);
        static_for<0, (1 << R)>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            sm[head_pos<T, LS>(b, (uint32_t)k)] = v[gi * (1 << R) + k];
        });
    });
}

// one word of a head pass back to its place (gi, k as in head_store)
template <class C, int R, int LS, class T>
HB_HD void head_store_word(uint32_t tid, T* sm, int gi, int k, T x) {
    using Gm = HeadGeom<C, R, LS>;
    sm[head_pos<T, LS>(Gm::base(tid + (uint32_t)gi * C::NT), (uint32_t)k)] = x;
}

// tail rows: thread tid owns rows tid + ri*NT of ROW contiguous words (128 bytes); with
// C::WARPTAIL warp w owns rows 64w + 32*ri + lane (a warp's 32 rows are consecutive either way)
template <class C>
HB_HD uint32_t tail_row(uint32_t tid, int ri) {
    if constexpr (C::WARPTAIL) return ((tid >> 5) << 6) + ((uint32_t)ri << 5) + (tid & 31u);
    else return tid + (uint32_t)ri * C::NT;
}
template <class C, class Xf>
HB_HD void tail_load(uint32_t tid, const typename C::elem* sm, typename C::elem* v, const Xf& xf) {
    constexpr int PER = 16 / sizeof(typename C::elem);   // words per 16-byte chunk
    static_for<0, C::E / C::ROW>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        const uint32_t row = tail_row<C>(tid, ri);
        static_for<0, 8>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            typename C::elem t[PER];
            ld_chunk(sm + row * C::ROW + (((uint32_t)c ^ (row & 7u)) * PER), t);
            static_for<0, PER>([&](auto ec) {
                constexpr int e = decltype(ec)::value;
                v[ri * C::ROW + c * PER + e] = xf(t[e]);
            });
        });
    });
}
template <class C>
HB_HD void tail_store(uint32_t tid, typename C::elem* sm, const typename C::elem* v) {
    constexpr int PER = 16 / sizeof(typename C::elem);
    static_for<0, C::E / C::ROW>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        const uint32_t row = tail_row<C>(tid, ri);
        static_for<0, 8>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            st_chunk(sm + row * C::ROW + (((uint32_t)c ^ (row & 7u)) * PER), v + ri * C::ROW + c * PER);
        });
    });
}

// ---------------------------------------------------------------------------
// forward transform pieces
// ---------------------------------------------------------------------------
template <class C, int P>
struct FwdPass {
    static constexpr int R = C::pass_r(P), S0 = C::pass_s0(P), LS = C::LOGN - S0 - R;
    static_assert(LS >= C::LOGROW, "head pass stride must cover a 128-byte row");
};

// butterflies of head pass P on registers v[E] (loaded by head_load)
template <class C, int P, class A, class Fin = NoFin>
HB_HD void fwd_head_compute(uint32_t tid, typename A::elem* v, const typename A::Tw* tw, const A& a,
                            const Fin& fin = Fin()) {
    using Ps = FwdPass<C, P>;
    using Gm = HeadGeom<C, Ps::R, Ps::LS>;
    static_for<0, Gm::G>([&](auto gc) {
        constexpr int gi = decltype(gc)::value;
        const uint32_t hi = Gm::hi(tid + gi * C::NT);
        fwd_group<Ps::R, 1, Ps::S0>(v + gi * (1 << Ps::R), tw + C::fwd_off(P) + (hi << Ps::R), a,
                            [&](int k0, int k1) { fin(gi, k0, k1); });
    });
}

// in-place head pass P > 0
template <class C, int P, class A>
HB_HD void fwd_head_pass(uint32_t tid, typename A::elem* sm, const typename A::Tw* tw, const A& a) {
    using Ps = FwdPass<C, P>;
    using T = typename A::elem;
    T v[C::E];
    auto ident = [](T x) { return x; };
    head_load<C, Ps::R, Ps::LS>(tid, sm, v, ident);
    // in place within the group: each pair goes back as soon as its last butterfly is done
    fwd_head_compute<C, P>(tid, v, tw, a, [&](int gi, int k0, int k1) {
        head_store_word<C, Ps::R, Ps::LS>(tid, sm, gi, k0, v[gi * (1 << Ps::R) + k0]);
        head_store_word<C, Ps::R, Ps::LS>(tid, sm, gi, k1, v[gi * (1 << Ps::R) + k1]);
    });
}

// tail: last LOGROW stages + final reduction on registers v[E] (from tail_load)
// after_row(ri) runs as soon as row ri is final: the kernels start that row's
// (asynchronous) stores there, so they drain while the next row is computed
template <class C, class A, class F>
HB_HD void fwd_tail_compute(uint32_t tid, typename A::elem* v, const typename A::Tw* tw, const A& a,
                            const F& after_row) {
    static_for<0, C::E / C::ROW>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        const uint32_t row = tail_row<C>(tid, ri);
        fwd_group<C::LOGROW, 32, C::HEAD>(v + ri * C::ROW, tw + C::fwd_off(C::NP) + tail_tw_base<C>(row), a);
        static_for<0, C::ROW>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            v[ri * C::ROW + k] = a.fwd_final(v[ri * C::ROW + k]);
        });
        after_row(ri);
    });
}
template <class C, class A>
HB_HD void fwd_tail_compute(uint32_t tid, typename A::elem* v, const typename A::Tw* tw, const A& a) {
    fwd_tail_compute<C>(tid, v, tw, a, [](int) {});
}

// ---------------------------------------------------------------------------
// inverse transform pieces
// ---------------------------------------------------------------------------
template <class C, int P>
struct InvPass {
    static constexpr int R = C::pass_r(P), U0 = C::inv_u0(P), LS = U0;
    static constexpr bool LAST = (P == C::NP - 1);
    static_assert(!LAST || (LS + R == C::LOGN), "pass schedule broken");
};

// first inverse pass: stages t = 1 .. ROW/2 on rows (registers from tail_load)
template <class C, class A>
HB_HD void inv_tail_compute(uint32_t tid, typename A::elem* v, const typename A::Tw* tw, const A& a) {
    static_for<0, C::E / C::ROW>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        const uint32_t row = tail_row<C>(tid, ri);
        // E0: the lazy policy's bound exponent (words enter below 2q); 0 for the others = "first stage of the transform"
        inv_group<C::LOGROW, false, 32, (A::kLazyInv ? 1 : 0)>(v + ri * C::ROW, tw + tail_tw_base<C>(row), a);
    });
    if constexpr (A::kLazyInv) {
        static_assert(InvLazy<C>::valid(), "lazy inverse bound schedule broken for this shape");
        if constexpr (InvLazy<C>::reduce_after(-1))
            static_for<0, C::E>([&](auto ec) { v[decltype(ec)::value] = a.reduce(v[decltype(ec)::value]); });
    }
}

template <class C, int P, class A, class Fin = NoFin>
HB_HD void inv_head_compute(uint32_t tid, typename A::elem* v, const typename A::Tw* tw, const A& a,
                            const Fin& fin = Fin()) {
    using Ps = InvPass<C, P>;
    using Gm = HeadGeom<C, Ps::R, Ps::LS>;
    static_for<0, Gm::G>([&](auto gc) {
        constexpr int gi = decltype(gc)::value;
        const uint32_t hi = Gm::hi(tid + gi * C::NT);
        inv_group<Ps::R, Ps::LAST, 1, (A::kLazyInv ? InvLazy<C>::e_in(P) : 1)>(
            v + gi * (1 << Ps::R), tw + C::inv_off(P) + (hi << Ps::R), a,
            [&](int k0, int k1) { fin(gi, k0, k1); });
    });
    if constexpr (A::kLazyInv) {
        if constexpr (InvLazy<C>::reduce_after(P))
            static_for<0, C::E>([&](auto ec) { v[decltype(ec)::value] = a.reduce(v[decltype(ec)::value]); });
    }
}

// in-place inverse head pass (not the last one)
template <class C, int P, class A>
HB_HD void inv_head_pass(uint32_t tid, typename A::elem* sm, const typename A::Tw* tw, const A& a) {
    using Ps = InvPass<C, P>;
    using T = typename A::elem;
    static_assert(!Ps::LAST, "the last pass goes to global memory");
    T v[C::E];
    auto ident = [](T x) { return x; };
    head_load<C, Ps::R, Ps::LS>(tid, sm, v, ident);
    if constexpr (A::kLazyInv) {      // the words may still need the mid-transform reduction
        inv_head_compute<C, P>(tid, v, tw, a);
        head_store<C, Ps::R, Ps::LS>(tid, sm, v);
    } else {
        inv_head_compute<C, P>(tid, v, tw, a, [&](int gi, int k0, int k1) {
            head_store_word<C, Ps::R, Ps::LS>(tid, sm, gi, k0, v[gi * (1 << Ps::R) + k0]);
            head_store_word<C, Ps::R, Ps::LS>(tid, sm, gi, k1, v[gi * (1 << Ps::R) + k1]);
        });
    }
}

// natural-order global index of register slot (gi,k) in the last inverse pass
template <class C>
HB_HD uint32_t inv_last_index(uint32_t tid, int gi, int k) {
    using Ps = InvPass<C, C::NP - 1>;
    using Gm = HeadGeom<C, Ps::R, Ps::LS>;
    return Gm::base(tid + gi * C::NT) + ((uint32_t)k << Ps::LS);
}

}  // namespace hb
