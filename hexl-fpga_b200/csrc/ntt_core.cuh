// ntt_core.cuh -- thread-level structure of the shared-memory negacyclic NTT.
//
// One polynomial (N = 2^LOGN uint64 words, N*8 bytes) lives in shared memory;
// a CTA of NT = N/E threads (E = 2^LOGE elements per thread) walks over the
// LOGN radix-2 stages in a few "passes".  In a pass every thread pulls E
// elements into registers, runs up to LOGE consecutive butterfly stages on
// them (register-resident radix-2^R groups), and puts them back in place, so
// a polynomial makes ceil((LOGN-4)/LOGE)+1 shared-memory round trips instead
// of LOGN.
//
//   forward (Cooley-Tukey, reference tests/test_utils/ntt.cpp:494-547):
//     head passes  stages 0 .. LOGN-5   (strides >= 16, lanes along `lo`)
//     tail pass    stages LOGN-4..LOGN-1 (16 contiguous words per thread row)
//   inverse (Gentleman-Sande, ntt.cpp:580-659): the mirror image.
//
// Shared-memory layout: 128-byte rows of 16 words; the 16-byte chunk c of row
// r is stored at chunk c ^ (r & 7).  That is exactly the TMA SWIZZLE_128B
// pattern; it makes both access shapes conflict-free: head passes touch one
// whole row per half warp (any permutation inside a row still covers all 32
// banks), the tail pass reads one 16-byte chunk of 8 different rows per
// quarter warp.
//
// All functions are host+device so tests/cpu_emul can replay them thread by
// thread on the CPU to validate the index math before GPU time is spent.
#pragma once
#include <type_traits>

#include "modarith.cuh"

namespace hb {

template <int LOGN_, int LOGE_>
struct NttCfg {
    static constexpr int LOGN = LOGN_, LOGE = LOGE_;
    static constexpr int N = 1 << LOGN, E = 1 << LOGE, NT = N / E;
    static constexpr int HEAD = LOGN - 4;                 // stages outside the 16-word tail
    static constexpr int NP = (HEAD + LOGE - 1) / LOGE;   // number of head passes
    static constexpr int BASE = HEAD / NP, REM = HEAD % NP;
    static_assert(LOGE >= 4 && LOGN >= LOGE + 4, "unsupported NTT shape");
    static constexpr int pass_r(int p) { return BASE + (p < REM ? 1 : 0); }
    static constexpr int pass_s0(int p) {
        int s = 0;
        for (int i = 0; i < p; ++i) s += pass_r(i);
        return s;
    }
};

HB_HD uint32_t swz(uint32_t idx) { return idx ^ (((idx >> 4) & 7u) << 1); }

// twiddle fetch: read-only path on the device.
HB_HD uint64_t ldtw(const uint64_t* p, uint32_t i) {
#if defined(__CUDA_ARCH__)
    return __ldg(p + i);
#else
    return p[i];
#endif
}

// ---------------------------------------------------------------------------
// register-resident groups
// ---------------------------------------------------------------------------

// R forward stages on 2^R registers.  Stage d (global stage s = S0+d) pairs
// k with k + 2^(R-1-d) inside blocks of 2^(R-d); its twiddle index is
// 2^s + (idx >> (LOGN-s)) = ((2^S0 + hi) << d) + blk      (ntt.cpp:494-500).
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

template <int R>
HB_HD void fwd_group(uint64_t (&v)[1 << R], const uint64_t* roots,
                     const uint64_t* precon, uint32_t m0_plus_hi, uint64_t q,
                     uint64_t twoq) {
    static_for<0, R>([&](auto dc) {
        constexpr int d = decltype(dc)::value;
        constexpr int half = 1 << (R - 1 - d);
        static_for<0, (1 << d)>([&](auto bc) {
            constexpr int blk = decltype(bc)::value;
            const uint32_t tw = (m0_plus_hi << d) + blk;
            const uint64_t w = ldtw(roots, tw), wp = ldtw(precon, tw);
            static_for<0, half>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                fwd_bfly(v[blk * 2 * half + j], v[blk * 2 * half + j + half], w,
                         wp, q, twoq);
            });
        });
    });
}

struct InvScale {
    uint64_t inv_n, inv_n_p, inv_n_w, inv_n_w_p;
};

// R inverse stages on 2^R registers.  Stage d (global stage u = U0+d, t = 2^u)
// pairs k with k + 2^d inside blocks of 2^(d+1); twiddle index
// 1 + N - (N >> u) + (idx >> (u+1)) = base_u + (hi << (R-d-1)) + blk
// (ntt.cpp:600-636).  When LAST, the final stage is the inv_n-fused one.
template <int LOGN, int R, int U0, bool LAST>
HB_HD void inv_group(uint64_t (&v)[1 << R], const uint64_t* inv_roots,
                     const uint64_t* precon_inv, uint32_t hi, uint64_t q,
                     uint64_t twoq, const InvScale& sc) {
    static_for<0, R>([&](auto dc) {
        constexpr int d = decltype(dc)::value;
        constexpr int half = 1 << d;
        constexpr int u = U0 + d;
        constexpr uint32_t base_u = 1u + (1u << LOGN) - ((1u << LOGN) >> u);
        static_for<0, (1 << (R - d - 1))>([&](auto bc) {
            constexpr int blk = decltype(bc)::value;
            if constexpr (LAST && d == R - 1) {
                static_for<0, half>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    inv_last_bfly(v[blk * 2 * half + j],
                                  v[blk * 2 * half + j + half], sc.inv_n,
                                  sc.inv_n_p, sc.inv_n_w, sc.inv_n_w_p, q, twoq);
                });
            } else {
                const uint32_t tw = base_u + (hi << (R - d - 1)) + blk;
                const uint64_t w = ldtw(inv_roots, tw),
                               wp = ldtw(precon_inv, tw);
                static_for<0, half>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    inv_bfly(v[blk * 2 * half + j], v[blk * 2 * half + j + half],
                             w, wp, q, twoq);
                });
            }
        });
    });
}

// ---------------------------------------------------------------------------
// 16-byte shared/global accessors (two words)
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

// ---------------------------------------------------------------------------
// forward passes
// ---------------------------------------------------------------------------

// Head pass P.  P == 0 reads the polynomial from `src` (global, coalesced along
// lo) through `xf`; later passes work in place in shared memory.
template <class C, int P, class Xf>
HB_HD void fwd_head_pass(uint32_t tid, uint64_t* sm, const uint64_t* src,
                         const Xf& xf, const uint64_t* roots,
                         const uint64_t* precon, uint64_t q, uint64_t twoq) {
    constexpr int R = C::pass_r(P), S0 = C::pass_s0(P), G = C::E >> R;
    constexpr int LS = C::LOGN - S0 - R;  // log2(stride)
    static_assert(LS >= 4, "head pass stride must cover a 128-byte row");
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        const uint32_t g = tid + gi * C::NT;
        const uint32_t lo = g & ((1u << LS) - 1);
        const uint32_t hi = (LS + R == C::LOGN) ? 0u : (g >> LS);
        const uint32_t base = (hi << (LS + R)) + lo;
        uint64_t v[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); ++k) {
            const uint32_t idx = base + ((uint32_t)k << LS);
            v[k] = (P == 0) ? xf(src[idx]) : sm[swz(idx)];
        }
        fwd_group<R>(v, roots, precon, (1u << S0) + hi, q, twoq);
#pragma unroll
        for (int k = 0; k < (1 << R); ++k)
            sm[swz(base + ((uint32_t)k << LS))] = v[k];
    }
}

// Tail pass: last four stages on rows of 16 contiguous words, final reduction
// to [0,q) (ntt.cpp:535-546) and 16-byte stores to `dst` (global).
template <class C, class Of>
HB_HD void fwd_tail_pass(uint32_t tid, const uint64_t* sm, uint64_t* dst,
                         const Of& of, const uint64_t* roots,
                         const uint64_t* precon, uint64_t q, uint64_t twoq) {
#pragma unroll
    for (int ri = 0; ri < C::E / 16; ++ri) {
        const uint32_t row = tid + ri * C::NT;
        uint64_t v[16];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            ld2(sm + row * 16 + (((uint32_t)c ^ (row & 7u)) << 1), v[2 * c],
                v[2 * c + 1]);
        fwd_group<4>(v, roots, precon, (1u << (C::LOGN - 4)) + row, q, twoq);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint64_t x = v[k];
            x -= (x >= twoq) ? twoq : 0;
            x -= (x >= q) ? q : 0;
            v[k] = x;
        }
        of(dst, row * 16, v);
    }
}

// ---------------------------------------------------------------------------
// inverse passes
// ---------------------------------------------------------------------------

// First inverse pass: stages t = 1,2,4,8 on rows of 16 contiguous words read
// straight from `src` (global, 16-byte loads), written to shared memory.
template <class C, class Xf>
HB_HD void inv_tail_pass(uint32_t tid, uint64_t* sm, const uint64_t* src,
                         const Xf& xf, const uint64_t* inv_roots,
                         const uint64_t* precon_inv, uint64_t q, uint64_t twoq) {
    InvScale none = {0, 0, 0, 0};
#pragma unroll
    for (int ri = 0; ri < C::E / 16; ++ri) {
        const uint32_t row = tid + ri * C::NT;
        uint64_t v[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            ld2(src + row * 16 + 2 * c, v[2 * c], v[2 * c + 1]);
            v[2 * c] = xf(v[2 * c]);
            v[2 * c + 1] = xf(v[2 * c + 1]);
        }
        inv_group<C::LOGN, 4, 0, false>(v, inv_roots, precon_inv, row, q, twoq,
                                        none);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            st2(sm + row * 16 + (((uint32_t)c ^ (row & 7u)) << 1), v[2 * c],
                v[2 * c + 1]);
    }
}

// Inverse head pass P (P = 0 .. NP-1) covers stages u0 .. u0+R-1 with
// u0 = 4 + sum of earlier pass sizes; the last one also applies inv_n /
// inv_n_w (ntt.cpp:640-657) and writes natural-order output to `dst`.
template <class C, int P, class Of>
HB_HD void inv_head_pass(uint32_t tid, uint64_t* sm, uint64_t* dst,
                         const Of& of, const uint64_t* inv_roots,
                         const uint64_t* precon_inv, uint64_t q, uint64_t twoq,
                         const InvScale& sc) {
    constexpr int R = C::pass_r(P), U0 = 4 + C::pass_s0(P), G = C::E >> R;
    constexpr int LS = U0;
    constexpr bool LAST = (P == C::NP - 1);
    static_assert(!LAST || (LS + R == C::LOGN), "pass schedule broken");
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        const uint32_t g = tid + gi * C::NT;
        const uint32_t lo = g & ((1u << LS) - 1);
        const uint32_t hi = LAST ? 0u : (g >> LS);
        const uint32_t base = (hi << (LS + R)) + lo;
        uint64_t v[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); ++k)
            v[k] = sm[swz(base + ((uint32_t)k << LS))];
        inv_group<C::LOGN, R, U0, LAST>(v, inv_roots, precon_inv, hi, q, twoq,
                                        sc);
#pragma unroll
        for (int k = 0; k < (1 << R); ++k) {
            const uint32_t idx = base + ((uint32_t)k << LS);
            if (LAST)
                of(dst, idx, v[k]);
            else
                sm[swz(idx)] = v[k];
        }
    }
}

}  // namespace hb
