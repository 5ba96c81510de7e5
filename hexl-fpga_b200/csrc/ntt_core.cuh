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

template <int LOGN_, int LOGE_>
struct NttCfg {
    static constexpr int LOGN = LOGN_, LOGE = LOGE_;
    static constexpr int N = 1 << LOGN, E = 1 << LOGE, NT = N / E;
    // CTAs per SM the register allocation must allow (launch bounds)
    static constexpr int MIN_CTAS = (NT * 128 <= 32768) ? (65536 / (NT * 128) > 4 ? 4 : 65536 / (NT * 128)) : 1;
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
    // ---- packed twiddle table geometry (entries of 16 bytes) ----
    // forward: head pass P owns 2^S0 groups of 2^R entries, then N/16 tail rows of 16
    static constexpr int fwd_off(int p) {  // p == NP -> tail
        int o = 0;
        for (int i = 0; i < p; ++i) o += 1 << (pass_s0(i) + pass_r(i));
        return o;
    }
    static constexpr int FWD_ENTRIES = fwd_off(NP) + N;
    // inverse: N/16 tail rows of 16 first, then head pass P: (N >> (U0+R)) groups of 2^R
    static constexpr int inv_u0(int p) { return 4 + pass_s0(p); }
    static constexpr int inv_off(int p) {
        int o = N;
        for (int i = 0; i < p; ++i) o += N >> inv_u0(i);
        return o;
    }
    static constexpr int INV_ENTRIES = inv_off(NP);
};

HB_HD uint32_t swz(uint32_t idx) { return idx ^ (((idx >> 4) & 7u) << 1); }

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
        const int r = (p == C::NP) ? 4 : C::pass_r(p);
        const uint32_t count = 1u << (s0 + r);
        if (e < count) {
            uint32_t hi = e >> r, slot = e & ((1u << r) - 1);
            if (p == C::NP) {  // tail: [row/32][slot][row%32] so a warp's loads coalesce
                slot = (e >> 5) & 15u;
                hi = ((e >> 9) << 5) | (e & 31u);
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
        const int r = (p < 0) ? 4 : C::pass_r(p);
        const uint32_t count = N >> u0;  // groups (N >> (u0+r)) * 2^r
        if (e < count) {
            uint32_t hi = e >> r, slot = e & ((1u << r) - 1);
            if (p < 0) {  // tail: [row/32][slot][row%32]
                slot = (e >> 5) & 15u;
                hi = ((e >> 9) << 5) | (e & 31u);
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
// arithmetic policies
// ---------------------------------------------------------------------------
struct InvScale {
    uint64_t inv_n, inv_n_p, inv_n_w, inv_n_w_p;
};

// reference op sequence, bit-exact even on out-of-range words
struct ExactArith {
    uint64_t q, twoq;
    HB_HD void fwd(uint64_t& X, uint64_t& Y, const TwPair& t) const { fwd_bfly(X, Y, t.w, t.wp, q, twoq); }
    HB_HD uint64_t fwd_final(uint64_t x) const {      // ntt.cpp:535-546
        x -= (x >= twoq) ? twoq : 0;
        x -= (x >= q) ? q : 0;
        return x;
    }
    HB_HD void inv(uint64_t& X, uint64_t& Y, const TwPair& t) const { inv_bfly(X, Y, t.w, t.wp, q, twoq); }
    HB_HD void inv_last(uint64_t& X, uint64_t& Y, const InvScale& s) const {
        inv_last_bfly(X, Y, s.inv_n, s.inv_n_p, s.inv_n_w, s.inv_n_w_p, q, twoq);
    }
};

// any-correct-algorithm path for in-contract inputs (modarith.cuh)
struct FastArith {
    FastMod m;
    HB_HD void fwd(uint64_t& X, uint64_t& Y, const TwPair& t) const { fwd_bfly_fast(X, Y, t.w, t.wp, m); }
    HB_HD uint64_t fwd_final(uint64_t x) const { return reduce_small_multiple(x, m); }
    HB_HD void inv(uint64_t& X, uint64_t& Y, const TwPair& t) const { inv_bfly_fast(X, Y, t.w, t.wp, m); }
    HB_HD void inv_last(uint64_t& X, uint64_t& Y, const InvScale& s) const {
        inv_last_bfly_fast(X, Y, s.inv_n, s.inv_n_p, s.inv_n_w, s.inv_n_w_p, m);
    }
};

// forward fast path is valid when every input word < 4q and 4q*(LOGN+1) < 2^64;
// inverse when every input word < 2q and 8q < 2^64.
HB_HD bool fwd_fast_modulus_ok(uint64_t q, int logn) {
    return q < (~(uint64_t)0) / (uint64_t)(4 * (logn + 2));
}
HB_HD bool inv_fast_modulus_ok(uint64_t q) { return q < ((uint64_t)1 << 60); }

// tail twiddles are stored [row/32][slot][row%32]: entry of (row, slot) is
// tail_tw_base(row) + 32*slot, so the 32 lanes of a warp (32 consecutive rows)
// read 512 contiguous bytes per slot
HB_HD uint32_t tail_tw_base(uint32_t row) { return ((row >> 5) << 9) | (row & 31u); }

// ---------------------------------------------------------------------------
// register-resident groups
// ---------------------------------------------------------------------------

// R forward stages on 2^R registers.  Stage d (global stage s = S0+d) pairs
// k with k + 2^(R-1-d) inside blocks of 2^(R-d); its twiddle
// roots[2^s + (idx >> (LOGN-s))] (ntt.cpp:494-500) is packed at slot 2^d + blk.
template <int R, int TS, class A>
HB_HD void fwd_group(uint64_t* v, const TwPair* g, const A& a) {
    static_for<0, R>([&](auto dc) {
        constexpr int d = decltype(dc)::value;
        constexpr int half = 1 << (R - 1 - d);
        if constexpr (TS > 1 && d + 1 < R) {   // tail: pull the next stage's twiddles into L1
            static_for<0, (2 << d)>([&](auto pc) { prefetch_pair(g + ((2 << d) + decltype(pc)::value) * TS); });
        }
        static_for<0, (1 << d)>([&](auto bc) {
            constexpr int blk = decltype(bc)::value;
            const TwPair t = ldpair(g + ((1 << d) + blk) * TS);
            static_for<0, half>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                a.fwd(v[blk * 2 * half + j], v[blk * 2 * half + j + half], t);
            });
        });
    });
}

// R inverse stages on 2^R registers.  Stage d (global stage u = U0+d, t = 2^u)
// pairs k with k + 2^d inside blocks of 2^(d+1); its twiddle
// inv_roots[1 + N - (N >> u) + (idx >> (u+1))] (ntt.cpp:600-636) is packed at
// slot 2^(R-1-d) + blk.  When LAST, the final stage is the inv_n-fused one.
template <int R, bool LAST, int TS, class A>
HB_HD void inv_group(uint64_t* v, const TwPair* g, const A& a, const InvScale& sc) {
    static_for<0, R>([&](auto dc) {
        constexpr int d = decltype(dc)::value;
        constexpr int half = 1 << d;
        if constexpr (TS > 1 && d + 1 < R) {   // tail: next stage has 2^(R-d-2) twiddles at slots 2^(R-d-2)..
            static_for<0, (1 << (R - d - 2))>(
                [&](auto pc) { prefetch_pair(g + ((1 << (R - d - 2)) + decltype(pc)::value) * TS); });
        }
        static_for<0, (1 << (R - d - 1))>([&](auto bc) {
            constexpr int blk = decltype(bc)::value;
            if constexpr (LAST && d == R - 1) {
                static_for<0, half>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    a.inv_last(v[blk * 2 * half + j], v[blk * 2 * half + j + half], sc);
                });
            } else {
                const TwPair t = ldpair(g + ((1 << (R - 1 - d)) + blk) * TS);
                static_for<0, half>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    a.inv(v[blk * 2 * half + j], v[blk * 2 * half + j + half], t);
                });
            }
        });
    });
}

// ---------------------------------------------------------------------------
// 16-byte accessors (two words)
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

// load all E words of a head pass from the (swizzled) shared buffer
template <class C, int R, int LS, class Xf>
HB_HD void head_load(uint32_t tid, const uint64_t* sm, uint64_t* v, const Xf& xf) {
    using Gm = HeadGeom<C, R, LS>;
    static_for<0, Gm::G>([&](auto gc) {
        constexpr int gi = decltype(gc)::value;
        const uint32_t b = Gm::base(tid + gi * C::NT);
        static_for<0, (1 << R)>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            v[gi * (1 << R) + k] = xf(sm[swz(b + ((uint32_t)k << LS))]);
        });
    });
}
template <class C, int R, int LS>
HB_HD void head_store(uint32_t tid, uint64_t* sm, const uint64_t* v) {
    using Gm = HeadGeom<C, R, LS>;
    static_for<0, Gm::G>([&](auto gc) {
        constexpr int gi = decltype(gc)::value;
        const uint32_t b = Gm::base(tid + gi * C::NT);
        static_for<0, (1 << R)>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            sm[swz(b + ((uint32_t)k << LS))] = v[gi * (1 << R) + k];
        });
    });
}

// tail rows: thread tid owns rows tid + ri*NT, 16 contiguous words each
template <class C, class Xf>
HB_HD void tail_load(uint32_t tid, const uint64_t* sm, uint64_t* v, const Xf& xf) {
    static_for<0, C::E / 16>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        const uint32_t row = tid + ri * C::NT;
        static_for<0, 8>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            uint64_t a, b;
            ld2(sm + row * 16 + (((uint32_t)c ^ (row & 7u)) << 1), a, b);
            v[ri * 16 + 2 * c] = xf(a);
            v[ri * 16 + 2 * c + 1] = xf(b);
        });
    });
}
template <class C>
HB_HD void tail_store(uint32_t tid, uint64_t* sm, const uint64_t* v) {
    static_for<0, C::E / 16>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        const uint32_t row = tid + ri * C::NT;
        static_for<0, 8>([&](auto cc) {
            constexpr int c = decltype(cc)::value;
            st2(sm + row * 16 + (((uint32_t)c ^ (row & 7u)) << 1), v[ri * 16 + 2 * c], v[ri * 16 + 2 * c + 1]);
        });
    });
}

// ---------------------------------------------------------------------------
// forward transform pieces
// ---------------------------------------------------------------------------
template <class C, int P>
struct FwdPass {
    static constexpr int R = C::pass_r(P), S0 = C::pass_s0(P), LS = C::LOGN - S0 - R;
    static_assert(LS >= 4, "head pass stride must cover a 128-byte row");
};

// butterflies of head pass P on registers v[E] (loaded by head_load)
template <class C, int P, class A>
HB_HD void fwd_head_compute(uint32_t tid, uint64_t* v, const TwPair* tw, const A& a) {
    using Ps = FwdPass<C, P>;
    using Gm = HeadGeom<C, Ps::R, Ps::LS>;
    static_for<0, Gm::G>([&](auto gc) {
        constexpr int gi = decltype(gc)::value;
        const uint32_t hi = Gm::hi(tid + gi * C::NT);
        fwd_group<Ps::R, 1>(v + gi * (1 << Ps::R), tw + C::fwd_off(P) + (hi << Ps::R), a);
    });
}

// in-place head pass P > 0
template <class C, int P, class A>
HB_HD void fwd_head_pass(uint32_t tid, uint64_t* sm, const TwPair* tw, const A& a) {
    using Ps = FwdPass<C, P>;
    uint64_t v[C::E];
    auto ident = [](uint64_t x) { return x; };
    head_load<C, Ps::R, Ps::LS>(tid, sm, v, ident);
    fwd_head_compute<C, P>(tid, v, tw, a);
    head_store<C, Ps::R, Ps::LS>(tid, sm, v);
}

// tail: last four stages + final reduction on registers v[E] (from tail_load)
template <class C, class A>
HB_HD void fwd_tail_compute(uint32_t tid, uint64_t* v, const TwPair* tw, const A& a) {
    static_for<0, C::E / 16>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        const uint32_t row = tid + ri * C::NT;
        fwd_group<4, 32>(v + ri * 16, tw + C::fwd_off(C::NP) + tail_tw_base(row), a);
        static_for<0, 16>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            v[ri * 16 + k] = a.fwd_final(v[ri * 16 + k]);
        });
    });
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

// first inverse pass: stages t = 1,2,4,8 on rows (registers from tail_load)
template <class C, class A>
HB_HD void inv_tail_compute(uint32_t tid, uint64_t* v, const TwPair* tw, const A& a) {
    const InvScale none = {0, 0, 0, 0};
    static_for<0, C::E / 16>([&](auto rc) {
        constexpr int ri = decltype(rc)::value;
        const uint32_t row = tid + ri * C::NT;
        inv_group<4, false, 32>(v + ri * 16, tw + tail_tw_base(row), a, none);
    });
}

template <class C, int P, class A>
HB_HD void inv_head_compute(uint32_t tid, uint64_t* v, const TwPair* tw, const A& a, const InvScale& sc) {
    using Ps = InvPass<C, P>;
    using Gm = HeadGeom<C, Ps::R, Ps::LS>;
    static_for<0, Gm::G>([&](auto gc) {
        constexpr int gi = decltype(gc)::value;
        const uint32_t hi = Gm::hi(tid + gi * C::NT);
        inv_group<Ps::R, Ps::LAST, 1>(v + gi * (1 << Ps::R), tw + C::inv_off(P) + (hi << Ps::R), a, sc);
    });
}

// in-place inverse head pass (not the last one)
template <class C, int P, class A>
HB_HD void inv_head_pass(uint32_t tid, uint64_t* sm, const TwPair* tw, const A& a) {
    using Ps = InvPass<C, P>;
    static_assert(!Ps::LAST, "the last pass goes to global memory");
    const InvScale none = {0, 0, 0, 0};
    uint64_t v[C::E];
    auto ident = [](uint64_t x) { return x; };
    head_load<C, Ps::R, Ps::LS>(tid, sm, v, ident);
    inv_head_compute<C, P>(tid, v, tw, a, none);
    head_store<C, Ps::R, Ps::LS>(tid, sm, v);
}

// natural-order global index of register slot (gi,k) in the last inverse pass
template <class C>
HB_HD uint32_t inv_last_index(uint32_t tid, int gi, int k) {
    using Ps = InvPass<C, C::NP - 1>;
    using Gm = HeadGeom<C, Ps::R, Ps::LS>;
    return Gm::base(tid + gi * C::NT) + ((uint32_t)k << Ps::LS);
}

}  // namespace hb
