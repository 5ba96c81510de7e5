// keyswitch_kernels.cu -- RNS keyswitch (sm_100a), staged version.
//
// Implements SURVEY.md Appendix A.4 == the reference's device pipeline
// (device/keyswitch/load.hpp -> intt_core.hpp -> intt1_redu.hpp -> ntt_core.hpp
// -> dyadmult.hpp -> intt2_redu.hpp -> ntt2.hpp -> ms.hpp) plus the host-side
// accumulate into `result` (host/src/fpga.cpp:441-475), for a chunk of items:
//
//   S1  U[b][j]      = INTT_{q_j}(t[b][j])                          items*D polys
//   S2  V[b][y]      = NTT_{q_idx(r)}(U[b][j] mod q_idx(r)), j != r items*D*D polys
//   S3  ACC[b][c][r] = sum_j op(b,r,j) (.) key[j][c][idx(r)]        elementwise
//                      (op = V entry, or t[b][j] itself when j == r:
//                       NTT(INTT(t_j)) == t_j)
//   S4  ACC[b][c][D] = INTT_{q_k}(ACC[b][c][D])                     items*2 polys
//   S5  w = NTT_{q_i}(round/convert(ACC[b][c][D]));
//       result[b][c][i] += (ACC[b][c][i] - w) * msf_i  (mod q_i)    items*2*D polys
//
// idx(r) = r for r < D and K-1 (the special prime) for r == D.  y enumerates
// the D*D (r, j) pairs with j != r: y = r*(D-1) + (j < r ? j : j-1) for r < D,
// y = D*(D-1) + j for r == D.  Every stage output is canonical in [0,q), so the
// result is the unique value the reference pipeline produces.  The four
// transform stages are the persistent TMA-fed kernels of ntt_block.cuh.
#include <type_traits>

#include "launch.h"

namespace hb {

// Stages S2 + S3 + S4 in one kernel with the sums in tensor memory (keyswitch_fused.cu), option "ks_fused".
// Bit-exact and without the V scratch, but MEASURED SLOWER than the staged kernels (72k vs 93k KeySwitch/s at
// N = 16384, D = 7: every transform has to pull its own 512 KiB of key quads from L2, where the staged
// multiply-accumulate shares one key load among four items), so it is off by default.
int g_ks_fused = 0;
// Items per S2 + S3 round (option "ks_sub_items", 0 = the whole chunk): small rounds keep the NTT'd digits
// (V, D*D polynomials per item) inside the 126 MB L2 between the transform that writes them and the
// multiply-accumulate that reads them, so they never travel to HBM.
int g_ks_sub_items = 0;
// Multiply-accumulate on the FP64 pipe (k_ks_mac_fp64) fed by S2 transforms that leave the raw doubles of their
// last stage in V (option "ks_mac_fp64", default on; needs every modulus <= 2^51 (1 + 1/32) and N = 16384)
int g_ks_mac_fp64 = 1;
// Stage S5 with its load transform and its modswitch / accumulate epilogue on the FP64 pipe, `result` read and
// written through the warp's staging slice by TMA (option "ks_s5_fp64", default on; same conditions as
// ks_mac_fp64 plus all moduli within 25 % of each other)
int g_ks_s5_fp64 = 1;
// S1 hands U over as non-negative doubles when S2 takes them as they are (the default FP64 path with same-size
// moduli): option "ks_u_fp64", default on
int g_ks_u_fp64 = 1;
int g_ks_mac_items = 4;   // MAC stage: 4 (default) or 8 items per key load; 1 = register-resident keys, 2 = 128-bit accumulators (both slower: latency bound)

HB_HD uint32_t ks_y(uint32_t D, uint32_t r, uint32_t j) {
    return r < D ? r * (D - 1) + (j < r ? j : j - 1) : D * (D - 1) + j;
}

// ---- S1 -------------------------------------------------------------------
// DCONV: the kernel produces canonical integers but U is to hold doubles (the exact pass behind an S1 that hands
// over doubles)
template <class C, bool DCONV = false>
struct JobIntt1 {
    static constexpr bool kOneModulus = false;
    // N = 16384, FP64 kernels: the grid walks over the polynomials modulus-major, so that a CTA keeps one modulus
    // for items / gridDim.x transforms in a row and its twiddles stay in shared / tensor memory (ntt_persistent)
    static constexpr bool kModulusRuns = C::LOGN == 14;
    KsDev ks;
    uint64_t* U;
    uint32_t B = 0;          // items of the chunk (0: walk in storage order)
    HB_D uint32_t order(uint32_t i) const {
        if (!B) return i;
        const uint32_t m = fdiv(i, B, ks.fB);
        return (i - m * B) * ks.D + m;
    }
    HB_D uint32_t src_row(uint32_t item) const { return item * (C::N / 16); }
    HB_D const ModTab& mod(uint32_t item) const { return ks.tabs[item - fdiv(item, ks.fD) * ks.D]; }
    HB_D XfIdent xf(uint32_t) const { return XfIdent(); }
    HB_D auto of(uint32_t item, const CUtensorMap*) const {
        if constexpr (DCONV) return OfWordsD{U + (size_t)item * C::N};
        else return OfWords{U + (size_t)item * C::N};
    }
};
template <class C, int MODE, int FP64 = 0, bool DCONV = false>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ks_intt1(const __grid_constant__ CUtensorMap tmap,
        const __grid_constant__ CUtensorMap smap, const JobIntt1<C, DCONV> job, uint32_t n_items, uint32_t* list) {
    ntt_persistent<C, false, MODE, JobIntt1<C, DCONV>, false, FP64>(&tmap, &smap, job, n_items, list);
}

// ---- S2 -------------------------------------------------------------------
// NR: every digit is already inside the forward contract under every target modulus (KsDev::s2_no_reduce, decided
// per plan): the load transform is the identity at COMPILE time.  As a run-time flag the kernel carried both
// versions of the 32-word load and every warp jumped over the unused one once per transform -- 2.7 % of the
// stage's samples were instruction-cache misses behind that jump.
// UFP (with NR): U holds doubles (S1 with FP64 = 4), no entry conversion either.
template <class C, bool NR = false, bool UFP = false>
struct JobNtt1 {
    static constexpr bool kOneModulus = false;
    static constexpr bool kModulusRuns = C::LOGN == 14;
    KsDev ks;
    uint64_t* V;
    uint32_t item0 = 0;     // first (item, pair) index of this launch (rounds of a few items)
    uint32_t B = 0;         // items of the chunk (0: walk in storage order)
    // modulus-major walk: output modulus r < D owns B * (D - 1) polynomials, the special prime B * D
    HB_D uint32_t order(uint32_t i) const {
        if (!B) return i;
        const uint32_t D = ks.D, lo = B * (D - 1);
        uint32_t b, y;
        if (i < D * lo) {
            const uint32_t r = fdiv(i, lo, ks.fBlo), rem = i - r * lo;
            b = fdiv(rem, ks.fDm1);
            y = r * (D - 1) + (rem - b * (D - 1));
        } else {
            const uint32_t rem = i - D * lo;
            b = fdiv(rem, ks.fD);
            y = D * (D - 1) + (rem - b * D);
        }
        return b * D * D + y;
    }
    HB_D void decode(uint32_t item, uint32_t& b, uint32_t& r, uint32_t& j) const {
        const uint32_t D = ks.D, per = D * D;
        item += item0;
        b = fdiv(item, ks.fDD);
        const uint32_t y = item - b * per;
        if (y < D * (D - 1)) {
            r = fdiv(y, ks.fDm1);
            const uint32_t jj = y - r * (D - 1);
            j = jj + (jj >= r ? 1 : 0);
        } else {
            r = D;
            j = y - D * (D - 1);
        }
    }
    HB_D uint32_t idx_of(uint32_t item) const {
        uint32_t b, r, j;
        decode(item, b, r, j);
        return r == ks.D ? ks.K - 1 : r;
    }
    HB_D uint32_t src_row(uint32_t item) const {
        uint32_t b, r, j;
        decode(item, b, r, j);
        return (b * ks.D + j) * (C::N / 16);
    }
    HB_D const ModTab& mod(uint32_t item) const { return ks.tabs[idx_of(item)]; }
    HB_D auto xf(uint32_t item) const {
        if constexpr (NR && UFP) {
            return XfIdentFp();
        } else if constexpr (NR) {
            return XfIdent();
        } else {
            const ModTab& t = ks.tabs[idx_of(item)];
            return XfReduce{t.q, t.mu, ks.s2_no_reduce};
        }
    }
    HB_D OfRows of(uint32_t item, const CUtensorMap* smap) const {
        item += item0;
        return OfRows{V + (size_t)item * C::N, smap, item * (C::N / 16)};
    }
};
template <class C, int MODE, int FP64 = 0, bool NR = false, bool UFP = false>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ks_ntt1(const __grid_constant__ CUtensorMap tmap,
        const __grid_constant__ CUtensorMap smap, const JobNtt1<C, NR, UFP> job, uint32_t n_items, uint32_t* list) {
    ntt_persistent<C, true, MODE, JobNtt1<C, NR, UFP>, false, FP64>(&tmap, &smap, job, n_items, list);
}

// ---- S3 -------------------------------------------------------------------
// grid: (N / (256*2), R, items); each thread owns two adjacent coefficients.
// Generic version (any modulus): exact 2-by-1 division per product.
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
                                      : V + ((size_t)b * ks.D * ks.D + ks_y(ks.D, r, j)) * N;
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

// Fast version (all moduli < 2^58): the keys carry their Shoup factors
// (keys_sh[j][c][i][l] = {key mod q_i, floor(key * 2^64 / q_i)}), every product is
// the 9-IMAD approximate Shoup product in [0,4q), the D products are summed
// without reduction (< 4q*D <= 60q) and reduced once.  Each thread serves
// kMacItems items with one load of the key words: the keys (29 MB per key set
// at D=7, K=8) would otherwise be re-read from L2 for every item.

// 32-byte load of two adjacent {key, Shoup factor} pairs, kept in L2 with
// priority (the key set is re-read by every group of items while V streams by)
HB_D void ld_keys2(const TwPair* p, TwPair& a, TwPair& b) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(a.w), "=l"(a.wp), "=l"(b.w), "=l"(b.wp)
                 : "l"(p));
}

template <int kMacItems>
__global__ void __launch_bounds__(256, kMacItems > 4 ? 2 : 3)
k_ks_mac_fast(KsDev ks, const uint64_t* __restrict__ t_target, const uint64_t* __restrict__ V,
              uint64_t* __restrict__ ACC, uint32_t items) {
    const uint32_t N = 1u << ks.logn;
    const uint32_t l = (blockIdx.x * 256 + threadIdx.x) * 2;
    const uint32_t r = blockIdx.y, b0 = blockIdx.z * kMacItems;
    const uint32_t idx = (r == ks.D) ? ks.K - 1 : r;
    const FastMod fm = ks.tabs[idx].fm;
    uint64_t a0[kMacItems][2], a1[kMacItems][2];
#pragma unroll
    for (int it = 0; it < kMacItems; ++it) a0[it][0] = a0[it][1] = a1[it][0] = a1[it][1] = 0;
    for (uint32_t j = 0; j < ks.D; ++j) {
        const TwPair* k0 = ks.keys_sh + (((size_t)j * 2 + 0) * ks.K + idx) * N + l;
        const TwPair* k1 = ks.keys_sh + (((size_t)j * 2 + 1) * ks.K + idx) * N + l;
        // every load of the digit is issued before the first product waits on one: the items past
        // the end of the batch re-read the last item (their sums are never stored), which keeps the
        // loads unconditional so that they can all be in flight together
        uint64_t x0[kMacItems], x1[kMacItems];
#pragma unroll
        for (int it = 0; it < kMacItems; ++it) {
            const uint32_t b = min(b0 + it, items - 1);
            const uint64_t* op = (j == r) ? t_target + ((size_t)b * ks.D + j) * N
                                          : V + ((size_t)b * ks.D * ks.D + ks_y(ks.D, r, j)) * N;
            ld2(op + l, x0[it], x1[it]);
        }
        TwPair u0, u1, w0, w1;
        ld_keys2(k0, u0, u1);
        ld_keys2(k1, w0, w1);
#pragma unroll
        for (int it = 0; it < kMacItems; ++it) {
            a0[it][0] += mul_shoup_approx(x0[it], u0.w, u0.wp, fm.nq);
            a0[it][1] += mul_shoup_approx(x1[it], u1.w, u1.wp, fm.nq);
            a1[it][0] += mul_shoup_approx(x0[it], w0.w, w0.wp, fm.nq);
            a1[it][1] += mul_shoup_approx(x1[it], w1.w, w1.wp, fm.nq);
        }
        // every term is below 4q and reduce_small_multiple takes sums below 64q: with more than 15 digits
        // (the reference stops at 6, plan_create admits up to 63) the sums are folded every 15 terms
        if ((j & 15u) == 14u) {
#pragma unroll
            for (int it = 0; it < kMacItems; ++it) {
                a0[it][0] = reduce_small_multiple(a0[it][0], fm);
                a0[it][1] = reduce_small_multiple(a0[it][1], fm);
                a1[it][0] = reduce_small_multiple(a1[it][0], fm);
                a1[it][1] = reduce_small_multiple(a1[it][1], fm);
            }
        }
    }
#pragma unroll
    for (int it = 0; it < kMacItems; ++it) {
        const uint32_t b = b0 + it;
        if (b < items) {
            st2(ACC + (((size_t)b * 2 + 0) * ks.R + r) * N + l, reduce_small_multiple(a0[it][0], fm),
                reduce_small_multiple(a0[it][1], fm));
            st2(ACC + (((size_t)b * 2 + 1) * ks.R + r) * N + l, reduce_small_multiple(a1[it][0], fm),
                reduce_small_multiple(a1[it][1], fm));
        }
    }
}

// FP64-pipe version (every modulus <= 2^51 (1 + 1/32), the moduli of the every-other-stage forward butterflies).
// The integer version above is bound by the integer multiplier (76 % busy, 18 IMAD per key product); here a key
// product is the six-instruction FP64 modular product of the butterflies, its twiddle the key itself:
// keys_fp[j][c][i][l] = {centred key mod q_i, fl(key / q_i)}.  And the stage in front does not have to produce
// canonical words any more: V holds the raw doubles of S2's last butterfly stage (|v| <= 1.92 q < 2^52, k_ks_ntt1
// with FP64 = 3: no exit conversion, 12 of a transform's ~170 scheduler cycles per word), read here as they are.
// The digit under its own modulus comes from the caller's t_target as integers (converted on the fly; a word
// above 2^52 -- out of contract, the integer kernels accept it -- is reduced first).  Every term is an exact
// integer with |r| <= q (1/2 + |y| 2^-54) <= 0.75 q; the sums take a full correction every fourth term, so they
// stay below 0.5 q + 4 * 0.75 q = 3.5 q < 2^53 and every addition is exact.
template <int kMacItems>
__global__ void __launch_bounds__(256, 3)
k_ks_mac_fp64(KsDev ks, const uint64_t* __restrict__ t_target, const uint64_t* __restrict__ V,
              uint64_t* __restrict__ ACC, uint32_t items) {
    const uint32_t N = 1u << ks.logn;
    const uint32_t l = (blockIdx.x * 256 + threadIdx.x) * 2;
    const uint32_t r = blockIdx.y, b0 = blockIdx.z * kMacItems;
    const uint32_t idx = (r == ks.D) ? ks.K - 1 : r;
    const Fp64Mod m = ks.tabs[idx].fd;
    const uint64_t q = ks.tabs[idx].q, mu = ks.tabs[idx].mu;
    double a0[kMacItems][2], a1[kMacItems][2];
#pragma unroll
    for (int it = 0; it < kMacItems; ++it) a0[it][0] = a0[it][1] = a1[it][0] = a1[it][1] = 0.0;
    for (uint32_t j = 0; j < ks.D; ++j) {
        const TwPair* k0 = ks.keys_fp + (((size_t)j * 2 + 0) * ks.K + idx) * N + l;
        const TwPair* k1 = ks.keys_fp + (((size_t)j * 2 + 1) * ks.K + idx) * N + l;
        // all loads of the digit in flight before the first product waits on one (cf. k_ks_mac_fast)
        uint64_t x0[kMacItems], x1[kMacItems];
#pragma unroll
        for (int it = 0; it < kMacItems; ++it) {
            const uint32_t b = min(b0 + it, items - 1);
            const uint64_t* op = (j == r) ? t_target + ((size_t)b * ks.D + j) * N
                                          : V + ((size_t)b * ks.D * ks.D + ks_y(ks.D, r, j)) * N;
            ld2(op + l, x0[it], x1[it]);
        }
        TwPair u0, u1, w0, w1;
        ld_keys2(k0, u0, u1);
        ld_keys2(k1, w0, w1);
        if (j == r) {   // caller integers -> doubles (uniform branch: r is the block's)
#pragma unroll
            for (int it = 0; it < kMacItems; ++it) {
                if ((x0[it] | x1[it]) >> 52) {
                    x0[it] = barrett_reduce64(x0[it], q, mu);
                    x1[it] = barrett_reduce64(x1[it], q, mu);
                }
                x0[it] = d2u(fp_from_int(x0[it]));
                x1[it] = d2u(fp_from_int(x1[it]));
            }
        }
#pragma unroll
        for (int it = 0; it < kMacItems; ++it) {
            a0[it][0] = fp_add(a0[it][0], fp_mulmod(u2d(x0[it]), u2d(u0.w), u2d(u0.wp), m));
            a0[it][1] = fp_add(a0[it][1], fp_mulmod(u2d(x1[it]), u2d(u1.w), u2d(u1.wp), m));
            a1[it][0] = fp_add(a1[it][0], fp_mulmod(u2d(x0[it]), u2d(w0.w), u2d(w0.wp), m));
            a1[it][1] = fp_add(a1[it][1], fp_mulmod(u2d(x1[it]), u2d(w1.w), u2d(w1.wp), m));
        }
        if ((j & 3u) == 3u) {
#pragma unroll
            for (int it = 0; it < kMacItems; ++it) {
                a0[it][0] = fp_cred_full(a0[it][0], m);
                a0[it][1] = fp_cred_full(a0[it][1], m);
                a1[it][0] = fp_cred_full(a1[it][0], m);
                a1[it][1] = fp_cred_full(a1[it][1], m);
            }
        }
    }
#pragma unroll
    for (int it = 0; it < kMacItems; ++it) {
        const uint32_t b = b0 + it;
        if (b < items) {
            st2(ACC + (((size_t)b * 2 + 0) * ks.R + r) * N + l, fp_to_canonical_full(a0[it][0], m),
                fp_to_canonical_full(a0[it][1], m));
            st2(ACC + (((size_t)b * 2 + 1) * ks.R + r) * N + l, fp_to_canonical_full(a1[it][0], m),
                fp_to_canonical_full(a1[it][1], m));
        }
    }
}

#ifdef HB_EXPERIMENTAL_VARIANTS
// Wide-accumulator version (all moduli < 2^58, D <= 16): no per-term reduction at
// all.  Every term is a plain 64x64 product of a reduced operand and a reduced
// key, accumulated as three partial sums by weight (2^0 as 96 bits, 2^32 and
// 2^64 as 64 bits -- operand and key high words are below 2^26, so 2*D cross
// products of < 2^58 and D high products of < 2^52 cannot overflow), assembled
// into 128 bits and reduced once with the 2-by-1 division of the dyadic kernel.
// 16 multiplier cycles per term instead of the Shoup product's 28, and the keys
// are read as plain words (8 bytes, not a {key, factor} pair).  MEASURED SLOWER
// than k_ks_mac_fast<4> (58.5k vs 68.9k KeySwitch/s): 28 accumulator registers per
// output leave 116 registers per thread and half the resident warps, and stage S3
// is latency bound (51 % multiplier, 52 % HBM), not multiplier bound.  Kept as
// option ks_mac_items=2.
struct WideAcc {
    uint64_t a0, a1, a2;
    uint32_t c0;
};
HB_D void wide_mac(WideAcc& a, uint64_t x, uint64_t k) {
    const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), k0 = (uint32_t)k, k1 = (uint32_t)(k >> 32);
    const uint64_t p00 = (uint64_t)x0 * k0;
    a.a0 += p00;
    a.c0 += (a.a0 < p00) ? 1u : 0u;
    a.a1 += (uint64_t)x0 * k1;
    a.a1 += (uint64_t)x1 * k0;
    a.a2 += (uint64_t)x1 * k1;
}
HB_D uint64_t wide_reduce(const WideAcc& a, const Divisor& dv) {
    const uint64_t lo = a.a0 + (a.a1 << 32);
    const uint64_t hi = (uint64_t)a.c0 + (a.a1 >> 32) + a.a2 + ((lo < a.a0) ? 1 : 0);
    // dv.s >= 6 here (q < 2^58); the sum is below D*q^2, so its shifted high word stays below d
    const uint64_t u1 = (hi << dv.s) | (lo >> (64 - dv.s));
    const uint64_t u0 = lo << dv.s;
    return rem_2by1(u1, u0, dv.d, dv.v) >> dv.s;
}
HB_D void ld_keys2_plain(const uint64_t* p, uint64_t& a, uint64_t& b) {
    // (the L2::evict_last policy of ld_keys2 exists for 256-bit loads only)
    asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}

template <int kMacItems>
__global__ void __launch_bounds__(256, 2)
k_ks_mac_wide(KsDev ks, const uint64_t* __restrict__ t_target, const uint64_t* __restrict__ V,
              uint64_t* __restrict__ ACC, uint32_t items) {
    const uint32_t N = 1u << ks.logn;
    const uint32_t l = (blockIdx.x * 256 + threadIdx.x) * 2;
    const uint32_t r = blockIdx.y, b0 = blockIdx.z * kMacItems;
    const uint32_t idx = (r == ks.D) ? ks.K - 1 : r;
    const Divisor dv = ks.divs[idx];
    WideAcc acc[kMacItems][2][2];   // [item][component][coefficient]
#pragma unroll
    for (int it = 0; it < kMacItems; ++it)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e) acc[it][c][e] = WideAcc{0, 0, 0, 0};
    for (uint32_t j = 0; j < ks.D; ++j) {
        // ks.keys holds the keys reduced mod q_i (k_ks_prepare_keys rewrites the device copy)
        uint64_t u[2], w[2];
        ld_keys2_plain(ks.keys + (((size_t)j * 2 + 0) * ks.K + idx) * N + l, u[0], u[1]);
        ld_keys2_plain(ks.keys + (((size_t)j * 2 + 1) * ks.K + idx) * N + l, w[0], w[1]);
#pragma unroll
        for (int it = 0; it < kMacItems; ++it) {
            const uint32_t b = b0 + it;
            if (b < items) {
                uint64_t x[2];
                if (j == r) {   // the digit's own modulus: caller data, reduce defensively
                    ld2(t_target + ((size_t)b * ks.D + j) * N + l, x[0], x[1]);
                    x[0] = mod64(x[0], dv);
                    x[1] = mod64(x[1], dv);
                } else {
                    ld2(V + ((size_t)b * ks.D * ks.D + ks_y(ks.D, r, j)) * N + l, x[0], x[1]);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    wide_mac(acc[it][0][e], x[e], u[e]);
                    wide_mac(acc[it][1][e], x[e], w[e]);
                }
            }
        }
    }
#pragma unroll
    for (int it = 0; it < kMacItems; ++it) {
        const uint32_t b = b0 + it;
        if (b < items) {
            st2(ACC + (((size_t)b * 2 + 0) * ks.R + r) * N + l, wide_reduce(acc[it][0][0], dv),
                wide_reduce(acc[it][0][1], dv));
            st2(ACC + (((size_t)b * 2 + 1) * ks.R + r) * N + l, wide_reduce(acc[it][1][0], dv),
                wide_reduce(acc[it][1][1], dv));
        }
    }
}

// Register-resident-key version: a thread owns ONE coefficient of one output
// modulus, keeps its 2*D {key, Shoup factor} pairs in registers and walks over
// a slice of the items, so the key set crosses L2 once per slice instead of
// once per item group (the per-item-group version above is L2-bandwidth bound:
// 29 MB of keys against 6 MB of operands).
template <int DMAX>
__global__ void __launch_bounds__(256, 2)
k_ks_mac_regkeys(KsDev ks, const uint64_t* __restrict__ t_target, const uint64_t* __restrict__ V,
                 uint64_t* __restrict__ ACC, uint32_t items, uint32_t items_per_slice) {
    const uint32_t N = 1u << ks.logn;
    const uint32_t l = blockIdx.x * 256 + threadIdx.x;
    const uint32_t r = blockIdx.y;
    const uint32_t idx = (r == ks.D) ? ks.K - 1 : r;
    const FastMod fm = ks.tabs[idx].fm;
    TwPair k0[DMAX], k1[DMAX];
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
        if (j < (int)ks.D) {
            k0[j] = ldpair(ks.keys_sh + (((size_t)j * 2 + 0) * ks.K + idx) * N + l);
            k1[j] = ldpair(ks.keys_sh + (((size_t)j * 2 + 1) * ks.K + idx) * N + l);
        }
    const uint32_t b_begin = blockIdx.z * items_per_slice;
    const uint32_t b_end = min(items, b_begin + items_per_slice);
    for (uint32_t b = b_begin; b < b_end; ++b) {
        uint64_t a0 = 0, a1 = 0;
#pragma unroll
        for (int j = 0; j < DMAX; ++j)
            if (j < (int)ks.D) {
                const uint64_t* op = ((uint32_t)j == r) ? t_target + ((size_t)b * ks.D + j) * N
                                                        : V + ((size_t)b * ks.D * ks.D + ks_y(ks.D, r, j)) * N;
                const uint64_t x = __ldg(op + l);
                a0 += mul_shoup_approx(x, k0[j].w, k0[j].wp, fm.nq);
                a1 += mul_shoup_approx(x, k1[j].w, k1[j].wp, fm.nq);
            }
        ACC[(((size_t)b * 2 + 0) * ks.R + r) * N + l] = reduce_small_multiple(a0, fm);
        ACC[(((size_t)b * 2 + 1) * ks.R + r) * N + l] = reduce_small_multiple(a1, fm);
    }
}

#endif  // HB_EXPERIMENTAL_VARIANTS

// one-time per plan: keys_sh from the raw keys
__global__ void k_ks_prepare_keys(KsDev ks, TwPair* __restrict__ out) {
    const uint32_t N = 1u << ks.logn;
    const size_t total = (size_t)ks.D * 2 * ks.K * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)((e / N) % ks.K);
        const uint64_t q = ks.tabs[i].q;
        const uint64_t k = ks.keys[e] % q;
        TwPair t;
        t.w = k;
        t.wp = (uint64_t)((((unsigned __int128)k) << 64) / q);
        out[e] = t;
        const_cast<uint64_t*>(ks.keys)[e] = k;   // the plan's own device copy: keep it reduced (k_ks_mac_wide)
    }
}

// one-time per plan: keys_fp from the raw keys (run after k_ks_prepare_keys: ks.keys is reduced by then, which
// this kernel does not rely on)
__global__ void k_ks_prepare_keys_fp64(KsDev ks, TwPair* __restrict__ out) {
    const uint32_t N = 1u << ks.logn;
    const size_t total = (size_t)ks.D * 2 * ks.K * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)((e / N) % ks.K);
        const uint64_t q = ks.tabs[i].q;
        const double kc = fp_centred(ks.keys[e] % q, q);
        TwPair t;
        t.w = d2u(kc);
        t.wp = d2u(fp_quot(kc, q));
        out[e] = t;
    }
}
cudaError_t launch_ks_prepare_keys_fp64(const KsDev& ks, TwPair* out, cudaStream_t st) {
    k_ks_prepare_keys_fp64<<<148 * 4, 256, 0, st>>>(ks, out);
    return cudaGetLastError();
}

cudaError_t launch_ks_prepare_keys(const KsDev& ks, TwPair* out, cudaStream_t st) {
    k_ks_prepare_keys<<<148 * 4, 256, 0, st>>>(ks, out);
    return cudaGetLastError();
}

// ---- S4 -------------------------------------------------------------------
template <class C>
struct JobIntt2 {
    static constexpr bool kOneModulus = true;    // the special prime
    static constexpr bool kModulusRuns = false;
    HB_D uint32_t order(uint32_t i) const { return i; }
    KsDev ks;
    uint64_t* ACC;
    HB_D uint32_t poly(uint32_t item) const { return item * ks.R + ks.D; }   // [b][c][D]
    HB_D uint32_t src_row(uint32_t item) const { return poly(item) * (C::N / 16); }
    HB_D const ModTab& mod(uint32_t) const { return ks.tabs[ks.K - 1]; }
    HB_D XfIdent xf(uint32_t) const { return XfIdent(); }
    HB_D OfWordsRound of(uint32_t item, const CUtensorMap*) const {
        const uint64_t qk = ks.tabs[ks.K - 1].q;
        return OfWordsRound{ACC + (size_t)poly(item) * C::N, qk, qk >> 1};
    }
};
template <class C, int MODE, int FP64 = 0>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ks_intt2(const __grid_constant__ CUtensorMap tmap,
        const __grid_constant__ CUtensorMap smap, const JobIntt2<C> job, uint32_t n_items, uint32_t* list) {
    ntt_persistent<C, false, MODE, JobIntt2<C>, false, FP64>(&tmap, &smap, job, n_items, list);
}

// ---- S5 -------------------------------------------------------------------
// tail-pass output functor: modswitch + accumulate into result
// (device/keyswitch/ms.hpp:68-83, host/src/fpga.cpp:453-468)
struct OfKsFinal {
    const uint64_t* acc;   // ACC[b][c][i]
    uint64_t* result;      // result[b][c][i]
    uint64_t q, msf, msf_p;
    HB_D uint64_t finish(uint64_t a, uint64_t w, uint64_t r) const {
        const uint64_t d = sub_mod(a, w, q);
        uint64_t o = mul_lazy(d, msf, msf_p, q);
        o -= (o >= q) ? q : 0;
        return add_mod(r, o, q);
    }
    // pull this thread's share of acc / result (one 128-byte line per row and
    // array) towards L2 while the transform runs; they come from HBM
    template <class C>
    HB_D void prefetch(uint32_t tid) const {
#pragma unroll
        for (int ri = 0; ri < C::E / 16; ++ri) {
            const uint32_t off = tail_row<C>(tid, ri) * 16;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(acc + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(result + off));
        }
    }
    template <class C>
    HB_D void store(uint32_t rw, const uint64_t* v) const {
        if constexpr (SmemPlan<C>::kStagedStore) {
            // transpose through the warp's staging slice so that acc / result are
            // touched with coalesced 8-byte accesses (lanes along the words)
            const uint32_t lane = threadIdx.x & 31u;
            uint64_t* slice = smem_poly<C>() + C::N + (threadIdx.x >> 5) * 512;
            const uint32_t base = (rw - lane) * 16;   // first word of the warp's 32 rows
            // all 32 global loads of the round are issued before anything waits on them
            // (one exposed L2 latency per round instead of one per word)
            uint64_t a[16], r[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                a[i] = __ldg(acc + base + lane + 32u * i);
                r[i] = result[base + lane + 32u * i];
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c)
                st2(slice + lane * 16 + (((uint32_t)c ^ (lane & 7u)) << 1), v[2 * c], v[2 * c + 1]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t w = lane + 32u * i;
                const uint32_t row = w >> 4, ch = (w >> 1) & 7u;
                const uint64_t x = slice[row * 16 + ((ch ^ (row & 7u)) << 1) + (w & 1u)];
                result[base + w] = finish(a[i], x, r[i]);
            }
        } else {
            const uint32_t off = rw * 16;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint64_t a[2], r[2];
                ld2(acc + off + 2 * c, a[0], a[1]);
                ld2(result + off + 2 * c, r[0], r[1]);
                st2(result + off + 2 * c, finish(a[0], v[2 * c], r[0]), finish(a[1], v[2 * c + 1], r[1]));
            }
        }
    }
};
// The same epilogue for the kernel whose transform hands over raw doubles (FP64 = 3): same data movement (the
// warp's rows transposed through its staging slice, acc / result touched with coalesced 8-byte accesses), the
// arithmetic on the FP64 pipe: a - w (|.| <= 2.92 q) takes the full correction, the product with msf is the
// six-instruction one, and its canonical value meets the caller's word in the reference's own integer add_mod
// (so a `result` word outside [0, q) gives the reference's wrap-around value, as before).
// Two other data paths were built and measured slower than this one (stage S5 of 1024 items at 7/8: 2307 us
// with the integer epilogue): `result` through the slice by TMA in both directions with acc read by the
// row's owner as 16-byte pieces (2493 us), and the same with 32-byte row accesses instead of the TMA store
// (2634 us) -- 32 different lines per warp instruction cost more than the transposition saves.
// The functor carries the job and the item only: pointers and constants are re-derived where they are used,
// so nothing of it stays in registers through the butterflies.
template <class C>
struct JobNtt2Fp;
template <class CC>
struct OfKsFinalFp {
    const JobNtt2Fp<CC>* job;
    uint32_t item;
    template <class C>
    HB_D void prefetch(uint32_t tid) const {
        const uint64_t* acc = job->acc_poly(item);
        const uint64_t* res = job->res_poly(item);
#pragma unroll
        for (int ri = 0; ri < C::E / 16; ++ri) {
            const uint32_t off = tail_row<C>(tid, ri) * 16;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(acc + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(res + off));
        }
    }
    template <class C>
    HB_D void store(uint32_t rw, const uint64_t* v) const {
        static_assert(SmemPlan<C>::kStagedStore, "the epilogue transposes through the staging slices");
        const uint32_t lane = threadIdx.x & 31u;
        uint64_t* slice = smem_poly<C>() + C::N + (threadIdx.x >> 5) * 512;
        const uint32_t base = (rw - lane) * 16;   // first word of the warp's 32 rows
        // 16-byte accesses: lane l owns the word pairs (2 l, 2 l + 1) + 64 i of the warp's 512 words -- half as many
        // load / store instructions in flight as with one word per lane (the epilogue's loads stall on the
        // load / store queue and on L2 latency), and a pair is one chunk of the swizzled staging slice
        const uint64_t* acc = job->acc_poly(item) + base + 2 * lane;
        uint64_t* res = job->res_poly(item) + base + 2 * lane;
        // The sums are asked for first; the caller's words only once the row has left its registers for the slice
        // (they then take those registers: no spills), and they are not needed before every product of the round is
        // done -- their latency hides behind the arithmetic instead of in front of it.
        uint64_t a[16], r[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const ulonglong2 ta = __ldg(reinterpret_cast<const ulonglong2*>(acc + 64 * i));
            a[2 * i] = ta.x;
            a[2 * i + 1] = ta.y;
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c)
            st2(slice + lane * 16 + (((uint32_t)c ^ (lane & 7u)) << 1), v[2 * c], v[2 * c + 1]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) ld2(res + 64 * i, r[2 * i], r[2 * i + 1]);
        const uint32_t i = item - fdiv(item, job->ks.fD) * job->ks.D;
        const ModTab* t = job->ks.tabs + i;
        Fp64Mod m;                   // the fields the epilogue's arithmetic reads
        m.q = t->fd.q;
        m.nq = t->fd.nq;
        m.inv_q = t->fd.inv_q;
        m.qi = t->q;
        const uint64_t q = t->q;
        const double msf_c = job->ks.msf_fp[i], msf_q = job->ks.msf_fp[job->ks.K + i];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t w = 2u * lane + 64u * k;
            const uint32_t row = w >> 4, ch = (w >> 1) & 7u;
            uint64_t x0, x1;
            ld2(slice + row * 16 + ((ch ^ (row & 7u)) << 1), x0, x1);
            const double d0 = fp_cred_full(fp_add(fp_from_int(a[2 * k]), -u2d(x0)), m);
            const double d1 = fp_cred_full(fp_add(fp_from_int(a[2 * k + 1]), -u2d(x1)), m);
            a[2 * k] = fp_canon_signed(fp_mulmod(d0, msf_c, msf_q, m), m);
            a[2 * k + 1] = fp_canon_signed(fp_mulmod(d1, msf_c, msf_q, m), m);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            st2(res + 64 * k, add_mod(r[2 * k], a[2 * k], q), add_mod(r[2 * k + 1], a[2 * k + 1], q));
    }
};
template <class C>
struct JobNtt2Fp {
    static constexpr bool kOneModulus = false;
    static constexpr bool kModulusRuns = C::LOGN == 14;
    KsDev ks;
    const uint64_t* ACC;
    uint64_t* result;
    CUtensorMap rmap;        // result: [items * 2 * D polynomials], box 32 rows
    uint32_t B2 = 0;         // 2 * items of the chunk (0: walk in storage order)
    HB_D uint32_t order(uint32_t i) const {
        if (!B2) return i;
        const uint32_t m = fdiv(i, B2, ks.fB2);
        return (i - m * B2) * ks.D + m;
    }
    HB_D uint32_t src_row(uint32_t item) const { return ((fdiv(item, ks.fD)) * ks.R + ks.D) * (C::N / 16); }
    HB_D const ModTab& mod(uint32_t item) const { return ks.tabs[(item - fdiv(item, ks.fD) * ks.D)]; }
    // item = bc * D + i:  ACC[bc][i], result[bc][i]
    HB_D const uint64_t* acc_poly(uint32_t item) const { return ACC + ((size_t)(fdiv(item, ks.fD)) * ks.R + (item - fdiv(item, ks.fD) * ks.D)) * C::N; }
    HB_D uint64_t* res_poly(uint32_t item) const { return result + (size_t)item * C::N; }
    HB_D uint32_t res_row(uint32_t item) const { return item * (C::N / 16); }
    HB_D XfKsConvertFp xf(uint32_t item) const {
        const ModTab& t = ks.tabs[(item - fdiv(item, ks.fD) * ks.D)];
        const uint64_t h = ks.tabs[ks.K - 1].q >> 1;
        const uint64_t hr = barrett_reduce64(h, t.q, t.mu);
        return XfKsConvertFp{fp_centred(hr ? t.q - hr : 0, t.q)};      // fix = -floor(qk / 2) mod q
    }
    HB_D OfKsFinalFp<C> of(uint32_t item, const CUtensorMap*) const { return OfKsFinalFp<C>{this, item}; }
};
template <class C>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ks_ntt2f(const __grid_constant__ CUtensorMap tmap,
        const __grid_constant__ CUtensorMap smap, const __grid_constant__ JobNtt2Fp<C> job, uint32_t n_items, uint32_t* list) {
    ntt_persistent<C, true, kFastTrust, JobNtt2Fp<C>, false, 3>(&tmap, &smap, job, n_items, list);
}

template <class C>
struct JobNtt2 {
    static constexpr bool kOneModulus = false;
    static constexpr bool kModulusRuns = false;
    HB_D uint32_t order(uint32_t i) const { return i; }
    KsDev ks;
    const uint64_t* ACC;
    uint64_t* result;
    HB_D uint32_t src_row(uint32_t item) const { return ((fdiv(item, ks.fD)) * ks.R + ks.D) * (C::N / 16); }
    HB_D const ModTab& mod(uint32_t item) const { return ks.tabs[(item - fdiv(item, ks.fD) * ks.D)]; }
    HB_D XfKsConvert xf(uint32_t item) const {
        const ModTab& t = ks.tabs[(item - fdiv(item, ks.fD) * ks.D)];
        const uint64_t qk = ks.tabs[ks.K - 1].q, h = qk >> 1;
        return XfKsConvert{t.q, t.mu, t.q - barrett_reduce64(h, t.q, t.mu), (qk < 2 * t.q) ? 1u : 0u};
    }
    HB_D OfKsFinal of(uint32_t item, const CUtensorMap*) const {
        const uint32_t i = (item - fdiv(item, ks.fD) * ks.D), bc = fdiv(item, ks.fD);
        return OfKsFinal{ACC + ((size_t)bc * ks.R + i) * C::N, result + ((size_t)bc * ks.D + i) * C::N,
                         ks.tabs[i].q, ks.msf[i], ks.msf_p[i]};
    }
};
template <class C, int MODE, int FP64 = 0>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS) k_ks_ntt2(const __grid_constant__ CUtensorMap tmap,
        const __grid_constant__ CUtensorMap smap, const JobNtt2<C> job, uint32_t n_items, uint32_t* list) {
    ntt_persistent<C, true, MODE, JobNtt2<C>, false, FP64>(&tmap, &smap, job, n_items, list);
}

static bool use_fused(const KsDev& ks) { return g_ks_fused && ks_fused_available(ks, ks.keys_fused); }

size_t ks_scratch_words_per_item(const KsDev& ks) {
    const size_t n = (size_t)1 << ks.logn;
    // U + V + ACC, plus room for the deferred list of stage S1 (D entries and the count); the fused
    // kernel has no V
    const size_t v = use_fused(ks) ? 0 : (size_t)ks.D * ks.D;
    return ((size_t)ks.D + v + 2 * (size_t)ks.R) * n + ks.D + 1;
}

template <class C>
struct KsWarpTailCfg {
    using type = C;
};
template <>
struct KsWarpTailCfg<NttCfg<14, 5, 4, 0>> {
    using type = NttCfg<14, 5, 4, 1>;
};

template <class K, class J>
static cudaError_t run_persistent(K kern, int threads, size_t smem, const CUtensorMap& tmap, const CUtensorMap& smap,
                                  const J& job, uint64_t n_items, uint32_t* list, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<persistent_grid((const void*)kern, threads, smem, n_items), threads, smem, st>>>(tmap, smap, job,
                                                                                           (uint32_t)n_items, list);
    return cudaGetLastError();
}

template <class C>
static cudaError_t ks_chunk(const KsDev& ks_in, uint64_t* result, const uint64_t* t_target, uint64_t items,
                            uint64_t* scratch, cudaStream_t st, int* launches) {
    const size_t smem = ntt_smem_bytes<C>();
    const uint64_t D = ks_in.D, R = ks_in.R;
    const uint32_t B = (uint32_t)items;
    KsDev ks = ks_in;            // + the prepared divisors of this chunk's index arithmetic
    ks.fB = make_fastdiv(B);
    ks.fBlo = make_fastdiv(B * (uint32_t)(D - 1));
    ks.fB2 = make_fastdiv(2 * B);
    const bool fused = use_fused(ks);
    uint64_t* U = scratch;
    uint64_t* V = U + items * D * C::N;
    uint64_t* ACC = V + (fused ? 0 : items * D * D * C::N);
    CUtensorMap m_t, m_u, m_acc, m_vs;
    cudaError_t e;
    if ((e = make_poly_tmap(&m_t, t_target, items * D, C::LOGN))) return e;
    if ((e = make_poly_tmap(&m_u, U, items * D, C::LOGN))) return e;
    if ((e = make_poly_tmap(&m_acc, ACC, items * 2 * R, C::LOGN))) return e;
    if (!fused && (e = make_poly_tmap(&m_vs, V, items * D * D, C::LOGN, 32))) return e;   // staged stores of S2
    uint32_t* list = reinterpret_cast<uint32_t*>(ACC + items * 2 * R * C::N);
    int nl = 0;
    // FP64 stages: tail rows dealt out by warp where the shape allows it (ntt_core.cuh, NttCfg::WARPTAIL)
    using CW = typename KsWarpTailCfg<C>::type;
    const size_t smemw = ntt_smem_bytes_fp64_plain<CW>();    // FP64 kernels: room for the head-pass twiddles
    // V as raw doubles + the multiply-accumulate on the FP64 pipe (not in the rounds of ks_sub_items, whose
    // multiply-accumulate is the integer one)
    const bool mac_fp64 = !fused && g_ks_mac_fp64 && ks.fast_ok && ks.fp64_ok && ks.fp64_alt_ok && ks.keys_fp &&
                          !std::is_same<CW, C>::value && !(g_ks_sub_items > 0 && (uint64_t)g_ks_sub_items < items);
    if (fused) {
        // S1, then S2 + S3 + S4 in one kernel (the sums in tensor memory), then S5
        if ((e = cudaMemsetAsync(list, 0, 8, st))) return e;
        if ((e = run_persistent(k_ks_intt1<CW, kFastVote, true>, C::NT, smemw, m_t, m_t, JobIntt1<CW>{ks, U, B}, items * D, list, st))) return e;
        if ((e = run_persistent(k_ks_intt1<C, kExactList>, C::NT, smem, m_t, m_t, JobIntt1<C>{ks, U}, items * D, list, st))) return e;
        if ((e = launch_ks_fused(ks, ks.keys_fused, t_target, U, ACC, items, st))) return e;
        if ((e = run_persistent(k_ks_ntt2<CW, kFastTrust, true>, C::NT, smemw, m_acc, m_acc, JobNtt2<CW>{ks, ACC, result}, items * 2 * D, list, st))) return e;
        if (launches) *launches = 4;
        return cudaSuccess;
    }
    if (ks.fast_ok && ks.fp64_ok) {
        // same stages with the butterflies on the FP64 pipe: every load transform hands over words in [0, 1.25q)
        if ((e = cudaMemsetAsync(list, 0, 8, st))) return e;
        // U as doubles when its only reader is the S2 kernel that takes them as they are (option "ks_u_fp64")
        const bool u_fp = mac_fp64 && ks.s2_no_reduce && g_ks_u_fp64;
        if (u_fp) {
            if ((e = run_persistent(k_ks_intt1<CW, kFastVote, 4>, C::NT, smemw, m_t, m_t, JobIntt1<CW>{ks, U, B}, items * D, list, st))) return e;
            if ((e = run_persistent(k_ks_intt1<C, kExactList, 0, true>, C::NT, smem, m_t, m_t, JobIntt1<C, true>{ks, U}, items * D, list, st))) return e;
        } else {
            if ((e = run_persistent(k_ks_intt1<CW, kFastVote, true>, C::NT, smemw, m_t, m_t, JobIntt1<CW>{ks, U, B}, items * D, list, st))) return e;
            if ((e = run_persistent(k_ks_intt1<C, kExactList>, C::NT, smem, m_t, m_t, JobIntt1<C>{ks, U}, items * D, list, st))) return e;
        }
        nl += 2;
        if (g_ks_sub_items > 0 && ks.keys_sh && (uint64_t)g_ks_sub_items < items) {
            // S2 + S3 in rounds of a few items: V is written and read back while it is still in L2
            for (uint64_t off = 0; off < items; off += (uint64_t)g_ks_sub_items) {
                const uint64_t cnt = items - off < (uint64_t)g_ks_sub_items ? items - off : (uint64_t)g_ks_sub_items;
                JobNtt1<CW> job{ks, V};
                job.item0 = (uint32_t)(off * D * D);
                if ((e = run_persistent(k_ks_ntt1<CW, kFastTrust, true>, C::NT, smemw, m_u, m_vs, job, cnt * D * D, list, st))) return e;
                dim3 gf(C::N / 512, ks.R, (unsigned)((cnt + 3) / 4));
                k_ks_mac_fast<4><<<gf, 256, 0, st>>>(ks, t_target + off * D * C::N, V + off * D * D * C::N,
                                                     ACC + off * 2 * R * C::N, (uint32_t)cnt);
                if ((e = cudaGetLastError())) return e;
                nl += 2;
            }
            if ((e = run_persistent(k_ks_intt2<CW, kFastTrust, true>, C::NT, smemw, m_acc, m_acc, JobIntt2<CW>{ks, ACC}, items * 2, list, st))) return e;
            if ((e = run_persistent(k_ks_ntt2<CW, kFastTrust, true>, C::NT, smemw, m_acc, m_acc, JobNtt2<CW>{ks, ACC, result}, items * 2 * D, list, st))) return e;
            if (launches) *launches = nl + 2;
            return cudaSuccess;
        }
        if (mac_fp64) {
            if (u_fp) {
                if ((e = run_persistent(k_ks_ntt1<CW, kFastTrust, 3, true, true>, C::NT, smemw, m_u, m_vs, JobNtt1<CW, true, true>{ks, V, 0, B}, items * D * D, list, st))) return e;
            } else if (ks.s2_no_reduce) {
                if ((e = run_persistent(k_ks_ntt1<CW, kFastTrust, 3, true>, C::NT, smemw, m_u, m_vs, JobNtt1<CW, true>{ks, V, 0, B}, items * D * D, list, st))) return e;
            } else {
                if ((e = run_persistent(k_ks_ntt1<CW, kFastTrust, 3>, C::NT, smemw, m_u, m_vs, JobNtt1<CW>{ks, V, 0, B}, items * D * D, list, st))) return e;
            }
        } else if (ks.fp64_alt_ok && !std::is_same<CW, C>::value) {
            if ((e = run_persistent(k_ks_ntt1<CW, kFastTrust, 2>, C::NT, smemw, m_u, m_vs, JobNtt1<CW>{ks, V, 0, B}, items * D * D, list, st))) return e;
        } else {
            if ((e = run_persistent(k_ks_ntt1<CW, kFastTrust, true>, C::NT, smemw, m_u, m_vs, JobNtt1<CW>{ks, V, 0, B}, items * D * D, list, st))) return e;
        }
        nl += 1;
    } else if (ks.fast_ok) {
        // S1 sees caller data: vote + deferred exact pass; the later stages read
        // words this pipeline produced (reduced by their load transforms)
        if ((e = cudaMemsetAsync(list, 0, 8, st))) return e;
        if ((e = run_persistent(k_ks_intt1<C, kFastVote>, C::NT, smem, m_t, m_t, JobIntt1<C>{ks, U}, items * D, list, st))) return e;
        if ((e = run_persistent(k_ks_intt1<C, kExactList>, C::NT, smem, m_t, m_t, JobIntt1<C>{ks, U}, items * D, list, st))) return e;
        if ((e = run_persistent(k_ks_ntt1<C, kFastTrust>, C::NT, smem, m_u, m_vs, JobNtt1<C>{ks, V}, items * D * D, list, st))) return e;
        nl += 3;
    } else {
        if ((e = run_persistent(k_ks_intt1<C, kExactAll>, C::NT, smem, m_t, m_t, JobIntt1<C>{ks, U}, items * D, list, st))) return e;
        if ((e = run_persistent(k_ks_ntt1<C, kExactAll>, C::NT, smem, m_u, m_vs, JobNtt1<C>{ks, V}, items * D * D, list, st))) return e;
        nl += 2;
    }
    dim3 g(C::N / 512, ks.R, (unsigned)items);
#ifdef HB_EXPERIMENTAL_VARIANTS
    if (ks.fast_ok && ks.keys_sh && g_ks_mac_items == 1 && ks.D <= 8) {
        // register-resident keys: ~8 item slices keep the grid several waves deep
        const uint32_t slices = (uint32_t)(items < 8 ? items : 8);
        const uint32_t per = (uint32_t)((items + slices - 1) / slices);
        dim3 gk(C::N / 256, ks.R, (unsigned)((items + per - 1) / per));
        k_ks_mac_regkeys<8><<<gk, 256, 0, st>>>(ks, t_target, V, ACC, (uint32_t)items, per);
    } else if (ks.fast_ok && ks.keys_sh && g_ks_mac_items == 2 && ks.D <= 16) {
        dim3 gw(C::N / 512, ks.R, (unsigned)((items + 1) / 2));
        k_ks_mac_wide<2><<<gw, 256, 0, st>>>(ks, t_target, V, ACC, (uint32_t)items);
    } else if (ks.fast_ok && ks.keys_sh && g_ks_mac_items == 8) {
        dim3 gf(C::N / 512, ks.R, (unsigned)((items + 7) / 8));
        k_ks_mac_fast<8><<<gf, 256, 0, st>>>(ks, t_target, V, ACC, (uint32_t)items);
    } else
#endif
    if (mac_fp64) {
        dim3 gf(C::N / 512, ks.R, (unsigned)((items + 3) / 4));
        k_ks_mac_fp64<4><<<gf, 256, 0, st>>>(ks, t_target, V, ACC, (uint32_t)items);
    } else if (ks.fast_ok && ks.keys_sh) {
        dim3 gf(C::N / 512, ks.R, (unsigned)((items + 3) / 4));
        k_ks_mac_fast<4><<<gf, 256, 0, st>>>(ks, t_target, V, ACC, (uint32_t)items);
    } else
        k_ks_mac<<<g, 256, 0, st>>>(ks, t_target, V, ACC);
    if ((e = cudaGetLastError())) return e;
    if (ks.fast_ok && ks.fp64_ok) {
        if ((e = run_persistent(k_ks_intt2<CW, kFastTrust, true>, C::NT, smemw, m_acc, m_acc, JobIntt2<CW>{ks, ACC}, items * 2, list, st))) return e;
        if constexpr (!std::is_same<CW, C>::value) {
            if (g_ks_s5_fp64 && ks.fp64_alt_ok && ks.s2_no_reduce && ks.msf_fp) {
                JobNtt2Fp<CW> job{ks, ACC, result, {}, 2 * B};
                if ((e = make_poly_tmap(&job.rmap, result, items * 2 * D, C::LOGN, 32))) return e;
                if ((e = run_persistent(k_ks_ntt2f<CW>, C::NT, smemw, m_acc, m_acc, job, items * 2 * D, list, st))) return e;
                nl += 3;
                if (launches) *launches = nl;
                return cudaSuccess;
            }
        }
        if (ks.fp64_alt_ok && !std::is_same<CW, C>::value) {
            if ((e = run_persistent(k_ks_ntt2<CW, kFastTrust, 2>, C::NT, smemw, m_acc, m_acc, JobNtt2<CW>{ks, ACC, result}, items * 2 * D, list, st))) return e;
        } else {
            if ((e = run_persistent(k_ks_ntt2<CW, kFastTrust, true>, C::NT, smemw, m_acc, m_acc, JobNtt2<CW>{ks, ACC, result}, items * 2 * D, list, st))) return e;
        }
    } else if (ks.fast_ok) {
        if ((e = run_persistent(k_ks_intt2<C, kFastTrust>, C::NT, smem, m_acc, m_acc, JobIntt2<C>{ks, ACC}, items * 2, list, st))) return e;
        if ((e = run_persistent(k_ks_ntt2<C, kFastTrust>, C::NT, smem, m_acc, m_acc, JobNtt2<C>{ks, ACC, result}, items * 2 * D, list, st))) return e;
    } else {
        if ((e = run_persistent(k_ks_intt2<C, kExactAll>, C::NT, smem, m_acc, m_acc, JobIntt2<C>{ks, ACC}, items * 2, list, st))) return e;
        if ((e = run_persistent(k_ks_ntt2<C, kExactAll>, C::NT, smem, m_acc, m_acc, JobNtt2<C>{ks, ACC, result}, items * 2 * D, list, st))) return e;
    }
    nl += 3;
    if (launches) *launches = nl;
    return cudaSuccess;
}

cudaError_t launch_ks_chunk(const KsDev& ks, uint64_t* result, const uint64_t* t_target, uint64_t items,
                            uint64_t* scratch, cudaStream_t st, int* launches) {
    if (launches) *launches = 0;
    if (items == 0) return cudaSuccess;
    if (items > 65535) return cudaErrorInvalidValue;  // gridDim.z of the MAC stage
    switch (ks.logn) {
        case 10: return ks_chunk<NttCfg<10, 4>>(ks, result, t_target, items, scratch, st, launches);
        case 11: return ks_chunk<NttCfg<11, 4>>(ks, result, t_target, items, scratch, st, launches);
        case 12: return ks_chunk<NttCfg<12, 4>>(ks, result, t_target, items, scratch, st, launches);
        case 13: return ks_chunk<NttCfg<13, 4>>(ks, result, t_target, items, scratch, st, launches);
        case 14: return ks_chunk<NttCfg<14, 5>>(ks, result, t_target, items, scratch, st, launches);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace hb
