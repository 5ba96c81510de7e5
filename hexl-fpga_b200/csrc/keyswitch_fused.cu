// keyswitch_fused.cu -- RNS keyswitch with stages S2 + S3 + S4 of keyswitch_kernels.cu in ONE kernel:
// the NTT'd digits (the D*D polynomials "V" of the staged version, 12.8 MB per item at N = 16384, D = 7)
// never exist in memory, and the multiply-accumulate with the keys rides in the epilogue of the transform
// that produced its operand -- the dataflow of the reference's on-chip pipeline
// (device/keyswitch/ntt1.hpp -> dyadmult.hpp:128-158 -> intt2_core.hpp -> intt2_redu.hpp:24-51), where the
// limbs go from engine to engine through pipes and never touch DDR.
//
// Work unit = (item b, output modulus r): one persistent CTA walks over the D digits of the item,
//     digit j != r :  NTT_{q_r}(U[b][j])  in shared memory (FP64-pipe butterflies, ntt_block.cuh);
//                     every output word x, still in the registers of the tail pass, is multiplied by
//                     key[j][0][r] and key[j][1][r] (Shoup products on the integer pipe, which the
//                     butterflies leave idle) and added to the two accumulator polynomials
//     digit j == r :  NTT(INTT(t_j)) = t_j: the words of t_target go straight into the products
// The two accumulators are 2 x 128 KiB -- more than the shared memory left next to the 128 KiB transform
// buffer, and more registers than a thread has.  They live in TENSOR MEMORY: the SM's 256 KiB of TMEM
// (512 columns x 128 lanes x 32 bit), allocated whole by the CTA and used as a plain scratchpad through
// tcgen05.st / tcgen05.ld -- every thread owns 128 columns of its lane (32 words x 2 components x 64 bit),
// which is exactly the accumulator state of its two tail rows.  Sums of up to 15 products below 4q stay
// below 2^64 and are reduced once, after the last digit:
//     r <  D       : ACC[b][c][r] leaves through staged TMA stores (for stage S5),
//     r == D (q_k) : the two sums go back into the transform buffer, through INTT_{q_k} and the rounding
//                    v = (x + floor(q_k / 2)) mod q_k  (stage S4), and out to ACC[b][c][D].
#include "launch.h"
#include "tmem.cuh"

namespace hb {

// {key mod q, Shoup factor} of both key components for one coefficient: one 32-byte load
struct alignas(32) KeyQuad {
    uint64_t k0, k0p, k1, k1p;
};

// fused-kernel key layout: [digit j][slot r][row / 32][word k][row % 32] -- the 32 lanes of a warp (32
// consecutive tail rows) read 1 KiB contiguous per word
HB_HD size_t key_quad_index(uint32_t R, uint32_t rows, uint32_t j, uint32_t r, uint32_t row, uint32_t k) {
    return ((((size_t)j * R + r) * (rows >> 5) + (row >> 5)) * 16 + k) * 32 + (row & 31u);
}

__global__ void k_ks_prepare_keys_fused(KsDev ks, KeyQuad* __restrict__ out) {
    const uint32_t N = 1u << ks.logn, rows = N / 16;
    const size_t total = (size_t)ks.D * ks.R * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const uint32_t l = (uint32_t)(e % N), r = (uint32_t)((e / N) % ks.R), j = (uint32_t)(e / N / ks.R);
        const uint32_t idx = (r == ks.D) ? ks.K - 1 : r;
        const uint64_t q = ks.tabs[idx].q;
        KeyQuad kq;
        kq.k0 = ks.keys[(((size_t)j * 2 + 0) * ks.K + idx) * N + l] % q;
        kq.k1 = ks.keys[(((size_t)j * 2 + 1) * ks.K + idx) * N + l] % q;
        kq.k0p = (uint64_t)((((unsigned __int128)kq.k0) << 64) / q);
        kq.k1p = (uint64_t)((((unsigned __int128)kq.k1) << 64) / q);
        out[key_quad_index(ks.R, rows, j, r, l >> 4, l & 15u)] = kq;
    }
}

cudaError_t launch_ks_prepare_keys_fused(const KsDev& ks, void* out, cudaStream_t st) {
    k_ks_prepare_keys_fused<<<148 * 4, 256, 0, st>>>(ks, reinterpret_cast<KeyQuad*>(out));
    return cudaGetLastError();
}
size_t ks_fused_key_bytes(const KsDev& ks) { return (size_t)ks.D * ks.R * ((size_t)1 << ks.logn) * sizeof(KeyQuad); }

HB_D void ld_quad(const KeyQuad* p, KeyQuad& k) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(k.k0), "=l"(k.k0p), "=l"(k.k1), "=l"(k.k1p)
                 : "l"(p));
}

// ---- accumulator state of one thread ---------------------------------------------------------------------
// columns of this thread: [tail row ri][component c][word k][lo, hi], starting at `col0` of lane
// 32 * (warp % 4) + lane (tcgen05.ld/st .32x32b: a warp reaches the 32 lanes of its own quarter only)
template <class C>
struct MacState {
    static constexpr uint32_t kColsPerThread = C::E * 4;          // E words x 2 components x 2 halves
    static_assert((C::NT / 32 + 3) / 4 * kColsPerThread <= 512, "accumulators do not fit tensor memory");
    uint32_t taddr;            // tensor-memory address of this thread's column 0 (lane field = the warp's quarter)
    const KeyQuad* keys;       // quads of (digit j, slot r): row block 0, word 0, lane 0
    uint64_t nq;               // 2^64 - q of the output modulus
    uint32_t first;            // first digit of the unit: the sums start here
    HB_D static uint32_t thread_taddr(uint32_t tmem_base) {
        const uint32_t warp = threadIdx.x >> 5;
        return tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * kColsPerThread;
    }
    // acc[c][row] += x (.) key[c]   for the 16 words of tail row `ri` (global row index `row`)
    HB_D void mac_row(int ri, uint32_t row, const uint64_t* x) const {
        const KeyQuad* kp = keys + ((size_t)(row >> 5) * 16) * 32 + (row & 31u);
        if (!first) tmem_wait_st();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint64_t a0[8], a1[8];
            const uint32_t c0 = taddr + (uint32_t)ri * 64u + (uint32_t)h * 16u;
            if (!first) {
                tmem_ld16(c0, a0);
                tmem_ld16(c0 + 32u, a1);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                KeyQuad kq;
                ld_quad(kp + (size_t)(h * 8 + k) * 32, kq);
                const uint64_t p0 = mul_shoup_approx(x[h * 8 + k], kq.k0, kq.k0p, nq);
                const uint64_t p1 = mul_shoup_approx(x[h * 8 + k], kq.k1, kq.k1p, nq);
                a0[k] = first ? p0 : a0[k] + p0;
                a1[k] = first ? p1 : a1[k] + p1;
            }
            tmem_st16(c0, a0);
            tmem_st16(c0 + 32u, a1);
        }
    }
    // the 16 reduced words of (row ri, component c)
    HB_D void read_row(int ri, int c, const FastMod& fm, uint64_t* out) const {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint64_t a[8];
            tmem_ld16(taddr + (uint32_t)ri * 64u + (uint32_t)c * 32u + (uint32_t)h * 16u, a);
#pragma unroll
            for (int k = 0; k < 8; ++k) out[h * 8 + k] = reduce_small_multiple(a[k], fm);
        }
    }
};

// tail-pass output functor of the fused transforms: multiply-accumulate instead of a store
template <class CC>
struct OfMac {
    MacState<CC> st;
    template <class C>
    HB_D void prefetch(uint32_t) const {}
    template <class C>
    HB_D void store(uint32_t row, const uint64_t* v) const {
        // rows of a thread: tail_row(tid, ri) = 64 * warp + 32 * ri + lane (WARPTAIL), or tid + ri * NT
        const int ri = C::WARPTAIL ? (int)((row >> 5) & 1u) : (int)(row / C::NT);
        st.mac_row(ri, row, v);
    }
};

template <class C>
struct FusedPlan {
    // shared memory: the ntt_block.cuh plan (polynomial, per-warp staging slices, barrier words) + one word
    // for the tensor-memory base address
    static constexpr uint32_t TMEM_WORD = SmemPlan<C>::CNT_WORD + 1;
    static constexpr size_t BYTES = (size_t)(TMEM_WORD + 1) * 8;
    static_assert(SmemPlan<C>::kStagedStore, "the fused kernel sends its sums out through the staging slices");
    static_assert(BYTES <= 227u * 1024u, "shared-memory plan does not fit");
};

// one CTA, units u = blockIdx.x, blockIdx.x + gridDim.x, ...;  u -> (item b = u / R, slot r = R - 1 - u % R)
template <class C>
__global__ void __launch_bounds__(C::NT, 1)
k_ks_fused(const __grid_constant__ CUtensorMap m_t, const __grid_constant__ CUtensorMap m_u,
           const __grid_constant__ CUtensorMap m_acc_store, const KsDev ks, const KeyQuad* __restrict__ keys_f,
           uint64_t* __restrict__ ACC, uint32_t n_items) {
    constexpr uint32_t ROWS = C::N / 16;
    uint64_t* W = smem_poly<C>();
    uint64_t* bar = W + SmemPlan<C>::BAR_WORD;
    const uint32_t tid = threadIdx.x;
    const uint32_t D = ks.D, R = ks.R;
    const uint32_t n_units = n_items * R;
    if (tid == 0) {
        if (smem_u32(W) & 1023u) __trap();
        mbar_init(bar, 1);
        W[SmemPlan<C>::FLAG_WORD] = 0;
        W[SmemPlan<C>::CNT_WORD] = 0;
        fence_barrier_init();
    }
    if (tid < 32) tmem_alloc_all(reinterpret_cast<uint32_t*>(W + FusedPlan<C>::TMEM_WORD));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(W + FusedPlan<C>::TMEM_WORD);
    const uint32_t taddr = MacState<C>::thread_taddr(tmem_base);

    // digit visited at step s of a unit: the unit's own digit first (no transform), then the others
    auto digit_at = [&](uint32_t r, uint32_t s) -> uint32_t {
        if (r >= D) return s;
        return s == 0 ? r : (s <= r ? s - 1 : s);
    };
    // where step s of unit (b, r) reads its polynomial: t_target[b][r] for the own digit, else U[b][j]
    auto source = [&](uint32_t b, uint32_t r, uint32_t s, const CUtensorMap*& map) -> uint32_t {
        const uint32_t j = digit_at(r, s);
        map = (j == r) ? &m_t : &m_u;
        return (b * D + j) * ROWS;
    };
    uint32_t u = blockIdx.x;
    if (u < n_units && tid == 0) {
        const CUtensorMap* map;
        const uint32_t row = source(u / R, R - 1 - u % R, 0, map);
        issue_poly_load<C>(W, map, bar, row);
    }
    uint32_t parity = 0;
    const bool leader = C::WARPTAIL ? (tid & 31u) == 0 : tid == 0;
    for (; u < n_units; u += gridDim.x) {
        const uint32_t b = u / R, r = R - 1 - u % R;
        const uint32_t idx = (r == D) ? ks.K - 1 : r;
        const ModTab& t = ks.tabs[idx];
        const uint32_t un = u + gridDim.x;
        for (uint32_t s = 0; s < D; ++s) {
            const uint32_t j = digit_at(r, s);
            // what lands in the buffer once this step has pulled its words into registers
            Prefetch pf;
            pf.map = &m_u;
            pf.row = kNoPrefetch;
            if (s + 1 < D) {
                const uint32_t row = source(b, r, s + 1, pf.map);
                if (leader) pf.row = row;
            } else if (r != D && un < n_units) {
                const uint32_t row = source(un / R, R - 1 - un % R, 0, pf.map);
                if (leader) pf.row = row;
            }
            MacState<C> ms;
            ms.taddr = taddr;
            ms.keys = keys_f + key_quad_index(R, ROWS, j, r, 0, 0);
            ms.nq = t.fm.nq;
            ms.first = (s == 0) ? 1u : 0u;
            mbar_wait(bar, parity);
            parity ^= 1;
            if (j == r) {
                uint64_t v[C::E];
                tail_load<C>(tid, W, v, XfIdent());
                if constexpr (C::WARPTAIL) {
                    pf.template issue_when_all_warps_done<C>();
                } else {
                    __syncthreads();
                    pf.template issue<C>();
                }
#pragma unroll
                for (int ri = 0; ri < C::E / 16; ++ri) ms.mac_row(ri, tail_row<C>(tid, ri), v + ri * 16);
            } else {
                const Fp64Arith a = {t.fd};
                ntt_fwd_cta<C, kFastTrust>(W, t, a, XfReduce{t.q, t.mu, ks.s2_no_reduce}, OfMac<C>{ms}, pf);
            }
        }
        // ---- the sums are complete ----
        MacState<C> ms;
        ms.taddr = taddr;
        ms.keys = nullptr;
        ms.nq = 0;
        ms.first = 0;
        tmem_wait_st();
        if (r != D) {
            // ACC[b][c][r], same bit-reversed order as the transform output
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const OfRows of = {nullptr, &m_acc_store, ((b * 2 + (uint32_t)c) * R + r) * ROWS};
#pragma unroll
                for (int ri = 0; ri < C::E / 16; ++ri) {
                    uint64_t o[16];
                    ms.read_row(ri, c, t.fm, o);
                    of.template store<C>(tail_row<C>(tid, ri), o);
                }
            }
        } else {
            // special prime: INTT + rounding of both sums (stage S4)
            const uint64_t qk = t.q;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                if (c == 1) __syncthreads();   // every warp has pulled the previous polynomial out of the buffer
                uint64_t o[C::E];
#pragma unroll
                for (int ri = 0; ri < C::E / 16; ++ri) ms.read_row(ri, c, t.fm, o + ri * 16);
                tail_store<C>(tid, W, o);
                Prefetch pf;
                pf.map = &m_u;
                pf.row = kNoPrefetch;
                if (c == 1 && un < n_units) {
                    const uint32_t row = source(un / R, R - 1 - un % R, 0, pf.map);
                    if (leader) pf.row = row;
                }
                const Fp64Arith a = {t.fd};
                const OfWordsRound of = {ACC + (size_t)((b * 2 + (uint32_t)c) * R + D) * C::N, qk, qk >> 1};
                ntt_inv_cta<C, kFastTrust>(W, t, a, XfIdent(), of, pf);
            }
        }
    }
    if (SmemPlan<C>::kStagedStore && (tid & 31u) == 0) tma_store_wait_read();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) tmem_dealloc_all(tmem_base);
}

// S2 + S3 + S4 of a chunk in one launch; false when the shape has to take the staged kernels
bool ks_fused_available(const KsDev& ks, const void* keys_f) {
    return keys_f && ks.logn == 14 && ks.fast_ok && ks.fp64_ok && ks.D <= 15;
}

cudaError_t launch_ks_fused(const KsDev& ks, const void* keys_f, const uint64_t* t_target, const uint64_t* U,
                            uint64_t* ACC, uint64_t items, cudaStream_t st) {
    using C = NttCfg<14, 5, 4, 1>;
    if (!ks_fused_available(ks, keys_f)) return cudaErrorInvalidValue;
    CUtensorMap m_t, m_u, m_as;
    cudaError_t e;
    if ((e = make_poly_tmap(&m_t, t_target, items * ks.D, C::LOGN))) return e;
    if ((e = make_poly_tmap(&m_u, U, items * ks.D, C::LOGN))) return e;
    if ((e = make_poly_tmap(&m_as, ACC, items * 2 * ks.R, C::LOGN, 32))) return e;
    auto kern = k_ks_fused<C>;
    const size_t smem = FusedPlan<C>::BYTES;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
    const uint64_t units = items * ks.R;
    kern<<<persistent_grid((const void*)kern, C::NT, smem, units), C::NT, smem, st>>>(
        m_t, m_u, m_as, ks, reinterpret_cast<const KeyQuad*>(keys_f), ACC, (uint32_t)items);
    return cudaGetLastError();
}

}  // namespace hb
