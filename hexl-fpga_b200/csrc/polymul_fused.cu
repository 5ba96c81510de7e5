// polymul_fused.cu -- negacyclic polynomial multiply  a * b mod (x^N + 1, q)  in ONE launch (SURVEY.md 8f
// row 4; the step either side of the standalone primitives, reference README.md:47).
//
// Per item a persistent CTA runs  NTT(a) -> NTT(b) -> pointwise product -> INTT  without any intermediate
// leaving the SM: HBM sees a, b and the result, 3 x 128 KiB at N = 16384, the algorithmic minimum.
//   1. a arrives by TMA, is transformed in shared memory (FP64-pipe butterflies); the output words, still
//      in the registers of the tail pass, are PARKED IN TENSOR MEMORY as centred doubles (tmem.cuh: 64
//      columns per thread = its 32 words) while b's TMA load, started when a had left the buffer, lands;
//   2. b is transformed the same way; in its tail pass every output word meets its partner from tensor
//      memory (the tail rows of a thread are the same for both transforms), the product is one FP64
//      modular multiplication, and the row goes back into the transform buffer;
//   3. the inverse transform starts from those rows (its first pass works on exactly the rows a thread
//      owns) and stores the result in natural order, while the next item's a is prefetched.
// Inputs are caller data: both forward transforms carry the range vote of the plain kernels, and an item
// with an out-of-contract word is left to the exact three-kernel path (deferred list, as everywhere).
// Shapes: N = 16384 and moduli 2^36 <= q <= 2^51 (1 + 1/32) (the FP64 butterflies that correct every other
// stage); everything else takes the three-launch version (capi.cu).
#include "ntt_launch.cuh"
#include "tmem.cuh"

namespace hb {

int g_polymul_fused = 1;   // option "polymul_fused"

// (Fp64ArithRaw: ntt_core.cuh)
// inverse transform whose input rows already hold centred doubles with |v| <= q/2 (1 + 2^-20)
struct Fp64ArithPre : Fp64Arith {
    HB_HD uint64_t enter_inv(uint64_t x) const { return x; }
};

// Tail output of the two forward transforms, ONE type so that the transform body exists once in the kernel
// (three unrolled transform bodies, ~210 KB of SASS, overflow the SM's instruction cache: ncu showed
// "no instruction" as the top stall, 3.7 warps per issue, profiles/r2_ncu_pm_summary.txt):
//   mul == 0 (NTT(a)): park the row in tensor memory (columns [ri][word][lo, hi] of the thread's lane)
//   mul == 1 (NTT(b)): multiply by the parked row and put the product row back into the transform buffer
template <class CC>
struct OfParkMul {
    uint32_t taddr;
    Fp64Mod m;
    double inv_q;
    uint64_t* W;
    uint32_t mul;
    template <class C>
    HB_D void prefetch(uint32_t) const {}
    template <class C>
    HB_D void store(uint32_t row, const uint64_t* v) const {
        const int ri = C::WARPTAIL ? (int)((row >> 5) & 1u) : (int)(row / C::NT);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t ta = taddr + (uint32_t)ri * 32u + (uint32_t)h * 16u;
            uint64_t w[8];
            if (!mul) {
#pragma unroll
                for (int k = 0; k < 8; ++k) w[k] = d2u(fp_cred_full(u2d(v[h * 8 + k]), m));   // |w| <= q/2: a valid "twiddle"
                tmem_st16(ta, w);
            } else {
                tmem_ld16(ta, w);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    // y * w (mod q): |y| <= 1.92 q < 2^52, |w| <= q/2; the quotient factor w/q is formed on the
                    // fly (two roundings instead of one: |c - y w / q| <= 1/2 + 0.5, so |r| <= q, still exact)
                    uint64_t p[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double wd = u2d(w[2 * c + e]);
                        const double r = fp_mulmod(u2d(v[h * 8 + 2 * c + e]), wd, fp_mul(wd, inv_q), m);
                        p[e] = d2u(fp_cred(r, m));
                    }
                    st_chunk(W + row * 16 + (((uint32_t)(h * 4 + c) ^ (row & 7u)) << 1), p);
                }
            }
        }
    }
};

template <class C>
struct PolymulPlan {
    static constexpr uint32_t TMEM_WORD = SmemPlan<C>::TMEM_WORD;
    static constexpr size_t BYTES = SmemPlan<C>::BYTES;
    static_assert(BYTES <= 227u * 1024u, "shared-memory plan does not fit");
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
k_polymul_fused(const __grid_constant__ CUtensorMap m_a, const __grid_constant__ CUtensorMap m_b, uint64_t* __restrict__ res,
                const ModTab t, uint32_t n_items, uint32_t* __restrict__ list) {
    static_assert(C::WARPTAIL, "a thread's tail rows must be its own in both directions");
    constexpr uint32_t ROWS = C::N / 16;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    uint64_t* W = smem_poly<C>();
    uint64_t* bar = W + SmemPlan<C>::BAR_WORD;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        if (smem_u32(W) & 1023u) __trap();
        mbar_init(bar, 1);
        W[SmemPlan<C>::FLAG_WORD] = 0;
        W[SmemPlan<C>::CNT_WORD] = 0;
        fence_barrier_init();
    }
    if (tid < 32) tmem_alloc_all(reinterpret_cast<uint32_t*>(W + PolymulPlan<C>::TMEM_WORD));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(W + PolymulPlan<C>::TMEM_WORD);
    // 64 columns per thread; the four warps that share a lane quarter take consecutive column ranges
    const uint32_t taddr = tmem_thread_addr(tmem_base, (tid >> 7) * 64u);
    const double inv_q = 1.0 / t.fd.q;
    const bool leader = (tid & 31u) == 0;

    uint32_t i = blockIdx.x;
    if (tid == 0 && i < n_items) issue_poly_load<C>(W, &m_a, bar, i * ROWS);
    uint32_t parity = 0;
    for (; i < n_items; i += gridDim.x) {
        const uint32_t next = i + gridDim.x;
        // the buffer must take the next item's a when this item is abandoned half way
        auto abandon = [&]() {
            __syncthreads();
            if (tid == 0) {
                defer_item(list, i);
                if (next < n_items) {
                    fence_proxy_async();
                    issue_poly_load<C>(W, &m_a, bar, next * ROWS);
                }
            }
        };
        Fp64ArithRaw a;
        a.m = t.fd;
        // ---- phase 0: NTT(a), parked in tensor memory, b lands behind it;  phase 1: NTT(b) (.) NTT(a) back
        //      into the buffer.  One call site: the forward transform body exists once. ----
        bool ok = true;
#pragma unroll 1
        for (uint32_t phase = 0; phase < 2 && ok; ++phase) {
            Prefetch pf;
            pf.map = &m_b;
            pf.row = (phase == 0 && leader) ? i * ROWS : kNoPrefetch;
            mbar_wait(bar, parity);
            parity ^= 1;
            tmem_wait_st();   // tensor-memory stores of the previous phase / item have drained
            ok = ntt_fwd_cta<C, kFastVote>(W, t, a, XfIdent(), OfParkMul<C>{taddr, t.fd, inv_q, W, phase}, pf);
            if (!ok && phase == 0) {
                // a is out of contract; the deferral path has started b's load (thread 0): let it land
                mbar_wait(bar, parity);
                parity ^= 1;
            }
        }
        if (!ok) {
            abandon();
            continue;
        }
        // ---- INTT of the product rows; the next item's a is prefetched once the buffer is in registers ----
        Fp64ArithPre ai;
        ai.m = t.fd;
        Prefetch pf;
        pf.map = &m_a;
        pf.row = (leader && next < n_items) ? next * ROWS : kNoPrefetch;
        ntt_inv_cta<C, kFastTrust>(W, t, ai, XfIdent(), OfWords{res + (size_t)i * C::N}, pf);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) tmem_dealloc_all(tmem_base);
}

bool polymul_fused_available(const ModTab& tab, uint32_t logn) {
    return g_polymul_fused && logn == 14 && tab.fp64_ok && tab.fp64_alt_ok && tab.fwd_fast_ok && tab.inv_fast_ok;
}

// res[i] <- a[i] * b[i]; items with out-of-contract words are appended to `list` (word 0 = count, zero on
// entry) and left untouched
cudaError_t launch_polymul_fused(uint64_t* res, const uint64_t* a, const uint64_t* b, const ModTab& tab, uint64_t batch,
                                 uint32_t* list, cudaStream_t st) {
    using C = NttCfg<14, 5, 4, 1>;
    CUtensorMap m_a, m_b;
    cudaError_t e;
    if ((e = make_poly_tmap(&m_a, a, batch, C::LOGN))) return e;
    if ((e = make_poly_tmap(&m_b, b, batch, C::LOGN))) return e;
    auto kern = k_polymul_fused<C>;
    const size_t smem = PolymulPlan<C>::BYTES;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
    return launch_dep(kern, (unsigned)persistent_grid((const void*)kern, C::NT, smem, batch), C::NT, smem, st, m_a, m_b, res,
                      tab, (uint32_t)batch, list);
}

// the exact three-kernel path over the items of the deferred list (exits at once when the list is empty)
cudaError_t launch_polymul_deferred(uint64_t* res, uint64_t* tb, const uint64_t* a, const uint64_t* b, const ModTab& tab,
                                    uint64_t batch, uint32_t* list, cudaStream_t st) {
    using C = NttCfg<14, 5>;
    CUtensorMap m_a, m_b, s_res, s_tb, m_res;
    cudaError_t e;
    if ((e = make_poly_tmap(&m_a, a, batch, C::LOGN))) return e;
    if ((e = make_poly_tmap(&m_b, b, batch, C::LOGN))) return e;
    if ((e = make_poly_tmap(&s_res, res, batch, C::LOGN, 32))) return e;
    if ((e = make_poly_tmap(&s_tb, tb, batch, C::LOGN, 32))) return e;
    if ((e = make_poly_tmap(&m_res, res, batch, C::LOGN))) return e;
    if ((e = launch_mode<C, true, kExactList>(m_a, s_res, res, tab, batch, list, st))) return e;
    if ((e = launch_mode<C, true, kExactList>(m_b, s_tb, tb, tab, batch, list, st))) return e;
    JobInvMul<C> job;
    job.data = res;
    job.tab = tab;
    job.other = tb;
    job.dv = make_divisor(tab.q);
    job.n_items = (uint32_t)batch;
    auto kern = k_ntt_inv_mul_list<C>;
    const size_t smem = ntt_smem_bytes<C>();
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
    kern<<<persistent_grid((const void*)kern, C::NT, smem, batch), C::NT, smem, st>>>(m_res, job, (uint32_t)batch, list);
    return cudaGetLastError();
}

}  // namespace hb
