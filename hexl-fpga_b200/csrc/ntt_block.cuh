// ntt_block.cuh -- CTA-level forward / inverse NTT of one polynomial in shared
// memory, built from the per-thread passes of ntt_core.cuh.  Device only.
#pragma once
#include <cuda_runtime.h>

#include "ntt_core.cuh"

namespace hb {

// Everything a kernel needs to transform under one modulus.  Tables follow the
// reference's caller-visible layout: roots/precon indexed m+i (bit-reversed
// powers, tests/test_utils/ntt.cpp:296-310), inv_roots/precon_inv in the
// 1-based stage order of ntt.cpp:312-324.
struct ModTab {
    uint64_t q, twoq;
    uint64_t mu;            // floor(2^64 / q), for barrett_reduce64
    InvScale sc;            // inv_n, inv_n_w and their Shoup factors
    const uint64_t* roots;
    const uint64_t* precon;
    const uint64_t* inv_roots;
    const uint64_t* precon_inv;
};

// ---- load transforms (applied to each word as it enters the transform) ----
struct XfIdent {
    HB_D uint64_t operator()(uint64_t x) const { return x; }
};
// base conversion of a coefficient-form word to this modulus
// (device/keyswitch/intt1_redu.hpp:36-38)
struct XfReduce {
    uint64_t q, mu;
    HB_D uint64_t operator()(uint64_t x) const { return barrett_reduce64(x, q, mu); }
};
// keyswitch rounding + base conversion (device/keyswitch/intt2_redu.hpp:24-51):
// v = (x + floor(qk/2)) mod qk ; out = (v mod qi + fix_i) mod qi
struct XfKsRound {
    uint64_t qk, qk_half, q, mu, fix;
    HB_D uint64_t operator()(uint64_t x) const {
        uint64_t v = x + qk_half;
        v -= (v >= qk) ? qk : 0;
        uint64_t r = barrett_reduce64(v, q, mu) + fix;
        return r - ((r >= q) ? q : 0);
    }
};

// ---- output functors ----
struct OfStore16 {  // 16 contiguous words -> 8 x 16-byte stores
    HB_D void operator()(uint64_t* dst, uint32_t off, const uint64_t (&v)[16]) const {
#pragma unroll
        for (int c = 0; c < 8; ++c) st2(dst + off + 2 * c, v[2 * c], v[2 * c + 1]);
    }
};
struct OfStore1 {
    HB_D void operator()(uint64_t* dst, uint32_t idx, uint64_t v) const { dst[idx] = v; }
};

template <class C, int P, class Xf>
HB_D void fwd_heads(uint32_t tid, uint64_t* sm, const uint64_t* src, const Xf& xf,
                    const ModTab& t) {
    if constexpr (P < C::NP) {
        fwd_head_pass<C, P>(tid, sm, src, xf, t.roots, t.precon, t.q, t.twoq);
        __syncthreads();
        fwd_heads<C, P + 1>(tid, sm, src, xf, t);
    }
}

// src (global, natural order) -> dst (global, bit-reversed order, [0,q)).
// src == dst is allowed (each CTA reads its whole polynomial before writing).
template <class C, class Xf, class Of>
HB_D void ntt_fwd_block(uint64_t* sm, const uint64_t* src, uint64_t* dst, const Xf& xf,
                        const Of& of, const ModTab& t) {
    const uint32_t tid = threadIdx.x;
    fwd_heads<C, 0>(tid, sm, src, xf, t);
    fwd_tail_pass<C>(tid, sm, dst, of, t.roots, t.precon, t.q, t.twoq);
}

template <class C, int P, class Of>
HB_D void inv_heads(uint32_t tid, uint64_t* sm, uint64_t* dst, const Of& of,
                    const ModTab& t) {
    if constexpr (P < C::NP) {
        __syncthreads();
        inv_head_pass<C, P>(tid, sm, dst, of, t.inv_roots, t.precon_inv, t.q, t.twoq, t.sc);
        inv_heads<C, P + 1>(tid, sm, dst, of, t);
    }
}

// src (global, bit-reversed order) -> dst (global, natural order, [0,q)).
template <class C, class Xf, class Of>
HB_D void ntt_inv_block(uint64_t* sm, const uint64_t* src, uint64_t* dst, const Xf& xf,
                        const Of& of, const ModTab& t) {
    const uint32_t tid = threadIdx.x;
    inv_tail_pass<C>(tid, sm, src, xf, t.inv_roots, t.precon_inv, t.q, t.twoq);
    inv_heads<C, 0>(tid, sm, dst, of, t);
}

}  // namespace hb
