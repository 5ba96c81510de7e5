// emul_ntt.cpp -- TEST INFRASTRUCTURE: replays the CUDA kernels' per-thread
// pass functions (hexl-fpga_b200/csrc/ntt_core.cuh, compiled as plain C++) on
// the CPU, one "thread" at a time with a barrier between passes, and compares
// against the oracle.  Validates the index math / swizzle / twiddle indexing
// without a GPU.  Built and run by tests/test_cpu_emul.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../hexl-fpga_b200/csrc/ntt_core.cuh"
#include "../../oracle/hexl_oracle.h"

using namespace hb;

struct XfIdent {
    uint64_t operator()(uint64_t x) const { return x; }
};
struct OfStore16 {
    void operator()(uint64_t* dst, uint32_t off, const uint64_t (&v)[16]) const {
        for (int k = 0; k < 16; ++k) dst[off + k] = v[k];
    }
};
struct OfStore1 {
    void operator()(uint64_t* dst, uint32_t idx, uint64_t v) const { dst[idx] = v; }
};

template <class C, int P>
void fwd_heads(std::vector<uint64_t>& sm, const uint64_t* src, const uint64_t* roots,
               const uint64_t* precon, uint64_t q) {
    if constexpr (P < C::NP) {
        for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid)
            fwd_head_pass<C, P>(tid, sm.data(), src, XfIdent(), roots, precon, q, 2 * q);
        fwd_heads<C, P + 1>(sm, src, roots, precon, q);
    }
}
template <class C, int P>
void inv_heads(std::vector<uint64_t>& sm, uint64_t* dst, const uint64_t* ir,
               const uint64_t* ip, uint64_t q, const InvScale& sc) {
    if constexpr (P < C::NP) {
        for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid)
            inv_head_pass<C, P>(tid, sm.data(), dst, OfStore1(), ir, ip, q, 2 * q, sc);
        inv_heads<C, P + 1>(sm, dst, ir, ip, q, sc);
    }
}

template <int LOGN, int LOGE>
int run(uint64_t q, int garbage) {
    using C = NttCfg<LOGN, LOGE>;
    const uint64_t n = C::N;
    uint64_t w = ho_min_primitive_root(2 * n, q);
    std::vector<uint64_t> roots(n), precon(n), ir(n), ip(n), a(n), ref(n), out(n), sm(n);
    ho_compute_roots(n, q, w, roots.data(), precon.data(), ir.data(), ip.data());
    ho_splitmix_fill(a.data(), n, 7 + LOGN, garbage ? 0 : q);
    if (garbage == 2) for (auto& x : a) x = ~(uint64_t)0;
    // forward
    ref = a;
    ho_fwd_ntt(ref.data(), n, q, roots.data(), precon.data());
    fwd_heads<C, 0>(sm, a.data(), roots.data(), precon.data(), q);
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid)
        fwd_tail_pass<C>(tid, sm.data(), out.data(), OfStore16(), roots.data(), precon.data(), q, 2 * q);
    int bad = 0;
    for (uint64_t i = 0; i < n; ++i) bad += out[i] != ref[i];
    // inverse of the (possibly garbage) input
    uint64_t inv_n = ho_inv_mod(n % q, q), inv_n_w = ho_mul_mod(inv_n, ir[n - 1], q);
    InvScale sc = {inv_n, ho_mult_factor64(inv_n, q), inv_n_w, ho_mult_factor64(inv_n_w, q)};
    ref = a;
    ho_inv_ntt(ref.data(), n, q, ir.data(), ip.data(), inv_n, inv_n_w);
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid)
        inv_tail_pass<C>(tid, sm.data(), a.data(), XfIdent(), ir.data(), ip.data(), q, 2 * q);
    inv_heads<C, 0>(sm, out.data(), ir.data(), ip.data(), q, sc);
    int badi = 0;
    for (uint64_t i = 0; i < n; ++i) badi += out[i] != ref[i];
    printf("LOGN=%d LOGE=%d q=%llu garbage=%d fwd_mismatch=%d inv_mismatch=%d\n", LOGN, LOGE,
           (unsigned long long)q, garbage, bad, badi);
    return bad + badi;
}

template <int LOGN, int LOGE>
int run_all() {
    uint64_t p[2];
    int rc = 0;
    size_t bits[] = {20, 51, 61};
    for (size_t b : bits) {
        ho_generate_primes(p, 1, b, (size_t)1 << LOGN);
        for (int g = 0; g < 3; ++g) rc += run<LOGN, LOGE>(p[0], g);
    }
    return rc;
}

int main() {
    int rc = 0;
    rc += run_all<10, 4>();
    rc += run_all<11, 4>();
    rc += run_all<12, 4>();
    rc += run_all<13, 4>();
    rc += run_all<14, 4>();
    rc += run_all<14, 5>();
    rc += run_all<13, 5>();
    rc += run_all<12, 5>();
    // divisor arithmetic
    uint64_t qs[] = {1, 2, 10, 20, 1000003, 2251799814045697ULL, (1ULL << 63) + 5, ~0ULL};
    uint64_t s = 99;
    std::vector<uint64_t> r(4);
    int dbad = 0;
    for (uint64_t q : qs) {
        Divisor dv = make_divisor(q);
        for (int it = 0; it < 20000; ++it) {
            s = ho_splitmix_fill(r.data(), 2, s, 0);
            uint64_t a = it & 1 ? r[0] : r[0] >> (it % 60), b = r[1];
            uint64_t am = mod64(a, dv), bm = mod64(b, dv);
            if (am != a % q || bm != b % q) ++dbad;
            uint64_t m = mulmod_reduced(am, bm, dv);
            if (m != (uint64_t)(((unsigned __int128)am * bm) % q)) ++dbad;
        }
    }
    printf("divisor mismatches=%d\n", dbad);
    rc += dbad;
    printf(rc ? "EMUL FAIL\n" : "EMUL OK\n");
    return rc != 0;
}
