// emul_ntt.cpp -- TEST INFRASTRUCTURE: replays the CUDA kernels' per-thread
// pass functions (hexl-fpga_b200/csrc/ntt_core.cuh, compiled as plain C++) on
// the CPU, one "thread" at a time with a barrier between passes, and compares
// against the oracle.  Validates the index math / swizzle / packed-twiddle
// layout, the exact arithmetic on garbage inputs and the fast arithmetic
// (approximate Shoup quotient, lazy forward, small-multiple reduction) on
// in-range inputs, without a GPU.  Built and run by tests/test_cpu_emul.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../hexl-fpga_b200/csrc/ntt_core.cuh"
#include "../../oracle/hexl_oracle.h"

using namespace hb;

struct Ident {
    template <class T>
    T operator()(T x) const { return x; }
};
static Ident ident;
struct Narrow {   // uint64 word -> uint32 register (small-modulus path)
    uint32_t operator()(uint64_t x) const { return (uint32_t)x; }
};

template <class C, int P, class A>
void fwd_mid(std::vector<typename A::elem>& sm, const typename A::Tw* tw, const A& a) {
    if constexpr (P < C::NP) {
        for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid) fwd_head_pass<C, P>(tid, sm.data(), tw, a);
        fwd_mid<C, P + 1>(sm, tw, a);
    }
}
template <class C, int P, class A>
void inv_mid(std::vector<typename A::elem>& sm, const typename A::Tw* tw, const A& a) {
    if constexpr (P < C::NP - 1) {
        for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid) inv_head_pass<C, P>(tid, sm.data(), tw, a);
        inv_mid<C, P + 1>(sm, tw, a);
    }
}

// forward transform of `in` exactly as the kernel sequences it.  C64 describes
// the uint64 landing buffer (TMA), C the working configuration (== C64 for the
// 64-bit paths; the uint32 configuration for the small-modulus path).
template <class C64, class C, class A>
void emul_fwd(const std::vector<uint64_t>& in, std::vector<uint64_t>& out, const typename A::Tw* tw, const A& a) {
    using T = typename A::elem;
    using P0 = FwdPass<C, 0>;
    std::vector<uint64_t> W(C::N);
    for (uint32_t i = 0; i < (uint32_t)C::N; ++i) W[swz(i)] = in[i];   // what the swizzled TMA load does
    std::vector<T> sm(C::N);
    std::vector<T> regs((size_t)C::NT * C::E);
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid) {
        if constexpr (sizeof(T) == 8) head_load<C, P0::R, P0::LS>(tid, W.data(), &regs[(size_t)tid * C::E], ident);
        else head_load<C, P0::R, P0::LS>(tid, W.data(), &regs[(size_t)tid * C::E], Narrow());
    }
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid) {
        fwd_head_compute<C, 0>(tid, &regs[(size_t)tid * C::E], tw, a);
        head_store<C, P0::R, P0::LS>(tid, sm.data(), &regs[(size_t)tid * C::E]);
    }
    fwd_mid<C, 1>(sm, tw, a);
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid)
        tail_load<C>(tid, sm.data(), &regs[(size_t)tid * C::E], ident);
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid) {
        T* v = &regs[(size_t)tid * C::E];
        fwd_tail_compute<C>(tid, v, tw, a);
        for (int ri = 0; ri < C::E / C::ROW; ++ri)
            for (int k = 0; k < C::ROW; ++k) out[tail_row<C>(tid, ri) * C::ROW + k] = v[ri * C::ROW + k];
    }
}

// inverse: the first pass reads rows.  For the small-modulus path the kernel
// reads each 32-word row as two 16-word rows of the uint64 landing buffer.
template <class C64, class C, class A>
void emul_inv(const std::vector<uint64_t>& in, std::vector<uint64_t>& out, const typename A::Tw* tw, const A& a) {
    using T = typename A::elem;
    using PL = InvPass<C, C::NP - 1>;
    std::vector<uint64_t> W(C::N);
    for (uint32_t i = 0; i < (uint32_t)C::N; ++i) W[swz(i)] = in[i];
    std::vector<T> sm(C::N);
    std::vector<T> regs((size_t)C::NT * C::E);
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid) {
        if constexpr (sizeof(T) == 8) {
            tail_load<C>(tid, W.data(), &regs[(size_t)tid * C::E], ident);
        } else {
            for (int ri = 0; ri < C::E / C::ROW; ++ri)
                for (int k = 0; k < C::ROW; ++k)
                    regs[(size_t)tid * C::E + ri * C::ROW + k] = (T)W[swz(tail_row<C>(tid, ri) * C::ROW + k)];
        }
    }
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid) {
        inv_tail_compute<C>(tid, &regs[(size_t)tid * C::E], tw, a);
        tail_store<C>(tid, sm.data(), &regs[(size_t)tid * C::E]);
    }
    inv_mid<C, 0>(sm, tw, a);
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid)
        head_load<C, PL::R, PL::LS>(tid, sm.data(), &regs[(size_t)tid * C::E], ident);
    for (uint32_t tid = 0; tid < (uint32_t)C::NT; ++tid) {
        T* v = &regs[(size_t)tid * C::E];
        inv_head_compute<C, C::NP - 1>(tid, v, tw, a);
        for (int gi = 0; gi < (C::E >> PL::R); ++gi)
            for (int k = 0; k < (1 << PL::R); ++k) out[inv_last_index<C>(tid, gi, k)] = v[gi * (1 << PL::R) + k];
    }
}

template <int LOGN, int LOGE>
int run(uint64_t q, int garbage) {
    using C = NttCfg<LOGN, LOGE>;
    const uint64_t n = C::N;
    uint64_t w = ho_min_primitive_root(2 * n, q);
    std::vector<uint64_t> roots(n), precon(n), ir(n), ip(n), a(n), ref(n), out(n);
    ho_compute_roots(n, q, w, roots.data(), precon.data(), ir.data(), ip.data());
    std::vector<TwPair> ftw(C::FWD_ENTRIES), itw(C::INV_ENTRIES);
    std::vector<int> used_f(n, 0), used_i(n, 0);
    for (uint32_t e = 0; e < (uint32_t)C::FWD_ENTRIES; ++e) {
        int s = fwd_pack_src<C>(e);
        ftw[e] = s < 0 ? TwPair{0, 0} : TwPair{roots[s], precon[s]};
        if (s >= 0) used_f[s]++;
    }
    for (uint32_t e = 0; e < (uint32_t)C::INV_ENTRIES; ++e) {
        int s = inv_pack_src<C>(e);
        itw[e] = s < 0 ? TwPair{0, 0} : TwPair{ir[s], ip[s]};
        if (s >= 0) used_i[s]++;
    }
    int cover = 0;
    for (uint64_t i = 1; i < n; ++i) cover += (used_f[i] != 1) + (used_i[i] != 1);
    ho_splitmix_fill(a.data(), n, 7 + LOGN, garbage ? 0 : q);
    if (garbage == 2) for (auto& x : a) x = ~(uint64_t)0;
    if (garbage == 3) for (uint64_t i = 0; i < n; ++i) a[i] = (i & 1) ? 4 * q - 1 - (i % 5) : q - 1;  // edge of fwd contract
    if (garbage == 4) for (uint64_t i = 0; i < n; ++i) a[i] = (i & 1) ? 2 * q - 1 - (i % 3) : q - 1;  // edge of inv contract
    if (garbage == 5) for (uint64_t i = 0; i < n; ++i) a[i] = 2 * q - 1;   // worst growth of the lazy inverse sums
    uint64_t inv_n = ho_inv_mod(n % q, q), inv_n_w = ho_mul_mod(inv_n, ir[n - 1], q);
    InvScale sc = {inv_n, ho_mult_factor64(inv_n, q), inv_n_w, ho_mult_factor64(inv_n_w, q)};
    ExactArith ex = {q, 2 * q, sc, 0};
    FastArith fa = {make_fastmod(q), sc};
    LazyInvArith la = {make_fastmod(q), sc};
    int bad = 0, badi = 0, badf = -1, badfi = -1, badli = -1;
    // exact path: always matches the oracle word for word
    ref = a;
    ho_fwd_ntt(ref.data(), n, q, roots.data(), precon.data());
    emul_fwd<C, C>(a, out, ftw.data(), ex);
    for (uint64_t i = 0; i < n; ++i) bad += out[i] != ref[i];
    const bool fwd_in_contract = (garbage == 0 || garbage >= 3) && fwd_fast_modulus_ok(q, LOGN);
    if (fwd_in_contract) {
        emul_fwd<C, C>(a, out, ftw.data(), fa);
        badf = 0;
        for (uint64_t i = 0; i < n; ++i) badf += out[i] != ref[i];
    }
    ref = a;
    ho_inv_ntt(ref.data(), n, q, ir.data(), ip.data(), inv_n, inv_n_w);
    emul_inv<C, C>(a, out, itw.data(), ex);
    for (uint64_t i = 0; i < n; ++i) badi += out[i] != ref[i];
    const bool inv_in_contract = (garbage == 0 || garbage >= 4) && inv_fast_modulus_ok(q);
    if (inv_in_contract) {
        emul_inv<C, C>(a, out, itw.data(), fa);
        badfi = 0;
        for (uint64_t i = 0; i < n; ++i) badfi += out[i] != ref[i];
        if (inv_lazy_modulus_ok(q)) {      // correction-free butterflies + mid-transform reduction
            emul_inv<C, C>(a, out, itw.data(), la);
            badli = 0;
            for (uint64_t i = 0; i < n; ++i) badli += out[i] != ref[i];
        }
    }
    printf("LOGN=%d LOGE=%d q=%llu in=%d exact fwd/inv mismatch=%d/%d fast fwd/inv mismatch=%d/%d lazy inv mismatch=%d "
           "pack_cover_err=%d\n",
           LOGN, LOGE, (unsigned long long)q, garbage, bad, badi, badf, badfi, badli, cover);
    return bad + badi + (badf > 0 ? badf : 0) + (badfi > 0 ? badfi : 0) + (badli > 0 ? badli : 0) + cover;
}

// small-modulus (uint32) path: configuration with 32-word rows
template <int LOGN, int LOGE>
int run_small(uint64_t q, int kind) {
    using C64 = NttCfg<LOGN, LOGE>;
    using C = NttCfg<LOGN, LOGE, 5>;
    const uint64_t n = C::N;
    uint64_t w = ho_min_primitive_root(2 * n, q);
    std::vector<uint64_t> roots(n), precon(n), ir(n), ip(n), a(n), ref(n), out(n);
    ho_compute_roots(n, q, w, roots.data(), precon.data(), ir.data(), ip.data());
    std::vector<Tw32> ftw(C::FWD_ENTRIES), itw(C::INV_ENTRIES);
    int cover = 0;
    std::vector<int> uf(n, 0), ui(n, 0);
    for (uint32_t e = 0; e < (uint32_t)C::FWD_ENTRIES; ++e) {
        int s = fwd_pack_src<C>(e);
        ftw[e] = s < 0 ? Tw32{0, 0} : Tw32{(uint32_t)roots[s], (uint32_t)(precon[s] >> 32)};
        if (s >= 0) uf[s]++;
    }
    for (uint32_t e = 0; e < (uint32_t)C::INV_ENTRIES; ++e) {
        int s = inv_pack_src<C>(e);
        itw[e] = s < 0 ? Tw32{0, 0} : Tw32{(uint32_t)ir[s], (uint32_t)(ip[s] >> 32)};
        if (s >= 0) ui[s]++;
    }
    for (uint64_t i = 1; i < n; ++i) cover += (uf[i] != 1) + (ui[i] != 1);
    ho_splitmix_fill(a.data(), n, 31 + LOGN, q);
    if (kind == 1) for (uint64_t i = 0; i < n; ++i) a[i] = (i & 1) ? 4 * q - 1 - (i % 5) : q - 1;   // edge of fwd contract
    if (kind == 2) for (uint64_t i = 0; i < n; ++i) a[i] = (i & 1) ? 2 * q - 1 - (i % 3) : q - 1;   // edge of inv contract
    uint64_t inv_n = ho_inv_mod(n % q, q), inv_n_w = ho_mul_mod(inv_n, ir[n - 1], q);
    InvScale sc = {inv_n, ho_mult_factor64(inv_n, q), inv_n_w, ho_mult_factor64(inv_n_w, q)};
    SmallArith sa = {make_small32(q, sc)};
    int bad = 0, badi = -1;
    ref = a;
    ho_fwd_ntt(ref.data(), n, q, roots.data(), precon.data());
    emul_fwd<C64, C>(a, out, ftw.data(), sa);
    for (uint64_t i = 0; i < n; ++i) bad += out[i] != ref[i];
    if (kind != 1) {
        ref = a;
        ho_inv_ntt(ref.data(), n, q, ir.data(), ip.data(), inv_n, inv_n_w);
        emul_inv<C64, C>(a, out, itw.data(), sa);
        badi = 0;
        for (uint64_t i = 0; i < n; ++i) badi += out[i] != ref[i];
    }
    printf("SMALL LOGN=%d LOGE=%d q=%llu in=%d fwd/inv mismatch=%d/%d pack_cover_err=%d\n", LOGN, LOGE,
           (unsigned long long)q, kind, bad, badi, cover);
    return bad + (badi > 0 ? badi : 0) + cover;
}

template <int LOGN, int LOGE>
int run_small_all() {
    uint64_t p[1];
    int rc = 0;
    size_t bits[] = {16, 20, 27, 29};
    for (size_t b : bits) {
        if (((size_t)1 << b) < ((size_t)2 << LOGN)) continue;
        if (ho_generate_primes(p, 1, b, (size_t)1 << LOGN) != 1) continue;
        for (int k = 0; k < 3; ++k) rc += run_small<LOGN, LOGE>(p[0], k);
    }
    rc += run_small<LOGN, LOGE>(136314881ULL % ((2ULL << LOGN)) == 1 ? 136314881ULL : p[0], 0);   // the reference's bench prime
    return rc;
}

// FP64-pipe arithmetic (modarith.cuh): words converted on entry as the kernels do after the vote
// Fp64AltArith that records the largest |word| / q leaving every stage
static double g_alt_max[32];
struct AltProbe : Fp64AltArith {
    template <int S> void fwd_at(uint64_t& X, uint64_t& Y, const TwPair& t) const {
        Fp64AltArith::fwd_at<S>(X, Y, t);
        const double v = fmax(fabs(u2d(X)), fabs(u2d(Y))) / m.q;
        if (v > g_alt_max[S]) g_alt_max[S] = v;
    }
};

template <int LOGN, int LOGE, int WT = 0>
int run_fp64(uint64_t q, int kind) {
    using C = NttCfg<LOGN, LOGE, 4, WT>;
    const uint64_t n = C::N;
    if (!fp64_modulus_ok(q)) return 0;
    uint64_t w = ho_min_primitive_root(2 * n, q);
    std::vector<uint64_t> roots(n), precon(n), ir(n), ip(n), a(n), ad(n), ref(n), out(n);
    ho_compute_roots(n, q, w, roots.data(), precon.data(), ir.data(), ip.data());
    std::vector<TwPair> ftw(C::FWD_ENTRIES), itw(C::INV_ENTRIES);
    auto entry = [&](uint64_t r) {
        const double ws = fp_centred(r, q);
        return TwPair{d2u(ws), d2u(fp_quot(ws, q))};
    };
    for (uint32_t e = 0; e < (uint32_t)C::FWD_ENTRIES; ++e) {
        int s = fwd_pack_src<C>(e);
        ftw[e] = s < 0 ? TwPair{0, 0} : entry(roots[s]);
    }
    for (uint32_t e = 0; e < (uint32_t)C::INV_ENTRIES; ++e) {
        int s = inv_pack_src<C>(e);
        itw[e] = s < 0 ? TwPair{0, 0} : entry(ir[s]);
    }
    ho_splitmix_fill(a.data(), n, 77 + LOGN + kind, q);
    const uint64_t top = q + (q >> 2) - 1;     // largest in-contract word
    if (kind == 1) for (auto& x : a) x = q - 1;
    if (kind == 2) for (uint64_t i = 0; i < n; ++i) a[i] = (i & 1) ? q - 1 : 0;
    if (kind == 3) for (uint64_t i = 0; i < n; ++i) a[i] = (i & 1) ? top - (i % 7) : (q >> 1) + (i % 3);
    if (kind == 4) for (uint64_t i = 0; i < n; ++i) a[i] = ((i >> (i % 14)) & 1) ? top : (q >> 1) + 1;
    if (kind == 5) for (auto& x : a) x = top;
    uint64_t inv_n = ho_inv_mod(n % q, q), inv_n_w = ho_mul_mod(inv_n, ir[n - 1], q);
    Fp64Arith fa = {make_fp64mod(q, inv_n, inv_n_w)};
    int bad = 0, badi = 0;
    ref = a;
    for (auto& x : ref) x %= q;                 // the oracle's contract; the result is the canonical residue either way
    ho_fwd_ntt(ref.data(), n, q, roots.data(), precon.data());
    for (uint64_t i = 0; i < n; ++i) ad[i] = fa.enter_fwd(a[i]);
    emul_fwd<C, C>(ad, out, ftw.data(), fa);
    for (uint64_t i = 0; i < n; ++i) bad += out[i] != ref[i];
    ref = a;
    for (auto& x : ref) x %= q;
    ho_inv_ntt(ref.data(), n, q, ir.data(), ip.data(), inv_n, inv_n_w);
    for (uint64_t i = 0; i < n; ++i) ad[i] = fa.enter_inv(a[i]);
    emul_inv<C, C>(ad, out, itw.data(), fa);
    for (uint64_t i = 0; i < n; ++i) badi += out[i] != ref[i];
    printf("FP64 LOGN=%d LOGE=%d warp-tail=%d q=%llu in=%d fwd/inv mismatch=%d/%d\n", LOGN, LOGE, WT,
           (unsigned long long)q, kind, bad, badi);
    // forward butterflies that correct every other stage (moduli up to 2^51 (1 + 1/32)): same canonical words,
    // and every intermediate word inside the bounds the analysis in modarith.cuh states
    int bada = 0;
    if (fp64_alt_modulus_ok(q)) {
        AltProbe pa;
        pa.m = fa.m;
        for (int i = 0; i < 32; ++i) g_alt_max[i] = 0.0;
        ref = a;
        for (auto& x : ref) x %= q;
        ho_fwd_ntt(ref.data(), n, q, roots.data(), precon.data());
        for (uint64_t i = 0; i < n; ++i) ad[i] = pa.enter_fwd(a[i]);
        emul_fwd<C, C>(ad, out, ftw.data(), pa);
        for (uint64_t i = 0; i < n; ++i) bada += out[i] != ref[i];
        double worst_a = 0, worst_b = 0;
        for (int sidx = 0; sidx < LOGN; ++sidx) ((sidx & 1) ? worst_b : worst_a) = fmax((sidx & 1) ? worst_b : worst_a, g_alt_max[sidx]);
        // after a correcting stage <= 1.26 q, after a plain one <= 1.92 q (and below 2^52 in any case)
        if (worst_a > 1.26 || worst_b > 1.92 || worst_b * (double)q >= 4503599627370496.0) ++bada;
        printf("FP64-alt LOGN=%d q=%llu in=%d fwd mismatch=%d max|v|/q after even/odd stages %.4f / %.4f\n", LOGN,
               (unsigned long long)q, kind, bada, worst_a, worst_b);
    }
    return bad + badi + bada;
}

template <int LOGN, int LOGE, int WT = 0>
int run_fp64_all() {
    int rc = 0;
    uint64_t p[1];
    size_t bits[] = {36, 44, 50, 51};
    for (size_t b : bits) {
        if (ho_generate_primes(p, 1, b, (size_t)1 << LOGN) != 1) continue;
        for (int k = 0; k < 6; ++k) rc += run_fp64<LOGN, LOGE, WT>(p[0], k);
    }
    // the largest modulus that takes the every-other-stage correction: the last NTT prime below 2^51 (1 + 1/32)
    for (uint64_t c = ((((uint64_t)1 << 51) + ((uint64_t)1 << 46)) / (2ull << LOGN)) * (2ull << LOGN) + 1;; c -= (2ull << LOGN))
        if (fp64_alt_modulus_ok(c) && ho_is_prime(c)) {
            for (int k = 0; k < 6; ++k) rc += run_fp64<LOGN, LOGE, WT>(c, k);
            break;
        }
    // the largest admissible modulus: the last NTT prime below 2^53 / 3
    for (uint64_t c = (((uint64_t)1 << 53) / 3 / (2ull << LOGN)) * (2ull << LOGN) + 1;; c -= (2ull << LOGN))
        if (c <= (((uint64_t)1 << 53) / 3) && ho_is_prime(c)) {
            for (int k = 0; k < 6; ++k) rc += run_fp64<LOGN, LOGE, WT>(c, k);
            break;
        }
    return rc;
}

template <int LOGN, int LOGE>
int run_all() {
    uint64_t p[2];
    int rc = 0;
    size_t bits[] = {12, 20, 30, 51, 57, 59, 61};
    for (size_t b : bits) {
        if (((size_t)1 << b) < ((size_t)2 << LOGN)) continue;
        if (ho_generate_primes(p, 1, b, (size_t)1 << LOGN) != 1) continue;
        for (int g = 0; g < 6; ++g) rc += run<LOGN, LOGE>(p[0], g);
    }
    return rc;
}

// unit properties of the FP64 modular product and the conditional correction at the edges
// of their stated ranges
static int fp64_properties() {
    int fbad = 0;
    uint64_t qs[] = {(1ULL << 36) + 1, 2251799814045697ULL, ((1ULL << 53) / 3 - 1) | 1};
    uint64_t s = 2024, r[3];
    for (uint64_t q : qs) {
        if (!fp64_modulus_ok(q)) { ++fbad; continue; }
        const Fp64Mod m = make_fp64mod(q, 1, 1);
        const uint64_t ymax = q + (q >> 1);            // 1.5 q: the largest multiplied word (inverse)
        for (int it = 0; it < 400000; ++it) {
            s = ho_splitmix_fill(r, 3, s, 0);
            uint64_t w = r[0] % q;
            if (it % 11 == 0) w = (q >> 1) + (it % 3);  // |centred w| at its maximum
            uint64_t ya = (it % 7 == 0) ? ymax - (r[1] % 5) : r[1] % (ymax + 1);
            const bool neg = r[2] & 1;
            const double y = neg ? -(double)(int64_t)ya : (double)(int64_t)ya;
            const double ws = fp_centred(w, q);
            const double rr = fp_mulmod(y, ws, fp_quot(ws, q), m);
            // exactness: rr is an integer congruent to y*w, within the stated bound
            const unsigned __int128 prod = (unsigned __int128)(ya % q) * w % q;
            uint64_t want = (uint64_t)prod;
            if (neg && want) want = q - want;
            const int64_t ri = (int64_t)rr;
            if ((double)ri != rr) ++fbad;
            const uint64_t got = (uint64_t)(((ri % (int64_t)q) + (int64_t)q) % (int64_t)q);
            if (got != want) ++fbad;
            const double bound = (double)q * (0.5 + (double)ya / 18014398509481984.0) + 1.0;
            if (fabs(rr) > bound) ++fbad;
            if (fp_to_canonical(rr, m) != want) ++fbad;
            // a product of a word below 1.5 q is below 0.75 q <= 2^51: the sign fix alone makes it canonical
            if (fabs(rr) > 2251799813685248.0 || fp_canon_signed(rr, m) != want) ++fbad;
            // full correction: any |x| < 2 q  ->  an integer congruent to x within q/2 (1 + 2^-40)
            {
                const uint64_t xa = r[1] % (2 * q);
                const double xf = neg ? -(double)(int64_t)xa : (double)(int64_t)xa;
                const double xc = fp_cred_full(xf, m);
                const int64_t xi = (int64_t)xc;
                if ((double)xi != xc || fabs(xc) > (double)q * 0.5 * (1.0 + 1e-12) + 1.0) ++fbad;
                uint64_t wantx = xa % q;
                if (neg && wantx) wantx = q - wantx;
                if ((uint64_t)(((xi % (int64_t)q) + (int64_t)q) % (int64_t)q) != wantx) ++fbad;
                if (fp_to_canonical_full(xf, m) != wantx) ++fbad;
            }
            // conditional correction: |x| <= 1.5 q -> at most max(q/2 (1 + 2^-20), |x| - q)
            const double x = neg ? -(double)(int64_t)ya : (double)(int64_t)ya;
            const double xr = fp_cred(x, m);
            const double lim = fmax((double)q * 0.5 * (1.0 + 1.0 / 1048576.0), fabs(x) - (double)q);
            if (fabs(xr) > lim) ++fbad;
            const int64_t xi = (int64_t)xr;
            if ((uint64_t)(((xi % (int64_t)q) + (int64_t)q) % (int64_t)q) != (neg ? (q - ya % q) % q : ya % q)) ++fbad;
            if (fp_from_int(ya) != (double)(int64_t)ya) ++fbad;
        }
    }
    printf("fp64 arithmetic property failures=%d\n", fbad);
    return fbad;
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
    rc += run_small_all<14, 5>();
    rc += run_small_all<13, 5>();
    rc += run_fp64_all<14, 5>();
    rc += run_fp64_all<14, 5, 1>();     // tail rows dealt out by warp
    rc += run_fp64_all<14, 4>();
    rc += run_fp64_all<12, 4>();
    rc += fp64_properties();
    // fast-arithmetic unit properties
    {
        uint64_t qs[] = {12289, 1073153, 2251799814045697ULL, (1ULL << 57) + 0x1234567ULL * 2 + 1, (1ULL << 60) - 93};
        uint64_t s = 4242, r[3];
        int fbad = 0;
        for (uint64_t q : qs) {
            FastMod m = make_fastmod(q);
            for (int it = 0; it < 200000; ++it) {
                s = ho_splitmix_fill(r, 3, s, 0);
                uint64_t w = r[0] % q, wp = ho_mult_factor64(w, q), y = it & 1 ? r[1] : r[1] % (60 * q > q ? 60 * q : q);
                uint64_t t = mul_shoup_approx(y, w, wp, m.nq);
                if (t >= 4 * q || t % q != ho_mul_mod(w, y % q, q)) ++fbad;
                uint64_t v = (q < (1ULL << 58)) ? r[2] % (60 * q) : r[2] % q;
                if (reduce_small_multiple(v, m) != v % q) ++fbad;
                if (q < (1ULL << 58)) {          // table form used by the forward kernels
                    uint64_t kq[64];
                    for (uint64_t k = 0; k < 64; ++k) kq[k] = k * q;
                    const uint64_t v3 = (it % 5 == 0) ? 64 * q - 1 - (r[2] % 64) : v;
                    if (reduce_by_table(v3, m, kq) != v3 % q) ++fbad;
                }
                if (inv_lazy_modulus_ok(q)) {
                    const uint64_t v2 = (it % 3 == 0) ? 1024 * q - 1 - (r[2] % 1000) : r[2] % (1024 * q);
                    if (reduce_mid(v2, m) != v2 % q) ++fbad;
                }
            }
        }
        printf("fast arithmetic property failures=%d\n", fbad);
        rc += fbad;
    }
    // divisor arithmetic (dyadic kernel)
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
            // the dyadic kernel's forms: normalisation shift on one operand, lazily summed cross term
            if (mulmod_preshifted(am << dv.s, bm, dv) != m) ++dbad;
            uint64_t hi, lo;
            mul_full(a, b, hi, lo);
            if ((((unsigned __int128)hi << 64) | lo) != (unsigned __int128)a * b) ++dbad;
            if ((q >> 63) == 0) {
                const uint64_t cm = mod64(r[0] ^ r[1], dv), em = (it % 7 == 0) ? q - 1 : mod64(r[1] * 3 + it, dv);
                const uint64_t want = (uint64_t)((((unsigned __int128)am * bm) % q + ((unsigned __int128)cm * em) % q) % q);
                if (mul2add_mod_preshifted(am << dv.s, bm, cm << dv.s, em, dv) != want) ++dbad;
                if (mul2add_mod_reduced(am, bm, cm, em, dv) != want) ++dbad;
            }
        }
    }
    printf("divisor mismatches=%d\n", dbad);
    rc += dbad;
    printf(rc ? "EMUL FAIL\n" : "EMUL OK\n");
    return rc != 0;
}
