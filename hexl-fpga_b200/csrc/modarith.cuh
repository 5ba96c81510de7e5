// modarith.cuh -- 64-bit modular arithmetic building blocks (host+device).
//
// Everything here is plain uint64 wrap-around arithmetic so that the exact
// Harvey butterflies reproduce the reference's words even for out-of-range
// inputs (reference: tests/test_utils/ntt.cpp:494-547, 618-657 and
// device/fwd_ntt.cpp:282-386, device/inv_ntt.cpp:149-437).
//
// The functions are __host__ __device__ so that tests/cpu_emul can run the very
// same index/arith code thread-by-thread on the CPU (test infrastructure, never
// a product fallback: the product entry points in capi.cu only launch kernels).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#define HB_D __device__ __forceinline__
#else
#define HB_HD inline
#define HB_D inline
#endif

namespace hb {

HB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

HB_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}

// Shoup/Harvey lazy product: w*x - floor(x*wp/2^64)*q  (mod 2^64), in [0,2q)
// when w < q and wp = floor(w*2^64/q).   tests/test_utils/ntt.hpp:87-101
HB_HD uint64_t mul_lazy(uint64_t x, uint64_t w, uint64_t wp, uint64_t q) {
    return w * x - mulhi64(x, wp) * q;
}

// Forward (Cooley-Tukey) Harvey butterfly, exact op order of
// tests/test_utils/ntt.cpp:519-546 / device/fwd_ntt.cpp:282-386.
HB_HD void fwd_bfly(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wp,
                    uint64_t q, uint64_t twoq) {
    uint64_t tx = X - ((X >= twoq) ? twoq : 0);
    uint64_t T = mul_lazy(Y, w, wp, q);
    X = tx + T;
    Y = tx + twoq - T;
}

// Inverse (Gentleman-Sande) Harvey butterfly, tests/test_utils/ntt.cpp:618-634
// / device/inv_ntt.cpp:149-398.
HB_HD void inv_bfly(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wp,
                    uint64_t q, uint64_t twoq) {
    uint64_t tx = X + Y;
    uint64_t ty = X + twoq - Y;
    X = tx - ((tx >= twoq) ? twoq : 0);
    Y = mul_lazy(ty, w, wp, q);
}

// Last inverse stage fused with the n^-1 scaling, ntt.cpp:640-657 /
// device/inv_ntt.cpp:400-437.  Outputs are fully reduced to [0,q).
// lazy_out: output_mod_factor = 2 of the reference (ntt.cpp:648-657): the words stay in [0, 2q)
HB_HD void inv_last_bfly(uint64_t& X, uint64_t& Y, uint64_t inv_n,
                         uint64_t inv_n_p, uint64_t inv_n_w, uint64_t inv_n_w_p,
                         uint64_t q, uint64_t twoq, bool lazy_out = false) {
    uint64_t tx = X + Y;
    tx -= (tx >= twoq) ? twoq : 0;
    uint64_t ty = X + twoq - Y;
    uint64_t x = mul_lazy(tx, inv_n, inv_n_p, q);
    uint64_t y = mul_lazy(ty, inv_n_w, inv_n_w_p, q);
    X = x - ((x >= q && !lazy_out) ? q : 0);
    Y = y - ((y >= q && !lazy_out) ? q : 0);
}

// ---------------------------------------------------------------------------
// "fast" arithmetic: any-correct-algorithm path for in-contract inputs.
//
// The exact butterflies above reproduce the reference word for word, which only
// matters for out-of-range inputs (the reference's ALL_MAX test).  When every
// input word is inside the algorithm's contract the output is the canonical
// residue, so ANY exact modular algorithm gives the same bits; the kernels
// check the range while loading and pick per polynomial.  The fast path trims
// the butterfly from ~28 to ~17-22 SASS instructions:
//   * Shoup quotient from 3 instead of 4 partial products:
//       Q'' = y1*p1 + hi32(y0*p1) + hi32(y1*p0)  in {Q-2, Q-1, Q}
//     so  T'' = w*y - Q''*q  lies in [0, 4q)   (T in [0,2q) for any y < 2^64);
//   * w*y + Q''*(-q) as one chain of 2 IMAD.WIDE + 4 IMAD (mod 2^64);
//   * forward: no per-stage correction at all (values grow by < 4q per stage,
//     bounded by 4q*(LOGN+1) < 2^64 for q < 2^58), one Barrett-style reduction
//     at the very end;
//   * inverse: values kept in [0,4q) with a single sign-test correction.
// ---------------------------------------------------------------------------
struct FastMod {
    uint64_t q;
    uint64_t nq;     // 2^64 - q
    uint64_t q4;     // 4q
    uint32_t kmul;   // floor(2^(shift+kb) / q), in (2^31, 2^32)
    uint32_t shift;  // max(bitlen(q) - 25, 0): (v >> shift) < 2^31 for v < 64q
    uint32_t kb;     // 56, or bitlen(q)+31 for tiny q
    uint32_t kmul3;  // table reduction (reduce_by_table): floor(2^(shift3+16) / q), at most 2^16
    // lazy inverse transform (q < 2^52): offsets 2^e * q and the constants of the
    // mid-transform reduction of values < 1024q
    uint64_t qsh[12];
    uint32_t kmul2;  // floor(2^(shift2+kb2) / q)
    uint32_t shift2; // max(bitlen(q) - 21, 0): (v >> shift2) < 2^31 for v < 1024q
    uint32_t kb2;    // 52, or bitlen(q)+31 for tiny q
    uint32_t shift3; // max(bitlen(q) - 6, 0): (v >> shift3) < 2^12 for v < 64q
    // Always 0.  Added as a third operand to two-operand 64-bit sums in the butterflies: ptxas
    // turns the high half of a plain a + b into IMAD.X (a*1 + b + carry) on the FMA-heavy pipe,
    // which is the pipe the transform is bound by; a three-operand sum stays an IADD3.X on the ALU.
    uint64_t zero64;
};

HB_HD FastMod make_fastmod(uint64_t q) {
    FastMod m;
    m.q = q;
    m.nq = (uint64_t)0 - q;
    m.q4 = q << 2;
    const int bl = 64 - clz64(q);
    m.shift = bl > 25 ? (uint32_t)(bl - 25) : 0u;
    m.kb = bl >= 25 ? 56u : (uint32_t)(bl + 31);
    unsigned __int128 k = (((unsigned __int128)1) << (m.shift + m.kb)) / q;
    m.kmul = k > 0xffffffffu ? 0xffffffffu : (uint32_t)k;
    m.shift3 = bl > 6 ? (uint32_t)(bl - 6) : 0u;
    m.kmul3 = (uint32_t)((((unsigned __int128)1) << (m.shift3 + 16)) / q);
    for (int e = 0; e < 12; ++e) m.qsh[e] = q << e;   // meaningful when q < 2^52 (lazy inverse only)
    m.shift2 = bl > 21 ? (uint32_t)(bl - 21) : 0u;
    m.kb2 = bl >= 21 ? 52u : (uint32_t)(bl + 31);
    unsigned __int128 k2 = (((unsigned __int128)1) << (m.shift2 + m.kb2)) / q;
    m.kmul2 = k2 > 0xffffffffu ? 0xffffffffu : (uint32_t)k2;
    m.zero64 = 0;
    return m;
}

// w*y - Q''*q (mod 2^64) in [0,4q) for ANY y; w < q, wp = floor(w*2^64/q).
HB_HD uint64_t mul_shoup_approx(uint64_t y, uint64_t w, uint64_t wp, uint64_t nq) {
    const uint32_t y0 = (uint32_t)y, y1 = (uint32_t)(y >> 32);
    const uint32_t p0 = (uint32_t)wp, p1 = (uint32_t)(wp >> 32);
    const uint32_t w0 = (uint32_t)w, w1 = (uint32_t)(w >> 32);
    const uint32_t n0 = (uint32_t)nq, n1 = (uint32_t)(nq >> 32);
#if defined(__CUDA_ARCH__)
    // Q'' with explicit 32-bit carry chains: keeps the two small additions on the
    // ALU pipe (ptxas otherwise zero-extends hi32(c) into a register pair to ride
    // on an IMAD.WIDE addend, i.e. two moves on the busier FMA pipe).
    uint32_t Q0, Q1, t0, t1;
    asm("{\n\t"
        ".reg .u32 c0, c1, d0, d1, e0, e1;\n\t"
        ".reg .u64 c, d, e, t;\n\t"
        "mul.wide.u32 c, %4, %7;\n\t"
        "mul.wide.u32 d, %5, %6;\n\t"
        "mul.wide.u32 e, %5, %7;\n\t"
        "mov.b64 {c0, c1}, c;\n\t"
        "mov.b64 {d0, d1}, d;\n\t"
        "mov.b64 {e0, e1}, e;\n\t"
        "add.cc.u32 e0, e0, c1;\n\t"
        "addc.u32 e1, e1, 0;\n\t"
        "add.cc.u32 %0, e0, d1;\n\t"
        "addc.u32 %1, e1, 0;\n\t"
        "mul.wide.u32 t, %8, %4;\n\t"
        "mad.wide.u32 t, %0, %9, t;\n\t"
        "mov.b64 {%2, %3}, t;\n\t"
        "}"
        : "=&r"(Q0), "=&r"(Q1), "=r"(t0), "=r"(t1)
        : "r"(y0), "r"(y1), "r"(p0), "r"(p1), "r"(w0), "r"(n0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t1) : "r"(w0), "r"(y1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t1) : "r"(w1), "r"(y0));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t1) : "r"(Q0), "r"(n1));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t1) : "r"(Q1), "r"(n0));
    return ((uint64_t)t1 << 32) | t0;
#else
    (void)w0; (void)w1; (void)n0; (void)n1;
    const uint64_t Q = (uint64_t)y1 * p1 + (((uint64_t)y0 * p1) >> 32) + (((uint64_t)y1 * p0) >> 32);
    return w * y + Q * nq;
#endif
}

// x - 4q if x >= 4q else x, for x < 2^63 + 4q (one sign test instead of a
// 64-bit compare).
HB_HD uint64_t csub(uint64_t x, uint64_t m) {
    const uint64_t d = x - m;
    return ((int64_t)d < 0) ? x : d;
}

// forward, lazy: X' = X + T'', Y' = X + 4q - T''   (no correction; see above)
HB_HD void fwd_bfly_fast(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wp, const FastMod& m) {
    const uint64_t T = mul_shoup_approx(Y, w, wp, m.nq);
    const uint64_t x = X;
    X = x + T + m.zero64;
    Y = x + m.q4 - T;
}

// v < 64q  ->  v mod q.  k = floor((v >> shift) * kmul / 2^kb) is floor(v/q) or
// one less: the truncations of v and kmul lose less than 2^-22 of the quotient,
// so r = v - k*q lies in [0, 2q) and one conditional subtraction finishes.
HB_HD uint64_t reduce_small_multiple(uint64_t v, const FastMod& m) {
    const uint32_t vh = (uint32_t)(v >> m.shift);
    const uint32_t k = (uint32_t)(((uint64_t)vh * m.kmul) >> m.kb);
    const uint64_t r = v - (uint64_t)k * m.q;
    return csub(r, m.q);
}

// Same reduction with the multiple of q looked up instead of multiplied: kq[k] = k*q
// for k < 64 (a 512-byte table in shared memory).  The estimate needs one 32-bit
// IMAD (12-bit by 17-bit product) instead of two IMAD.WIDE and an IMAD, i.e. 2
// instead of 10 cycles of the multiplier pipe the forward transform is bound by:
// k = floor((v >> shift3) * kmul3 / 2^16) is floor(v/q) or one less (the two
// truncations cost less than 0.1 of a unit), so r = v - kq[k] lies in [0, 2q).
HB_HD uint64_t reduce_by_table(uint64_t v, const FastMod& m, const uint64_t* kq) {
    const uint32_t vh = (uint32_t)(v >> m.shift3);
    const uint32_t k = (vh * m.kmul3) >> 16;
    return csub(v - kq[k], m.q);
}

// inverse, values in [0,4q):  X' = (X+Y) csub 4q,  Y' = T''(X + 4q - Y)
HB_HD void inv_bfly_fast(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wp, const FastMod& m) {
    const uint64_t tx = X + Y + m.zero64;
    const uint64_t ty = X + m.q4 - Y;
    X = csub(tx, m.q4);
    Y = mul_shoup_approx(ty, w, wp, m.nq);
}

// ---- lazy inverse (q < 2^52) ------------------------------------------------
// Gentleman-Sande without the per-stage correction of the sum: with inputs below
// 2^E * q,  X' = X + Y < 2^(E+1) q  and  Y' = T''(X + 2^E q - Y) in [0,4q)  (the
// Shoup product accepts ANY 64-bit argument), so the bound doubles per stage and
// one real reduction (reduce_mid) half way keeps everything below 2^64:
// 2^(E+1) q < 2^64 needs E <= 11 for q < 2^52.  Saves the compare / select /
// subtract of every butterfly (about a fifth of the inverse instruction stream).
template <int E>
HB_HD void inv_bfly_lazy(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wp, const FastMod& m) {
    static_assert(E >= 1 && E <= 11, "lazy inverse bound out of range");
    const uint64_t tx = X + Y;
    const uint64_t ty = X + m.qsh[E] - Y;
    X = tx;
    Y = mul_shoup_approx(ty, w, wp, m.nq);
}
template <int E>
HB_HD void inv_last_bfly_lazy(uint64_t& X, uint64_t& Y, uint64_t inv_n, uint64_t inv_n_p, uint64_t inv_n_w,
                              uint64_t inv_n_w_p, const FastMod& m) {
    static_assert(E >= 1 && E <= 11, "lazy inverse bound out of range");
    const uint64_t tx = X + Y;
    const uint64_t ty = X + m.qsh[E] - Y;
    uint64_t x = mul_shoup_approx(tx, inv_n, inv_n_p, m.nq);   // [0,4q)
    uint64_t y = mul_shoup_approx(ty, inv_n_w, inv_n_w_p, m.nq);
    x = csub(x, m.q << 1);
    y = csub(y, m.q << 1);
    X = csub(x, m.q);
    Y = csub(y, m.q);
}
// v < 1024q -> v mod q; same construction as reduce_small_multiple with wider constants:
// the estimate never exceeds floor(v/q) and falls short of it by at most one.
HB_HD uint64_t reduce_mid(uint64_t v, const FastMod& m) {
    const uint32_t vh = (uint32_t)(v >> m.shift2);
    const uint32_t k = (uint32_t)(((uint64_t)vh * m.kmul2) >> m.kb2);
    const uint64_t r = v - (uint64_t)k * m.q;
    return csub(r, m.q);
}
HB_HD bool inv_lazy_modulus_ok(uint64_t q) { return q < ((uint64_t)1 << 52); }

// last inverse stage with the n^-1 scaling, canonical outputs
HB_HD void inv_last_bfly_fast(uint64_t& X, uint64_t& Y, uint64_t inv_n, uint64_t inv_n_p,
                              uint64_t inv_n_w, uint64_t inv_n_w_p, const FastMod& m) {
    const uint64_t tx = X + Y;            // < 8q: fine for the Shoup product
    const uint64_t ty = X + m.q4 - Y;
    uint64_t x = mul_shoup_approx(tx, inv_n, inv_n_p, m.nq);   // [0,4q)
    uint64_t y = mul_shoup_approx(ty, inv_n_w, inv_n_w_p, m.nq);
    x = csub(x, m.q << 1);
    y = csub(y, m.q << 1);
    X = csub(x, m.q);
    Y = csub(y, m.q);
}

// ---------------------------------------------------------------------------
// FP64-pipe arithmetic for 2^36 <= q <= 2^53 / 3  (the 50..52-bit primes of the
// HE parameter sets), in-contract inputs only.
//
// The 64-bit integer butterfly is bound by the integer multiplier (28 FMA-heavy
// pipe cycles per warp, modarith.cuh above); the B200's FP64 pipe issues a
// DFMA/DADD/DMUL every ~2.1 cycles per warp and is otherwise idle.  Here the
// words of a transform are integer-valued doubles in a centred, slightly
// redundant range and a modular product is six FP64 instructions:
//     c = rint(y * (w/q))        fma(y, wi, M) - M,  M = 1.5 * 2^52
//     h + l = y * w  exactly     h = y*w,  l = fma(y, w, -h)
//     r = (h - c*q) + l          fma(-c, q, h) is exact: an integer below 2^53
// With |w| <= q/2 (centred twiddle), wi = fl(w/q) and |y| <= 2^52:
//     |c - y*w/q| <= 1/2 + |y| * 2^-54   =>   |r| <= q * (1/2 + |y| * 2^-54).
// A forward butterfly keeps every word in |v| <= 1.25 q: x is first brought to
// |x| <= q/2 (1 + 2^-20) by one conditional -+q (the comparison and the selection
// of the correction run on the ALU pipe, on the high word of the double), then
// X' = x + r, Y' = x - r with |r| <= q (1/2 + 1.25 q 2^-54) <= 0.735 q.  An
// inverse butterfly keeps |v| <= 0.75 q: s = x + y (<= 1.5 q) gets the same
// conditional correction, u = x - y (<= 1.5 q <= 2^52) goes through the product,
// |r| <= q (1/2 + 1.5 q 2^-54) <= 0.75 q.  Every sum is an integer below 2^53, so
// all of this is exact and the canonical residue comes out bit-identical to the
// integer paths: 9 FP64 + ~4 ALU instructions per butterfly.
// ---------------------------------------------------------------------------
HB_HD double u2d(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
HB_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
// explicitly rounded operations: never contracted or re-associated by the compiler
HB_HD double fp_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
HB_HD double fp_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
HB_HD double fp_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

constexpr uint64_t kFpMagicBits = 0x4338000000000000ull;   // 1.5 * 2^52
constexpr uint64_t kFpTwo52Bits = 0x4330000000000000ull;   // 2^52

struct Fp64Mod {
    double q, nq;          // the modulus and its negative
    uint32_t half_hi;      // high word of (double)(q/2): the threshold of the conditional correction
    uint32_t q_hi, q_lo;   // bit pattern of (double)q
    uint32_t pad;
    uint64_t qi;           // the modulus as an integer
    uint64_t vote;         // in-contract inputs are below q + q/4
    // n^-1 and n^-1 * w of the last inverse stage: centred residues and their quotients by q
    double inv_n, inv_n_q, inv_n_w, inv_n_w_q;
    double inv_q;          // fl(1 / q), for the full reduction fp_cred_full
};
HB_HD bool fp64_modulus_ok(uint64_t q) {
    return q >= ((uint64_t)1 << 36) && q <= (((uint64_t)1 << 53) / 3) && (q & 1);
}
// centred representative of a residue as a double, and its correctly rounded quotient by q
HB_HD double fp_centred(uint64_t w, uint64_t q) {
    return (w > (q >> 1)) ? -(double)(int64_t)(q - w) : (double)(int64_t)w;
}
HB_HD double fp_quot(double ws, uint64_t q) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(ws, (double)(int64_t)q);
#else
    return ws / (double)(int64_t)q;
#endif
}
HB_HD Fp64Mod make_fp64mod(uint64_t q, uint64_t inv_n, uint64_t inv_n_w) {
    Fp64Mod m;
    m.q = (double)(int64_t)q;
    m.nq = -m.q;
    m.half_hi = (uint32_t)(d2u(m.q * 0.5) >> 32);
    m.q_hi = (uint32_t)(d2u(m.q) >> 32);
    m.q_lo = (uint32_t)d2u(m.q);
    m.pad = 0;
    m.qi = q;
    m.vote = q + (q >> 2);
    m.inv_n = fp_centred(inv_n, q);
    m.inv_n_q = fp_quot(m.inv_n, q);
    m.inv_n_w = fp_centred(inv_n_w, q);
    m.inv_n_w_q = fp_quot(m.inv_n_w, q);
    m.inv_q = fp_quot(1.0, q);
    return m;
}

// |x| <= 1.5 q  ->  |x'| <= max(q/2 (1 + 2^-20), |x| - q),  x' = x (mod q)
HB_HD double fp_cred(double x, const Fp64Mod& m) {
#if defined(__CUDA_ARCH__)
    const uint32_t hi = (uint32_t)__double2hiint(x);
    const bool big = (hi & 0x7fffffffu) > m.half_hi;
    const uint32_t chi = big ? (m.q_hi | (hi & 0x80000000u)) : 0u;
    const uint32_t clo = big ? m.q_lo : 0u;
    return __dadd_rn(x, -__hiloint2double((int)chi, (int)clo));
#else
    const uint32_t hi = (uint32_t)(d2u(x) >> 32);
    const bool big = (hi & 0x7fffffffu) > m.half_hi;
    const uint32_t chi = big ? (m.q_hi | (hi & 0x80000000u)) : 0u;
    const uint32_t clo = big ? m.q_lo : 0u;
    return fp_add(x, -u2d(((uint64_t)chi << 32) | clo));
#endif
}
// FULL reduction  x - q * rint(x / q):  |x| < 2^52  ->  |x'| <= q/2 (1 + 2^-49), x' = x (mod q).
// Three FP64 instructions (six cycles of the warp scheduler: every FP64 instruction takes two,
// tools/ubench5.cu) against the conditional correction's one FP64 + five ALU (seven) -- and it takes ANY
// magnitude, which is what lets the forward butterflies below correct only every other stage.
// rint(x / q) is at most 2 in magnitude here, so c * q and the difference are exact.
HB_HD double fp_cred_full(double x, const Fp64Mod& m) {
    const double magic = u2d(kFpMagicBits);
    const double c = fp_add(fp_fma(x, m.inv_q, magic), -magic);
    return fp_fma(c, m.nq, x);
}
// integer-valued double with |v| < q (and |v| <= 2^51)  ->  canonical residue in [0, q) as an integer:
// v + (v < 0 ? q : 0).  The bit pattern of v + 1.5 * 2^52 is that of the constant plus v as a two's-complement
// integer, so the sign of v is one comparison of the high word, and the constant's offset and the conditional
// + q are folded into one selected 64-bit addend: one DADD + 5 ALU instructions.
HB_HD uint64_t fp_canon_signed(double v, const Fp64Mod& m) {
    const double t = fp_add(v, u2d(kFpMagicBits));
    const uint64_t b = d2u(t);
#if defined(__CUDA_ARCH__)
    // the high word through an opaque move: the compiler otherwise widens the test to a 64-bit compare (two ISETP)
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(t));
    (void)lo;
#else
    const uint32_t hi = (uint32_t)(b >> 32);
#endif
    const bool neg = hi < (uint32_t)(kFpMagicBits >> 32);
    return b + (neg ? m.qi - kFpMagicBits : (uint64_t)0 - kFpMagicBits);
}
// integer-valued double with |v| < q < 2^52  ->  canonical residue in [0, q) as an integer:  v + (v < 0 ? q : 0).
// The conditional + q rides on the conversion itself: t = v + (v < 0 ? q + 2^52 : 2^52) is an integer in
// [2^52, 2^53), i.e. exact, and its bit pattern is that of 2^52 with the residue in the mantissa.  The sign of v is
// one unsigned comparison of its high word (> 0x80000000: -0.0 counts as zero), the addend two selects, and the
// result stays in the register pair the DADD wrote (only its high word loses the exponent): one DADD + 4 ALU
// instructions.  For results that leave through 16-byte stores (forward final, keyswitch sums): with fp_canon_signed
// below the two halves of the integer sum land in unrelated registers, four moves in front of every store.  Words
// that leave one by one (last inverse stage, keyswitch S5) keep fp_canon_signed: its chain has one FP64 instruction
// instead of two behind the product, and the inverse kernel measured 3 % slower with this one.
HB_HD uint64_t fp_canon_pair(double v, const Fp64Mod& m) {
#if defined(__CUDA_ARCH__)
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(v));
    (void)lo;
#else
    const uint32_t hi = (uint32_t)(d2u(v) >> 32);
#endif
    const bool neg = hi > 0x80000000u;
    const double two52 = u2d(kFpTwo52Bits);
    const double t = fp_add(v, neg ? fp_add(m.q, two52) : two52);
    return d2u(t) ^ kFpTwo52Bits;
}
// |v| < 2^52  ->  canonical residue in [0, q) as an integer (full reduction first)
HB_HD uint64_t fp_to_canonical_full(double v, const Fp64Mod& m) { return fp_canon_pair(fp_cred_full(v, m), m); }
// y * w (mod q) for |y| <= 2^52, in |r| <= q (1/2 + |y| 2^-54)
HB_HD double fp_mulmod(double y, double w, double wi, const Fp64Mod& m) {
    const double magic = u2d(kFpMagicBits);
    const double c = fp_add(fp_fma(y, wi, magic), -magic);
    const double h = fp_mul(y, w);
    const double l = fp_fma(y, w, -h);
    const double d = fp_fma(c, m.nq, h);
    return fp_add(d, l);
}
// integer word below 2^52 -> double
HB_HD double fp_from_int(uint64_t x) { return fp_add(u2d(x | kFpTwo52Bits), -u2d(kFpTwo52Bits)); }
// |v| <= 1.5 q  ->  canonical residue in [0, q) as an integer
HB_HD uint64_t fp_to_canonical(double v, const Fp64Mod& m) { return fp_canon_pair(fp_cred(v, m), m); }
HB_HD void fwd_bfly_fp64(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wi, const Fp64Mod& m) {
    const double x = fp_cred(u2d(X), m);
    const double r = fp_mulmod(u2d(Y), u2d(w), u2d(wi), m);
    X = d2u(fp_add(x, r));
    Y = d2u(fp_add(x, -r));
}
// Forward butterflies that correct every OTHER stage (moduli up to 2^51 (1 + 1/32), see fp64_alt_modulus_ok):
//   stage A (even):  x <- fp_cred_full(x)  (|x| <= q/2),  X' = x + r,  Y' = x - r     |out| <= q (1 + k b)
//   stage B (odd) :  no correction at all,                X' = x + r,  Y' = x - r     |out| <= b' + q (1/2 + k b')
// with k = q 2^-54 <= 0.1328 and |r| <= q (1/2 + k |y|).  From inputs below 1.25 q the bounds settle at
// b' <= 1.26 q after an A stage and b <= 1.92 q after a B stage: every y stays below 2^52 (the product's
// contract) and every word far below 2^53 (exact integers).  3 FP64 instead of 1 FP64 + 5 ALU per correction
// and half as many corrections: 19.5 instead of 23 scheduler cycles per butterfly on average.
HB_HD void fwd_bfly_fp64_a(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wi, const Fp64Mod& m) {
    const double x = fp_cred_full(u2d(X), m);
    const double r = fp_mulmod(u2d(Y), u2d(w), u2d(wi), m);
    X = d2u(fp_add(x, r));
    Y = d2u(fp_add(x, -r));
}
HB_HD void fwd_bfly_fp64_b(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wi, const Fp64Mod& m) {
    const double x = u2d(X);
    const double r = fp_mulmod(u2d(Y), u2d(w), u2d(wi), m);
    X = d2u(fp_add(x, r));
    Y = d2u(fp_add(x, -r));
}
// the bounds above need  1.92 q <= 2^52:  q <= 2^51 (1 + 1/32)
HB_HD bool fp64_alt_modulus_ok(uint64_t q) {
    return fp64_modulus_ok(q) && q <= (((uint64_t)1 << 51) + ((uint64_t)1 << 46));
}
// HB_INV_CRED_FULL: correct the sum with x - q rint(x / q) (3 FP64 = 6 scheduler cycles) instead of the conditional
// +-q (1 FP64 + 5 ALU = 7): the inverse kernels are issue-bound with both pipes half idle
#ifndef HB_INV_CRED_FULL
#define HB_INV_CRED_FULL 1
#endif
HB_HD void inv_bfly_fp64(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wi, const Fp64Mod& m) {
    const double x = u2d(X), y = u2d(Y);
    const double s = fp_add(x, y), u = fp_add(x, -y);
#if HB_INV_CRED_FULL
    X = d2u(fp_cred_full(s, m));
#else
    X = d2u(fp_cred(s, m));
#endif
    Y = d2u(fp_mulmod(u, u2d(w), u2d(wi), m));
}
// First inverse stage: the words come straight from fp_from_int, x, y in [0, 1.25 q) (the contract of the
// range vote), no centring on entry.  The sum (< 2.5 q) takes the full correction, which accepts any
// magnitude; the difference |u| < 1.25 q is inside the product's contract as it is.
HB_HD void inv_bfly_fp64_first(uint64_t& X, uint64_t& Y, uint64_t w, uint64_t wi, const Fp64Mod& m) {
    const double x = u2d(X), y = u2d(Y);
    const double s = fp_add(x, y), u = fp_add(x, -y);
    X = d2u(fp_cred_full(s, m));
    Y = d2u(fp_mulmod(u, u2d(w), u2d(wi), m));
}
// Last inverse stage with the n^-1 scaling.  |s|, |u| <= 1.5 q, so both products come out with
// |r| <= q (1/2 + 1.5 q 2^-54) <= 0.75 q <= 2^51 (q <= 2^53 / 3): inside fp_canon_signed's range without a correction.
HB_HD void inv_last_bfly_fp64(uint64_t& X, uint64_t& Y, const Fp64Mod& m) {
    const double x = u2d(X), y = u2d(Y);
    const double s = fp_add(x, y), u = fp_add(x, -y);
    X = fp_canon_signed(fp_mulmod(s, m.inv_n, m.inv_n_q, m), m);
    Y = fp_canon_signed(fp_mulmod(u, m.inv_n_w, m.inv_n_w_q, m), m);
}

// The same stage for a consumer on the FP64 pipe (keyswitch S1 -> S2): the results stay doubles, as the non-negative
// representative v + (v < 0 ? q : 0) in [0, q) that the reference's base conversion starts from -- no integer
// canonicalisation here, no integer-to-double conversion in each of the transforms that read the word.
HB_HD double fp_nonneg(double v, const Fp64Mod& m) {
#if defined(__CUDA_ARCH__)
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(v));
    (void)lo;
#else
    const uint32_t hi = (uint32_t)(d2u(v) >> 32);
#endif
    return fp_add(v, hi > 0x80000000u ? m.q : 0.0);     // -0.0 counts as zero and comes out as +0.0
}
HB_HD void inv_last_bfly_fp64_d(uint64_t& X, uint64_t& Y, const Fp64Mod& m) {
    const double x = u2d(X), y = u2d(Y);
    const double s = fp_add(x, y), u = fp_add(x, -y);
    X = d2u(fp_nonneg(fp_mulmod(s, m.inv_n, m.inv_n_q, m), m));
    Y = d2u(fp_nonneg(fp_mulmod(u, m.inv_n_w, m.inv_n_w_q, m), m));
}

// x mod q for any x < 2^64 with mu = floor(2^64/q)   (q < 2^63).
// device/keyswitch/intt1_redu.hpp:36-38 computes the same canonical value.
HB_HD uint64_t barrett_reduce64(uint64_t x, uint64_t q, uint64_t mu) {
    uint64_t r = x - mulhi64(x, mu) * q;  // in [0, 2q)
    return r - ((r >= q) ? q : 0);
}

HB_HD uint64_t add_mod(uint64_t a, uint64_t b, uint64_t q) {  // a,b < q < 2^63
    uint64_t s = a + b;
    return s - ((s >= q) ? q : 0);
}
HB_HD uint64_t sub_mod(uint64_t a, uint64_t b, uint64_t q) {  // a,b < q
    return (a >= b) ? a - b : a + q - b;
}

// ---- generic 128-by-64 remainder (any modulus >= 1), Moeller-Granlund
// "division by invariant integers" 2-by-1 step with a precomputed reciprocal.
struct Divisor {
    uint64_t d;   // modulus shifted left until its top bit is set
    uint64_t v;   // floor((2^128-1)/d) - 2^64
    uint32_t s;   // shift
    uint32_t pad;
    uint64_t q;   // the modulus itself
};

HB_HD Divisor make_divisor(uint64_t q) {
    Divisor dv;
    dv.q = q;
    dv.pad = 0;
    if (q == 0) {  // undefined in the reference; keep the arithmetic total
        dv.d = (uint64_t)1 << 63;
        dv.s = 63;
        dv.v = ~(uint64_t)0;
        return dv;
    }
    dv.s = (uint32_t)clz64(q);
    dv.d = q << dv.s;
    unsigned __int128 all = ~(unsigned __int128)0;
    dv.v = (uint64_t)(all / dv.d);  // low 64 bits == floor((2^128-1)/d) - 2^64
    return dv;
}

// remainder of (u1:u0) by d, d normalised, u1 < d.
HB_HD uint64_t rem_2by1(uint64_t u1, uint64_t u0, uint64_t d, uint64_t v) {
    // (q1:q0) = v*u1 + (u1:u0)
    uint64_t q0 = v * u1;
    uint64_t q1 = mulhi64(v, u1);
    uint64_t t0 = q0 + u0;
    q1 += u1 + ((t0 < q0) ? 1 : 0);
    q0 = t0;
    q1 += 1;
    uint64_t r = u0 - q1 * d;
    if (r > q0) r += d;
    if (r >= d) r -= d;
    return r;
}

// x mod q for one 64-bit word.
HB_HD uint64_t mod64(uint64_t x, const Divisor& dv) {
    if (x < dv.q) return x;
    uint64_t u1 = dv.s ? (x >> (64 - dv.s)) : 0;
    uint64_t u0 = x << dv.s;
    return rem_2by1(u1, u0, dv.d, dv.v) >> dv.s;
}

// (a*b) mod q with a,b < q (product's high word is then < q).
HB_HD uint64_t mulmod_reduced(uint64_t a, uint64_t b, const Divisor& dv) {
    uint64_t lo = a * b;
    uint64_t hi = mulhi64(a, b);
    uint64_t u1 = dv.s ? ((hi << dv.s) | (lo >> (64 - dv.s))) : hi;
    uint64_t u0 = lo << dv.s;
    return rem_2by1(u1, u0, dv.d, dv.v) >> dv.s;
}

// full 64x64 -> 128 product from four 32x32 partial products (4 IMAD.WIDE)
HB_HD void mul_full(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    const uint64_t p00 = (uint64_t)a0 * b0, p01 = (uint64_t)a0 * b1, p10 = (uint64_t)a1 * b0, p11 = (uint64_t)a1 * b1;
    const uint64_t mid = p01 + (p00 >> 32);              // < 2^64: (2^32-1)^2 + 2^32 - 1
    const uint64_t mid2 = p10 + (uint32_t)mid;
    hi = p11 + (mid >> 32) + (mid2 >> 32);
    lo = (mid2 << 32) | (uint32_t)p00;
}

// Products with the normalisation shift applied to ONE OPERAND instead of the
// 128-bit product: for a, b < q,  (a << s) * b = (a*b) << s  has its high word
// below d = q << s, which is what rem_2by1 needs; the remainder comes back
// shifted by s.  as_ = a << s.
HB_HD uint64_t mulmod_preshifted(uint64_t as_, uint64_t b, const Divisor& dv) {
    uint64_t u1, u0;
    mul_full(as_, b, u1, u0);
    return rem_2by1(u1, u0, dv.d, dv.v) >> dv.s;
}
// (a*b + c*e) mod q for a,b,c,e < q < 2^63 from pre-shifted a, c: the sum of
// the two shifted products stays below d * 2^64 (2*q^2*2^s <= q*2^s*2^64).
HB_HD uint64_t mul2add_mod_preshifted(uint64_t as_, uint64_t b, uint64_t cs_, uint64_t e, const Divisor& dv) {
    uint64_t h1, l1, h2, l2;
    mul_full(as_, b, h1, l1);
    mul_full(cs_, e, h2, l2);
    const uint64_t u0 = l1 + l2;
    const uint64_t u1 = h1 + h2 + ((u0 < l1) ? 1 : 0);
    return rem_2by1(u1, u0, dv.d, dv.v) >> dv.s;
}

// (a*b + c*d) mod q with a,b,c,d < q < 2^63: one reduction of the 128-bit sum.
HB_HD uint64_t mul2add_mod_reduced(uint64_t a, uint64_t b, uint64_t c, uint64_t d, const Divisor& dv) {
    const uint64_t lo1 = a * b, lo2 = c * d;
    const uint64_t lo = lo1 + lo2;
    const uint64_t hi = mulhi64(a, b) + mulhi64(c, d) + ((lo < lo1) ? 1 : 0);
    const uint64_t u1 = dv.s ? ((hi << dv.s) | (lo >> (64 - dv.s))) : hi;
    const uint64_t u0 = lo << dv.s;
    return rem_2by1(u1, u0, dv.d, dv.v) >> dv.s;
}

}  // namespace hb
