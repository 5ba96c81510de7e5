// oracle/sycl_shim/ac_int_shim.hpp -- TEST INFRASTRUCTURE ONLY.
//
// Unsigned arbitrary-width integers with the bit-growth rules of Intel's ac_int, as far as the
// reference's device code relies on them (device/mod_ops.hpp:17-129, device/keyswitch/dyadmult.hpp:
// 37-60, the 512-bit lane shifts of ntt_core.hpp / intt_core.hpp):
//   a * b  -> W1 + W2 bits     a + b, a - b -> max(W1, W2) + 1 bits (two's complement wrap)
//   a << n, a >> n -> W1 bits  a | b, a & b -> max(W1, W2) bits     a / b -> W1 bits
// and assignment / construction truncates to the destination width.  Storage is an array of 64-bit
// limbs, little endian, and nothing else, so that the reference's reinterpret casts of uint64 arrays
// and packed key structs to ac_int<256/512> see the layout the FPGA tools give them.
#pragma once
#include <cstdint>
#include <type_traits>

template <int W, bool S = false>
class ac_int {
    static_assert(!S, "only unsigned ac_int is modelled");
    static_assert(W >= 1, "width");

public:
    static constexpr int width = W;
    static constexpr int L = (W + 63) / 64;
    uint64_t limb[L];

    ac_int() {
        for (int i = 0; i < L; ++i) limb[i] = 0;
    }
    template <class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
    ac_int(I x) {
        const uint64_t ext = (std::is_signed<I>::value && x < 0) ? ~(uint64_t)0 : 0;
        limb[0] = (uint64_t)x;
        for (int i = 1; i < L; ++i) limb[i] = ext;
        trim();
    }
    ac_int(unsigned __int128 x) {
        limb[0] = (uint64_t)x;
        if (L > 1) limb[1 % L] = (uint64_t)(x >> 64);
        for (int i = 2; i < L; ++i) limb[i] = 0;
        trim();
    }
    template <int W2>
    ac_int(const ac_int<W2, false>& o) {
        for (int i = 0; i < L; ++i) limb[i] = i < ac_int<W2, false>::L ? o.limb[i] : 0;
        trim();
    }
    void trim() {
        if (W % 64) limb[L - 1] &= (((uint64_t)1 << (W % 64)) - 1);
    }
    uint64_t to_uint64() const { return limb[0]; }
    unsigned to_uint() const { return (unsigned)limb[0]; }
    bool is_zero() const {
        for (int i = 0; i < L; ++i)
            if (limb[i]) return false;
        return true;
    }

    template <int W2>
    ac_int& operator-=(const ac_int<W2, false>& o) {
        *this = ac_int(*this - o);
        return *this;
    }
    template <int W2>
    ac_int& operator+=(const ac_int<W2, false>& o) {
        *this = ac_int(*this + o);
        return *this;
    }
    ac_int& operator-=(uint64_t o) { return *this -= ac_int<64, false>(o); }
    ac_int& operator+=(uint64_t o) { return *this += ac_int<64, false>(o); }
};

template <int A, int B>
ac_int<A + B, false> operator*(const ac_int<A, false>& a, const ac_int<B, false>& b) {
    ac_int<A + B, false> r;
    constexpr int LR = ac_int<A + B, false>::L;
    for (int i = 0; i < ac_int<A, false>::L; ++i) {
        unsigned __int128 carry = 0;
        for (int j = 0; j < ac_int<B, false>::L && i + j < LR; ++j) {
            const unsigned __int128 t = (unsigned __int128)a.limb[i] * b.limb[j] + r.limb[i + j] + carry;
            r.limb[i + j] = (uint64_t)t;
            carry = t >> 64;
        }
        for (int k = i + ac_int<B, false>::L; carry && k < LR; ++k) {
            const unsigned __int128 t = (unsigned __int128)r.limb[k] + carry;
            r.limb[k] = (uint64_t)t;
            carry = t >> 64;
        }
    }
    r.trim();
    return r;
}
template <int A, int B>
ac_int<(A > B ? A : B) + 1, false> operator+(const ac_int<A, false>& a, const ac_int<B, false>& b) {
    ac_int<(A > B ? A : B) + 1, false> r;
    unsigned __int128 carry = 0;
    for (int i = 0; i < r.L; ++i) {
        const uint64_t x = i < ac_int<A, false>::L ? a.limb[i] : 0, y = i < ac_int<B, false>::L ? b.limb[i] : 0;
        const unsigned __int128 t = (unsigned __int128)x + y + carry;
        r.limb[i] = (uint64_t)t;
        carry = t >> 64;
    }
    r.trim();
    return r;
}
template <int A, int B>
ac_int<(A > B ? A : B) + 1, false> operator-(const ac_int<A, false>& a, const ac_int<B, false>& b) {
    ac_int<(A > B ? A : B) + 1, false> r;
    uint64_t borrow = 0;
    for (int i = 0; i < r.L; ++i) {
        const uint64_t x = i < ac_int<A, false>::L ? a.limb[i] : 0, y = i < ac_int<B, false>::L ? b.limb[i] : 0;
        const uint64_t d = x - y - borrow;
        borrow = (x < y) || (x == y && borrow) ? 1 : 0;
        r.limb[i] = d;
    }
    r.trim();
    return r;
}
template <int A>
ac_int<A, false> operator>>(const ac_int<A, false>& a, unsigned n) {
    ac_int<A, false> r;
    constexpr int L = ac_int<A, false>::L;
    const unsigned ws = n / 64, bs = n % 64;
    for (int i = 0; i < L; ++i) {
        const unsigned s = i + ws;
        uint64_t v = s < (unsigned)L ? a.limb[s] >> bs : 0;
        if (bs && s + 1 < (unsigned)L) v |= a.limb[s + 1] << (64 - bs);
        r.limb[i] = v;
    }
    return r;
}
template <int A>
ac_int<A, false> operator<<(const ac_int<A, false>& a, unsigned n) {
    ac_int<A, false> r;
    constexpr int L = ac_int<A, false>::L;
    const unsigned ws = n / 64, bs = n % 64;
    for (int i = L - 1; i >= 0; --i) {
        uint64_t v = 0;
        if ((unsigned)i >= ws) {
            v = a.limb[i - ws] << bs;
            if (bs && (unsigned)i >= ws + 1) v |= a.limb[i - ws - 1] >> (64 - bs);
        }
        r.limb[i] = v;
    }
    r.trim();
    return r;
}
template <int A, class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
ac_int<A, false> operator>>(const ac_int<A, false>& a, I n) { return a >> (unsigned)n; }
template <int A, class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
ac_int<A, false> operator<<(const ac_int<A, false>& a, I n) { return a << (unsigned)n; }
template <int A, int B>
ac_int<A, false> operator>>(const ac_int<A, false>& a, const ac_int<B, false>& n) { return a >> (unsigned)n.limb[0]; }
template <int A, int B>
ac_int<A, false> operator<<(const ac_int<A, false>& a, const ac_int<B, false>& n) { return a << (unsigned)n.limb[0]; }

template <int A, int B>
ac_int<(A > B ? A : B), false> operator|(const ac_int<A, false>& a, const ac_int<B, false>& b) {
    ac_int<(A > B ? A : B), false> r;
    for (int i = 0; i < r.L; ++i)
        r.limb[i] = (i < ac_int<A, false>::L ? a.limb[i] : 0) | (i < ac_int<B, false>::L ? b.limb[i] : 0);
    return r;
}
template <int A, int B>
ac_int<(A > B ? A : B), false> operator&(const ac_int<A, false>& a, const ac_int<B, false>& b) {
    ac_int<(A > B ? A : B), false> r;
    for (int i = 0; i < r.L; ++i)
        r.limb[i] = (i < ac_int<A, false>::L ? a.limb[i] : 0) & (i < ac_int<B, false>::L ? b.limb[i] : 0);
    return r;
}
template <int A, int B>
int ac_cmp(const ac_int<A, false>& a, const ac_int<B, false>& b) {
    constexpr int LA = ac_int<A, false>::L, LB = ac_int<B, false>::L, LM = LA > LB ? LA : LB;
    for (int i = LM - 1; i >= 0; --i) {
        const uint64_t x = i < LA ? a.limb[i] : 0, y = i < LB ? b.limb[i] : 0;
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}
// schoolbook binary long division (only MultiplyUIntModLazy3 divides, mod_ops.hpp:135-151)
template <int A, int B>
ac_int<A, false> operator/(const ac_int<A, false>& a, const ac_int<B, false>& b) {
    ac_int<A, false> q;
    ac_int<A + 1, false> rem;
    for (int bit = A - 1; bit >= 0; --bit) {
        rem = ac_int<A + 1, false>(rem << 1u);
        rem.limb[0] |= (a.limb[bit / 64] >> (bit % 64)) & 1u;
        if (ac_cmp(rem, b) >= 0) {
            rem = ac_int<A + 1, false>(rem - b);
            q.limb[bit / 64] |= (uint64_t)1 << (bit % 64);
        }
    }
    return q;
}

#define AC_SHIM_CMP(op)                                                                                     \
    template <int A, int B>                                                                                 \
    bool operator op(const ac_int<A, false>& a, const ac_int<B, false>& b) { return ac_cmp(a, b) op 0; }    \
    template <int A, class I, class = typename std::enable_if<std::is_integral<I>::value>::type>            \
    bool operator op(const ac_int<A, false>& a, I b) { return ac_cmp(a, ac_int<64, false>((uint64_t)b)) op 0; } \
    template <int A, class I, class = typename std::enable_if<std::is_integral<I>::value>::type>            \
    bool operator op(I a, const ac_int<A, false>& b) { return ac_cmp(ac_int<64, false>((uint64_t)a), b) op 0; }
AC_SHIM_CMP(<)
AC_SHIM_CMP(<=)
AC_SHIM_CMP(>)
AC_SHIM_CMP(>=)
AC_SHIM_CMP(==)
AC_SHIM_CMP(!=)
#undef AC_SHIM_CMP

// mixed operands: a built-in integer behaves as a 64-bit ac_int
#define AC_SHIM_MIXED(op)                                                                                   \
    template <int A, class I, class = typename std::enable_if<std::is_integral<I>::value>::type>            \
    auto operator op(const ac_int<A, false>& a, I b) { return a op ac_int<64, false>((uint64_t)b); }        \
    template <int A, class I, class = typename std::enable_if<std::is_integral<I>::value>::type>            \
    auto operator op(I a, const ac_int<A, false>& b) { return ac_int<64, false>((uint64_t)a) op b; }
AC_SHIM_MIXED(*)
AC_SHIM_MIXED(+)
AC_SHIM_MIXED(-)
AC_SHIM_MIXED(|)
AC_SHIM_MIXED(&)
AC_SHIM_MIXED(/)
#undef AC_SHIM_MIXED
