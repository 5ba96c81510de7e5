#include "number_theory.h"

namespace hexl_b200 {
namespace nt {

using u128 = unsigned __int128;

uint64_t mul_mod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((u128)a * b % q); }

uint64_t pow_mod(uint64_t b, uint64_t e, uint64_t q) {
    uint64_t acc = 1 % q;
    for (b %= q; e; e >>= 1, b = mul_mod(b, b, q))
        if (e & 1) acc = mul_mod(acc, b, q);
    return acc;
}

uint64_t inv_mod(uint64_t a, uint64_t q) {
    // extended Euclid on (q, a mod q) tracking only the coefficient of a
    __int128 r0 = q, r1 = a % q, s0 = 0, s1 = 1;
    while (r1) {
        __int128 k = r0 / r1, t;
        t = r0 - k * r1, r0 = r1, r1 = t;
        t = s0 - k * s1, s0 = s1, s1 = t;
    }
    if (r0 != 1) return 0;
    return (uint64_t)(s0 < 0 ? s0 + q : s0);
}

uint64_t shoup(uint64_t x, uint64_t q) { return (uint64_t)(((u128)x << 64) / q); }

uint64_t barrett_mu(uint64_t q) { return (uint64_t)((((u128)1) << 64) / q); }

bool is_primitive_root(uint64_t r, uint64_t degree, uint64_t q) {
    return r && pow_mod(r, degree / 2, q) == q - 1;
}

// deterministic Miller-Rabin for 64-bit integers (the first twelve primes as bases)
bool is_prime(uint64_t q) {
    if (q < 2) return false;
    for (uint64_t p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (q == p) return true;
        if (q % p == 0) return false;
    }
    uint64_t d = q - 1;
    int s = 0;
    while (!(d & 1)) d >>= 1, ++s;
    for (uint64_t a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        uint64_t x = pow_mod(a, d, q);
        if (x == 1 || x == q - 1) continue;
        bool composite = true;
        for (int i = 1; i < s && composite; ++i) {
            x = mul_mod(x, x, q);
            if (x == q - 1) composite = false;
        }
        if (composite) return false;
    }
    return true;
}

uint64_t min_primitive_root(uint64_t degree, uint64_t q) {
    if (degree < 2 || (q - 1) % degree) return 0;
    // Z_q^* is cyclic only for a prime q: a composite modulus has no use here, and the search below
    // would not terminate in any reasonable time (half of all candidates work when q is prime)
    if (!is_prime(q)) return 0;
    const uint64_t cofactor = (q - 1) / degree;
    uint64_t g = 0;
    for (uint64_t c = 2; c < q && c < 4096 && !g; ++c) {
        uint64_t r = pow_mod(c, cofactor, q);
        if (is_primitive_root(r, degree, q)) g = r;
    }
    if (!g) return 0;
    // the primitive roots are exactly the odd powers of g
    const uint64_t g2 = mul_mod(g, g, q);
    uint64_t best = g, cur = g;
    for (uint64_t i = 1; i < degree / 2; ++i) {
        cur = mul_mod(cur, g2, q);
        if (cur < best) best = cur;
    }
    return best;
}

static uint32_t bitrev(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i, x >>= 1) r = (r << 1) | (x & 1);
    return r;
}

static void finish(Tables& t, uint64_t n) {
    const uint64_t q = t.q;
    t.precon.resize(n);
    t.precon_inv.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        t.precon[i] = shoup(t.roots[i], q);
        t.precon_inv[i] = shoup(t.inv_roots[i], q);
    }
    t.inv_n = inv_mod(n % q, q);
    t.inv_n_w = mul_mod(t.inv_n, t.inv_roots[n - 1], q);
}

Tables make_tables(uint64_t n, uint64_t q) {
    Tables t;
    t.q = q;
    t.root = min_primitive_root(2 * n, q);
    if (!t.root) return t;
    int bits = 0;
    while ((1ull << bits) < n) ++bits;
    const uint64_t winv = inv_mod(t.root, q);
    std::vector<uint64_t> inv_br(n);
    t.roots.assign(n, 0);
    uint64_t pw = 1, ipw = 1;
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t r = bitrev((uint32_t)i, bits);
        t.roots[r] = pw;
        inv_br[r] = ipw;
        pw = mul_mod(pw, t.root, q);
        ipw = mul_mod(ipw, winv, q);
    }
    // stage order: block sizes n/2, n/4, ..., 1 of the bit-reversed inverse powers
    t.inv_roots.assign(n, 0);
    t.inv_roots[0] = 1;
    uint64_t at = 1;
    for (uint64_t m = n >> 1; m; m >>= 1)
        for (uint64_t i = 0; i < m; ++i) t.inv_roots[at++] = inv_br[m + i];
    finish(t, n);
    return t;
}

Tables tables_from_keyswitch_block(uint64_t n, uint64_t q, const uint64_t* blk) {
    Tables t;
    t.q = q;
    t.root = 0;  // unknown; not needed
    t.roots.assign(blk + 2 * n, blk + 3 * n);
    t.roots[0] = 1;
    t.inv_roots.assign(n, 0);
    // Two layouts of the inverse table are in circulation: the FPGA's 0-based
    // one (host/src/twiddle-factors.cpp:46-55: first stage twiddle at [0],
    // [n-1] == 0) and intel-hexl's 1-based one ([0] == 1, what
    // GetInvRootOfUnityPowers returns).  [0] == 1 can only be the latter.
    if (blk[0] == 1 && blk[n - 1] != 0) {
        t.inv_roots.assign(blk, blk + n);
    } else {
        t.inv_roots[0] = 1;
        for (uint64_t i = 0; i + 1 < n; ++i) t.inv_roots[i + 1] = blk[i];
    }
    finish(t, n);
    return t;
}

}  // namespace nt
}  // namespace hexl_b200
