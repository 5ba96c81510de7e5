// number_theory.h -- host-side number theory of the product path: what the
// keyswitch plan needs when the caller passes no twiddle tables (reference:
// host/src/number_theory_util.cpp:12-154, host/src/twiddle-factors.cpp:16-62,
// host/src/fpga.cpp:1039-1109).  Independent of oracle/ (which is test-only).
#pragma once
#include <cstdint>
#include <vector>

namespace hexl_b200 {
namespace nt {

uint64_t mul_mod(uint64_t a, uint64_t b, uint64_t q);
uint64_t pow_mod(uint64_t b, uint64_t e, uint64_t q);
// modular inverse, 0 if gcd(a,q) != 1
uint64_t inv_mod(uint64_t a, uint64_t q);
// floor(x * 2^64 / q), x < q  (the 64-bit Shoup factor)
uint64_t shoup(uint64_t x, uint64_t q);
// floor(2^64 / q)
uint64_t barrett_mu(uint64_t q);
bool is_primitive_root(uint64_t r, uint64_t degree, uint64_t q);
// smallest primitive degree-th root of unity mod prime q (degree a power of
// two dividing q-1); 0 if none.  Same value as the reference's
// MinimalPrimitiveRoot (number_theory_util.cpp:95-116) whatever generator is
// found first.
bool is_prime(uint64_t q);
uint64_t min_primitive_root(uint64_t degree, uint64_t q);

struct Tables {
    // hexl layouts: roots[bitrev(i)] = w^i; inv_roots 1-based stage order.
    std::vector<uint64_t> roots, precon, inv_roots, precon_inv;
    uint64_t q = 0, root = 0, inv_n = 0, inv_n_w = 0;
};
// tables for x^n + 1 over Z_q from the minimal primitive 2n-th root
Tables make_tables(uint64_t n, uint64_t q);
// tables from a caller-supplied block in the keyswitch 4-table format
// [inv_roots(0-based) | precon_inv | roots | precon_roots], each n words
// (host/src/fpga.cpp:1102-1109, tests/test_keyswitch.cpp:73-90)
Tables tables_from_keyswitch_block(uint64_t n, uint64_t q, const uint64_t* block4n);

}  // namespace nt
}  // namespace hexl_b200
