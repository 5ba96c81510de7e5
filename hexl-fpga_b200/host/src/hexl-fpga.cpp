// hexl-fpga.cpp -- the C++ drop-in: intel::hexl::* and intel::hexl::fpga::*
// (reference host/src/hexl-fpga.cpp:18-94 and host/src/{dyadic_multiply,
// keyswitch,ntt,intt,fpga_context}.cpp) forwarded to the C ABI of
// libhexl_b200.so.  The reference API has no error channel: on failure we
// report and abort(), which is what its FPGA_ASSERT does (fpga_assert.h:19-38).
#include "hexl-fpga.h"

#include <cstdio>
#include <cstdlib>

#include "hexl_b200.h"

namespace {
void check(int rc, const char* what) {
    if (rc == 0) return;
    std::fprintf(stderr, "hexl-fpga (B200): %s failed (%d): %s\n", what, rc, hexl_b200_last_error());
    std::abort();
}
}  // namespace

namespace intel {
namespace hexl {
namespace fpga {

void acquire_FPGA_resources() { check(hexl_b200_host_acquire(), "acquire_FPGA_resources"); }
void release_FPGA_resources() { check(hexl_b200_host_release(), "release_FPGA_resources"); }

void set_worksize_DyadicMultiply(uint64_t ws) {
    check(hexl_b200_host_set_worksize_dyadic_multiply(ws), "set_worksize_DyadicMultiply");
}
void DyadicMultiply(uint64_t* results, const uint64_t* operand1, const uint64_t* operand2,
                    uint64_t n, const uint64_t* moduli, uint64_t n_moduli) {
    check(hexl_b200_host_dyadic_multiply(results, operand1, operand2, n, moduli, n_moduli),
          "DyadicMultiply");
}
bool DyadicMultiplyCompleted() {
    check(hexl_b200_host_dyadic_multiply_completed(), "DyadicMultiplyCompleted");
    return true;
}

void set_worksize_KeySwitch(uint64_t ws) {
    check(hexl_b200_host_set_worksize_keyswitch(ws), "set_worksize_KeySwitch");
}
void KeySwitch(uint64_t* result, const uint64_t* t_target_iter_ptr, uint64_t n,
               uint64_t decomp_modulus_size, uint64_t key_modulus_size,
               uint64_t rns_modulus_size, uint64_t key_component_count, const uint64_t* moduli,
               const uint64_t** k_switch_keys, const uint64_t* modswitch_factors,
               const uint64_t* twiddle_factors) {
    check(hexl_b200_host_keyswitch(result, t_target_iter_ptr, n, decomp_modulus_size,
                                   key_modulus_size, rns_modulus_size, key_component_count, moduli,
                                   k_switch_keys, modswitch_factors, twiddle_factors),
          "KeySwitch");
}
bool KeySwitchCompleted() {
    check(hexl_b200_host_keyswitch_completed(), "KeySwitchCompleted");
    return true;
}

void set_worksize_NTT(uint64_t ws) { check(hexl_b200_host_set_worksize_ntt(ws), "set_worksize_NTT"); }
void NTT(uint64_t* operand, const uint64_t* root_of_unity_powers,
         const uint64_t* precon_root_of_unity_powers, uint64_t coeff_modulus, uint64_t n) {
    check(hexl_b200_host_ntt(operand, root_of_unity_powers, precon_root_of_unity_powers,
                             coeff_modulus, n),
          "NTT");
}
bool NTTCompleted() {
    check(hexl_b200_host_ntt_completed(), "NTTCompleted");
    return true;
}

void set_worksize_INTT(uint64_t ws) { check(hexl_b200_host_set_worksize_intt(ws), "set_worksize_INTT"); }
void INTT(uint64_t* operand, const uint64_t* inv_root_of_unity_powers,
          const uint64_t* precon_inv_root_of_unity_powers, uint64_t coeff_modulus, uint64_t inv_n,
          uint64_t inv_n_w, uint64_t n) {
    check(hexl_b200_host_intt(operand, inv_root_of_unity_powers, precon_inv_root_of_unity_powers,
                              coeff_modulus, inv_n, inv_n_w, n),
          "INTT");
}
bool INTTCompleted() {
    check(hexl_b200_host_intt_completed(), "INTTCompleted");
    return true;
}

}  // namespace fpga

// public wrappers, reference host/src/hexl-fpga.cpp:18-94
void acquire_FPGA_resources() { fpga::acquire_FPGA_resources(); }
void release_FPGA_resources() { fpga::release_FPGA_resources(); }
void set_worksize_DyadicMultiply(uint64_t ws) { fpga::set_worksize_DyadicMultiply(ws); }
void DyadicMultiply(uint64_t* results, const uint64_t* operand1, const uint64_t* operand2,
                    uint64_t n, const uint64_t* moduli, uint64_t n_moduli) {
    fpga::DyadicMultiply(results, operand1, operand2, n, moduli, n_moduli);
}
bool DyadicMultiplyCompleted() { return fpga::DyadicMultiplyCompleted(); }
void set_worksize_KeySwitch(uint64_t ws) { fpga::set_worksize_KeySwitch(ws); }
void KeySwitch(uint64_t* result, const uint64_t* t_target_iter_ptr, uint64_t n,
               uint64_t decomp_modulus_size, uint64_t key_modulus_size,
               uint64_t rns_modulus_size, uint64_t key_component_count, const uint64_t* moduli,
               const uint64_t** k_switch_keys, const uint64_t* modswitch_factors,
               const uint64_t* twiddle_factors) {
    fpga::KeySwitch(result, t_target_iter_ptr, n, decomp_modulus_size, key_modulus_size,
                    rns_modulus_size, key_component_count, moduli, k_switch_keys,
                    modswitch_factors, twiddle_factors);
}
bool KeySwitchCompleted() { return fpga::KeySwitchCompleted(); }
void _set_worksize_NTT(uint64_t ws) { fpga::set_worksize_NTT(ws); }
void _NTT(uint64_t* operand, const uint64_t* root_of_unity_powers,
          const uint64_t* precon_root_of_unity_powers, uint64_t coeff_modulus, uint64_t n) {
    fpga::NTT(operand, root_of_unity_powers, precon_root_of_unity_powers, coeff_modulus, n);
}
bool _NTTCompleted() { return fpga::NTTCompleted(); }
void _set_worksize_INTT(uint64_t ws) { fpga::set_worksize_INTT(ws); }
void _INTT(uint64_t* operand, const uint64_t* inv_root_of_unity_powers,
           const uint64_t* precon_inv_root_of_unity_powers, uint64_t coeff_modulus, uint64_t inv_n,
           uint64_t inv_n_w, uint64_t n) {
    fpga::INTT(operand, inv_root_of_unity_powers, precon_inv_root_of_unity_powers, coeff_modulus,
               inv_n, inv_n_w, n);
}
bool _INTTCompleted() { return fpga::INTTCompleted(); }

}  // namespace hexl
}  // namespace intel
