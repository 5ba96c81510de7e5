// hexl-fpga.h -- public C++ API of the B200 drop-in for intel/hexl-fpga.
//
// Same names, namespaces, argument order/meaning and asynchronous protocol as
// the reference's host/inc/hexl-fpga.h:15-161 (and its intel::hexl::fpga::*
// twin, host/inc/{dyadic_multiply,keyswitch,ntt,intt,fpga_context}.h), so code
// written against the reference compiles and links against lib/libhexl-fpga.so
// unchanged.  The functions forward to the C ABI in include/hexl_b200.h.
//
// Protocol (reference host/src/fpga_int.cpp:171-537):
//   acquire_FPGA_resources();                      // once
//   set_worksize_X(ws);  X(...) x ws;  XCompleted();   // async batch
//   X(...) with worksize 1 (default) is synchronous
//   release_FPGA_resources();                      // once
// Differences, all supersets: n may be any power of two in [1024,16384] for
// _NTT/_INTT/KeySwitch (reference: 16384 only for _NTT/_INTT); KeySwitch
// accepts key_modulus_size > 7 and moduli < 2^61; batch items need not be
// contiguous in host memory; failures print a message and abort() (the
// reference aborts only when built with FPGA_DEBUG, otherwise misbehaves).
#ifndef HEXL_FPGA_B200_HEXL_FPGA_H_
#define HEXL_FPGA_B200_HEXL_FPGA_H_

#include <cstdint>

namespace intel {
namespace hexl {

void acquire_FPGA_resources();
void release_FPGA_resources();

void set_worksize_DyadicMultiply(uint64_t ws);
void DyadicMultiply(uint64_t* results, const uint64_t* operand1, const uint64_t* operand2,
                    uint64_t n, const uint64_t* moduli, uint64_t n_moduli);
bool DyadicMultiplyCompleted();

void set_worksize_KeySwitch(uint64_t ws);
void KeySwitch(uint64_t* result, const uint64_t* t_target_iter_ptr, uint64_t n,
               uint64_t decomp_modulus_size, uint64_t key_modulus_size,
               uint64_t rns_modulus_size, uint64_t key_component_count, const uint64_t* moduli,
               const uint64_t** k_switch_keys, const uint64_t* modswitch_factors,
               const uint64_t* twiddle_factors = nullptr);
bool KeySwitchCompleted();

// deprecated in the reference since v1.1 but still exported (hexl-fpga.h:101-161)
void _set_worksize_NTT(uint64_t ws);
void _NTT(uint64_t* operand, const uint64_t* root_of_unity_powers,
          const uint64_t* precon_root_of_unity_powers, uint64_t coeff_modulus, uint64_t n);
bool _NTTCompleted();

void _set_worksize_INTT(uint64_t ws);
void _INTT(uint64_t* operand, const uint64_t* inv_root_of_unity_powers,
           const uint64_t* precon_inv_root_of_unity_powers, uint64_t coeff_modulus,
           uint64_t inv_n, uint64_t inv_n_w, uint64_t n);
bool _INTTCompleted();

namespace fpga {
void acquire_FPGA_resources();
void release_FPGA_resources();
void set_worksize_DyadicMultiply(uint64_t ws);
void DyadicMultiply(uint64_t* results, const uint64_t* operand1, const uint64_t* operand2,
                    uint64_t n, const uint64_t* moduli, uint64_t n_moduli);
bool DyadicMultiplyCompleted();
void set_worksize_KeySwitch(uint64_t ws);
void KeySwitch(uint64_t* result, const uint64_t* t_target_iter_ptr, uint64_t n,
               uint64_t decomp_modulus_size, uint64_t key_modulus_size,
               uint64_t rns_modulus_size, uint64_t key_component_count, const uint64_t* moduli,
               const uint64_t** k_switch_keys, const uint64_t* modswitch_factors,
               const uint64_t* twiddle_factors = nullptr);
bool KeySwitchCompleted();
void set_worksize_NTT(uint64_t ws);
void NTT(uint64_t* operand, const uint64_t* root_of_unity_powers,
         const uint64_t* precon_root_of_unity_powers, uint64_t coeff_modulus, uint64_t n);
bool NTTCompleted();
void set_worksize_INTT(uint64_t ws);
void INTT(uint64_t* operand, const uint64_t* inv_root_of_unity_powers,
          const uint64_t* precon_inv_root_of_unity_powers, uint64_t coeff_modulus, uint64_t inv_n,
          uint64_t inv_n_w, uint64_t n);
bool INTTCompleted();
}  // namespace fpga

}  // namespace hexl
}  // namespace intel
#endif
