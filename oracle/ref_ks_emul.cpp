// oracle/ref_ks_emul.cpp -- TEST INFRASTRUCTURE ONLY: the reference's OWN keyswitch device code on the CPU.
//
// Compiles, unmodified and where they lie under /root/reference,
//     device/keyswitch.cpp  (+ device/keyswitch/*.hpp, device/mod_ops.hpp, common/types.hpp)
//     host/src/twiddle-factors.cpp, host/src/number_theory_util.cpp
// against oracle/sycl_shim (threads + FIFOs standing in for the oneAPI FPGA emulator the reference
// uses for RUN_CHOICE=1) and drives them the way the reference host does.  This file restates only
// the HOST glue of host/src/fpga.cpp that cannot be compiled here (it needs the SYCL runtime,
// dlopen'd bitstreams and USM):
//     build_modulus_meta / build_invn_meta      fpga.cpp:1039-1089
//     KeySwitch_load_twiddles                   fpga.cpp:1091-1123
//     KeySwitch_load_keys (3 x 256-bit packing) fpga.cpp:1167-1248 + host/inc/fpga.h:38-68
//     enqueue_input_data_KeySwitch              fpga.cpp:1250-1319
//     process_output_KeySwitch / read_output    fpga.cpp:1517-1572
//     FPGAObject_KeySwitch::fill_out_data       fpga.cpp:441-475  (host-side accumulate into result)
// All modular arithmetic of the keyswitch itself is executed by the reference's kernels.
//
// Usage:  ks_ref_emul <problem.bin> <result.bin>
//   problem.bin: u64 header {n, D, K, batch}, moduli[K], modswitch_factors[K], keys[D][2*K*n],
//                t_target[batch][D*n], result[batch][2*D*n]          (all u64, little endian)
//   result.bin : result[batch][2*D*n] after the accumulate
// Limits are the bitstream's: K == 7 (special prime on engine 6), D <= 6, n in {1024..16384}
// (device/keyswitch/params.hpp:33-35, load.hpp:80-84, dyadmult.hpp:144-146).
// The process ends with _Exit: the autorun kernels never return.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "number_theory_util.h"   // reference host/inc

namespace intel { namespace hexl { namespace fpga {
void ComputeRootOfUnityPowers(uint64_t m_q, uint64_t m_degree, uint64_t m_degree_bits, uint64_t m_w,
                              uint64_t* inv_root_of_unity_powers, uint64_t* precon64_inv_root_of_unity_powers,
                              uint64_t* root_of_unity_powers, uint64_t* precon64_root_of_unity_powers);
}}}

#include "device/keyswitch.cpp"   // the reference's device translation unit (-I/root/reference)

using namespace intel::hexl::fpga;

static uint64_t precompute_modulus_k(uint64_t modulus) {   // fpga.cpp:1039-1047
    uint64_t k = 0;
    for (uint64_t i = 64; i > 0; i--)
        if ((1UL << i) >= modulus) k = i;
    return k;
}

static std::vector<uint64_t> read_all(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) {
        perror(path);
        exit(2);
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint64_t> v(sz / 8);
    if (fread(v.data(), 8, v.size(), f) != v.size()) exit(2);
    fclose(f);
    return v;
}

// set `bits` bits of a 256-bit little-endian word at bit offset `pos`
static void put_bits(uint64_t* w, unsigned pos, unsigned bits, uint64_t val) {
    for (unsigned b = 0; b < bits; ++b)
        if ((val >> b) & 1) w[(pos + b) / 64] |= (uint64_t)1 << ((pos + b) % 64);
}

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s problem.bin result.bin\n", argv[0]);
        return 2;
    }
    std::vector<uint64_t> in = read_all(argv[1]);
    const uint64_t n = in[0], D = in[1], K = in[2], batch = in[3];
    if (K != MAX_KEY_MODULUS_SIZE || D < 1 || D > MAX_DECOMP_MODULUS_SIZE || n < 1024 || n > MAX_COFF_COUNT ||
        (n & (n - 1)) || in.size() != 4 + 2 * K + D * 2 * K * n + batch * 3 * D * n) {
        fprintf(stderr, "unsupported shape n=%lu D=%lu K=%lu batch=%lu\n", n, D, K, batch);
        return 3;
    }
    const uint64_t* moduli = &in[4];
    const uint64_t* msf = moduli + K;
    const uint64_t* keys = msf + K;                 // [D][2*K*n]
    const uint64_t* t_target = keys + D * 2 * K * n;
    std::vector<uint64_t> result(t_target + batch * D * n, t_target + batch * D * n + batch * 2 * D * n);

    // ---- KeySwitch_load_twiddles, caller passes no table (fpga.cpp:1097-1109) ----
    std::vector<uint64_t> tw(MAX_KEY_MODULUS_SIZE * 4 * n, 0);
    for (uint64_t i = 0; i < K; ++i) {
        const uint64_t w = MinimalPrimitiveRoot(2 * n, moduli[i]);
        ComputeRootOfUnityPowers(moduli[i], n, Log2(n), w, &tw[i * n * 4], &tw[i * n * 4 + n], &tw[i * n * 4 + 2 * n],
                                 &tw[i * n * 4 + 3 * n]);
    }
    // ---- build_modulus_meta / build_invn_meta (fpga.cpp:1049-1089) ----
    moduli_t modulus_meta;
    invn_t invn;
    for (uint64_t i = 0; i < K; ++i) {
        sycl::ulong4 m;
        m.s0() = moduli[i];
        m.s1() = MultiplyFactor(1, 64, moduli[i]).BarrettFactor();
        uint64_t modulus = moduli[i], twice = 2 * modulus, four = 4 * modulus;
        m.s2() = ReduceMod<8>(msf[i], modulus, &twice, &four);
        const uint64_t k = precompute_modulus_k(moduli[i]);
        __int128 a = 1;
        const uint64_t r = (a << (2 * k)) / moduli[i];
        m.s3() = (r << 8) | k;
        modulus_meta.data[i] = m;
        sycl::ulong4 v;
        const uint64_t inv_n = InverseUIntMod(n, moduli[i]);
        const uint64_t W_op = tw[i * n * 4 + n - 1];
        const uint64_t inv_nw = MultiplyUIntMod(inv_n, W_op, moduli[i]);
        v.s0() = inv_n;
        v.s1() = (r << 8) | k;
        v.s2() = DivideUInt128UInt64Lo(inv_n, 0, moduli[i]);
        v.s3() = DivideUInt128UInt64Lo(inv_nw, 0, moduli[i]);
        invn.data[i] = v;
    }
    // ---- KeySwitch_load_keys: 14 x 52-bit values per (digit, coefficient) in 3 x 256 bits ----
    // (fpga.cpp:1182-1225 with the bit-field structs of host/inc/fpga.h:38-68; the unpacking is the
    //  reference's own, device/keyswitch/dyadmult.hpp:37-60)
    std::vector<uint256_t> kv1(D * n), kv2(D * n), kv3(D * n);
    for (uint64_t k = 0; k < D; ++k)
        for (uint64_t j = 0; j < n; ++j) {
            uint64_t w1[4] = {0, 0, 0, 0}, w2[4] = {0, 0, 0, 0}, w3[4] = {0, 0, 0, 0};
            for (uint64_t i = 0; i < K; ++i) {
                const uint64_t key1 = keys[k * 2 * K * n + i * n + j];
                const uint64_t key2 = keys[k * 2 * K * n + (i + K) * n + j];
                switch (i) {
                    case 0: put_bits(w1, 0, 52, key1); put_bits(w1, 52, 52, key2); break;
                    case 1: put_bits(w1, 104, 52, key1); put_bits(w1, 156, 52, key2); break;
                    case 2:
                        put_bits(w1, 208, 48, key1 & BIT_MASK(48));
                        put_bits(w2, 0, 4, (key1 >> 48) & BIT_MASK(4));
                        put_bits(w2, 4, 52, key2);
                        break;
                    case 3: put_bits(w2, 56, 52, key1); put_bits(w2, 108, 52, key2); break;
                    case 4:
                        put_bits(w2, 160, 52, key1);
                        put_bits(w2, 212, 44, key2 & BIT_MASK(44));
                        put_bits(w3, 0, 8, (key2 >> 44) & BIT_MASK(8));
                        break;
                    case 5: put_bits(w3, 8, 52, key1); put_bits(w3, 60, 52, key2); break;
                    case 6: put_bits(w3, 112, 52, key1); put_bits(w3, 164, 52, key2); break;
                }
            }
            memcpy(&kv1[k * n + j], w1, 32);
            memcpy(&kv2[k * n + j], w2, 32);
            memcpy(&kv3[k * n + j], w3, 32);
        }

    // ---- enqueue (fpga.cpp:684-688, 1119-1121, 1273-1301) and store (fpga.cpp:1558-1565) ----
    sycl::queue q_load, q_store;
    launchAllAutoRunKernels(q_load);
    sycl::buffer<uint64_t> b_tw(tw.data(), tw.size());
    launchConfigurableKernels(q_load, &b_tw, (unsigned)n, true);
    sycl::buffer<uint256_t> b_k1(kv1.data(), kv1.size()), b_k2(kv2.data(), kv2.size()), b_k3(kv3.data(), kv3.size());
    launchStoreSwitchKeys(q_load, b_k1, b_k2, b_k3, (int)batch);
    std::vector<uint64_t> tt(t_target, t_target + batch * D * n);
    sycl::buffer<uint64_t> b_t(tt.data(), tt.size());
    std::vector<sycl::ulong2> out(batch * D * n);
    sycl::buffer<sycl::ulong2> b_out(out.data(), out.size());
    sycl::event e_load = load(q_load, nullptr, b_t, modulus_meta, n, D, batch, invn, 1);
    sycl::event e_store = store(q_store, nullptr, b_out, batch, n, D, modulus_meta, 1, 1);
    e_load.wait();
    e_store.wait();

    // ---- FPGAObject_KeySwitch::fill_out_data: accumulate into result (fpga.cpp:441-475) ----
    for (uint64_t b = 0; b < batch; ++b) {
        const size_t size_out = b * D * n * 2;
        uint64_t* res = &result[b * 2 * D * n];
        const uint64_t* output = reinterpret_cast<const uint64_t*>(out.data());
        for (size_t i = 0; i < D; ++i) {
            const uint64_t modulus = moduli[i];
            for (size_t j = 0; j < n; ++j) {
                const size_t k = i * n + j;
                res[k] += output[size_out + 2 * k];
                res[k] = (res[k] >= modulus) ? res[k] - modulus : res[k];
                res[k + n * D] += output[size_out + 2 * k + 1];
                res[k + n * D] = (res[k + n * D] >= modulus) ? res[k + n * D] - modulus : res[k + n * D];
            }
        }
    }
    FILE* f = fopen(argv[2], "wb");
    if (!f || fwrite(result.data(), 8, result.size(), f) != result.size()) {
        perror(argv[2]);
        _Exit(2);
    }
    fclose(f);
    fflush(stdout);
    _Exit(0);   // the autorun kernel threads never return
}
