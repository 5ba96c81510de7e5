// tests/cpp/dropin_driver.cpp -- a CALLER of the reference API, compiled against the reference's OWN,
// unmodified header (host/inc/hexl-fpga.h, found with -I/root/reference/host/inc in the build container)
// and linked to our libhexl-fpga.so: the drop-in boundary EXECUTED, not just its symbol table.
//
// Each mode reproduces the call sequence of the matching reference test:
//   ntt      tests/test_fwd_ntt.cpp:97-117        _set_worksize_NTT / _NTT x batch / _NTTCompleted
//   intt     tests/test_inv_ntt.cpp:97-125        _set_worksize_INTT / _INTT x batch / _INTTCompleted
//   dyadic   tests/test_dyadic_multiply.cpp:88-109 set_worksize_DyadicMultiply / DyadicMultiply x batch / Completed
//   keyswitch tests/test_keyswitch.cpp:119-146    set_worksize_KeySwitch / KeySwitch x batch / KeySwitchCompleted
// bracketed by acquire_FPGA_resources / release_FPGA_resources as tests/fpga_context.h:8-14 does.
// Inputs and outputs are raw little-endian uint64 files written / checked by tests/test_gpu_cxx_dropin.py
// (the checker there is the oracle; this program contains no arithmetic).
//
//   dropin_driver <mode> <in.bin> <out.bin>
//   ntt      : {batch, n, q} roots[n] precon[n] data[batch][n]
//   intt     : {batch, n, q, inv_n, inv_n_w} inv_roots[n] precon_inv[n] data[batch][n]
//   dyadic   : {batch, n, M} moduli[batch][M] op1[batch][2Mn] op2[batch][2Mn]           -> res[batch][3Mn]
//   keyswitch: {batch, n, D, K} moduli[K] msf[K] keys[D][2Kn] t[batch][Dn] result[batch][2Dn] -> result
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "hexl-fpga.h"

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

int main(int argc, char** argv) {
    if (argc != 4) return 2;
    const std::string mode = argv[1];
    std::vector<uint64_t> in = read_all(argv[2]);
    std::vector<uint64_t> out;
    intel::hexl::acquire_FPGA_resources();
    if (mode == "ntt") {
        const uint64_t batch = in[0], n = in[1], q = in[2];
        const uint64_t* roots = &in[3];
        const uint64_t* precon = roots + n;
        out.assign(precon + n, precon + n + batch * n);
        intel::hexl::_set_worksize_NTT(batch);
        for (uint64_t b = 0; b < batch; ++b) intel::hexl::_NTT(&out[b * n], roots, precon, q, n);
        if (!intel::hexl::_NTTCompleted()) return 5;
    } else if (mode == "intt") {
        const uint64_t batch = in[0], n = in[1], q = in[2], inv_n = in[3], inv_n_w = in[4];
        const uint64_t* roots = &in[5];
        const uint64_t* precon = roots + n;
        out.assign(precon + n, precon + n + batch * n);
        intel::hexl::_set_worksize_INTT(batch);
        for (uint64_t b = 0; b < batch; ++b) intel::hexl::_INTT(&out[b * n], roots, precon, q, inv_n, inv_n_w, n);
        if (!intel::hexl::_INTTCompleted()) return 5;
    } else if (mode == "dyadic") {
        const uint64_t batch = in[0], n = in[1], M = in[2];
        const uint64_t* moduli = &in[3];
        const uint64_t* op1 = moduli + batch * M;
        const uint64_t* op2 = op1 + batch * 2 * M * n;
        out.assign(batch * 3 * M * n, 0);
        intel::hexl::set_worksize_DyadicMultiply(batch);
        for (uint64_t b = 0; b < batch; ++b)
            intel::hexl::DyadicMultiply(&out[b * 3 * M * n], op1 + b * 2 * M * n, op2 + b * 2 * M * n, n,
                                        moduli + b * M, M);
        if (!intel::hexl::DyadicMultiplyCompleted()) return 5;
    } else if (mode == "keyswitch") {
        const uint64_t batch = in[0], n = in[1], D = in[2], K = in[3];
        const uint64_t* moduli = &in[4];
        const uint64_t* msf = moduli + K;
        const uint64_t* keys = msf + K;
        const uint64_t* t = keys + D * 2 * K * n;
        const uint64_t* res = t + batch * D * n;
        std::vector<const uint64_t*> key_ptrs(D);
        for (uint64_t j = 0; j < D; ++j) key_ptrs[j] = keys + j * 2 * K * n;
        out.assign(res, res + batch * 2 * D * n);
        intel::hexl::set_worksize_KeySwitch(batch);
        for (uint64_t b = 0; b < batch; ++b)
            intel::hexl::KeySwitch(&out[b * 2 * D * n], t + b * D * n, n, D, K, D + 1, 2, moduli, key_ptrs.data(), msf);
        if (!intel::hexl::KeySwitchCompleted()) return 5;
    } else {
        return 2;
    }
    intel::hexl::release_FPGA_resources();
    FILE* f = fopen(argv[3], "wb");
    if (!f || fwrite(out.data(), 8, out.size(), f) != out.size()) return 6;
    fclose(f);
    return 0;
}
