// oracle/ref_dev_emul.cpp -- TEST INFRASTRUCTURE ONLY: the reference's OWN forward-NTT, inverse-NTT and
// dyadic-multiply DEVICE code on the CPU (cf. ref_ks_emul.cpp for the keyswitch).
//
// One of the reference's device translation units is compiled UNMODIFIED where it lies, against
// oracle/sycl_shim, selected at build time (they define clashing macros and non-inline helpers, so
// each gets its own binary):
//     -DEMUL_FWD_NTT   device/fwd_ntt.cpp          -> oracle/_ref/dev_ref_emul_ntt
//     -DEMUL_INV_NTT   device/inv_ntt.cpp          -> oracle/_ref/dev_ref_emul_intt
//     -DEMUL_DYADIC    device/dyadic_multiply.cpp  -> oracle/_ref/dev_ref_emul_dyadic
// The host glue restated here is what host/src/fpga.cpp does around the launchers:
//     NTT    fill_in_data fpga.cpp:392-413,  enqueue :1014-1021, output :1406-1418
//     INTT   fill_in_data fpga.cpp:415-439,  enqueue :986-993,   output (intt_output)
//     dyadic fill_in_data fpga.cpp:355-390 (Barrett constants len, barr_lo), enqueue :959-966,
//            non-blocking output poll :1327-1404
//
// Usage: dev_ref_emul_X <in.bin> <out.bin>   (u64 little endian)
//   ntt   : {batch, q} roots[16384] precon[16384] data[batch][16384]            -> data
//   intt  : {batch, q, inv_n, inv_n_w} inv_roots[16384] precon_inv[16384] data   -> data
//   dyadic: {batch, n, M} moduli[batch][M] op1[batch][2*M*n] op2[batch][2*M*n]   -> res[batch][3*M*n]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#if defined(EMUL_FWD_NTT)
#include "device/fwd_ntt.cpp"
#elif defined(EMUL_INV_NTT)
#include "device/inv_ntt.cpp"
#elif defined(EMUL_DYADIC)
#include "device/dyadic_multiply.cpp"
#else
#error "pick one device translation unit"
#endif

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
static void write_all(const char* path, const uint64_t* p, size_t n) {
    FILE* f = fopen(path, "wb");
    if (!f || fwrite(p, 8, n, f) != n) {
        perror(path);
        _Exit(2);
    }
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    std::vector<uint64_t> in = read_all(argv[1]);
    sycl::queue q_in, q_out, q_auto;
#if defined(EMUL_FWD_NTT)
    const uint64_t batch = in[0];
    uint64_t modulus = in[1];
    const size_t N = FPGA_NTT_SIZE;
    if (in.size() != 2 + 2 * N + batch * N) return 3;
    uint64_t* roots = &in[2];
    uint64_t* precon = roots + N;
    uint64_t* data = precon + N;
    std::vector<uint64_t> out(batch * N);
    fwd_ntt(q_auto);
    sycl::event e0 = ntt_input(q_in, (unsigned)batch, data, data, &modulus, roots, precon);   // fpga.cpp:1014-1021
    sycl::event e1 = ntt_output(q_out, (int)batch, out.data());
    e0.wait();
    e1.wait();
    write_all(argv[2], out.data(), out.size());
#elif defined(EMUL_INV_NTT)
    const uint64_t batch = in[0];
    uint64_t modulus = in[1], inv_n = in[2], inv_n_w = in[3];
    const size_t N = FPGA_INTT_SIZE;
    if (in.size() != 4 + 2 * N + batch * N) return 3;
    uint64_t* roots = &in[4];
    uint64_t* precon = roots + N;
    uint64_t* data = precon + N;
    std::vector<uint64_t> out(batch * N);
    inv_ntt(q_auto);
    sycl::event e0 = intt_input(q_in, (unsigned)batch, data, &modulus, &inv_n, &inv_n_w, roots, precon);
    sycl::event e1 = intt_output(q_out, (unsigned)batch, out.data());
    e0.wait();
    e1.wait();
    write_all(argv[2], out.data(), out.size());
#else
    const uint64_t batch = in[0], n = in[1], M = in[2];
    if (in.size() != 3 + batch * M + 2 * batch * 2 * M * n) return 3;
    uint64_t* moduli = &in[3];
    uint64_t* op1 = moduli + batch * M;
    uint64_t* op2 = op1 + batch * 2 * M * n;
    // FPGAObject_DyadicMultiply::fill_in_data (fpga.cpp:366-373): per (item, modulus) Barrett constants
    std::vector<moduli_info_t> info(batch * M);
    for (uint64_t i = 0; i < batch * M; ++i) {
        const uint64_t modulus = moduli[i];
        const uint64_t len = uint64_t(floorl(log2l(modulus)) - 1);
        const unsigned __int128 nn = (unsigned __int128)1 << (len + 64);
        info[i] = (moduli_info_t){modulus, len, uint64_t(nn / modulus)};
    }
    std::vector<uint64_t> ddr_in(batch * M * n * 4), ddr_out(batch * M * n * 3), out(batch * M * n * 3);
    submit_autorun_kernels(q_auto);
    sycl::event e0 = input_fifo_usm(q_in, op1, op2, n, info.data(), M, 7, ddr_in.data(), ddr_out.data(), batch);
    e0.wait();
    int tag = -1, valid = 0;
    while (!valid) {   // fpga.cpp:1327-1404 polls the non-blocking output kernel until the tag comes back
        sycl::event e1 = output_nb_fifo_usm(q_out, out.data(), &tag, &valid);
        e1.wait();
    }
    if (tag != 7) return 4;
    write_all(argv[2], out.data(), out.size());
#endif
    fflush(stdout);
    _Exit(0);   // the autorun kernel threads never return
}
