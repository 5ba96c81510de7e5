// internal.h -- shared between capi.cu and host/src/runtime.cpp (not public).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "ntt_block.cuh"

namespace hexl_b200 {
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
extern std::atomic<uint64_t> g_launches, g_h2d, g_d2h;
hb::ModTab make_modtab(uint64_t q, uint64_t inv_n, uint64_t inv_n_w, const hb::TwPair* ftw,
                       const hb::TwPair* itw, int logn, const hb::Tw32* ftw32 = nullptr,
                       const hb::Tw32* itw32 = nullptr);
}  // namespace hexl_b200
