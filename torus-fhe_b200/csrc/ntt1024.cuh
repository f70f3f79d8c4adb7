// ntt1024.cuh -- per-thread passes of the warp-level 1024-point negacyclic NTT
// (one warp per polynomial, 32 coefficients per thread in registers; see ntt32.cuh
// for the decomposition).  Shared by the CUDA kernels and tests/host_emu.
//
// Register/lane layouts (lane = threadIdx.x & 31):
//   coefficient layout : thread lane holds a[32*i1 + lane] in x[i1]
//   NTT-domain layout  : thread lane holds A[brev5(lane) + 32*brev5(r)] in x[r]
// Between pass 1 and pass 2 the warp transposes its 32x32 tile through a padded
// (row stride 33) scratch tile:
//   forward : store x[r] -> tile[r*33 + lane];  load x[j] = tile[lane*33 + j]
//   inverse : store x[j] -> tile[lane*33 + j];  load x[r] = tile[r*33 + lane]
#pragma once
#include "ntt32.cuh"

namespace ntt {

constexpr int N = 1024;
constexpr int TILE_STRIDE = 33;               // u64 elements per padded tile row
constexpr int TILE_ELEMS = 32 * TILE_STRIDE;  // 1056 u64 = 8448 B

MK_HD u64 ld_tab(const u64* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const unsigned long long*>(p));
#else
    return *p;
#endif
}

// tw_fwd[r*32 + i0] = psi^(i0 * (2*brev5(r) + 1))
MK_HD void fwd_pass1(u64 (&x)[32], const u64* tw_fwd, int lane) {
    twist32(x);
    dif32(x);
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = gl::mul(x[r], ld_tab(tw_fwd + r * 32 + lane));
}
MK_HD void fwd_pass2(u64 (&x)[32]) { dif32(x); }

// tw_inv[j*32 + lane] = psi^(-j * (2*brev5(lane) + 1)) / 1024
MK_HD void inv_pass1(u64 (&x)[32], const u64* tw_inv, int lane) {
    dit32_inv(x);
#pragma unroll
    for (int j = 0; j < 32; j++) x[j] = gl::mul(x[j], ld_tab(tw_inv + j * 32 + lane));
}
MK_HD void inv_pass2(u64 (&x)[32]) {
    dit32_inv(x);
    untwist32(x);
}

}  // namespace ntt
