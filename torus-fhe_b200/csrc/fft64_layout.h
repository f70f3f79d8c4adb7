// fft64_layout.h -- constants of the FP64 FFT channel shared by the kernels (fft64.cuh) and the host table generator (tables_fft.h).
#pragma once
#include <stddef.h>

#if defined(__CUDACC__)
#define MKF_HD __host__ __device__
#else
#define MKF_HD
#endif

namespace mkf {

constexpr int N = 1024;        // ring degree (real negacyclic polynomials mod X^N + 1)
constexpr int M = 512;         // complex points: C[X] / (X^M - i), a~[j] = a[j] + i a[j + M]
// key limbs: a Torus64 key word = l0 + l1 2^22 + l2 2^43 (three balanced limbs of 22 / 21 / 21 bits); a Torus32 key word (Torus32 mode:
// unshifted 32-bit keys, products added as R << 32) = l0 + l1 2^16 (two balanced limbs of 16 bits)
constexpr int LIMBS = 3, LIMBS_T32 = 2;
MKF_HD inline constexpr int limb_shift(int nl, int limb) { return nl == 3 ? (limb == 0 ? 0 : limb == 1 ? 22 : 43) : 16 * limb; }

// twiddle table, complex-double entries (staged in shared memory by the kernels)
constexpr int TF_A = 0;                 // forward stages d = 1..4 (warp-half-uniform): TF_A + 2^d - 2 + g, g = group = pos >> (9 - d); the kernels load
                                        // g = h 2^(d-1) and multiply by compile-time constants for the other groups of a half-warp
constexpr int TF_B = 32;                // forward stages d = 5..8 (per lane): TF_B + 32 (d - 5) + lane = s(d, lane 2^(d-5)); the other groups of a
                                        // lane are this value times a compile-time constant (fft64_core.cuh, fwd_passB)
constexpr int TI_A = TF_B + 128;        // inverse pass A', row span rs = 2^t, t = 0..3: TI_A + 16 t + l16 = exp(-2 pi i l16 / (32 rs)); the factor
                                        // exp(-2 pi i (r mod rs) / (2 rs)) of the other rows is a compile-time constant
constexpr int T_WJ = TI_A + 64;        // last inverse stage: exp(-2 pi i j / 512), j < 256
constexpr int T_UT = T_WJ + 256;        // untwist zeta^-j, j < 256 (zeta^-(j + 256) = zeta^-j exp(-i pi / 4))
constexpr int T_ENTRIES = T_UT + 256;   // 736 entries = 11 776 bytes
constexpr int TW_BYTES = T_ENTRIES * 16;

// bootstrapping key in the FFT layout: complex double [elem][s = src l + q][out 2][limb nl][512 points], point (c, lane) at c 32 + lane
// (c = register index, lane = thread of the transform's output layout); values forward(fold(limb)) / 512
MKF_HD inline constexpr size_t bsk_elem_cpx(int l, int nl = LIMBS) { return (size_t)2 * l * 2 * nl * M; }

}  // namespace mkf
