// ntt_f64.cuh -- PROTOTYPE (round-2 candidate, not part of the shipped library): one RNS channel of the 1024-point negacyclic NTT
// carried in doubles, so that its butterflies run on the FP64 pipe of sm_100 instead of the fmaheavy (IMAD) pipe that binds the
// blind-rotate kernels (go / no-go: tools/fp64_pipe_ubench.cu).
//
// Same prime, same merged Cooley-Tukey / Gentleman-Sande index structure and the same 31-entry twiddle tables as the u32 channel
// (torus-fhe_b200/csrc/ntt_rns.cuh), so a gate could run primes 0 and 1 in u32 and prime 2 in doubles and feed the same Garner lift.
// Every value is an exact integer held in a double:
//   mulmod(y, w):  h = y*w (rounded), l = fma(y, w, -h) (the rounding error, exact), q = rint(h / p) by the 1.5*2^52 trick,
//                  t = fma(-q, p, h) + l  -- exact whenever |y*w| < 2^100 or so, and |t| <= p/2 + |y| 2^-24 (p < 2^28, |y| < 2^40)
//   butterflies:   no range correction at all.  Forward: the untouched operand grows by |t| < 0.51 p per stage (10 stages: < 6.1 p from
//                  |x| < p).  Inverse: sums double per stage (10 stages: < 2^10 * 0.51 p * 2l < 2^41), differences go through mulmod.
//   int <-> double without conversion instructions (the XU pipe is 16 lanes wide): the 2^52 + 2^31 bias trick, one DADD each way.
// 8 FP64 instructions per butterfly (DMUL, 3 DFMA, 4 DADD) against 4 fmaheavy slots + 3 alu slots for the u32 Harvey butterfly.
// All functions are __host__ __device__: tools/f64_channel/emu.cpp runs them on the CPU against the u32 channel and the O(N^2)
// definition (tests/test_host_emu.py::test_f64_channel_prototype).  Host build needs -ffp-contract=off (explicit fma only).
#pragma once
#include <math.h>
#include <string.h>
#include "ntt_rns.cuh"

namespace rnsf {

// u32 / u64 are the global typedefs of rns.cuh

struct Mod {
    double p, pinv;
};

constexpr double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: x + MAGIC - MAGIC == rint(x) for |x| < 2^51

MK_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// y * w mod p as an exact integer in about (-p/2, p/2); y, w exact integers with |y * w / p| < 2^51 (the quotient goes through the
// 1.5 * 2^52 rounding trick): any |y| < 2^51 for a residue w < p
MK_HD double mulmod(double y, double w, const Mod& m) {
    const double h = y * w;
    const double l = fma_(y, w, -h);
    const double q = fma_(h, m.pinv, MAGIC) - MAGIC;
    return fma_(-q, m.p, h) + l;
}

// forward (Cooley-Tukey): (X, Y) -> (X + wY, X - wY)
MK_HD void ct_bfly(double& X, double& Y, double w, const Mod& m) {
    const double t = mulmod(Y, w, m);
    const double x = X;
    X = x + t;
    Y = x - t;
}
// inverse (Gentleman-Sande): (X, Y) -> (X + Y, (X - Y) w)
MK_HD void gs_bfly(double& X, double& Y, double w, const Mod& m) {
    const double s = X + Y, d = X - Y;
    X = s;
    Y = mulmod(d, w, m);
}

// 32-point in-register networks, same stage / block / table-entry structure as rns::ct32 and rns::gs32
template <class TW>
MK_HD void ct32(double (&x)[32], TW tw, const Mod& m) {
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int g = 16 >> k;
#pragma unroll
        for (int b = 0; b < (1 << k); b++) {
            const double w = tw((1 << k) - 1 + b);
#pragma unroll
            for (int j = 0; j < g; j++) ct_bfly(x[2 * g * b + j], x[2 * g * b + j + g], w, m);
        }
    }
}
template <class TW>
MK_HD void gs32(double (&x)[32], TW tw, const Mod& m) {
#pragma unroll
    for (int k = 4; k >= 0; k--) {
        const int g = 16 >> k;
#pragma unroll
        for (int b = 0; b < (1 << k); b++) {
            const double w = tw((1 << k) - 1 + b);
#pragma unroll
            for (int j = 0; j < g; j++) gs_bfly(x[2 * g * b + j], x[2 * g * b + j + g], w, m);
        }
    }
}
struct TwUniform {        // pass A: 31 entries, uniform across the warp
    const double* t;
    MK_HD double operator()(int e) const { return t[e]; }
};
struct TwLane {           // pass B: t points at column `lane` of a [31][32] table
    const double* t;
    MK_HD double operator()(int e) const { return t[e * 32]; }
};

// signed 32-bit integer <-> double through the mantissa: no I2F / F2I instruction
MK_HD double from_i32(int32_t v) {
    const u64 bits = 0x4330000000000000ull | (u64)((u32)v ^ 0x80000000u);      // 2^52 + (v + 2^31)
    double d;
#if defined(__CUDA_ARCH__)
    d = __longlong_as_double((long long)bits);
#else
    memcpy(&d, &bits, 8);
#endif
    return d - 4503601774854144.0;                                             // 2^52 + 2^31
}
// exact integer |t| < 2^31 -> its residue in [0, p) as u32 (t in (-p, p) after one mulmod by the last twiddle, or reduce first)
MK_HD u32 to_residue(double t, const Mod& m) {
    const double r = t - m.p * (fma_(t, m.pinv, MAGIC) - MAGIC);               // (-p/2, p/2]-ish
    const double s = r + 4503601774854144.0;                                    // 2^52 + 2^31 + r: low word = r + 2^31
    u64 bits;
#if defined(__CUDA_ARCH__)
    bits = (u64)__double_as_longlong(s);
#else
    memcpy(&bits, &s, 8);
#endif
    const int32_t v = (int32_t)((u32)bits ^ 0x80000000u);
    return v < 0 ? (u32)(v + (int32_t)m.p) : (u32)v;
}

}  // namespace rnsf
