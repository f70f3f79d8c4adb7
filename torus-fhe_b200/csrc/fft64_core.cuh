// fft64_core.cuh -- arithmetic core of the FP64 FFT channel (fft64.cuh): butterflies, the per-thread passes of the 512-point transforms,
// the integer <-> double bit tricks.  Everything is __host__ __device__ so that tests/host_emu/fft64_emu.cpp runs the very same code on the
// CPU for 32 emulated lanes (std::fma is the correctly rounded fused multiply-add the DFMA instruction computes).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "fft64_layout.h"

#if defined(__CUDACC__)
#define MKF_FN __host__ __device__ __forceinline__
#else
#define MKF_FN inline
#ifndef __restrict__
#define __restrict__
#endif
#endif

namespace mkf {

struct alignas(16) cpx { double x, y; };

constexpr double INV_SQRT2 = 0.70710678118654752440;
// (a, b) -> (a + w b, a - w b): 6 DFMA-pipe instructions
MKF_FN void ct(cpx& a, cpx& b, const cpx w) {
    const double tr = fma(w.x, b.x, fma(-w.y, b.y, a.x));
    const double ti = fma(w.x, b.y, fma(w.y, b.x, a.y));
    b.x = fma(2.0, a.x, -tr);
    b.y = fma(2.0, a.y, -ti);
    a.x = tr;
    a.y = ti;
}
// w = 1 and w = -i: 4 DADD
MKF_FN void ct_one(cpx& a, cpx& b) {
    const double ar = a.x, ai = a.y;
    a.x = ar + b.x; a.y = ai + b.y;
    b.x = ar - b.x; b.y = ai - b.y;
}
MKF_FN void ct_negi(cpx& a, cpx& b) {   // w b = -i b = (b.y, -b.x)
    const double ar = a.x, ai = a.y, br = b.x, bi = b.y;
    a.x = ar + bi; a.y = ai - br;
    b.x = ar - bi; b.y = ai + br;
}
MKF_FN cpx cmul(const cpx a, const cpx b) { return {fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)}; }

// exact small integer -> double without a conversion instruction: bits 0x43300000'u = 2^52 + u
MKF_FN double u2d(unsigned u, double magic) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, (int)u) - magic;
#else
    const uint64_t bits = 0x4330000000000000ull | (uint64_t)u;
    double d;
    memcpy(&d, &bits, 8);
    return d - magic;
#endif
}

// stage 0 of the forward transform straight from the digit words: word j (j < 256) holds the biased digit bytes of coefficients
// j, j + 256, j + 512, j + 768; a~[j] = d0 + i d2, a~[j + 256] = d1 + i d3; out = a~[j] +- exp(i pi / 4) a~[j + 256] (+ for h = 0)
MKF_FN void fwd_stage0_digits(cpx (&v)[16], const uint32_t* __restrict__ dig_s, int lane, int half_bg) {
    const int h = lane >> 4, l16 = lane & 15;
    const double cs = h ? -INV_SQRT2 : INV_SQRT2;
    const double two52 = 4503599627370496.0;
    const double m0 = two52 + (double)half_bg, mP = two52 + 256.0, mQ = two52 + (double)(2 * half_bg);
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const uint32_t word = dig_s[16 * r + l16];
        const uint32_t b0 = word & 0xffu, b1 = (word >> 8) & 0xffu, b2 = (word >> 16) & 0xffu, b3 = word >> 24;
        const double x0 = u2d(b0, m0), y0 = u2d(b2, m0);
        const double P = u2d(b1 - b3 + 256u, mP), Q = u2d(b1 + b3, mQ);      // x1 - y1, x1 + y1
        v[r].x = fma(cs, P, x0);
        v[r].y = fma(cs, Q, y0);
    }
}
// the same from 16-bit digit fields (Torus32 mode, gadget digits of up to 13 bits): wa[j] = (coefficient j, j + 512), wb[j] = (j + 256, j + 768)
MKF_FN void fwd_stage0_digits_wide(cpx (&v)[16], const uint32_t* __restrict__ wa, const uint32_t* __restrict__ wb, int lane, int half_bg) {
    const int h = lane >> 4, l16 = lane & 15;
    const double cs = h ? -INV_SQRT2 : INV_SQRT2;
    const double two52 = 4503599627370496.0;
    const double m0 = two52 + (double)half_bg, mP = two52 + 65536.0, mQ = two52 + (double)(2 * half_bg);
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const uint32_t a = wa[16 * r + l16], b = wb[16 * r + l16];
        const uint32_t b0 = a & 0xffffu, b2 = a >> 16, b1 = b & 0xffffu, b3 = b >> 16;
        const double x0 = u2d(b0, m0), y0 = u2d(b2, m0);
        const double P = u2d(b1 - b3 + 65536u, mP), Q = u2d(b1 + b3, mQ);
        v[r].x = fma(cs, P, x0);
        v[r].y = fma(cs, Q, y0);
    }
}
// the same from the four real coefficients (key transform): a0 = a[j], a1 = a[j + 256], a2 = a[j + 512], a3 = a[j + 768]
MKF_FN cpx fwd_stage0_real(double a0, double a1, double a2, double a3, int h) {
    const double cs = h ? -INV_SQRT2 : INV_SQRT2;
    return {fma(cs, a1 - a3, a0), fma(cs, a1 + a3, a2)};
}

// forward stages 1..4 in the row layout (register r = position bits 7..4): group g = h 2^(d-1) + sub, and
// s(d, g) = s(d, h 2^(d-1)) * exp(i pi rev(sub) / 2^(d-1)): one (warp-half-uniform) table load per stage times compile-time constants
MKF_FN cpx mul_i(const cpx a) { return {-a.y, a.x}; }
MKF_FN void fwd_passA(cpx (&v)[16], const cpx* __restrict__ tw, int h) {
    constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173;   // cos, sin of pi / 8
    {
        const cpx w = tw[TF_A + 0 + h];
#pragma unroll
        for (int r = 0; r < 8; r++) ct(v[r], v[r + 8], w);
    }
    {
        cpx w[2];
        w[0] = tw[TF_A + 2 + 2 * h];
        w[1] = mul_i(w[0]);
#pragma unroll
        for (int g = 0; g < 2; g++)
#pragma unroll
            for (int r = 0; r < 4; r++) ct(v[8 * g + r], v[8 * g + r + 4], w[g]);
    }
    {
        cpx w[4];
        w[0] = tw[TF_A + 6 + 4 * h];
        w[1] = mul_i(w[0]);                                      // rev2(1) = 2
        w[2] = cmul(w[0], cpx{INV_SQRT2, INV_SQRT2});            // rev2(2) = 1
        w[3] = mul_i(w[2]);                                      // rev2(3) = 3
#pragma unroll
        for (int g = 0; g < 4; g++)
#pragma unroll
            for (int r = 0; r < 2; r++) ct(v[4 * g + r], v[4 * g + r + 2], w[g]);
    }
    {
        cpx w[8];
        w[0] = tw[TF_A + 14 + 8 * h];
        w[4] = cmul(w[0], cpx{C1, S1});                          // rev3(4) = 1: exp(i pi / 8)
        w[2] = cmul(w[0], cpx{INV_SQRT2, INV_SQRT2});            // rev3(2) = 2
        w[6] = cmul(w[0], cpx{S1, C1});                          // rev3(6) = 3
        w[1] = mul_i(w[0]);                                      // rev3(1) = 4
        w[5] = mul_i(w[4]);
        w[3] = mul_i(w[2]);
        w[7] = mul_i(w[6]);
#pragma unroll
        for (int g = 0; g < 8; g++) ct(v[2 * g], v[2 * g + 1], w[g]);
    }
}
// forward stages 5..8 in the column layout (register c = position bits 3..0), per-lane twiddles.
// s(d, g), g = lane 2^(d-5) + sub, factors as s(d, lane 2^(d-5)) * exp(i pi rev(sub) / 2^(d-5)): one table load per stage and compile-time
// constants -- the load/store data path, not the FP64 pipe, binds the kernel (loading all fifteen twiddles was 1.2 % slower).  Computing
// all fourteen before the first use is what ptxas allocates best: interleaving them with the butterflies cost 4 %.
MKF_FN void fwd_passB(cpx (&v)[16], const cpx* __restrict__ tw, int lane) {
    {
        const cpx w = tw[TF_B + lane];
#pragma unroll
        for (int c = 0; c < 8; c++) ct(v[c], v[c + 8], w);
    }
    constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173;   // cos, sin of pi / 8
    cpx w6[2], w7[4], w8[8];
    w6[0] = tw[TF_B + 32 + lane];
    w6[1] = mul_i(w6[0]);
    w7[0] = tw[TF_B + 64 + lane];
    w7[1] = mul_i(w7[0]);                                    // rev2(1) = 2: exp(i pi / 2)
    w7[2] = cmul(w7[0], cpx{INV_SQRT2, INV_SQRT2});          // rev2(2) = 1: exp(i pi / 4)
    w7[3] = mul_i(w7[2]);                                    // rev2(3) = 3
    w8[0] = tw[TF_B + 96 + lane];
    w8[4] = cmul(w8[0], cpx{C1, S1});                        // rev3(4) = 1: exp(i pi / 8)
    w8[2] = cmul(w8[0], cpx{INV_SQRT2, INV_SQRT2});          // rev3(2) = 2
    w8[6] = cmul(w8[0], cpx{S1, C1});                        // rev3(6) = 3: exp(3 i pi / 8)
    w8[1] = mul_i(w8[0]);                                    // rev3(1) = 4
    w8[5] = mul_i(w8[4]);                                    // rev3(5) = 5
    w8[3] = mul_i(w8[2]);                                    // rev3(3) = 6
    w8[7] = mul_i(w8[6]);                                    // rev3(7) = 7
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
        for (int c = 0; c < 4; c++) ct(v[8 * g + c], v[8 * g + c + 4], w6[g]);
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
        for (int c = 0; c < 2; c++) ct(v[4 * g + c], v[4 * g + c + 2], w7[g]);
#pragma unroll
    for (int g = 0; g < 8; g++) ct(v[2 * g], v[2 * g + 1], w8[g]);
}
// inverse spans 1, 2, 4, 8 in the column layout: twiddle exp(-2 pi i (c mod sp) / (2 sp)), compile-time constants
MKF_FN void inv_passB(cpx (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 16; c += 2) ct_one(v[c], v[c + 1]);
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
        ct_one(v[c], v[c + 2]);
        ct_negi(v[c + 1], v[c + 3]);
    }
#pragma unroll
    for (int c = 0; c < 16; c += 8) {
        ct_one(v[c], v[c + 4]);
        ct(v[c + 1], v[c + 5], cpx{INV_SQRT2, -INV_SQRT2});
        ct_negi(v[c + 2], v[c + 6]);
        ct(v[c + 3], v[c + 7], cpx{-INV_SQRT2, -INV_SQRT2});
    }
    constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173;   // cos, sin of pi / 8
    ct_one(v[0], v[8]);
    ct(v[1], v[9], cpx{C1, -S1});
    ct(v[2], v[10], cpx{INV_SQRT2, -INV_SQRT2});
    ct(v[3], v[11], cpx{S1, -C1});
    ct_negi(v[4], v[12]);
    ct(v[5], v[13], cpx{-S1, -C1});
    ct(v[6], v[14], cpx{-INV_SQRT2, -INV_SQRT2});
    ct(v[7], v[15], cpx{-C1, -S1});
}
// inverse spans 16, 32, 64, 128 in the row layout: twiddle exp(-2 pi i ((r mod rs) 16 + l16) / (32 rs)) = exp(-2 pi i l16 / (32 rs)) (one table
// load per stage) times exp(-2 pi i (r mod rs) / (2 rs)) (compile-time constants): 4 loads instead of 15, +1.6 %
MKF_FN cpx mul_negi(const cpx a) { return {a.y, -a.x}; }
MKF_FN void inv_passA(cpx (&v)[16], const cpx* __restrict__ tw, int l16) {
    constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173;   // cos, sin of pi / 8
    {
        const cpx w = tw[TI_A + l16];
#pragma unroll
        for (int r0 = 0; r0 < 16; r0 += 2) ct(v[r0], v[r0 + 1], w);
    }
    {
        cpx w[2];
        w[0] = tw[TI_A + 16 + l16];
        w[1] = mul_negi(w[0]);
#pragma unroll
        for (int e = 0; e < 2; e++)
#pragma unroll
            for (int r0 = 0; r0 < 16; r0 += 4) ct(v[r0 + e], v[r0 + e + 2], w[e]);
    }
    {
        cpx w[4];
        w[0] = tw[TI_A + 32 + l16];
        w[1] = cmul(w[0], cpx{INV_SQRT2, -INV_SQRT2});
        w[2] = mul_negi(w[0]);
        w[3] = mul_negi(w[1]);
#pragma unroll
        for (int e = 0; e < 4; e++)
#pragma unroll
            for (int r0 = 0; r0 < 16; r0 += 8) ct(v[r0 + e], v[r0 + e + 4], w[e]);
    }
    {
        cpx w[8];
        w[0] = tw[TI_A + 48 + l16];
        w[1] = cmul(w[0], cpx{C1, -S1});
        w[2] = cmul(w[0], cpx{INV_SQRT2, -INV_SQRT2});
        w[3] = cmul(w[0], cpx{S1, -C1});
#pragma unroll
        for (int e = 0; e < 4; e++) w[4 + e] = mul_negi(w[e]);
#pragma unroll
        for (int e = 0; e < 8; e++) ct(v[e], v[e + 8], w[e]);
    }
}

// bits(r + 1.5 * 2^52) - bits(1.5 * 2^52) = rint(r) for |r| < 2^51: the limb results are recombined on these bit patterns
constexpr double ROUND_MAGIC = 6755399441055744.0;
constexpr uint64_t ROUND_MAGIC_BITS = 0x4338000000000000ull;
MKF_FN uint64_t round_bits(double r) {
    const double t = r + ROUND_MAGIC;
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(t);
#else
    uint64_t b;
    memcpy(&b, &t, 8);
    return b;
#endif
}
// Last inverse stage (span 256), untwist and rounding of one limb for the point pair (j, j + 256): lo = Y[j], hi = Y[j + 256];
// adds the limb's contribution (shifted bit patterns; the caller subtracts round_k(nl) once) to R[b] = coefficient j + 256 b
MKF_FN constexpr uint64_t round_k(int nl) {
    return nl == 3 ? ROUND_MAGIC_BITS + (ROUND_MAGIC_BITS << limb_shift(3, 1)) + (ROUND_MAGIC_BITS << limb_shift(3, 2))
                   : ROUND_MAGIC_BITS + (ROUND_MAGIC_BITS << limb_shift(2, 1));
}
MKF_FN void recombine_limb(uint64_t (&R)[4], cpx lo, cpx hi, const cpx wj, const cpx ut, int sh) {
    ct(lo, hi, wj);
    const cpx e = cmul(lo, ut), g8 = cmul(hi, ut);      // zeta^-(j + 256) = zeta^-j exp(-i pi / 4)
    const cpx f = {(g8.x + g8.y) * INV_SQRT2, (g8.y - g8.x) * INV_SQRT2};
    R[0] += round_bits(e.x) << sh;     // coefficient j
    R[1] += round_bits(f.x) << sh;     // j + 256
    R[2] += round_bits(e.y) << sh;     // j + 512
    R[3] += round_bits(f.y) << sh;     // j + 768
}
// balanced limbs of a key word.  nl = 3 (Torus64): k = l0 + l1 2^22 + l2 2^43 (mod 2^64), |l0| <= 2^21, |l1|, |l2| <= 2^20;
// nl = 2 (Torus32 mode, k holds a 32-bit value): k = l0 + l1 2^16 (mod 2^32), |l0|, |l1| <= 2^15
MKF_FN double key_limb(int64_t k, int limb, int nl = LIMBS) {
    if (nl == 2) {
        const int32_t v = (int32_t)k;
        const int32_t l0 = ((v + (1 << 15)) & 0xFFFF) - (1 << 15);
        const int32_t l1 = (int32_t)((uint32_t)v - (uint32_t)l0) >> 16;
        return (double)(limb == 0 ? l0 : l1);
    }
    const int64_t l0 = ((k + (1ll << 21)) & ((1ll << 22) - 1)) - (1ll << 21);
    const int64_t k1 = (int64_t)((uint64_t)k - (uint64_t)l0) >> 22;
    const int64_t l1 = ((k1 + (1ll << 20)) & ((1ll << 21) - 1)) - (1ll << 20);
    const int64_t l2 = (k1 - l1) >> 21;
    return (double)(limb == 0 ? l0 : limb == 1 ? l1 : l2);
}
// slots of the swizzled transpose buffer: element (half h, row r, column c16) of a warp's 512 points
MKF_FN int transpose_slot(int h, int r, int c16) { return h * 256 + r * 16 + (c16 ^ r); }

}  // namespace mkf
