// rns.cuh -- exact arithmetic of the external product: a residue number system of three 28-bit
// NTT-friendly primes with lazy (Harvey) butterflies, one warp per (polynomial, prime).
//
// Why three 32-bit-word primes instead of one 64-bit prime (DESIGN.md §"Arithmetic"):
// the first build of this repo used the Goldilocks prime 2^64-2^32+1 with two 32-bit key limbs.  It was
// bit-exact but ncu showed it bound by the ALU pipe (64 % busy, fma pipe 10 %): every 64-bit modular
// add/sub/shift is 6-12 IADD3/ISETP/SEL on a 32-bit datapath.  A 28-bit prime makes a whole butterfly
// 1 IMAD.HI + 2 IMAD + 2 IADD3 + 1 VIADDMNMX (3 fma-pipe + 3 alu-pipe instructions), and the exact
// integer result (|R| <= 2l*N*(Bg/2)*2^63 < 2^82) is recovered from the three residues by Garner's CRT
// (M = p0*p1*p2 ~ 2^84), then reduced mod 2^64 -- exactly the Torus64 wrap the reference computes.
//
// All functions are __host__ __device__ so tests/host_emu runs the very same code on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MK_HD __host__ __device__ __forceinline__
#else
#define MK_HD inline
#endif

typedef uint64_t u64;
typedef uint32_t u32;

namespace rns {

constexpr int NP = 3;                       // primes
constexpr int LOGN = 10, N = 1 << LOGN;     // ring degree
// the three largest primes p = 1 (mod 2N) below 2^28; 16p < 2^32 leaves room for lazy sums of 8 products
constexpr u32 PRIME0 = 268369921u, PRIME1 = 268367873u, PRIME2 = 268361729u;

struct alignas(8) uint2_ { u32 x, y; };                // (w, floor(w * 2^32 / p)) twiddle pair; same layout as CUDA's uint2

MK_HD u32 mulhi32(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}
MK_HD u32 umin32(u32 a, u32 b) { return a < b ? a : b; }
// a + b on the alu pipe: ptxas likes to turn two-input adds into IMAD.IADD "to balance the pipes", but here the fma pipe
// is the binding one (IMAD.HI is half rate).  max(a + b, 0) is a single VIADDMNMX.U32 and cannot be moved.
MK_HD u32 alu_add(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __viaddmax_u32(a, b, 0u);
#else
    return a + b;
#endif
}
// Keeps a loop-invariant multiple of p in its own register: without this ptxas rematerialises `x - 4p` as
// IMAD(p, -4, x), spending a slot of the binding fma pipe on what should be an IADD3 / VIADDMNMX of the alu pipe.
MK_HD u32 keep_in_register(u32 v) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(v));
#endif
    return v;
}

// Shoup multiplication by a constant w (wp = floor(w 2^32 / p)): any y < 2^32 -> y*w mod p in [0, 2p)
MK_HD u32 shoup_mul(u32 y, u32 w, u32 wp, u32 p) { return y * w - mulhi32(y, wp) * p; }

// Harvey butterflies with lazy reduction.  16p < 2^32 (p < 2^28), Shoup products accept any 32-bit input and return
// [0, 2p), so sums may grow for several stages before a conditional subtraction is needed.
// forward (Cooley-Tukey): (X, Y) -> (X + wY, X - wY).  X < B on input  =>  both outputs < B + 2p.
//   REDUCE: first bring X from [0, 4p) to [0, 2p)   (VIADDMNMX.U32)
template <bool REDUCE>
MK_HD void ct_bfly(u32& X, u32& Y, u32 w, u32 wp, u32 p, u32 p2) {
    const u32 x = REDUCE ? umin32(X, X - p2) : X;
    const u32 t = shoup_mul(Y, w, wp, p);         // [0, 2p)
    X = alu_add(x, t);
    Y = x - t + p2;
}
// inverse (Gentleman-Sande): (X, Y) -> (X + Y, (X - Y) w).  X, Y < B on input (B <= 4p)  =>  X out < 2B (REDUCE: brought
// back below 4p when 2B = 8p), Y out in [0, 2p).  `off` is a multiple of p that is >= B.
template <bool REDUCE>
MK_HD void gs_bfly(u32& X, u32& Y, u32 w, u32 wp, u32 p, u32 off, u32 p4) {
    const u32 s = alu_add(X, Y), d = X - Y + off;
    X = REDUCE ? umin32(s, s - p4) : s;           // VIADDMNMX.U32 when p4 lives in a register
    Y = shoup_mul(d, w, wp, p);
}

// Montgomery product d * k * 2^-32 mod p in [0, 2p): d < 2^32, k < p, pinv_neg = -p^-1 mod 2^32
MK_HD u32 mont_mul(u32 d, u32 k, u32 p, u32 pinv_neg) {
    const u64 prod = (u64)d * k;
    const u32 m = (u32)prod * pinv_neg;
    return (u32)((prod + (u64)m * p) >> 32);
}

// (d1 * k1 + d2 * k2) * 2^-32 mod p with ONE reduction: d < 2^32, k < p < 2^28, so the sum is < 2^61; result < 2.75p when
// d1, d2 < 14p (the forward transform's output range)
MK_HD u32 mont_mul2(u32 d1, u32 k1, u32 d2, u32 k2, u32 p, u32 pinv_neg) {
    const u64 prod = (u64)d1 * k1 + (u64)d2 * k2;
    const u32 m = (u32)prod * pinv_neg;
    return (u32)((prod + (u64)m * p) >> 32);
}

// 32-point in-register networks.  x[j] are the 32 elements a thread owns; stage k (k = 0..4) pairs
// elements at gap g = 16 >> k inside blocks of 2g; the twiddle of block b is entry e = 2^k - 1 + b of a
// 31-entry table that `tw(e)` returns (uniform across the warp in the pass over the high index bits,
// per-lane in the pass over the low index bits -- see ntt_rns.cuh).
//
// ct32: inputs < B with B <= 4p; no reduction inside (the bound grows by 2p per stage); outputs < B + 10p <= 14p < 2^32.
template <class TW>
MK_HD void ct32(u32 (&x)[32], TW tw, u32 p) {
    const u32 p2 = keep_in_register(2 * p);
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int g = 16 >> k;
#pragma unroll
        for (int b = 0; b < (1 << k); b++) {
            const uint2_ w = tw((1 << k) - 1 + b);
#pragma unroll
            for (int j = 0; j < g; j++) ct_bfly<false>(x[2 * g * b + j], x[2 * g * b + j + g], w.x, w.y, p, p2);
        }
    }
}
// ct32 whose first stage is already half done: x[16..31] hold w0 * y (the stage-0 twiddle products, in [0, p)) instead of y.
// Used for the gadget digits, whose 7-bit values make w0 * y a table lookup instead of a multiplication.
template <class TW>
MK_HD void ct32_pre(u32 (&x)[32], TW tw, u32 p) {
    const u32 p2 = keep_in_register(2 * p);
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const u32 X = x[j], t = x[j + 16];
        x[j] = alu_add(X, t);
        x[j + 16] = X - t + p2;
    }
#pragma unroll
    for (int k = 1; k < 5; k++) {
        const int g = 16 >> k;
#pragma unroll
        for (int b = 0; b < (1 << k); b++) {
            const uint2_ w = tw((1 << k) - 1 + b);
#pragma unroll
            for (int j = 0; j < g; j++) ct_bfly<false>(x[2 * g * b + j], x[2 * g * b + j + g], w.x, w.y, p, p2);
        }
    }
}
// [0, 16p) -> [0, 4p): between the two passes of a forward transform
MK_HD u32 reduce_to_4p(u32 v, u32 p4) {           // p4 = 4p (held in a register by the caller)
    v = umin32(v, v - 2 * p4);
    return umin32(v, v - p4);
}
// gs32: inputs in [0, 4p); the sum output is reduced every stage (it doubles otherwise); outputs in [0, 4p).
template <class TW>
MK_HD void gs32(u32 (&x)[32], TW tw, u32 p) {
    const u32 p4 = keep_in_register(4 * p);
#pragma unroll
    for (int k = 4; k >= 0; k--) {
        const int g = 16 >> k;
#pragma unroll
        for (int b = 0; b < (1 << k); b++) {
            const uint2_ w = tw((1 << k) - 1 + b);
#pragma unroll
            for (int j = 0; j < g; j++) gs_bfly<true>(x[2 * g * b + j], x[2 * g * b + j + g], w.x, w.y, p, p4, p4);
        }
    }
}

// ---- plain modular helpers (table generation, key transform; not on the hot path) ----
MK_HD u32 mulmod(u32 a, u32 b, u32 p) { return (u32)(((u64)a * b) % p); }
MK_HD u32 powmod(u32 a, u64 e, u32 p) {
    u32 r = 1;
    while (e) { if (e & 1) r = mulmod(r, a, p); a = mulmod(a, a, p); e >>= 1; }
    return r;
}
MK_HD u32 shoup_of(u32 w, u32 p) { return (u32)(((u64)w << 32) / p); }
// signed 64-bit integer -> residue in [0, p)
MK_HD u32 residue_i64(int64_t v, u32 p) {
    const int64_t r = v % (int64_t)p;
    return (u32)(r < 0 ? r + (int64_t)p : r);
}

// CRT constants.  Two interchangeable lifts: Garner's mixed radix (MK_CRT_FLOAT = 0, the default) and the floating-point-corrected
// sum (MK_CRT_FLOAT = 1, see crt_lift): exact and 5 IMAD slots shorter per coefficient, but its four int<->float conversions run on
// the 16-lane XU pipe and the kernel measured 6.5 % slower (profiles/ab_r1.txt).
#ifndef MK_CRT_FLOAT
#define MK_CRT_FLOAT 0
#endif
struct Crt {
    u32 p[3];
    u32 c01, c01s, c02, c02s, c12, c12s;   // p0^-1 mod p1, p0^-1 mod p2, p1^-1 mod p2 with Shoup companions
    u64 p01;                               // p0 * p1
    u64 m_mod64;                           // p0 * p1 * p2 mod 2^64
    u64 C[3];                              // M / p_i (< 2^56, exact)
    float inv_p[3];                        // 1 / p_i
    u32 yscale[3];                         // (M / p_i)^-1 mod p_i, folded into the stored key when MK_CRT_FLOAT
};
#if MK_CRT_FLOAT
// Residues y_i = R * (M / p_i)^-1 mod p_i, lazily reduced (any representative below 2^32 / 3), of an integer R with |R| < M/4
// -> R mod 2^64 (two's complement).  R = sum_i y_i (M / p_i) - kappa M with kappa = round(sum_i y_i / p_i): since |R| / M < 1/4 the
// sum is within 1/4 of the integer kappa, so a float32 estimate (error < 2^-18 here) always rounds to it.  Three independent
// 32 x 64-bit products instead of Garner's three dependent modular ones; the factor (M / p_i)^-1 rides in the key scaling.
MK_HD u64 crt_lift(u32 y0, u32 y1, u32 y2, const Crt& c) {
    const float x = (float)y0 * c.inv_p[0] + (float)y1 * c.inv_p[1] + (float)y2 * c.inv_p[2];
#if defined(__CUDA_ARCH__)
    const u32 kappa = __float2uint_rn(x);
#else
    const u32 kappa = (u32)(x + 0.5f);
#endif
    return (u64)y0 * c.C[0] + (u64)y1 * c.C[1] + (u64)y2 * c.C[2] - (u64)kappa * c.m_mod64;
}
#else
// residues r_i in [0, 4 p_i) of an integer R with |R| < M/4 -> R mod 2^64 (two's complement)
MK_HD u64 crt_lift(u32 r0, u32 r1, u32 r2, const Crt& c) {
    const u32 p0 = c.p[0], p1 = c.p[1], p2 = c.p[2];
    u32 v0 = umin32(r0, r0 - 2 * p0);
    v0 = umin32(v0, v0 - p0);                                          // [0, p0)
    u32 v1 = shoup_mul(r1 + 2 * p1 - v0, c.c01, c.c01s, p1);           // v0 < p0 < 2 p1
    v1 = umin32(v1, v1 - p1);                                          // [0, p1)
    const u32 u = shoup_mul(r2 + 2 * p2 - v0, c.c02, c.c02s, p2);      // [0, 2 p2)
    u32 v2 = shoup_mul(u + 2 * p2 - v1, c.c12, c.c12s, p2);
    v2 = umin32(v2, v2 - p2);                                          // [0, p2)
    u64 R = (u64)v0 + (u64)p0 * v1 + c.p01 * v2;
    if (v2 > p2 / 2) R -= c.m_mod64;                                   // negative representative
    return R;
}
#endif

}  // namespace rns
