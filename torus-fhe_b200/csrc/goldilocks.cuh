// goldilocks.cuh -- arithmetic in F_p, p = 2^64 - 2^32 + 1 ("Goldilocks").
//
// Why this prime: 2^96 = -1 (mod p), so 2 has order 192 and
// 8 = 2^3 is a primitive 64th root of unity -- every twiddle inside a size-32
// sub-transform is a shift-and-fold on the integer ALU, not a multiplication.
//
// All functions are __host__ __device__ so tests/host_emu can check the very
// same code on the CPU.  Values are canonical, i.e. in [0, p), unless stated.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MK_HD __host__ __device__ __forceinline__
#else
#define MK_HD inline
#endif

typedef uint64_t u64;
typedef uint32_t u32;

namespace gl {

constexpr u64 P = 0xFFFFFFFF00000001ull;
constexpr u64 EPS = 0xFFFFFFFFull;  // 2^64 mod p = 2^32 - 1

MK_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

// a, b canonical -> canonical
MK_HD u64 add(u64 a, u64 b) {
    u64 s = a + b;
    u64 t = s + EPS;  // s - p (mod 2^64)
    return (s < a || s >= P) ? t : s;
}
// a, b canonical -> canonical (a - b + p never needs a second correction)
MK_HD u64 sub(u64 a, u64 b) {
    u64 d = a - b;
    return (a < b) ? d - EPS : d;  // d + p (mod 2^64)
}
MK_HD u64 neg(u64 a) { return a ? P - a : 0; }

// (hi:lo) 128-bit -> canonical.  2^64 = 2^32 - 1, 2^96 = -1.
MK_HD u64 reduce128(u64 lo, u64 hi) {
    u64 h0 = hi & EPS, h1 = hi >> 32;
    u64 t = lo - h1;
    if (lo < h1) t -= EPS;
    u64 m = (h0 << 32) - h0;
    u64 r = t + m;
    if (r < t) r += EPS;
    if (r >= P) r -= P;
    return r;
}
MK_HD u64 mul(u64 a, u64 b) { return reduce128(a * b, mulhi64(a, b)); }

// x * 2^s mod p for 0 <= s < 96 (x canonical).  With s a compile-time constant
// after unrolling, every branch below folds away.
MK_HD u64 mul_pow2_lt96(u64 x, int s) {
    if (s == 0) return x;
    if (s < 64) return reduce128(x << s, x >> (64 - s));
    // s = 64 + r: x*2^r = w0 + w1*2^32 + w2*2^64 (w2 < 2^r), times 2^64:
    //   w0*2^64 + w1*2^96 + w2*2^128 = w0*(2^32-1) - w1 - w2*2^32
    int r = s - 64;
    u64 lo = x << r, w2 = r ? (x >> (64 - r)) : 0;
    u64 w0 = lo & EPS, w1 = lo >> 32;
    u64 a = w0 << 32;                 // < p
    u64 b = w0 + w1 + (w2 << 32);     // < 2^63 + 2^33 < p
    return sub(a, b);
}
// x * 2^s mod p for 0 <= s < 192
MK_HD u64 mul_pow2(u64 x, int s) { return s < 96 ? mul_pow2_lt96(x, s) : neg(mul_pow2_lt96(x, s - 96)); }

// signed small integer -> field element (|v| < p)
MK_HD u64 from_i64(int64_t v) { return v >= 0 ? (u64)v : P - (u64)(-v); }
// centred lift, as a two's-complement 64-bit value (exact when the true integer has |.| < p/2)
MK_HD u64 lift(u64 v) { return v > P / 2 ? v - P : v; }

MK_HD u64 pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) { if (e & 1) r = mul(r, a); a = mul(a, a); e >>= 1; }
    return r;
}

}  // namespace gl
