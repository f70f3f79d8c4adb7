// ntt32.cuh -- in-register size-32 transforms over Goldilocks whose twiddles are
// all powers of two (W = 2^6 is a primitive 32nd root of unity, 2^3 a 64th).
//
// The 1024-point negacyclic NTT  A[k] = sum_i a_i psi^(i(2k+1))  (psi^32 = 2^3)
// is split as i = 32*i1 + i0, k = k0 + 32*k1:
//   F1  Y[i0][k0] = sum_i1 a[32 i1+i0] * 2^(3 i1) * W^(i1 k0)      (twist + dif32, shifts only)
//   F2  Z[i0][k0] = Y[i0][k0] * psi^(i0 (2 k0 + 1))                 (1024 general multiplications)
//   F3  A[k0+32k1] = sum_i0 Z[i0][k0] * W^(i0 k1)                   (dif32, shifts only)
// and the inverse runs I3 (dit32_inv), I2 (times psi^-(..)/1024), I1 (dit32_inv + untwist).
// dif32 maps natural order -> bit-reversed order, dit32_inv maps bit-reversed ->
// natural, so no explicit permutation is ever executed.
#pragma once
#include "goldilocks.cuh"

namespace ntt {

MK_HD int brev5(int r) { return ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | ((r & 8) >> 2) | ((r & 16) >> 4); }

// x[i] *= 2^(3 i)
MK_HD void twist32(u64 (&x)[32]) {
#pragma unroll
    for (int i = 1; i < 32; i++) x[i] = gl::mul_pow2_lt96(x[i], 3 * i);
}
// x[i] *= 2^(-3 i) = -2^(96 - 3 i)
MK_HD void untwist32(u64 (&x)[32]) {
#pragma unroll
    for (int i = 1; i < 32; i++) x[i] = gl::neg(gl::mul_pow2_lt96(x[i], 96 - 3 * i));
}

// forward cyclic size-32 NTT, root W = 2^6, decimation in frequency:
// in natural order, out[r] = X[brev5(r)], X[k] = sum_i x[i] W^(ik)
MK_HD void dif32(u64 (&x)[32]) {
#pragma unroll
    for (int len = 16; len >= 1; len >>= 1) {
#pragma unroll
        for (int blk = 0; blk < 32; blk += 2 * len) {
#pragma unroll
            for (int j = 0; j < len; j++) {
                u64 u = x[blk + j], v = x[blk + j + len];
                x[blk + j] = gl::add(u, v);
                x[blk + j + len] = gl::mul_pow2_lt96(gl::sub(u, v), 6 * j * (16 / len));
            }
        }
    }
}

// inverse cyclic size-32 NTT (unscaled: returns 32 * x), root W^-1, decimation in
// time: in[r] = X[brev5(r)], out natural.  W^-e = -2^(96 - 6e) for e > 0.
MK_HD void dit32_inv(u64 (&x)[32]) {
#pragma unroll
    for (int len = 1; len <= 16; len <<= 1) {
#pragma unroll
        for (int blk = 0; blk < 32; blk += 2 * len) {
#pragma unroll
            for (int j = 0; j < len; j++) {
                int e = j * (16 / len);
                u64 u = x[blk + j], v = x[blk + j + len];
                if (e == 0) {
                    x[blk + j] = gl::add(u, v);
                    x[blk + j + len] = gl::sub(u, v);
                } else {
                    v = gl::mul_pow2_lt96(v, 96 - 6 * e);  // = -(v * W^-e)
                    x[blk + j] = gl::sub(u, v);
                    x[blk + j + len] = gl::add(u, v);
                }
            }
        }
    }
}

}  // namespace ntt
