// ntt_rns.cuh -- per-thread passes of the warp-level 1024-point negacyclic NTT over one 28-bit prime
// (one warp per polynomial, 32 coefficients per thread in registers).  Shared by the CUDA kernels and
// tests/host_emu.
//
// Merged negacyclic Cooley-Tukey NTT (no separate psi-twist): stage with m blocks pairs j and j + t
// (t = N / 2m) with twiddle psi^brev(m + i), i = j / 2t; the output is in bit-reversed index order,
// which never matters because the key is transformed by the same routine and the inverse undoes it.
//
//   pass A (t = 512..32): thread `lane` holds a[32 r + lane] in x[r]; the twiddle depends on r only
//                         -> the 31-entry table is uniform across the warp (__constant__ memory)
//   transpose through a padded 32x33 tile:  store x[r] -> tile[33 r + lane];  load x[c] = tile[33 lane + c]
//   (forward: values are lazily reduced -- inputs < 2p, < 12p after pass A, brought below 4p on the tile load, < 14p after pass B)
//   pass B (t = 16..1)  : thread `lane` holds a[32 lane + c] in x[c]; the twiddle depends on (lane, c)
//                         -> 31 entries per lane (table [31][32] staged in shared memory)
// The inverse (Gentleman-Sande, inverse twiddles) runs pass B, the transposed move, then pass A; its 1/N
// and the Montgomery factor of the pointwise products are folded into the stored key (Consts::key_scale).
#pragma once
#include "rns.cuh"

namespace rns {

struct Consts {
    u32 p[NP];
    u32 pinv_neg[NP];     // -p^-1 mod 2^32
    u32 key_scale[NP];    // N^-1 * 2^32 mod p
    uint2_ twA[NP][2][31];// [prime][0 = forward, 1 = inverse][entry]
    Crt crt;
};

constexpr int TILE_STRIDE = 33;
constexpr int TILE_WORDS = 32 * TILE_STRIDE;   // 1056 u32 = 4224 B

struct TwUniform {        // pass A
    const uint2_* t;
    MK_HD uint2_ operator()(int e) const { return t[e]; }
};
struct TwLane {           // pass B: t already points at column `lane` of a [31][32] table
    const uint2_* t;
    MK_HD uint2_ operator()(int e) const { return t[e * 32]; }
};

MK_HD void fwd_passA(u32 (&x)[32], const uint2_* twA_fwd, u32 p) { ct32(x, TwUniform{twA_fwd}, p); }
MK_HD void fwd_passA_pre(u32 (&x)[32], const uint2_* twA_fwd, u32 p) { ct32_pre(x, TwUniform{twA_fwd}, p); }
MK_HD void fwd_passB(u32 (&x)[32], const uint2_* twB_fwd_lane, u32 p) { ct32(x, TwLane{twB_fwd_lane}, p); }
MK_HD void inv_passB(u32 (&x)[32], const uint2_* twB_inv_lane, u32 p) { gs32(x, TwLane{twB_inv_lane}, p); }
MK_HD void inv_passA(u32 (&x)[32], const uint2_* twA_inv, u32 p) { gs32(x, TwUniform{twA_inv}, p); }

// position (in the transformed, bit-reversed-order polynomial) of element c of thread `lane` after the
// forward transform, and its slot in the streamed key layout: a warp-wide 128-bit load of slots
// [4q .. 4q+3] of every lane is one contiguous 512-byte segment.
MK_HD int ntt_pos(int lane, int c) { return 32 * lane + c; }
MK_HD int key_slot(int lane, int c) { return (c >> 2) * 128 + lane * 4 + (c & 3); }

}  // namespace rns
