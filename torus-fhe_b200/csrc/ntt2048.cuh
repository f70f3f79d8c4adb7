// ntt2048.cuh -- arithmetic core of the N = 2048 parameter sets (mktfhe_parameters_{16..128}party_3gen, mk_api.jl:214-298), used by
// kernels2k.cuh.  Verified on the CPU by tests/host_emu/ntt2048_emu.cpp, which runs these very functions for 32 emulated lanes
// against the O(N^2) definition and an exact schoolbook product.
//
// What changes against the N = 1024 path (rns.cuh / ntt_rns.cuh):
//   * gadget digits are up to 26 bits (l = 1, Bg = 2^26 at 16 / 32 parties), so the exact integer result of an external product
//     reaches 2l * N * (Bg/2) * 2^63 = 2^100: FOUR 28-bit primes p = 1 (mod 4096), M ~ 2^112, four-prime Garner lift;
//   * one warp still owns one polynomial, now 64 coefficients per thread: pass A is a 64-point in-register network over the six high
//     index bits (63 warp-uniform twiddles), pass B two independent 32-point networks over the five low bits (per-lane twiddles);
//   * thread `lane` holds a[32 r + lane], r < 64, before the transpose and the two 32-element runs 32 lane + c and 1024 + 32 lane + c
//     after it (rows lane and lane + 32 of a padded 64 x 33 tile: conflict-free both ways).
// The butterflies, Shoup / Montgomery products and their lazy ranges are those of rns.cuh.
#pragma once
#include "rns.cuh"

namespace rns2k {

using rns::uint2_;

constexpr int NP = 4;
constexpr int LOGN = 11, N = 1 << LOGN;
// the four largest primes p = 1 (mod 2N) below 2^28
constexpr u32 PRIMES[NP] = {268369921u, 268361729u, 268271617u, 268238849u};

// 2^LOGM-point in-register Cooley-Tukey network, lazy: stage k pairs elements at gap (2^LOGM / 2) >> k inside blocks of twice the
// gap; block b of stage k uses twiddle entry 2^k - 1 + b.  Inputs < B (B <= 4p); outputs < B + 2 LOGM p (<= 16p for LOGM = 6).
template <int LOGM, class TW>
MK_HD void ct_net(u32 (&x)[1 << LOGM], TW tw, u32 p) {
    const u32 p2 = rns::keep_in_register(2 * p);
#pragma unroll
    for (int k = 0; k < LOGM; k++) {
        const int g = (1 << (LOGM - 1)) >> k;
#pragma unroll
        for (int b = 0; b < (1 << k); b++) {
            const uint2_ w = tw((1 << k) - 1 + b);
#pragma unroll
            for (int j = 0; j < g; j++) rns::ct_bfly<false>(x[2 * g * b + j], x[2 * g * b + j + g], w.x, w.y, p, p2);
        }
    }
}
// A warp-uniform multiple of p that ptxas cannot derive from p again.  rns::keep_in_register is opaque to the front end only: under
// this kernel's register pressure (64 coefficients per thread) ptxas rematerialised `s - 4p` as IMAD(p, -4, s) + VIMNMX -- one slot of the
// binding fma pipe and one extra issue slot per inverse butterfly (profiles/ncu_r1_o_2k_opmix.txt: 2.7 IMAD per IMAD.HI instead of 2).
// A value that went through a SHFL has to stay in its register, and `min(s, s - p4)` becomes one VIADDMNMX.U32 with a negated operand.
// Call with all 32 lanes converged; the value must be the same in every lane (it is read back from lane 0).
#ifndef MK2K_OPAQUE
#define MK2K_OPAQUE 1
#endif
MK_HD u32 opaque_multiple(u32 v) {
#if defined(__CUDA_ARCH__) && MK2K_OPAQUE
    return __shfl_sync(0xffffffffu, v, 0);
#else
    return rns::keep_in_register(v);
#endif
}
// Gentleman-Sande inverse of the same network: inputs and outputs in [0, 4p) (the sum output is reduced every stage); p4 = 4p
template <int LOGM, class TW>
MK_HD void gs_net(u32 (&x)[1 << LOGM], TW tw, u32 p, u32 p4) {
#pragma unroll
    for (int k = LOGM - 1; k >= 0; k--) {
        const int g = (1 << (LOGM - 1)) >> k;
#pragma unroll
        for (int b = 0; b < (1 << k); b++) {
            const uint2_ w = tw((1 << k) - 1 + b);
#pragma unroll
            for (int j = 0; j < g; j++) rns::gs_bfly<true>(x[2 * g * b + j], x[2 * g * b + j + g], w.x, w.y, p, p4, p4);
        }
    }
}

struct Consts {
    u32 p[NP];
    u32 pinv_neg[NP];        // -p^-1 mod 2^32
    u32 key_scale[NP];       // N^-1 * 2^32 mod p (inverse-transform scaling and Montgomery factor, folded into the stored key)
    uint2_ twA[NP][2][63];   // pass A, [prime][0 = forward, 1 = inverse][entry]: psi^brev(entry + 1)
    // Garner: inv[i][j] = p_i^-1 mod p_j (i < j) with Shoup companions, partial products mod 2^64
    u32 ginv[NP][NP], ginv_s[NP][NP];
    u64 pp[NP];              // pp[i] = p_0 * ... * p_{i-1} mod 2^64 (pp[0] = 1)
    u64 m_mod64;             // p_0 p_1 p_2 p_3 mod 2^64
    // quotient-corrected lift (crt4_lift_kappa): the stored key carries the extra factor yscale[i] = (M / p_i)^-1 mod p_i
    u32 yscale[NP];
    u32 key_scale_k[NP];     // key_scale[i] * yscale[i] mod p_i
    u64 cm[NP];              // (M / p_i) mod 2^64
    u32 kf[NP];              // floor(2^59 / p_i): y * kf / 2^59 = y / p_i to 2^-29
};

constexpr int TILE_STRIDE = 33;
constexpr int TILE_WORDS = 64 * TILE_STRIDE;   // 2112 u32 = 8448 B per warp

struct TwUniform {   // pass A
    const uint2_* t;
    MK_HD uint2_ operator()(int e) const { return t[e]; }
};
struct TwLane {      // pass B: t points at this lane's column of a [31][32] table (one table per run of 32)
    const uint2_* t;
    MK_HD uint2_ operator()(int e) const { return t[e * 32]; }
};

// per-lane pass-B tables: twB[prime][dir][half h][31 entries][32 lanes]; the run of thread `lane` in half h is block q = lane + 32 h
// of 32 consecutive coefficients, whose stage-k twiddles are psi^brev(2^k (64 + q) + b)
MK_HD size_t twB_index(int prime, int dir, int half, int entry, int lane) { return ((((size_t)prime * 2 + dir) * 2 + half) * 31 + entry) * 32 + lane; }

// x[r] = a[32 r + lane] in [0, 2p)  ->  after pass A in [0, 14p)
MK_HD void fwd_passA64(u32 (&x)[64], const uint2_* twA_fwd, u32 p) { ct_net<6>(x, TwUniform{twA_fwd}, p); }
// y[c] = the run of 32 consecutive coefficients of block q, in [0, 4p)  ->  positions 32 q + c of the transformed polynomial, < 14p
MK_HD void fwd_passB32(u32 (&y)[32], const uint2_* twB_fwd_lane, u32 p) { rns::ct32(y, TwLane{twB_fwd_lane}, p); }
// the inverse passes take p4 = opaque_multiple(4p), made once per transform (gs_net<5> is rns::gs32 with that parameter)
MK_HD void inv_passB32(u32 (&y)[32], const uint2_* twB_inv_lane, u32 p, u32 p4) { gs_net<5>(y, TwLane{twB_inv_lane}, p, p4); }
MK_HD void inv_passA64(u32 (&x)[64], const uint2_* twA_inv, u32 p, u32 p4) { gs_net<6>(x, TwUniform{twA_inv}, p, p4); }

// [0, 16p) -> [0, 4p) between the passes of the forward transform; p8 = 8p, p4 = 4p, both opaque_multiple()s
MK_HD u32 reduce_to_4p(u32 v, u32 p8, u32 p4) {
    v = rns::umin32(v, v - p8);
    return rns::umin32(v, v - p4);
}

// residues r_i in [0, 4 p_i) of an integer R with |R| < M/4 -> R mod 2^64 (two's complement); Garner's mixed radix over four primes
MK_HD u64 crt4_lift(const u32 (&r)[NP], const Consts& c) {
    u32 v[NP];
    v[0] = rns::umin32(r[0], r[0] - 2 * c.p[0]);
    v[0] = rns::umin32(v[0], v[0] - c.p[0]);                                       // [0, p0)
#pragma unroll
    for (int j = 1; j < NP; j++) {
        const u32 p = c.p[j];
        u32 u = r[j];                                                              // [0, 4p)
#pragma unroll
        for (int i = 0; i < j; i++) {
            // u <- (u - v_i) * p_i^-1 mod p_j, lazily in [0, 2p): v_i < p_i < 2p (the primes differ by < 2^-11), u < 4p
            u = rns::shoup_mul(u + 2 * p - v[i], c.ginv[i][j], c.ginv_s[i][j], p);
        }
        v[j] = rns::umin32(u, u - p);                                              // [0, p)
    }
    u64 R = v[0];
#pragma unroll
    for (int j = 1; j < NP; j++) R += c.pp[j] * v[j];
    if (v[NP - 1] > c.p[NP - 1] / 2) R -= c.m_mod64;                               // negative representative
    return R;
}

// The lift the N = 2048 kernels use (round 2).  Inputs are y_i = R * (M / p_i)^-1 mod p_i in ANY lazy representative below 2^30
// (the factor rides in the stored key, Consts::key_scale_k).  R = sum_i y_i (M / p_i) - kappa M with kappa = round(sum_i y_i / p_i):
// since |R| < M / 4 the sum is within 1/4 of an integer, and a 59-bit fixed-point estimate (error < 2^-27) rounds to it; a lazy
// representative y_i + a p_i adds the integer a to the sum and a M to the first term, which cancel.  Everything is taken mod 2^64.
// 4 IMAD.WIDE for kappa + 5 products of a 64-bit constant by a 32-bit value: ~23 fma-pipe slots against ~36 for Garner's four-prime
// chain (six dependent Shoup products); the CRT phase was 14 % of the kernel's fma-pipe slots (profiles/ncu_r2_f_*).
#ifndef MK2K_CRT_KAPPA
#define MK2K_CRT_KAPPA 1
#endif
MK_HD u64 crt4_lift_kappa(const u32 (&y)[NP], const Consts& c) {
    u64 frac = (u64)1 << 58;                                                       // rounding
    u64 R = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        frac += (u64)y[i] * c.kf[i];                                               // < 4 * 2^30 * 2^32 + 2^58 < 2^64
        R += c.cm[i] * (u64)y[i];
    }
    return R - (frac >> 59) * c.m_mod64;
}
MK_HD u64 lift4(const u32 (&r)[NP], const Consts& c) {
#if MK2K_CRT_KAPPA
    return crt4_lift_kappa(r, c);
#else
    return crt4_lift(r, c);
#endif
}
MK_HD u32 stored_key_scale(const Consts& c, int i) {
#if MK2K_CRT_KAPPA
    return c.key_scale_k[i];
#else
    return c.key_scale[i];
#endif
}

}  // namespace rns2k
