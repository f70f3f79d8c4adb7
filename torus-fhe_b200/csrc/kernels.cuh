// kernels.cuh -- sm_100a kernels of the 3gen MK-TFHE bootstrapped-gate path.
//
//   blind_rotate_kernel    gate prologue + mod-switch + k*n mux-rotate steps + sample extraction + (fused) multi-key key switch:
//                          throughput shape, two six-warp gates per CTA
//   blind_rotate_lat_kernel  the same body with one gate per CTA on 6 l warps: batches (and tails of batches) of at most one gate per SM
//   keyswitch_kernel       stand-alone key switch (parity hook; parameter sets the fused epilogue does not cover)
//   bsk_transform_kernel   one-time: int64 key polynomials -> three NTT-domain residue polynomials, streaming layout
//   extprod_kernel / negacyclic_mul_kernel   parity hooks / key-generation primitive built from the same device code
//
// Reference semantics (3-gen-mk-tfhe/src/): 3gen_mk_internals.jl:59-116, tgsw_3gen.jl:102-113, tgsw.jl:112-138,
// rlwe.jl:70-74, keyswitch.jl:45-80, mk_internals.jl:730-744, numeric-functions.jl:70-73,109-111.
//
// Work decomposition of the blind rotation: one gate = 6 warps, one per (RNS prime, output polynomial);
// a CTA holds 2 gates (12 warps at 168 registers, one CTA per SM) plus one shared-memory copy of the twiddle tables.  Per gate,
// resident in shared memory for all k*n steps: the Torus64 accumulator (2 x 1024 x 8 B), the current gadget digits
// (2l x 1024 bytes) and one padded 32x33 word tile per warp (NTT transposes, the exchange of transformed digits between the
// two warps of a prime, then the inverse transforms' residues for the CRT).  The latency shape (lat_wpg below) spends 6 l warps on one
// gate: one warp per (prime, output polynomial, digit pair).
#pragma once
#include <cuda_runtime.h>
#include "ntt_rns.cuh"

namespace mk {

using rns::uint2_;

constexpr int N = rns::N;

// ---- tuning knobs; every default below was chosen by an A/B run recorded in profiles/ab_r1.txt ---------------------------
// Warps per gate.  6: one warp per (prime, output polynomial): the two warps of a prime split the forward transforms by parity
// of the digit polynomial, exchange the transformed digits through their tiles, and each accumulates and inverse-transforms ONE
// output (64 live registers).  3: one warp per prime doing both outputs (96 live registers; 4 gates per CTA).
#ifndef MK_WPG
#define MK_WPG 6
#endif
#ifndef MK_MAX_GPC
#define MK_MAX_GPC (MK_WPG == 6 ? 2 : 4)            // gates per CTA: 12 warps per SM so that each thread gets 168 registers
#endif
// register cap: the largest multiple of 8 with ceil(warps / 4) * 4 * 32 * regs <= 65536 (the register file is allocated in
// units of 4 warps); 12 warps -> 168.  More registers measurably help even without spills (128: -7 %).
#ifndef MK_MAXNREG
#define MK_MAXNREG 168
#endif
#ifndef MK_S_UNROLL
#define MK_S_UNROLL (MK_WPG == 6 ? 1 : 2)           // unroll factor of the loop over digit polynomials
#endif
#ifndef MK_ACC64
#define MK_ACC64 1          // 1: the 2L products of a point accumulate in 64 bits, one Montgomery reduction; 0: one per pair
#endif
#ifndef MK_CRT_ILP
#define MK_CRT_ILP (MK_WPG == 6 ? 11 : 6)   // Garner chains interleaved per thread in the CRT phase; 11 = all of a thread's coefficients at once
#endif
#ifndef MK_LUT_REPL
#define MK_LUT_REPL 1       // 32: per-lane replica of the digit table (conflict-free, 48 KB) -- slower: it shrinks the L1
#endif
// cache policy of the two key streams: __ldg = read-only path, allocating in L1 (the gates of a CTA share key lines there);
// __ldcs = streaming
#ifndef MK_KEY_LD
#define MK_KEY_LD __ldg
#endif
#ifndef MK_KS_LD
#define MK_KS_LD __ldg
#endif
#ifndef MK_KS_UNROLL
#define MK_KS_UNROLL 4      // coefficients per iteration of the fused key-switch gather (x t row loads in flight per thread)
#endif
#ifndef MK_LOCKSTEP
#define MK_LOCKSTEP 0       // n > 0: CTA-wide barrier every n steps to keep the gates on the same key element (measured: loses)
#endif
#ifndef MK_ANTIPHASE
#define MK_ANTIPHASE 0      // experiment (profiles/ab_r2.txt): the two gates of a CTA run half a step apart, held there by two CTA-wide barriers per
#endif                      // step, so that one gate's decomposition / CRT overlaps the other's transforms; MK_AP_SPLIT picks the half-step boundary
#ifndef MK_AP_SPLIT
#define MK_AP_SPLIT 0       // 0: after the last forward transform; 1: between its two passes
#endif
#ifndef MK_PHASE_TRACE
#define MK_PHASE_TRACE 0
#endif
#ifndef MK_KEY_FIXED
#define MK_KEY_FIXED 0
#endif
#ifndef MK_STAGGER_NS
#define MK_STAGGER_NS 0     // start odd gate slots this many ns late to interleave the IMAD-free phases (measured: loses)
#endif
// ----------------------------------------------------------------------------------------------------------------------------
constexpr int WPG = MK_WPG;
constexpr int TPG = 32 * WPG;                        // threads per gate
// Latency launch (batches that fit on the SMs one gate each, l >= 2): 6 l warps per gate, one per (prime, output polynomial, digit
// pair), so the l forward transforms a throughput warp runs back to back run side by side; the partial sums of the digit pairs are
// combined through the tiles before the inverse transform.  Same arithmetic, bit-identical results.
__host__ __device__ constexpr int lat_wpg(int l) { return 6 * l; }
// Warp numbering of the latency launch: gw = 2 l w + l o + slot, slot in [0, l).  The six warps with digit pair i = 0 alone run the
// inverse transforms; lat_inv_slot picks their slot so that they spread over the four schedulers of the SM (gw mod 4) instead of
// piling onto two of them (measured on l = 2: 8.3 ms -> 7.3 ms per launch).  Digit pair of a warp: i = (slot - inv_slot + l) mod l.
__host__ __device__ constexpr int lat_inv_slot(int l, int w, int o) {
    return l == 2 ? (w & 1) : l == 3 ? (w == 2 ? 2 - o : 0) : l == 4 ? ((2 * w + o) & 3) : 0;
}
__host__ __device__ constexpr int lat_warp(int l, int w, int o, int i) { return 2 * l * w + l * o + (i + lat_inv_slot(l, w, o)) % l; }
constexpr int MAX_GPC = MK_MAX_GPC;
constexpr int S_UNROLL = MK_S_UNROLL;

// shared-memory tables, staged once per CTA
constexpr int TWB_WORDS = rns::NP * 2 * 31 * 32 * 2; // per-lane pass-B twiddles [prime][dir][31][32] of (w, w') = 47616 B
constexpr int TWA_WORDS = rns::NP * 2 * 32 * 2;      // warp-uniform pass-A twiddles [prime][dir][32 (31 used)] of (w, w') = 1536 B
// stage-0 products w0 * (digit - Bg/2) mod p for every biased digit byte: lut[prime][byte] (or [prime][byte][lane] when replicated)
constexpr int LUT_REPL = MK_LUT_REPL;
constexpr int LUT_BYTES_MAX = LUT_REPL == 32 ? 128 : 256;
constexpr int LUT_WORDS = rns::NP * LUT_BYTES_MAX * LUT_REPL;
constexpr int TW_SMEM_BYTES = (TWB_WORDS + TWA_WORDS + LUT_WORDS) * 4;

__constant__ rns::Consts c_rns;

// BSK streaming layout (u32): [elem = party*n + j][prime][s = src*l + q][out][key slot]
//   out : 0 = mask', 1 = body'   (accumulator polynomial written)
//   src : 0 = body digits (c0), 1 = mask digits (c1)
//   (out, src) -> reference part: body<-body part_1, body<-mask part_2, mask<-mask part_3, mask<-body part_4
//   key slot: rns::key_slot(lane, c) of transformed position 32*lane + c; values NTT(K mod p) * N^-1 * 2^32 mod p
__host__ __device__ inline size_t bsk_elem_words(int l) { return (size_t)rns::NP * 2 * l * 2 * N; }
// KSK device layout: int32 [party][N][t][B-1][ks_row_stride(n)] -- rows (a[0..n-1], b) padded to a multiple of 4 words so that a
// row is read with 16-byte loads -- followed by one all-zero row
__host__ __device__ constexpr int ks_row_stride(int n) { return (n + 1 + 3) & ~3; }
// per gate: Torus64 accumulator, packed digits, one padded tile per warp
__host__ __device__ constexpr size_t gate_smem_bytes(int l, int wpg = WPG, bool wide = false) {   // wide: 16-bit digit fields (Torus32 mode)
    return (size_t)2 * N * 8 + (size_t)2 * l * N * (wide ? 2 : 1) + (size_t)wpg * rns::TILE_WORDS * 4;
}
// gates per CTA: as many as fit in the 227 KB of shared memory beside the twiddle tables (at most MAX_GPC)
__host__ __device__ constexpr int gpc_for(int l, bool wide = false) {
    int g = MAX_GPC;
    while (g > 1 && (size_t)TW_SMEM_BYTES + (size_t)g * gate_smem_bytes(l, WPG, wide) > 227 * 1024) g--;
    return g;
}
__host__ __device__ constexpr size_t cta_smem_bytes(int l, bool wide = false) {
    return (size_t)TW_SMEM_BYTES + (size_t)gpc_for(l, wide) * gate_smem_bytes(l, WPG, wide);
}

struct GateLinear {   // temp = mu0 + cx*x + cy*y + cz*z   (3gen_mk_gates.jl:8-74)
    int32_t mu0, cx, cy, cz;
};

struct BlindRotateArgs {
    int G, n, k, bgbit;
    int g0;             // first gate of this launch (a batch may be split into a throughput launch and a tail launch); gates g0 .. G-1
    const u32* bsk;
    const uint2_* twB;
    const int32_t *xa, *xb, *ya, *yb, *za, *zb;
    GateLinear lin;
    const int32_t* gate_ids;   // [G] per-gate ids (mixed batches) or nullptr: every gate uses `lin`
    int64_t mu;
    int32_t* ext_out;   // [G][N+1] extracted samples, or nullptr
    int64_t* acc_out;   // [G][2][N] or nullptr
    // fused key switch (epilogue of the same kernel): ksk [k][N][t][B-1][ks_row_stride(n)] followed by one all-zero row
    const int32_t* ksk; // nullptr: no key switch in this launch
    int ks_t, ks_basebit;
    int32_t *oa, *ob;   // [G][k][n], [G]
};

// linear prologue constants of the five bootstrapped gates, 3gen_mk_gates.jl:8-74 (encode_message(m, S) = m << (32 - log2 S))
__host__ __device__ inline GateLinear gate_linear(int gate) {
    switch (gate) {
    case 0: return {(int32_t)(1u << 29), -1, -1, 0};            // NAND   +1/8 - x - y
    case 1: return {(int32_t)(1u << 29), 1, 1, 0};              // OR     +1/8 + x + y
    case 2: return {(int32_t)(0u - (1u << 29)), 1, 1, 0};       // AND    -1/8 + x + y
    case 3: return {(int32_t)(1u << 30), 2, 2, 0};              // XOR    +1/4 + 2x + 2y
    case 4: return {(int32_t)(0u - (1u << 30)), 1, 1, 1};       // 3AND   -1/4 + x + y + z
    default: return {0, 0, 0, 0};
    }
}

template <int W = WPG>
__device__ __forceinline__ void gate_barrier(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(32 * W) : "memory"); }
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// decode_message(x, 2N), numeric-functions.jl:70-73 (wrapping add, arithmetic shift)
__device__ __forceinline__ int mod_switch_2N(int32_t x) {
    return (int32_t)((uint32_t)x + (1u << 20)) >> 21;   // N = 1024: 32 - log2(2N) = 21
}
// t64tot32, numeric-functions.jl:109-111: Int64 -> Float64 (RN), /2^32 (exact), trunc toward zero.
// The reference throws when the quotient is 2^31 (p ~ 2^-54); __double2int_rz saturates.
__device__ __forceinline__ int32_t t64tot32(int64_t v) {
    return __double2int_rz(__ll2double_rn(v) * (1.0 / 4294967296.0));
}

// per-lane pass-B table from global memory, uniform pass-A table from __constant__ memory (a broadcast LDS is cheaper than an
// LDC whose address the compiler cannot prove warp-uniform)
__device__ __forceinline__ void stage_twiddles(uint2_* twB_s, const uint2_* twB_g) {
    const uint4* src = reinterpret_cast<const uint4*>(twB_g);
    uint4* dst = reinterpret_cast<uint4*>(twB_s);
    for (int i = threadIdx.x; i < TWB_WORDS / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    uint2_* twA_s = twB_s + TWB_WORDS / 2;
    for (int i = threadIdx.x; i < rns::NP * 2 * 31; i += blockDim.x) twA_s[(i / 31) * 32 + i % 31] = c_rns.twA[i / 62][(i / 31) & 1][i % 31];
}
// the first forward stage multiplies coefficients 512..1023 by one twiddle w0; for gadget digits (Bg <= 256 values) that product
// comes from this table: lut[prime][byte] = w0 * (byte - Bg/2) mod p
__device__ __forceinline__ void stage_digit_lut(u32* lut, int bgbit) {
    const int half = 1 << (bgbit - 1);
    for (int i = threadIdx.x; i < rns::NP * LUT_BYTES_MAX; i += blockDim.x) {
        const int pi = i / LUT_BYTES_MAX, byte = i % LUT_BYTES_MAX;
        const u32 p = c_rns.p[pi];
        const int d = byte - half;
        const u32 r = d >= 0 ? (u32)d : p - (u32)(-d);
        const u32 v = byte < 2 * half ? rns::mulmod(r, c_rns.twA[pi][0][0].x, p) : 0u;
#pragma unroll
        for (int l = 0; l < LUT_REPL; l++) lut[i * LUT_REPL + l] = v;
    }
}

// forward transform of the 32 elements of this thread (coefficients 32 r + lane, in [0, 2p) -> positions 32 lane + c, in [0, 14p))
__device__ __forceinline__ void warp_ntt_fwd(u32 (&x)[32], u32* tile, const uint2_* twA, const uint2_* twB_lane, u32 p, int lane) {
    const u32 p4 = rns::keep_in_register(4 * p);
    rns::fwd_passA(x, twA, p);
#pragma unroll
    for (int r = 0; r < 32; r++) tile[r * rns::TILE_STRIDE + lane] = x[r];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 32; c++) x[c] = rns::reduce_to_4p(tile[lane * rns::TILE_STRIDE + c], p4);  // pass A leaves [0, 12p)
    __syncwarp();
    rns::fwd_passB(x, twB_lane, p);                                                               // -> [0, 14p)
}
// same for gadget digits: x[0..15] = digit + (p - Bg/2), x[16..31] = lut[digit byte] (first-stage products)
__device__ __forceinline__ void warp_ntt_fwd_digits(u32 (&x)[32], u32* tile, const uint2_* twA, const uint2_* twB_lane, u32 p, int lane) {
    const u32 p4 = rns::keep_in_register(4 * p);
    rns::fwd_passA_pre(x, twA, p);
#pragma unroll
    for (int r = 0; r < 32; r++) tile[r * rns::TILE_STRIDE + lane] = x[r];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 32; c++) x[c] = rns::reduce_to_4p(tile[lane * rns::TILE_STRIDE + c], p4);
    __syncwarp();
    rns::fwd_passB(x, twB_lane, p);
}
// inverse transform (positions 32 lane + c -> coefficients 32 r + lane), scaled by N
__device__ __forceinline__ void warp_ntt_inv(u32 (&x)[32], u32* tile, const uint2_* twA, const uint2_* twB_lane, u32 p, int lane) {
    rns::inv_passB(x, twB_lane, p);
#pragma unroll
    for (int c = 0; c < 32; c++) tile[lane * rns::TILE_STRIDE + c] = x[c];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = tile[r * rns::TILE_STRIDE + lane];
    __syncwarp();
    rns::inv_passA(x, twA, p);
}

// Phase 1 of a step: rotate-subtract and gadget-decompose (tgsw.jl:112-138); digits biased to [0, Bg) and packed 4 per word
// in the order the NTT warps read them (dig: [2L][8][32] words).
template <int L, bool MUX, int W = WPG, bool WIDE = false, bool BODY = false>
__device__ __forceinline__ void decompose_phase(const u64* __restrict__ acc, u32* __restrict__ dig, int a, int bgbit, int gtid) {
    u64 off = 0;
#pragma unroll
    for (int q = 1; q <= L; q++) off += ((u64)1 << (64 - q * bgbit)) << (bgbit - 1);   // tgsw.jl:24-30
    const u32 dmask = (1u << bgbit) - 1;
    // byte fields: 4 coefficients per word, dig [2L][8][32]; WIDE (gadget digits of 9..16 bits): 2 per word, dig [2L][16][32]
    constexpr int PER = WIDE ? 2 : 4, ROWS = 32 / PER, FIELD = 32 / PER;
    // BODY: only the body polynomial (acc[1]) is decomposed -- the mask operand is known to be zero (CCS hybrid product)
    for (int task = gtid; task < (BODY ? 1 : 2) * ROWS * 32; task += 32 * W) {
        const int c = BODY ? 1 : task / (ROWS * 32), rh = (task >> 5) % ROWS, ln = task & 31;
        const u64* poly = acc + c * N;
        u32 packed[L];
#pragma unroll
        for (int q = 0; q < L; q++) packed[q] = 0;
#pragma unroll
        for (int b = 0; b < PER; b++) {
            const int i = 32 * (PER * rh + b) + ln;
            u64 t;
            if (MUX) {
                const int idx = (i - a) & (2 * N - 1);        // (X^a * p)[i] = +-p[(i - a) mod 2N]
                u64 v = poly[idx & (N - 1)];
                if (idx & N) v = 0 - v;
                t = v - poly[i];
            } else {
                t = poly[i];
            }
            t += off;
#pragma unroll
            for (int q = 0; q < L; q++) packed[q] |= ((u32)(t >> (64 - (q + 1) * bgbit)) & dmask) << (FIELD * b);
        }
        const int src = 1 - c;   // src 0 = body = acc[1]
#pragma unroll
        for (int q = 0; q < L; q++) dig[((src * L + q) * ROWS + rh) * 32 + ln] = packed[q];
    }
}

// digit polynomial s of this gate -> the 32 coefficients 32 r + lane of this thread: r < 16 as residues digit + (p - Bg/2) in
// [0, 2p); r >= 16 (the half the first NTT stage multiplies by w0) directly as the products w0 * digit from the table
__device__ __forceinline__ void load_digits(u32 (&x)[32], const u32* __restrict__ dig, const u32* __restrict__ lut, int s, int lane, u32 bias) {
#pragma unroll
    for (int rh = 0; rh < 8; rh++) {
        const u32 word = dig[(s * 8 + rh) * 32 + lane];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const u32 byte = __byte_perm(word, 0, 0x4440 + b);
            x[4 * rh + b] = rh < 4 ? rns::alu_add(byte, bias) : lut[byte * LUT_REPL];
        }
    }
}

// the same for 16-bit digit fields (Torus32 mode, gadget digits of up to 16 bits): all 32 coefficients as residues digit + (p - Bg/2);
// no table stage -- the transform runs the plain first stage
__device__ __forceinline__ void load_digits_wide(u32 (&x)[32], const u32* __restrict__ dig, int s, int lane, u32 bias) {
#pragma unroll
    for (int rh = 0; rh < 16; rh++) {
        const u32 word = dig[(s * 16 + rh) * 32 + lane];
        x[2 * rh] = rns::alu_add(word & 0xFFFFu, bias);
        x[2 * rh + 1] = rns::alu_add(word >> 16, bias);
    }
}

// One external product (tgsw_extern_mul_3gen, tgsw_3gen.jl:102-113) on the accumulator held in shared memory, by the TPG
// threads of one gate.
//   MUX = true : acc += ExtProd(X^a * acc - acc, key)   (mk_mux_rotate_3gen, 3gen_mk_internals.jl:59-62)
//   MUX = false: acc  = ExtProd(acc, key)
// acc: [2][N] u64, [0] = mask, [1] = body.  tiles: [WPG][TILE_WORDS].  bar_id: gate barrier; pbar_id: first pair barrier.
// Garner's lift of both output polynomials by all threads of the gate, added to (MUX) or stored in the accumulator.  The residues
// of output oo sit in coefficient order at tiles + oo * out_stride (+ off1 / off2 words for primes 1 / 2).  The lift is one long dependent
// chain per coefficient: CRT_ILP coefficients are interleaved per thread.
template <bool MUX, int W, int CRT_ILP, int SHIFT = 0>
__device__ __forceinline__ void crt_phase(u64* __restrict__ acc, const u32* __restrict__ tiles, int out_stride, int off1, int off2, int gtid) {
    constexpr int T = 32 * W, CRT_ROUNDS = (2 * N + T * CRT_ILP - 1) / (T * CRT_ILP);
#pragma unroll 1
    for (int round = 0; round < CRT_ROUNDS; round++) {
        u32 r0[CRT_ILP], r1[CRT_ILP], r2[CRT_ILP];
        u64 old[CRT_ILP];
#pragma unroll
        for (int j = 0; j < CRT_ILP; j++) {
            const int idx = gtid + (round * CRT_ILP + j) * T;
            const int cl = idx < 2 * N ? idx : 0;           // clamp: out-of-range slots compute on coefficient 0 and are dropped
            const int oo = cl >> 10, i = cl & (N - 1);
            const u32* rt = tiles + oo * out_stride;
            r0[j] = rt[i]; r1[j] = rt[off1 + i]; r2[j] = rt[off2 + i];
            old[j] = MUX ? acc[cl] : 0;
        }
#pragma unroll
        for (int j = 0; j < CRT_ILP; j++) {
            const int idx = gtid + (round * CRT_ILP + j) * T;
            const u64 R = rns::crt_lift(r0[j], r1[j], r2[j], c_rns.crt);
            if (idx < 2 * N) acc[idx] = old[j] + (SHIFT ? R << SHIFT : R);      // Torus32 mode: the product of unshifted keys goes to the top half
        }
    }
}

// the same with an arbitrary tile per (prime, output) (latency launch); 2048 / (32 W) coefficients per thread, all chains at once
template <bool MUX, int W>
__device__ __forceinline__ void crt_phase_lat(u64* __restrict__ acc, const u32* __restrict__ tiles, int t00, int t01, int t10, int t11, int t20,
                                              int t21, int gtid) {
    constexpr int T = 32 * W, ILP = (2 * N + T - 1) / T;
    u32 r0[ILP], r1[ILP], r2[ILP];
    u64 old[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) {
        const int idx = gtid + j * T;
        const int cl = idx < 2 * N ? idx : 0;
        const int oo = cl >> 10, i = cl & (N - 1);
        r0[j] = tiles[(oo ? t01 : t00) * rns::TILE_WORDS + i];
        r1[j] = tiles[(oo ? t11 : t10) * rns::TILE_WORDS + i];
        r2[j] = tiles[(oo ? t21 : t20) * rns::TILE_WORDS + i];
        old[j] = MUX ? acc[cl] : 0;
    }
#pragma unroll
    for (int j = 0; j < ILP; j++) {
        const int idx = gtid + j * T;
        const u64 R = rns::crt_lift(r0[j], r1[j], r2[j], c_rns.crt);
        if (idx < 2 * N) acc[idx] = old[j] + R;
    }
}

template <int L, bool MUX, int W = WPG, bool WIDE = false, bool BODY = false>
__device__ __forceinline__ void extprod_step(u64* __restrict__ acc, u32* __restrict__ dig, u32* __restrict__ tiles,
                                             const uint2_* __restrict__ twB, const u32* __restrict__ key, int a, int bgbit,
                                             int bar_id, int pbar_id, int gtid, bool ap = false, int cta_threads = 0) {
    static_assert(W == 3 || W == 6 || (L >= 2 && W == lat_wpg(L)), "warps per gate: 3, 6, or 6 l (latency launch)");
    constexpr bool LAT = L >= 2 && W == lat_wpg(L);
    static_assert(!WIDE || (W == 6 && !LAT), "16-bit digit fields (Torus32 mode) are served by the six-warp shape only");
    static_assert(!BODY || WIDE, "the body-only product exists in Torus32 mode");
    const int gw = gtid >> 5, lane = gtid & 31;
    const int w = LAT ? gw / (2 * L) : W == 6 ? gw >> 1 : gw;          // prime of this warp
    decompose_phase<L, MUX, W, WIDE, BODY>(acc, dig, a, bgbit, gtid);
    gate_barrier<W>(bar_id);
    const u32 p = c_rns.p[w], pinv = c_rns.pinv_neg[w];
    const u32 p4 = rns::keep_in_register(4 * p);
    u32* tile = tiles + gw * rns::TILE_WORDS;
    const uint2_* twBf = twB + ((size_t)(w * 2 + 0) * 31) * 32 + lane;
    const uint2_* twBi = twB + ((size_t)(w * 2 + 1) * 31) * 32 + lane;
    const uint2_* twAf = twB + TWB_WORDS / 2 + (w * 2 + 0) * 32;
    const uint2_* twAi = twAf + 32;
    const u32* lut = reinterpret_cast<const u32*>(twB) + TWB_WORDS + TWA_WORDS + w * (LUT_BYTES_MAX * LUT_REPL) + (LUT_REPL == 32 ? lane : 0);
    const uint4* kp = reinterpret_cast<const uint4*>(key + (size_t)w * (2 * L * 2 * N)) + lane;   // K[prime][s][out][slot]
    const u32 bias = p - (1u << (bgbit - 1));
    if constexpr (LAT) {
        // ---- phase 2 (6 l warps): warp (prime w, output o, digit pair i) transforms digit polynomial 2i + o, reads 2i + 1 - o from
        // its partner (w, 1 - o, i), and forms the partial sum of output o over this pair
        const int o = (gw / L) & 1, slot = gw % L;
        const int i = (slot - lat_inv_slot(L, w, o) + L) % L;
        const u32* ptile = tiles + lat_warp(L, w, 1 - o, i) * rns::TILE_WORDS;
        const int qb = pbar_id + w;      // one named barrier per prime: its 2 l warps exchange tiles among themselves only
        auto prime_barrier = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(qb), "r"(64 * L) : "memory"); };
        const int s_own = 2 * i + o, s_for = 2 * i + 1 - o;
        u32 x[32];
        load_digits(x, dig, lut, s_own, lane, bias);
        warp_ntt_fwd_digits(x, tile, twAf, twBf, p, lane);
#pragma unroll
        for (int c = 0; c < 32; c++) tile[c * 32 + lane] = x[c];
        prime_barrier();
        const uint4* k_own = kp + (size_t)(s_own * 2 + o) * (N / 4);
        const uint4* k_for = kp + (size_t)(s_for * 2 + o) * (N / 4);
#pragma unroll
        for (int q4 = 0; q4 < 8; q4++) {
            const uint4 ka = MK_KEY_LD(k_own + q4 * 32), kb = MK_KEY_LD(k_for + q4 * 32);
            x[4 * q4 + 0] = rns::mont_mul2(x[4 * q4 + 0], ka.x, ptile[(4 * q4 + 0) * 32 + lane], kb.x, p, pinv);   // < 2.75p
            x[4 * q4 + 1] = rns::mont_mul2(x[4 * q4 + 1], ka.y, ptile[(4 * q4 + 1) * 32 + lane], kb.y, p, pinv);
            x[4 * q4 + 2] = rns::mont_mul2(x[4 * q4 + 2], ka.z, ptile[(4 * q4 + 2) * 32 + lane], kb.z, p, pinv);
            x[4 * q4 + 3] = rns::mont_mul2(x[4 * q4 + 3], ka.w, ptile[(4 * q4 + 3) * 32 + lane], kb.w, p, pinv);
        }
        prime_barrier();                                       // the partner is done with this warp's tile
        // ---- phase 3: the i > 0 warps hand their partial sums to the i = 0 warp of the same (prime, output), which adds them,
        // inverse-transforms and leaves the residues in its tile; then the CRT of both outputs by the whole gate
        if (i != 0) {
#pragma unroll
            for (int c = 0; c < 32; c++) tile[c * 32 + lane] = x[c];
            prime_barrier();
        } else {
            prime_barrier();
#pragma unroll
            for (int c = 0; c < 32; c++) {
                u32 v = x[c];
#pragma unroll
                for (int ii = 1; ii < L; ii++) v += tiles[lat_warp(L, w, o, ii) * rns::TILE_WORDS + c * 32 + lane];   // < 2.75 l p <= 11p
                if (L > 2) v = rns::umin32(v, v - 2 * p4);
                x[c] = rns::umin32(v, v - p4);                  // [0, 4p)
            }
            warp_ntt_inv(x, tile, twAi, twBi, p, lane);
#pragma unroll
            for (int r = 0; r < 32; r++) tile[32 * r + lane] = x[r];   // residues in coefficient order
        }
        gate_barrier<W>(bar_id);
        // residues of (prime w', output oo) are in the tile of warp lat_warp(L, w', oo, 0)
        constexpr int t00 = lat_warp(L, 0, 0, 0), t01 = lat_warp(L, 0, 1, 0), t10 = lat_warp(L, 1, 0, 0), t11 = lat_warp(L, 1, 1, 0),
                      t20 = lat_warp(L, 2, 0, 0), t21 = lat_warp(L, 2, 1, 0);
        crt_phase_lat<MUX, W>(acc, tiles, t00, t01, t10, t11, t20, t21, gtid);
        gate_barrier<W>(bar_id);
    } else if constexpr (W == 6) {
        // ---- phase 2 (6 warps): this warp transforms the digit polynomials of its parity, reads the partner's from the
        // partner's tile, and accumulates output polynomial `o` for both, two products per Montgomery reduction
        const int o = gw & 1;
        const u32* ptile = tiles + (gw ^ 1) * rns::TILE_WORDS;
        const int pb = pbar_id + w;
#if MK_ACC64
        // all 2L products of a point accumulate in 64 bits (each < 2^60, 2L <= 8) and are Montgomery-reduced once
        u64 acc64[32];
#pragma unroll
        for (int c = 0; c < 32; c++) acc64[c] = 0;
#else
        u32 accv[32];
#pragma unroll
        for (int c = 0; c < 32; c++) accv[c] = 0;
#endif
        // BODY (zero mask operand): only the L body digit polynomials s < L exist; the pairs (2i, 2i + 1) with 2i < L are processed and
        // a warp whose own or partner's polynomial is s >= L skips that term (barriers are kept)
        constexpr int PAIRS = BODY ? (L + 1) / 2 : L;
#pragma unroll S_UNROLL
        for (int i = 0; i < PAIRS; i++) {
            const int s_own = 2 * i + o, s_for = 2 * i + 1 - o;
            const bool have_own = !BODY || s_own < L;      // a missing polynomial is parked as zeros: its products vanish by themselves
            u32 x[32];
            if constexpr (WIDE) {
                if (have_own) {
                    load_digits_wide(x, dig, s_own, lane, bias);
                    warp_ntt_fwd(x, tile, twAf, twBf, p, lane);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; c++) x[c] = 0;
                }
            } else {
            load_digits(x, dig, lut, s_own, lane, bias);
#if MK_ANTIPHASE && MK_AP_SPLIT == 1
            {   // the forward transform written out, with the half-step boundary between its passes
                const u32 p4f = rns::keep_in_register(4 * p);
                rns::fwd_passA_pre(x, twAf, p);
#pragma unroll
                for (int r = 0; r < 32; r++) tile[r * rns::TILE_STRIDE + lane] = x[r];
                __syncwarp();
                if (ap && i == L - 1) asm volatile("bar.sync 15, %0;" ::"r"(cta_threads) : "memory");
#pragma unroll
                for (int c = 0; c < 32; c++) x[c] = rns::reduce_to_4p(tile[lane * rns::TILE_STRIDE + c], p4f);
                __syncwarp();
                rns::fwd_passB(x, twBf, p);
            }
#else
            warp_ntt_fwd_digits(x, tile, twAf, twBf, p, lane);
#if MK_ANTIPHASE
            if (ap && i == L - 1) asm volatile("bar.sync 15, %0;" ::"r"(cta_threads) : "memory");
#endif
#endif
            }
#pragma unroll
            for (int c = 0; c < 32; c++) tile[c * 32 + lane] = x[c];
            pair_barrier(pb);
            const uint4* k_own = kp + (size_t)(s_own * 2 + o) * (N / 4);
            const uint4* k_for = kp + (size_t)(s_for * 2 + o) * (N / 4);
#pragma unroll
            for (int q4 = 0; q4 < 8; q4++) {
                const uint4 ka = MK_KEY_LD(k_own + q4 * 32), kb = MK_KEY_LD(k_for + q4 * 32);
#if MK_ACC64
                acc64[4 * q4 + 0] += (u64)x[4 * q4 + 0] * ka.x + (u64)ptile[(4 * q4 + 0) * 32 + lane] * kb.x;
                acc64[4 * q4 + 1] += (u64)x[4 * q4 + 1] * ka.y + (u64)ptile[(4 * q4 + 1) * 32 + lane] * kb.y;
                acc64[4 * q4 + 2] += (u64)x[4 * q4 + 2] * ka.z + (u64)ptile[(4 * q4 + 2) * 32 + lane] * kb.z;
                acc64[4 * q4 + 3] += (u64)x[4 * q4 + 3] * ka.w + (u64)ptile[(4 * q4 + 3) * 32 + lane] * kb.w;
#else
                accv[4 * q4 + 0] = rns::alu_add(accv[4 * q4 + 0], rns::mont_mul2(x[4 * q4 + 0], ka.x, ptile[(4 * q4 + 0) * 32 + lane], kb.x, p, pinv));
                accv[4 * q4 + 1] = rns::alu_add(accv[4 * q4 + 1], rns::mont_mul2(x[4 * q4 + 1], ka.y, ptile[(4 * q4 + 1) * 32 + lane], kb.y, p, pinv));
                accv[4 * q4 + 2] = rns::alu_add(accv[4 * q4 + 2], rns::mont_mul2(x[4 * q4 + 2], ka.z, ptile[(4 * q4 + 2) * 32 + lane], kb.z, p, pinv));
                accv[4 * q4 + 3] = rns::alu_add(accv[4 * q4 + 3], rns::mont_mul2(x[4 * q4 + 3], ka.w, ptile[(4 * q4 + 3) * 32 + lane], kb.w, p, pinv));
#endif
            }
            pair_barrier(pb);                                  // both warps are done with each other's tile
        }
        // ---- phase 3 (6 warps): inverse transform of this warp's residue polynomial; then CRT of both outputs by the whole gate
        u32 x[32];
#pragma unroll
        for (int c = 0; c < 32; c++) {
#if MK_ACC64
            const u32 m = (u32)acc64[c] * pinv;
            u32 v = (u32)((acc64[c] + (u64)m * p) >> 32);      // < 2L * 14p * p / 2^32 + p <= 8p
#else
            u32 v = accv[c];                                   // sum of L double products, each < 2.75p
#endif
            if (L > 2) v = rns::umin32(v, v - 2 * p4);
            x[c] = rns::umin32(v, v - p4);                     // [0, 4p)
        }
        warp_ntt_inv(x, tile, twAi, twBi, p, lane);
#pragma unroll
        for (int r = 0; r < 32; r++) tile[32 * r + lane] = x[r];   // residues in coefficient order
        gate_barrier<W>(bar_id);
        crt_phase<MUX, W, MK_CRT_ILP, WIDE ? 32 : 0>(acc, tiles, rns::TILE_WORDS, 2 * rns::TILE_WORDS, 4 * rns::TILE_WORDS, gtid);   // warp (w, o) = 2 w + o
        gate_barrier<W>(bar_id);
    } else {
        // ---- phase 2 (3 warps): per prime, forward NTT of each digit polynomial and multiply-accumulate with the key
        u32 acc0[32], acc1[32];
#pragma unroll
        for (int c = 0; c < 32; c++) acc0[c] = acc1[c] = 0;
#pragma unroll S_UNROLL
        for (int s = 0; s < 2 * L; s++) {
            u32 x[32];
            load_digits(x, dig, lut, s, lane, bias);
            warp_ntt_fwd_digits(x, tile, twAf, twBf, p, lane);
            const uint4* k0 = kp + (size_t)(s * 2) * (N / 4);
#pragma unroll
            for (int q4 = 0; q4 < 8; q4++) {
                const uint4 kv = __ldg(k0 + q4 * 32);
                acc0[4 * q4 + 0] = rns::alu_add(acc0[4 * q4 + 0], rns::mont_mul(x[4 * q4 + 0], kv.x, p, pinv));
                acc0[4 * q4 + 1] = rns::alu_add(acc0[4 * q4 + 1], rns::mont_mul(x[4 * q4 + 1], kv.y, p, pinv));
                acc0[4 * q4 + 2] = rns::alu_add(acc0[4 * q4 + 2], rns::mont_mul(x[4 * q4 + 2], kv.z, p, pinv));
                acc0[4 * q4 + 3] = rns::alu_add(acc0[4 * q4 + 3], rns::mont_mul(x[4 * q4 + 3], kv.w, p, pinv));
            }
            const uint4* k1 = k0 + N / 4;
#pragma unroll
            for (int q4 = 0; q4 < 8; q4++) {
                const uint4 kv = __ldg(k1 + q4 * 32);
                acc1[4 * q4 + 0] = rns::alu_add(acc1[4 * q4 + 0], rns::mont_mul(x[4 * q4 + 0], kv.x, p, pinv));
                acc1[4 * q4 + 1] = rns::alu_add(acc1[4 * q4 + 1], rns::mont_mul(x[4 * q4 + 1], kv.y, p, pinv));
                acc1[4 * q4 + 2] = rns::alu_add(acc1[4 * q4 + 2], rns::mont_mul(x[4 * q4 + 2], kv.z, p, pinv));
                acc1[4 * q4 + 3] = rns::alu_add(acc1[4 * q4 + 3], rns::mont_mul(x[4 * q4 + 3], kv.w, p, pinv));
            }
        }
        // ---- phase 3 (3 warps): per output polynomial, inverse NTT of the three residue polynomials, Garner CRT, update
#pragma unroll 1
        for (int out = 0; out < 2; out++) {
            u32 x[32];
#pragma unroll
            for (int c = 0; c < 32; c++) {
                u32 v = out ? acc1[c] : acc0[c];                 // sum of 2L products, each in [0, 2p)
                if (L > 2) v = rns::umin32(v, v - 2 * p4);
                x[c] = rns::umin32(v, v - p4);                   // [0, 4p)
            }
            warp_ntt_inv(x, tile, twAi, twBi, p, lane);
#pragma unroll
            for (int r = 0; r < 32; r++) tile[32 * r + lane] = x[r];   // residues in coefficient order
            gate_barrier<W>(bar_id);
            u64* ap = acc + out * N;
            for (int i = gtid; i < N; i += 32 * W) {
                const u64 R = rns::crt_lift(tiles[i], tiles[rns::TILE_WORDS + i], tiles[2 * rns::TILE_WORDS + i], c_rns.crt);
                ap[i] = MUX ? ap[i] + R : R;
            }
            gate_barrier<W>(bar_id);
        }
    }
}

// Sample extraction + multi-key LWE key switch of ONE gate by its TPG threads, as the epilogue of the blind rotation
// (rlwe_extract_sample_64 rlwe.jl:70-74, mk_keyswitch_3gen mk_internals.jl:730-744, keyswitch keyswitch.jl:45-80).
// The gather of k*N*t rows (11.2 MB from L2 at the 2-party parameters) is latency/LSU work; fused here it overlaps with the
// IMAD-bound steps of the other gate on the SM instead of running as a separate 4 %-of-step kernel.
// Thread `gtid` owns the four output columns 4*gtid .. 4*gtid+3 (one LDG.128 per row); zero digits read the all-zero row
// instead of branching, so the 4 * T row reads of four consecutive coefficients are all in flight together.
__host__ __device__ constexpr bool ks_fusable(int n, int t) { return ks_row_stride(n) <= 4 * TPG && (t == 3 || t == 5); }
template <int T, int W = WPG>
__device__ __forceinline__ void fused_keyswitch(const u64* __restrict__ acc, u32* __restrict__ s_a, const BlindRotateArgs& p, int g, int gtid,
                                                int bar_id) {
    const int n = p.n, stride = ks_row_stride(n), bb = p.ks_basebit, B1 = (1 << bb) - 1;
    const uint32_t prec_offset = 1u << (32 - (1 + bb * T));   // keyswitch.jl:58
    for (int i = gtid; i < N; i += 32 * W) {
        const u64 v = i == 0 ? acc[0] : (u64)0 - acc[N - i];
        const int32_t ai = t64tot32((int64_t)v);
        if (p.ext_out) p.ext_out[(size_t)g * (N + 1) + i] = ai;
        s_a[i] = (uint32_t)ai + prec_offset;                  // :59
    }
    const int32_t eb = t64tot32((int64_t)acc[N]);
    if (p.ext_out && gtid == 0) p.ext_out[(size_t)g * (N + 1) + N] = eb;
    gate_barrier<W>(bar_id);
    if constexpr (W < 12) {
        // six-warp gates (throughput launch): one group of stride / 4 threads does the whole gather
        const int col0 = 4 * gtid;
        if (col0 >= stride) return;                               // no barrier below: idle threads may leave
        const size_t party_words = (size_t)N * T * B1 * stride;
        const uint4* zero_row = reinterpret_cast<const uint4*>(p.ksk + (size_t)p.k * party_words) + gtid;
        uint32_t bsum = 0;
        for (int party = 0; party < p.k; party++) {
            uint4 out = make_uint4(0, 0, 0, 0);
            const int32_t* rows = p.ksk + (size_t)party * party_words;
#pragma unroll 1
            for (int i = 0; i < N; i += MK_KS_UNROLL) {
#pragma unroll
                for (int ii = 0; ii < MK_KS_UNROLL; ii++) {
                    const uint32_t ai = s_a[i + ii];
#pragma unroll
                    for (int j = 1; j <= T; j++) {
                        const uint32_t d = (ai >> (32 - j * bb)) & (uint32_t)B1;   // :65-67
                        const uint4* r = d ? reinterpret_cast<const uint4*>(rows + (((size_t)(i + ii) * T + (j - 1)) * B1 + (d - 1)) * stride) + gtid
                                           : zero_row;                              // :74-76
                        const uint4 v = MK_KS_LD(r);
                        out.x -= v.x; out.y -= v.y; out.z -= v.z; out.w -= v.w;
                    }
                }
            }
            const uint32_t o4[4] = {out.x, out.y, out.z, out.w};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int col = col0 + c;
                if (col < n) p.oa[((size_t)g * p.k + party) * n + col] = (int32_t)o4[c];
                if (col == n) bsum += o4[c];
            }
        }
        if (col0 <= n && n < col0 + 4) p.ob[g] = (int32_t)((uint32_t)eb + bsum);   // thread owning column n
        return;
    }
    // A row needs stride / 4 threads (131 at n = 520).  With the 6 l warps of the latency launch several such groups fit: group q takes
    // the coefficients i = 4q, 4q + 1, .. (mod 4 * groups) and the partial sums meet in shared memory behind s_a.
    const int tpr = stride / 4;                               // threads per row (stride is a multiple of 4)
    constexpr int MAXG = 4;
    const int groups = min(MAXG, 32 * W / tpr);
    const int grp = gtid / tpr, t = gtid - grp * tpr;
    const bool active = grp < groups;
    const int col0 = 4 * t;
    const size_t party_words = (size_t)N * T * B1 * stride;
    const uint4* zero_row = reinterpret_cast<const uint4*>(p.ksk + (size_t)p.k * party_words) + t;
    uint4* part = reinterpret_cast<uint4*>(s_a + N);          // [2 (party parity)][groups - 1][tpr] partial sums
    uint32_t bsum = 0;
    for (int party = 0; party < p.k; party++) {
        uint4 out = make_uint4(0, 0, 0, 0);
        const int32_t* rows = p.ksk + (size_t)party * party_words;
        if (active) {
#pragma unroll 1
            for (int i = 4 * grp; i < N; i += 4 * groups) {
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    const uint32_t ai = s_a[i + ii];
#pragma unroll
                    for (int j = 1; j <= T; j++) {
                        const uint32_t d = (ai >> (32 - j * bb)) & (uint32_t)B1;   // :65-67
                        const uint4* r = d ? reinterpret_cast<const uint4*>(rows + (((size_t)(i + ii) * T + (j - 1)) * B1 + (d - 1)) * stride) + t
                                           : zero_row;                              // :74-76
                        const uint4 v = MK_KS_LD(r);
                        out.x -= v.x; out.y -= v.y; out.z -= v.z; out.w -= v.w;
                    }
                }
            }
        }
        if (groups > 1) {
            uint4* pp = part + (size_t)(party & 1) * (MAXG - 1) * tpr;
            if (active && grp > 0) pp[(grp - 1) * tpr + t] = out;
            gate_barrier<W>(bar_id);
            if (grp == 0)
                for (int q = 1; q < groups; q++) {
                    const uint4 v = pp[(q - 1) * tpr + t];
                    out.x += v.x; out.y += v.y; out.z += v.z; out.w += v.w;
                }
        }
        if (grp == 0) {
            const uint32_t o4[4] = {out.x, out.y, out.z, out.w};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int col = col0 + c;
                if (col < n) p.oa[((size_t)g * p.k + party) * n + col] = (int32_t)o4[c];
                if (col == n) bsum += o4[c];
            }
        }
    }
    if (grp == 0 && col0 <= n && n < col0 + 4) p.ob[g] = (int32_t)((uint32_t)eb + bsum);   // thread owning column n
}

// GPC gates per CTA, WPG warps per gate.  Accumulators resident in shared memory for all k*n steps.
template <int L, int GPC, int W, bool WIDE = false>
__device__ __forceinline__ void blind_rotate_body(const BlindRotateArgs& p) {
    constexpr int TPG = 32 * W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2_* twB = reinterpret_cast<uint2_*>(smem_raw);
    stage_twiddles(twB, p.twB);
    stage_digit_lut(reinterpret_cast<u32*>(twB) + TWB_WORDS + TWA_WORDS, p.bgbit);
    __syncthreads();
    static_assert(W == 3 || (1 + rns::NP) * GPC < 16, "named barriers: one per gate plus one per (gate, prime)");
    const int slot = threadIdx.x / TPG, gtid = threadIdx.x - slot * TPG, bar_id = 1 + slot, pbar_id = 1 + GPC + slot * rns::NP;
    const int g = p.g0 + blockIdx.x * GPC + slot;
    if (g >= p.G) return;   // no CTA-wide barrier below this line
    unsigned char* base = smem_raw + TW_SMEM_BYTES + (size_t)slot * gate_smem_bytes(L, W, WIDE);
    u64* acc = reinterpret_cast<u64*>(base);
    u32* dig = reinterpret_cast<u32*>(base + 2 * N * 8);
    u32* tiles = dig + 2 * L * (N / (WIDE ? 2 : 4));
    const int kn = p.k * p.n;
    // gate prologue (3gen_mk_gates.jl) + mod switch (3gen_mk_internals.jl:102-103); every thread of the gate computes
    // the same rotation amounts from broadcast loads
    const GateLinear lin = p.gate_ids ? gate_linear(__ldg(p.gate_ids + g)) : p.lin;
    auto rotation = [&](const int32_t* xs, const int32_t* ys, const int32_t* zs, size_t idx, uint32_t mu0) {
        uint32_t t = mu0 + (uint32_t)lin.cx * (uint32_t)__ldg(xs + idx);
        if (lin.cy) t += (uint32_t)lin.cy * (uint32_t)__ldg(ys + idx);
        if (lin.cz) t += (uint32_t)lin.cz * (uint32_t)__ldg(zs + idx);
        return mod_switch_2N((int32_t)t);
    };
    const int barb = rotation(p.xb, p.yb, p.zb, g, (uint32_t)lin.mu0);
    // acc = (0, X^{-barb} * testvect), testvect = mu * (1 + X + ... + X^{N-1})  (:88-92, rlwe.jl:113-119)
    {
        const int s = (-barb) & (2 * N - 1);
        for (int i = gtid; i < N; i += TPG) {
            const int idx = (i - s) & (2 * N - 1);
            acc[i] = 0;
            acc[N + i] = (idx & N) ? (u64)0 - (u64)p.mu : (u64)p.mu;
        }
    }
    gate_barrier<W>(bar_id);
#if MK_STAGGER_NS
    // The gates of a CTA execute identical instruction streams and would run in phase (both in the IMAD-free decompose / CRT
    // phases at the same time).  Start every other gate a fraction of a step later so that those phases interleave.
    if (slot & 1) __nanosleep(MK_STAGGER_NS);
#endif
    // mk_blind_rotate_3gen: parties outer, coefficients inner (:66-84); element index = party*n + j
    const size_t estride = bsk_elem_words(L);
    const size_t abase = (size_t)g * kn;
    // the mask words of step it + 1 are loaded (raw) while step it runs and only combined / mod-switched when needed, so the
    // load latency never sits on the critical path
    int32_t rx = __ldg(p.xa + abase), ry = lin.cy ? __ldg(p.ya + abase) : 0, rz = lin.cz ? __ldg(p.za + abase) : 0;
#if MK_LOCKSTEP
    const bool cta_full = (blockIdx.x + 1) * GPC <= p.G;      // every gate slot of this CTA is active
#endif
#if MK_ANTIPHASE
    // both gate slots of the CTA active, two six-warp gates: run them half a step apart (gate 1 waits one barrier more at the start, gate 0
    // one more at the end: 2 kn + 1 CTA-wide barriers each)
    const bool ap = GPC == 2 && W == 6 && (blockIdx.x + 1) * GPC + p.g0 <= p.G;
    if (ap && slot == 1) asm volatile("bar.sync 15, %0;" ::"r"(GPC * TPG) : "memory");
#endif
    for (int it = 0; it < kn; it++) {
        const int a = mod_switch_2N((int32_t)((uint32_t)lin.cx * (uint32_t)rx + (uint32_t)lin.cy * (uint32_t)ry + (uint32_t)lin.cz * (uint32_t)rz));
        if (it + 1 < kn) {
            rx = __ldg(p.xa + abase + it + 1);
            if (lin.cy) ry = __ldg(p.ya + abase + it + 1);
            if (lin.cz) rz = __ldg(p.za + abase + it + 1);
        }
#if MK_LOCKSTEP
        // keep the gates of a CTA on the same key element: their key loads then hit the same L1 lines (measured: running them
        // out of phase costs 3 %)
        if (cta_full && (it % MK_LOCKSTEP) == 0) asm volatile("bar.sync 15, %0;" ::"r"(GPC * TPG) : "memory");
#endif
#if MK_ANTIPHASE
        if (ap) asm volatile("bar.sync 15, %0;" ::"r"(GPC * TPG) : "memory");          // start of the first half-step
        if (a == 0) { if (ap) asm volatile("bar.sync 15, %0;" ::"r"(GPC * TPG) : "memory"); continue; }
#else
        if (a == 0) continue;   // :69 (uniform across the gate)
#endif
#if MK_PHASE_TRACE   // diagnostic: SM clock at the start of every step of every gate of CTA 0 and CTA 1, dumped through acc_out
        if (p.acc_out && gtid == 0 && blockIdx.x < 2 && it < 2 * N) p.acc_out[(size_t)g * 2 * N + it] = (int64_t)clock64();
#endif
#if MK_KEY_FIXED     // diagnostic only (NOT exact): every step reads key element 0, which stays in L1 -- the upper bound of what any
        extprod_step<L, true, W>(acc, dig, tiles, twB, p.bsk, a, p.bgbit, bar_id, pbar_id, gtid);            // key-staging scheme (TMA, smem) could gain
#elif MK_ANTIPHASE
        extprod_step<L, true, W>(acc, dig, tiles, twB, p.bsk + (size_t)it * estride, a, p.bgbit, bar_id, pbar_id, gtid, ap, GPC * TPG);
#else
        extprod_step<L, true, W, WIDE>(acc, dig, tiles, twB, p.bsk + (size_t)it * estride, a, p.bgbit, bar_id, pbar_id, gtid);
#endif
    }
#if MK_ANTIPHASE
    if (ap && slot == 0) asm volatile("bar.sync 15, %0;" ::"r"(GPC * TPG) : "memory");
#endif
#if !MK_PHASE_TRACE
    if (p.acc_out) {
        int64_t* ao = p.acc_out + (size_t)g * 2 * N;
        for (int i = gtid; i < 2 * N; i += TPG) ao[i] = (int64_t)acc[i];
    }
#endif
    if (p.ksk) {   // fused extraction + key switch (the host only sets ksk when ks_fusable(n, t))
        if (p.ks_t == 3) fused_keyswitch<3, W>(acc, tiles, p, g, gtid, bar_id);
        else fused_keyswitch<5, W>(acc, tiles, p, g, gtid, bar_id);
        return;
    }
    // rlwe_extract_sample_64 (rlwe.jl:70-74): a'_0 = mask_0, a'_i = -mask_{N-i}, b' = body_0
    int32_t* ext = p.ext_out + (size_t)g * (N + 1);
    for (int i = gtid; i < N; i += TPG) {
        const u64 v = i == 0 ? acc[0] : (u64)0 - acc[N - i];
        ext[i] = t64tot32((int64_t)v);
    }
    if (gtid == 0) ext[N] = t64tot32((int64_t)acc[N]);
}

// GPC gates per CTA, W warps per gate, 12 warps per SM at 168 registers (throughput: <L, 2, 6>; small batches of l = 1: <1, 1, 6>)
template <int L, int GPC, int W = WPG>
__global__ void __maxnreg__(MK_MAXNREG) blind_rotate_kernel(BlindRotateArgs p) { blind_rotate_body<L, GPC, W>(p); }

// Torus32 mode (mktfhe_params.flags & MKTFHE_FLAG_TORUS32): gadget digits of up to 16 bits in 16-bit fields, keys loaded UNSHIFTED
// (32-bit signed values) and every external product added as R << 32 -- the exact integer R then stays below 2^59 whatever the
// gadget base, where the v << 32 embedding of the default kernels would leave the CRT range for Bg > 2^8.  Serves the reference's
// tfhe_parameters_80 (Bg = 2^10, api.jl:76-91).  Separate entry points: the default kernels' machine code does not change.
template <int L, int GPC>
__global__ void __maxnreg__(MK_MAXNREG) blind_rotate_t32_kernel(BlindRotateArgs p) { blind_rotate_body<L, GPC, WPG, true>(p); }

// latency launch: one gate per CTA, 6 l warps (168 registers at l = 2, 112 at l = 3, 80 at l = 4: a latency warp keeps no wide accumulators)
template <int L>
__global__ void __launch_bounds__(32 * lat_wpg(L), 1) blind_rotate_lat_kernel(BlindRotateArgs p) { blind_rotate_body<L, 1, lat_wpg(L)>(p); }

// parity hook: acc_out[g] = ExtProd(acc_in[g], bsk[elem[g]])
template <int L, int GPC>
__global__ void __maxnreg__(MK_MAXNREG) extprod_kernel(int G, const u32* bsk, const uint2_* twB_g, int bgbit, const int32_t* elem,
                                                               const int64_t* acc_in, int64_t* acc_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2_* twB = reinterpret_cast<uint2_*>(smem_raw);
    stage_twiddles(twB, twB_g);
    stage_digit_lut(reinterpret_cast<u32*>(twB) + TWB_WORDS + TWA_WORDS, bgbit);
    __syncthreads();
    const int slot = threadIdx.x / TPG, gtid = threadIdx.x - slot * TPG, bar_id = 1 + slot, pbar_id = 1 + GPC + slot * rns::NP;
    const int g = blockIdx.x * GPC + slot;
    if (g >= G) return;
    unsigned char* base = smem_raw + TW_SMEM_BYTES + (size_t)slot * gate_smem_bytes(L);
    u64* acc = reinterpret_cast<u64*>(base);
    u32* dig = reinterpret_cast<u32*>(base + 2 * N * 8);
    u32* tiles = dig + 2 * L * (N / 4);
    for (int i = gtid; i < 2 * N; i += TPG) acc[i] = (u64)acc_in[(size_t)g * 2 * N + i];
    gate_barrier(bar_id);
    extprod_step<L, false>(acc, dig, tiles, twB, bsk + (size_t)elem[g] * bsk_elem_words(L), 0, bgbit, bar_id, pbar_id, gtid);
    for (int i = gtid; i < 2 * N; i += TPG) acc_out[(size_t)g * 2 * N + i] = (int64_t)acc[i];
}

// the same in Torus32 mode: acc_in / acc_out hold Torus32 values in the TOP half of their words (v << 32)
// BODY: the mask operand is zero and is neither read nor decomposed (the rounds of the CCS hybrid product)
template <int L, int GPC, bool BODY = false>
__global__ void __maxnreg__(MK_MAXNREG) extprod_t32_kernel(int G, const u32* bsk, const uint2_* twB_g, int bgbit, const int32_t* elem,
                                                            const int64_t* acc_in, int64_t* acc_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2_* twB = reinterpret_cast<uint2_*>(smem_raw);
    stage_twiddles(twB, twB_g);
    __syncthreads();
    const int slot = threadIdx.x / TPG, gtid = threadIdx.x - slot * TPG, bar_id = 1 + slot, pbar_id = 1 + GPC + slot * rns::NP;
    const int g = blockIdx.x * GPC + slot;
    if (g >= G) return;
    unsigned char* base = smem_raw + TW_SMEM_BYTES + (size_t)slot * gate_smem_bytes(L, WPG, true);
    u64* acc = reinterpret_cast<u64*>(base);
    u32* dig = reinterpret_cast<u32*>(base + 2 * N * 8);
    u32* tiles = dig + 2 * L * (N / 2);
    for (int i = gtid + (BODY ? N : 0); i < 2 * N; i += TPG) acc[i] = (u64)acc_in[(size_t)g * 2 * N + i];
    gate_barrier(bar_id);
    extprod_step<L, false, WPG, true, BODY>(acc, dig, tiles, twB, bsk + (size_t)elem[g] * bsk_elem_words(L), 0, bgbit, bar_id, pbar_id, gtid);
    for (int i = gtid; i < 2 * N; i += TPG) acc_out[(size_t)g * 2 * N + i] = (int64_t)acc[i];
}

// One warp per (polynomial, prime): raw int64 key -> NTT-domain residues in the streaming layout.
// raw: [n][4 parts][l][N] int64 of one party; task = ((j*4 + part)*l + q)*3 + prime.
constexpr int XF_WARPS = 4;
__global__ void __launch_bounds__(XF_WARPS * 32) bsk_transform_kernel(const int64_t* __restrict__ raw, u32* __restrict__ bsk, int n, int l, int party,
                                                                        const uint2_* __restrict__ twB, int ntasks) {
    __shared__ u32 tiles[XF_WARPS * rns::TILE_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * XF_WARPS + warp;
    if (task >= ntasks) return;
    const int pi = task % 3, pq = task / 3;
    const int q = pq % l, part = (pq / l) & 3, j = pq / (4 * l);
    // part_1: body<-body, part_2: body<-mask, part_3: mask<-mask, part_4: mask<-body  (tgsw_3gen.jl:109-110)
    const int out = part < 2 ? 1 : 0;
    const int src = (part == 0 || part == 3) ? 0 : 1;
    const u32 p = c_rns.p[pi];
    const int64_t* poly = raw + (size_t)pq * N;
    u32 x[32];
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = rns::residue_i64(poly[32 * r + lane], p);
    warp_ntt_fwd(x, tiles + warp * rns::TILE_WORDS, c_rns.twA[pi][0], twB + ((size_t)(pi * 2 + 0) * 31) * 32 + lane, p, lane);
    const size_t e = (size_t)party * n + j;
    u32* dst = bsk + e * bsk_elem_words(l) + ((size_t)(pi * 2 * l + (src * l + q)) * 2 + out) * N;
#pragma unroll
    for (int c = 0; c < 32; c++) dst[rns::key_slot(lane, c)] = rns::mulmod(x[c] % p, c_rns.key_scale[pi], p);
}

// parity hook and key-generation primitive: exact c = a * b mod (X^N + 1, 2^64) for |a_i| <= 2^8; 3 warps = 3 primes
constexpr int NM_THREADS = 32 * rns::NP;
__global__ void __launch_bounds__(NM_THREADS) negacyclic_mul_kernel(const int64_t* __restrict__ a, const int64_t* __restrict__ b, int64_t* __restrict__ c,
                                                              const uint2_* __restrict__ twB) {
    __shared__ u32 tiles[rns::NP * rns::TILE_WORDS];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t g = blockIdx.x;
    const u32 p = c_rns.p[w], pinv = c_rns.pinv_neg[w];
    u32* tile = tiles + w * rns::TILE_WORDS;
    const uint2_* twBf = twB + ((size_t)(w * 2 + 0) * 31) * 32 + lane;
    const uint2_* twBi = twB + ((size_t)(w * 2 + 1) * 31) * 32 + lane;
    u32 x[32], y[32];
#pragma unroll
    for (int r = 0; r < 32; r++) {
        x[r] = rns::residue_i64(a[g * N + 32 * r + lane], p);
        y[r] = rns::residue_i64(b[g * N + 32 * r + lane], p);
    }
    warp_ntt_fwd(x, tile, c_rns.twA[w][0], twBf, p, lane);
    warp_ntt_fwd(y, tile, c_rns.twA[w][0], twBf, p, lane);
#pragma unroll
    for (int cc = 0; cc < 32; cc++) {
        const u32 ks = rns::mulmod(y[cc] % p, c_rns.key_scale[w], p);
        x[cc] = rns::mont_mul(x[cc], ks, p, pinv);
    }
    warp_ntt_inv(x, tile, c_rns.twA[w][1], twBi, p, lane);
#pragma unroll
    for (int r = 0; r < 32; r++) tile[32 * r + lane] = x[r];
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += NM_THREADS)
        c[g * N + i] = (int64_t)rns::crt_lift(tiles[i], tiles[rns::TILE_WORDS + i], tiles[2 * rns::TILE_WORDS + i], c_rns.crt);
}

// Multi-key LWE key switch (mk_keyswitch_3gen mk_internals.jl:730-744, keyswitch keyswitch.jl:45-80).
// One CTA per sample; thread c owns output columns c, c+KS_THREADS, ... of the (n+1)-wide rows.
// ksk: int32 [k][N][t][B-1][ks_row_stride(n)] (rows padded to 16 bytes).
constexpr int KS_THREADS = 256;
constexpr int KS_MAXCOLS = 4;   // n + 1 <= 1024
__global__ void __launch_bounds__(KS_THREADS) keyswitch_kernel(int n, int k, int t, int basebit, const int32_t* __restrict__ ksk,
                                                                const int32_t* __restrict__ ext, int32_t* __restrict__ oa, int32_t* __restrict__ ob) {
    __shared__ uint32_t s_a[N];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int B1 = (1 << basebit) - 1, row = n + 1, stride = ks_row_stride(n);
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));   // keyswitch.jl:58
    const int32_t* e = ext + (size_t)g * (N + 1);
    for (int i = tid; i < N; i += KS_THREADS) s_a[i] = (uint32_t)e[i] + prec_offset;   // :59
    __syncthreads();
    const int ncols = (row + KS_THREADS - 1) / KS_THREADS;
    uint32_t bsum = 0;
    for (int p = 0; p < k; p++) {
        uint32_t acc[KS_MAXCOLS];
#pragma unroll
        for (int c = 0; c < KS_MAXCOLS; c++) acc[c] = 0;
        const int32_t* rows = ksk + (size_t)p * N * t * B1 * stride;
        for (int i = 0; i < N; i++) {
            const uint32_t ai = s_a[i];
            for (int j = 1; j <= t; j++) {
                const uint32_t d = (ai >> (32 - j * basebit)) & (uint32_t)B1;   // :65-67
                if (d != 0) {                                                 // :74-76
                    const int32_t* r = rows + (((size_t)i * t + (j - 1)) * B1 + (d - 1)) * stride;
#pragma unroll
                    for (int c = 0; c < KS_MAXCOLS; c++) {
                        const int col = tid + c * KS_THREADS;
                        if (c < ncols && col < row) acc[c] -= (uint32_t)__ldg(r + col);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < KS_MAXCOLS; c++) {
            const int col = tid + c * KS_THREADS;
            if (c < ncols && col < n) oa[((size_t)g * k + p) * n + col] = (int32_t)acc[c];
            if (c < ncols && col == n) bsum += acc[c];
        }
    }
    if (tid == n % KS_THREADS) ob[g] = (int32_t)((uint32_t)e[N] + bsum);   // thread owning column n
}

// Multi-key key switch of the CCS scheme (mk_keyswitch, mk_internals.jl:703-719): the extracted sample carries ONE MASK PER PARTY
// (ext_a [G][k][N]; the 3gen extraction has a single mask), party p's mask goes through party p's key: a_out[:, p] = keyswitch(ks[p],
// (a[:, p], 0)).a, b_out = b + sum_p keyswitch(...).b.  Same row layout and thread mapping as keyswitch_kernel.
__global__ void __launch_bounds__(KS_THREADS) mk_keyswitch_kernel(int n, int k, int t, int basebit, const int32_t* __restrict__ ksk,
                                                                   const int32_t* __restrict__ ext_a, const int32_t* __restrict__ ext_b,
                                                                   int32_t* __restrict__ oa, int32_t* __restrict__ ob) {
    __shared__ uint32_t s_a[N];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int B1 = (1 << basebit) - 1, row = n + 1, stride = ks_row_stride(n);
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));   // keyswitch.jl:58
    const int ncols = (row + KS_THREADS - 1) / KS_THREADS;
    uint32_t bsum = 0;
    for (int p = 0; p < k; p++) {
        __syncthreads();
        const int32_t* e = ext_a + ((size_t)g * k + p) * N;
        for (int i = tid; i < N; i += KS_THREADS) s_a[i] = (uint32_t)e[i] + prec_offset;
        __syncthreads();
        uint32_t acc[KS_MAXCOLS];
#pragma unroll
        for (int c = 0; c < KS_MAXCOLS; c++) acc[c] = 0;
        const int32_t* rows = ksk + (size_t)p * N * t * B1 * stride;
        for (int i = 0; i < N; i++) {
            const uint32_t ai = s_a[i];
            for (int j = 1; j <= t; j++) {
                const uint32_t d = (ai >> (32 - j * basebit)) & (uint32_t)B1;
                if (d != 0) {
                    const int32_t* r = rows + (((size_t)i * t + (j - 1)) * B1 + (d - 1)) * stride;
#pragma unroll
                    for (int c = 0; c < KS_MAXCOLS; c++) {
                        const int col = tid + c * KS_THREADS;
                        if (c < ncols && col < row) acc[c] -= (uint32_t)__ldg(r + col);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < KS_MAXCOLS; c++) {
            const int col = tid + c * KS_THREADS;
            if (c < ncols && col < n) oa[((size_t)g * k + p) * n + col] = (int32_t)acc[c];
            if (c < ncols && col == n) bsum += acc[c];
        }
    }
    if (tid == n % KS_THREADS) ob[g] = (int32_t)((uint32_t)ext_b[g] + bsum);
}

// ---- CCS multi-key blind rotation (mk_internals.jl:793-850) as a composition: the elementwise kernels around the two rounds of
// G (k+1) external products per step (extprod_t32_kernel).  acc: u64 [G][k+1][N] = (b, a_1 .. a_k), Torus32 values in the top halves;
// xin / xout: the external products' operands u64 [G (k+1)][2][N] ([0] = mask, [1] = body); elem: their key elements.
constexpr int CCS_THREADS = 256;
// acc = (X^-barb * testvect, 0, .., 0)   (mk_blind_rotate_and_extract :826-832)
__global__ void __launch_bounds__(CCS_THREADS) ccs_init_kernel(u64* __restrict__ acc, const int32_t* __restrict__ b_in, int k, int64_t mu) {
    const int g = blockIdx.x;
    const int s = (-mod_switch_2N(b_in[g])) & (2 * N - 1);
    u64* a = acc + (size_t)g * (k + 1) * N;
    for (int i = threadIdx.x; i < (k + 1) * N; i += CCS_THREADS) {
        const int c = i & (N - 1);
        const int idx = (c - s) & (2 * N - 1);
        a[i] = i < N ? ((idx & N) ? (u64)0 - (u64)mu : (u64)mu) : 0;
    }
}
// round 1 operands of the step of key bit (party, j): body_i = X^a acc_i - acc_i with a = decode_message(a_in[g][party][j], 2N)
// (mk_mux_rotate :793-800); element (party (k+1) + i) n + j
__global__ void __launch_bounds__(CCS_THREADS) ccs_round1_kernel(const u64* __restrict__ acc, const int32_t* __restrict__ a_in, u64* __restrict__ xin,
                                                                  int32_t* __restrict__ elem, int k, int n, int party, int j) {
    const int g = blockIdx.x, i = blockIdx.y;
    const int a = mod_switch_2N(a_in[((size_t)g * k + party) * n + j]);
    const u64* p = acc + ((size_t)g * (k + 1) + i) * N;
    u64* body = xin + (((size_t)g * (k + 1) + i) * 2 + 1) * N;
    for (int c = threadIdx.x; c < N; c += CCS_THREADS) {
        const int idx = (c - a) & (2 * N - 1);
        u64 v = p[idx & (N - 1)];
        if (idx & N) v = 0 - v;
        body[c] = v - p[c];
    }
    if (threadIdx.x == 0) elem[g * (k + 1) + i] = (party * (k + 1) + i) * n + j;
}
// after round 1 (xout = (v_i, u_i)): acc_i += u_i; round 2 operands body_i = v_i; element (k (k+1) + party) n + j
__global__ void __launch_bounds__(CCS_THREADS) ccs_round2_kernel(u64* __restrict__ acc, const u64* __restrict__ xout, u64* __restrict__ xin,
                                                                  int32_t* __restrict__ elem, int k, int n, int party, int j) {
    const int g = blockIdx.x, i = blockIdx.y;
    const size_t e = (size_t)g * (k + 1) + i;
    u64* p = acc + e * N;
    for (int c = threadIdx.x; c < N; c += CCS_THREADS) {
        p[c] += xout[(e * 2 + 1) * N + c];
        xin[(e * 2 + 1) * N + c] = xout[(e * 2) * N + c];
    }
    if (threadIdx.x == 0) elem[e] = (k * (k + 1) + party) * n + j;
}
// after round 2 (xout = (w1_i, w0_i)): b += sum_i w0_i, a_party += sum_i w1_i   (UniProduct :529-531)
__global__ void __launch_bounds__(CCS_THREADS) ccs_accumulate_kernel(u64* __restrict__ acc, const u64* __restrict__ xout, int k, int party) {
    const int g = blockIdx.x;
    u64* b = acc + (size_t)g * (k + 1) * N;
    u64* ap = b + (size_t)(1 + party) * N;
    for (int c = threadIdx.x; c < N; c += CCS_THREADS) {
        u64 w0 = 0, w1 = 0;
        for (int i = 0; i <= k; i++) {
            const size_t e = (size_t)g * (k + 1) + i;
            w1 += xout[(e * 2) * N + c];
            w0 += xout[(e * 2 + 1) * N + c];
        }
        b[c] += w0;
        ap[c] += w1;
    }
}
// mk_rlwe_extract_sample (mk_internals.jl:145-154): per party a'_0 = a_0, a'_m = -a_{N-m}; b' = b_0 (Torus32 = top halves)
__global__ void __launch_bounds__(CCS_THREADS) ccs_extract_kernel(const u64* __restrict__ acc, int32_t* __restrict__ ext_a, int32_t* __restrict__ ext_b, int k) {
    const int g = blockIdx.x;
    const u64* b = acc + (size_t)g * (k + 1) * N;
    for (int i = threadIdx.x; i < k * N; i += CCS_THREADS) {
        const int p = i / N, m = i & (N - 1);
        const u64* a = b + (size_t)(1 + p) * N;
        const u64 v = m == 0 ? a[0] : (u64)0 - a[N - m];
        ext_a[((size_t)g * k + p) * N + m] = (int32_t)(v >> 32);
    }
    if (threadIdx.x == 0) ext_b[g] = (int32_t)(b[0] >> 32);
}

// ---- key generation on the device: the key-switching key (keyswitch.jl:14-41) -------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter-based: row and column of a key word are its counter, so the key is a pure function
// of (seed, party) and no generator state is kept.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const u32 hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const u32 hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
// standard normal from two 32-bit words (Box-Muller, double precision)
__device__ __forceinline__ double philox_normal(u32 a, u32 b) {
    const double u1 = ((double)a + 1.0) * (1.0 / 4294967296.0), u2 = (double)b * (1.0 / 4294967296.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
// pass 1: one Gaussian per row (row = (i * t + j) * B1 + h) and their sum, for the re-centring of keyswitch.jl:28-29
__global__ void ksk_noise_kernel(double* __restrict__ noise, double* __restrict__ sum, size_t rows, double sigma, uint2 key) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (r < rows) {
        const uint4 x = philox4x32(make_uint4((u32)r, (u32)(r >> 32), 0u, 0xE0150000u), key);
        v = philox_normal(x.x, x.y) * sigma;
        noise[r] = v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, v);
}
// pass 2: one warp per row: a = n uniform words, b = message(i, j, h) + dtot32(noise - mean) + <a, s>  (keyswitch.jl:35-39, lwe.jl:47-53);
// rows are written in the device layout of the fused key switch (padded to ks_row_stride(n) words)
constexpr int KG_WARPS = 8;
__global__ void __launch_bounds__(32 * KG_WARPS) ksk_generate_kernel(int32_t* __restrict__ rows_out, const double* __restrict__ noise,
                                                                      const double* __restrict__ sum, const int32_t* __restrict__ s,
                                                                      const int64_t* __restrict__ z, int n, int t, int basebit, size_t rows, uint2 key) {
    const size_t r = (size_t)blockIdx.x * KG_WARPS + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31, B1 = (1 << basebit) - 1, stride = ks_row_stride(n);
    const int h = (int)(r % B1) + 1, j = (int)((r / B1) % t) + 1;
    const size_t i = r / ((size_t)B1 * t);
    int32_t* row = rows_out + r * stride;
    u32 dot = 0;
    for (int c0 = 4 * lane; c0 < stride; c0 += 128) {
        const uint4 x = philox4x32(make_uint4((u32)r, (u32)(r >> 32), (u32)(c0 >> 2), 0xA0000000u), key);
        const u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = c0 + q;
            if (c < n) {
                row[c] = (int32_t)w[q];
                dot += w[q] * (u32)s[c];
            } else if (c > n) {
                row[c] = 0;                                    // padding
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) {
        const double e = noise[r] - *sum / (double)rows;                                        // keyswitch.jl:29
        const u32 msg = (u32)((int32_t)z[i] * h) << (32 - j * basebit);                           // :35 (Int32 wrap)
        row[n] = (int32_t)(msg + (u32)__double2int_rz(e * 4294967296.0) + dot);                   // dtot32: trunc(Int32, d * 2^32)
    }
}

}  // namespace mk
