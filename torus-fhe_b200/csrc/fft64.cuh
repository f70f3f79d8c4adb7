// fft64.cuh -- the external product of the blind rotation on the FP64 pipe: every N = 1024 parameter set (Torus64 keys with byte-sized gadget
// digits; Torus32 mode with digits of up to 13 bits).
//
//   blind_rotate_fft_kernel<L, GPC>       gate prologue + mod-switch + k n mux-rotate steps + extraction + fused key switch; two six-warp gates per
//                                         CTA (throughput) or one (batches and tails of at most one gate per SM)
//   blind_rotate_fft_t32_kernel<L, GPC>   the same in Torus32 mode (two 16-bit key limbs, 16-bit digit fields, products added as R << 32)
//   extprod_fft_kernel<L, GPC, T32, BODY> one external product per gate slot (parity hook; BODY: the rounds of the CCS hybrid product)
//   bsk_transform_fft_kernel              one-time: int64 key polynomials -> limb spectra in the streaming layout
//
// Same exact result as the three-prime NTT kernels of kernels.cuh (every product mod (X^N + 1, 2^64), bit for bit), computed the way
// the reference computes it -- a folded complex FFT (3-gen-mk-tfhe/src/polynomials.jl:208-242) -- but made EXACT: the Torus64 key word
// is split in three balanced limbs of 22 / 21 / 21 bits, so every limb product sum_s digit_s * limb_s is an integer below
// 2 l N (Bg / 2) 2^21 <= 2^39, which a double-precision FFT of 512 complex points reproduces to within 2^-13 of an integer (measured
// worst case, tools/fft_channel/proto.py; the norm-wise rounding bound of the three transforms is about 2^-7) -- rounding recovers it exactly and
// R = r0 + (r1 << 22) + (r2 << 43) mod 2^64 is the wrap the reference's Int64 arithmetic performs (tgsw_3gen.jl:102-113).
//
// Why: per gate and step the NTT formulation needs 12 forward + 6 inverse 1024-point transforms at 4 IMAD-pipe slots per butterfly
// (446 k slots); this one needs 4 forward + 6 inverse 512-point complex transforms at 6 DFMA-pipe slots per butterfly (about 212 k
// slots), on a pipe of the same width (DFMA 62.5 lanes / clk / SM against IMAD 62.4, profiles/fp64_pipe_ubench_r2.txt), with no
// modular corrections, no CRT, and the integer / load-store pipes left for the decomposition and the limb recombination.
//
// Transform networks (tools/fft_channel/proto.py is the numpy twin, tables_fft.h the table generator):
//   forward  Cooley-Tukey, natural -> bit-reversed, on C[X] / (X^512 - i): the negacyclic twist is merged into the group twiddles, stage 0
//            has the single twiddle exp(i pi / 4) and is formed straight from the digit bytes
//   inverse  decimation in time, bit-reversed -> natural: pass B' has compile-time twiddles; the last stage and the untwist are done
//            by the threads that round, recombine the limbs and update the accumulator
// A warp holds a 512-point transform as 16 complex values per thread: position = h 256 + r 16 + l16 (lane = 16 h + l16, register r) in the
// row layout, lane 16 + c (register c) in the column layout; one swizzled shared-memory transpose between the two.  Of the 15 twiddles a
// pass needs per thread, one per stage is loaded and the rest are compile-time multiples of it (fft64_core.cuh).
// The arithmetic core lives in fft64_core.cuh (__host__ __device__: tests/host_emu/fft64_emu.cpp runs it on the CPU); measured decisions are in
// DESIGN.md section 4d and profiles/ab_r2.txt.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "fft64_core.cuh"

namespace mkf {

constexpr int WPG = 6, TPG = 32 * WPG;

// per gate: Torus64 accumulator, packed digits (one byte per coefficient), and max(2l, 6) buffers of 512 complex values: the 2l digit spectra
// during the forward transforms and the multiply-accumulate, then (one gate barrier later) the 6 limb outputs.  Separate buffers for the
// two roles would fit at l = 2 and save that barrier, but measured 0.9 % slower: the 64 KB they cost are worth more as L1 for the key stream.
// Torus32 mode (wide = true): 16-bit digit fields (2 bytes per coefficient) and two key limbs, i.e. max(2l, 4) buffers.
__host__ __device__ constexpr int nbuf(int l, int nl = LIMBS) { return 2 * l > 2 * nl ? 2 * l : 2 * nl; }
__host__ __device__ constexpr size_t gate_bytes(int l, bool wide = false) {
    return (size_t)2 * N * 8 + (size_t)2 * l * N * (wide ? 2 : 1) + (size_t)nbuf(l, wide ? LIMBS_T32 : LIMBS) * M * 16;
}
__host__ __device__ constexpr int gpc_for(int l, bool wide = false) {
    int g = 2;
    while (g > 1 && (size_t)TW_BYTES + (size_t)g * gate_bytes(l, wide) > 227 * 1024) g--;
    return g;
}
__host__ __device__ constexpr size_t cta_bytes(int l, int gpc, bool wide = false) { return (size_t)TW_BYTES + (size_t)gpc * gate_bytes(l, wide); }

// row layout -> column layout and back through a 512-entry buffer, XOR-swizzled so that both sides are conflict-free
__device__ __forceinline__ void rows_to_cols(cpx (&v)[16], cpx* __restrict__ buf, int lane) {
    const int h = lane >> 4, l16 = lane & 15;
#pragma unroll
    for (int r = 0; r < 16; r++) buf[transpose_slot(h, r, l16)] = v[r];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 16; c++) v[c] = buf[transpose_slot(h, l16, c)];
    __syncwarp();
}
__device__ __forceinline__ void cols_to_rows(cpx (&v)[16], cpx* __restrict__ buf, int lane) {
    const int h = lane >> 4, l16 = lane & 15;
#pragma unroll
    for (int c = 0; c < 16; c++) buf[transpose_slot(h, l16, c)] = v[c];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = buf[transpose_slot(h, r, l16)];
    __syncwarp();
}

__device__ __forceinline__ void stage_tables(cpx* tw_s, const cpx* __restrict__ tw_g) {
    const double2* src = reinterpret_cast<const double2*>(tw_g);
    double2* dst = reinterpret_cast<double2*>(tw_s);
    for (int i = threadIdx.x; i < T_ENTRIES; i += blockDim.x) dst[i] = __ldg(src + i);
}

// Phase 1 of a step: rotate-subtract and gadget-decompose (tgsw.jl:112-138); digits biased to [0, Bg).  Byte fields: dig [2L][256] words, word j of
// a polynomial = the bytes of coefficients j, j + 256, j + 512, j + 768.  WIDE (16-bit fields): dig [2L][2][256] words, word (0, j) = coefficients
// (j, j + 512), word (1, j) = (j + 256, j + 768).  BODY: the mask operand is zero; only the body polynomial acc[1] is decomposed (s < L).
template <int L, bool MUX, bool WIDE = false, bool BODY = false>
__device__ __forceinline__ void decompose(const u64* __restrict__ acc, u32* __restrict__ dig, int a, int bgbit, int gtid) {
    u64 off = 0;
#pragma unroll
    for (int q = 1; q <= L; q++) off += ((u64)1 << (64 - q * bgbit)) << (bgbit - 1);   // tgsw.jl:24-30
    const u32 dmask = (1u << bgbit) - 1;
    for (int task = gtid; task < (BODY ? 256 : 512); task += TPG) {
        const int c = BODY ? 1 : task >> 8, j = task & 255;
        const u64* poly = acc + c * N;
        u32 packed[L], packed_b[WIDE ? L : 1];
#pragma unroll
        for (int q = 0; q < L; q++) packed[q] = 0;
#pragma unroll
        for (int q = 0; q < (WIDE ? L : 1); q++) packed_b[q] = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int i = j + 256 * b;
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
            for (int q = 0; q < L; q++) {
                const u32 d = (u32)(t >> (64 - (q + 1) * bgbit)) & dmask;
                if (WIDE) {
                    if (b & 1) packed_b[q] |= d << (8 * (b - 1));     // b = 1 -> low field, b = 3 -> high field of word (1, j)
                    else packed[q] |= d << (8 * b);                    // b = 0 -> low field, b = 2 -> high field of word (0, j)
                } else {
                    packed[q] |= d << (8 * b);
                }
            }
        }
        const int src = 1 - c;   // src 0 = body = acc[1]
#pragma unroll
        for (int q = 0; q < L; q++) {
            if (WIDE) {
                dig[((src * L + q) * 2 + 0) * 256 + j] = packed[q];
                dig[((src * L + q) * 2 + 1) * 256 + j] = packed_b[q];
            } else {
                dig[(src * L + q) * 256 + j] = packed[q];
            }
        }
    }
}

// ---- phases of one external product (tgsw_extern_mul_3gen, tgsw_3gen.jl:102-113) on the accumulator held in shared memory -----------------
//   MUX = true : acc += ExtProd(X^a * acc - acc, key)   (mk_mux_rotate_3gen, 3gen_mk_internals.jl:59-62)
//   MUX = false: acc  = ExtProd(acc, key)
// Per gate: buf = max(2L, 6) buffers [512]: the 2L digit spectra, later the 6 limb outputs (index lo = 2 limb + out).

// forward transforms of the 2L digit polynomials by the warps gw = s, s + 6, ..; spectrum point (c, lane) at c 32 + lane
template <int NS, bool WIDE>
__device__ __forceinline__ void forward_phase(const u32* __restrict__ dig, cpx* __restrict__ spec, const cpx* __restrict__ tw, int bgbit, int gw, int lane) {
#pragma unroll 1
    for (int s = gw; s < NS; s += WPG) {
        cpx v[16];
        if (WIDE) fwd_stage0_digits_wide(v, dig + s * 512, dig + s * 512 + 256, lane, 1 << (bgbit - 1));
        else fwd_stage0_digits(v, dig + s * 256, lane, 1 << (bgbit - 1));
        fwd_passA(v, tw, lane >> 4);
        rows_to_cols(v, spec + s * M, lane);
        fwd_passB(v, tw, lane);
#pragma unroll
        for (int c = 0; c < 16; c++) spec[s * M + c * 32 + lane] = v[c];
    }
}
// inverse transform of the 16 spectrum points of this thread up to the last stage; result in natural position order in buf
__device__ __forceinline__ void inverse_to_buffer(cpx (&v)[16], cpx* __restrict__ buf, const cpx* __restrict__ tw, int lane) {
    inv_passB(v);
    cols_to_rows(v, buf, lane);
    inv_passA(v, tw, lane & 15);
    const int h = lane >> 4, l16 = lane & 15;
#pragma unroll
    for (int r = 0; r < 16; r++) buf[h * 256 + r * 16 + l16] = v[r];
}
// last inverse stage, untwist, rounding, limb recombination, accumulator update: task (out, j) -> coefficients j + 256 b.
// NL = 2 (Torus32 mode): the exact product of the unshifted 32-bit keys is added as R << 32.
template <bool MUX, int NL>
__device__ __forceinline__ void recombine_task(u64* __restrict__ acc, const cpx* __restrict__ ybuf, const cpx wj, const cpx ut, int out, int j) {
    uint64_t R[4] = {0, 0, 0, 0};
#pragma unroll
    for (int limb = 0; limb < NL; limb++) {
        const cpx* Y = ybuf + (limb * 2 + out) * M;
        recombine_limb(R, Y[j], Y[j + 256], wj, ut, limb_shift(NL, limb));
    }
    u64* ap = acc + out * N + j;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const u64 r = R[b] - round_k(NL);
        ap[256 * b] = (MUX ? ap[256 * b] : 0) + (NL == 2 ? r << 32 : r);
    }
}
// 512 tasks on 192 threads: thread t takes j = t for both outputs (one pair of twiddle loads), and the j = 192 .. 255 left over go to
// threads 0 .. 127 as (j = 192 + t / 2, out = t & 1)
template <bool MUX, int NL>
__device__ __forceinline__ void recombine_phase(u64* __restrict__ acc, const cpx* __restrict__ ybuf, const cpx* __restrict__ tw, int gtid) {
    {
        const cpx wj = tw[T_WJ + gtid], ut = tw[T_UT + gtid];
        recombine_task<MUX, NL>(acc, ybuf, wj, ut, 0, gtid);
        recombine_task<MUX, NL>(acc, ybuf, wj, ut, 1, gtid);
    }
    if (gtid < 128) {
        const int j = 192 + (gtid >> 1);
        recombine_task<MUX, NL>(acc, ybuf, tw[T_WJ + j], tw[T_UT + j], gtid & 1, j);
    }
}

// smem pointers of one gate slot
struct GateMem {
    u64* acc;
    u32* dig;
    cpx* buf;      // max(2l, 6) x [512]: digit spectra, then limb outputs (index 2 limb + out)
};
template <int L, bool WIDE = false>
__device__ __forceinline__ GateMem gate_mem(unsigned char* smem_raw, int slot) {
    unsigned char* base = smem_raw + TW_BYTES + (size_t)slot * gate_bytes(L, WIDE);
    GateMem m;
    m.acc = reinterpret_cast<u64*>(base);
    m.dig = reinterpret_cast<u32*>(base + 2 * N * 8);
    m.buf = reinterpret_cast<cpx*>(base + 2 * N * 8 + 2 * L * N * (WIDE ? 2 : 1));
    return m;
}

// One step of ONE gate by its 192 threads (every warp streams its own key polynomials: each key value is loaded once per gate and each
// spectrum value six times per gate).
template <int L, bool MUX, int NL = LIMBS, bool WIDE = false, bool BODY = false>
__device__ __forceinline__ void extprod_step(const GateMem& m, const cpx* __restrict__ tw, const cpx* __restrict__ key, int a, int bgbit, int bar_id,
                                             int gtid) {
    constexpr int NS = BODY ? L : 2 * L;                            // digit polynomials (BODY: the body's only; they are s = 0 .. L-1)
    const int gw = gtid >> 5, lane = gtid & 31;
    const int limb = gw >> 1, out = gw & 1;                         // warps 0 .. 2 NL - 1 each own one limb output
    const double2* kp = reinterpret_cast<const double2*>(key) + ((size_t)out * NL + limb) * M + lane;   // [s][out][limb][512]
    decompose<L, MUX, WIDE, BODY>(m.acc, m.dig, a, bgbit, gtid);
    mk::gate_barrier<WPG>(bar_id);
    forward_phase<NS, WIDE>(m.dig, m.buf, tw, bgbit, gw, lane);
    mk::gate_barrier<WPG>(bar_id);
    {   // warp (limb, out): multiply-accumulate over the spectra, inverse transform up to the last stage
        cpx v[16];
        if (NL == LIMBS || gw < 2 * NL) {
#pragma unroll
            for (int c = 0; c < 16; c++) v[c] = cpx{0.0, 0.0};
            // every key register is refilled for the next digit polynomial as soon as it is consumed: the phase waits on L2 once per step
            double2 kreg[16];
#pragma unroll
            for (int c = 0; c < 16; c++) kreg[c] = __ldg(kp + c * 32);
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const double2* kn = kp + (size_t)(s + 1) * (2 * NL * M);
                const cpx* xs = m.buf + s * M + lane;
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    const double2 k = kreg[c];
                    if (s + 1 < NS) kreg[c] = __ldg(kn + c * 32);
                    const cpx x = xs[c * 32];
                    v[c].x = fma(x.x, k.x, fma(-x.y, k.y, v[c].x));
                    v[c].y = fma(x.x, k.y, fma(x.y, k.x, v[c].y));
                }
            }
        }
        mk::gate_barrier<WPG>(bar_id);                              // every warp is done reading the spectra: the buffers change roles
        if (NL == LIMBS || gw < 2 * NL) inverse_to_buffer(v, m.buf + gw * M, tw, lane);
    }
    mk::gate_barrier<WPG>(bar_id);
    recombine_phase<MUX, NL>(m.acc, m.buf, tw, gtid);
    mk::gate_barrier<WPG>(bar_id);
}

// GPC gates per CTA, six warps per gate; prologue, k n steps, extraction and key switch as in mk::blind_rotate_body
template <int L, int GPC, int NL = LIMBS, bool WIDE = false>
__device__ __forceinline__ void blind_rotate_body(const mk::BlindRotateArgs& p, const cpx* __restrict__ key_fft, const cpx* __restrict__ tw_g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* tw = reinterpret_cast<cpx*>(smem_raw);
    stage_tables(tw, tw_g);
    __syncthreads();
    const int slot = threadIdx.x / TPG, gtid = threadIdx.x - slot * TPG, bar_id = 1 + slot;
    const int g = p.g0 + blockIdx.x * GPC + slot;
    if (g >= p.G) return;   // no CTA-wide barrier below this line
    const GateMem m = gate_mem<L, WIDE>(smem_raw, slot);
    u64* acc = m.acc;
    const int kn = p.k * p.n;
    const mk::GateLinear lin = p.gate_ids ? mk::gate_linear(__ldg(p.gate_ids + g)) : p.lin;
    uint32_t tb = (uint32_t)lin.mu0 + (uint32_t)lin.cx * (uint32_t)__ldg(p.xb + g);
    if (lin.cy) tb += (uint32_t)lin.cy * (uint32_t)__ldg(p.yb + g);
    if (lin.cz) tb += (uint32_t)lin.cz * (uint32_t)__ldg(p.zb + g);
    const int barb = mk::mod_switch_2N((int32_t)tb);
    // acc = (0, X^{-barb} * testvect), testvect = mu * (1 + X + ... + X^{N-1})  (3gen_mk_internals.jl:88-92, rlwe.jl:113-119)
    {
        const int s = (-barb) & (2 * N - 1);
        for (int i = gtid; i < N; i += TPG) {
            const int idx = (i - s) & (2 * N - 1);
            acc[i] = 0;
            acc[N + i] = (idx & N) ? (u64)0 - (u64)p.mu : (u64)p.mu;
        }
    }
    mk::gate_barrier<WPG>(bar_id);
    const size_t estride = bsk_elem_cpx(L, NL);
    const size_t abase = (size_t)g * kn;
    int32_t rx = __ldg(p.xa + abase), ry = lin.cy ? __ldg(p.ya + abase) : 0, rz = lin.cz ? __ldg(p.za + abase) : 0;
    for (int it = 0; it < kn; it++) {
        const int a = mk::mod_switch_2N((int32_t)((uint32_t)lin.cx * (uint32_t)rx + (uint32_t)lin.cy * (uint32_t)ry + (uint32_t)lin.cz * (uint32_t)rz));
        if (it + 1 < kn) {
            rx = __ldg(p.xa + abase + it + 1);
            if (lin.cy) ry = __ldg(p.ya + abase + it + 1);
            if (lin.cz) rz = __ldg(p.za + abase + it + 1);
        }
        if (a == 0) continue;   // 3gen_mk_internals.jl:69 (uniform across the gate)
        extprod_step<L, true, NL, WIDE>(m, tw, key_fft + (size_t)it * estride, a, p.bgbit, bar_id, gtid);
    }
    if (p.acc_out) {
        int64_t* ao = p.acc_out + (size_t)g * 2 * N;
        for (int i = gtid; i < 2 * N; i += TPG) ao[i] = (int64_t)acc[i];
    }
    if (p.ksk) {   // fused extraction + key switch (the host only sets ksk when mk::ks_fusable(n, t)); scratch: the spectra
        if (p.ks_t == 3) mk::fused_keyswitch<3, WPG>(acc, reinterpret_cast<u32*>(m.buf), p, g, gtid, bar_id);
        else mk::fused_keyswitch<5, WPG>(acc, reinterpret_cast<u32*>(m.buf), p, g, gtid, bar_id);
        return;
    }
    int32_t* ext = p.ext_out + (size_t)g * (N + 1);
    for (int i = gtid; i < N; i += TPG) {
        const u64 v = i == 0 ? acc[0] : (u64)0 - acc[N - i];
        ext[i] = mk::t64tot32((int64_t)v);
    }
    if (gtid == 0) ext[N] = mk::t64tot32((int64_t)acc[N]);
}

#ifndef MKF_MAXNREG
#define MKF_MAXNREG 168
#endif
template <int L, int GPC>
__global__ void __maxnreg__(MKF_MAXNREG) blind_rotate_fft_kernel(mk::BlindRotateArgs p, const cpx* __restrict__ key_fft, const cpx* __restrict__ tw_g) {
    blind_rotate_body<L, GPC>(p, key_fft, tw_g);
}
// Torus32 mode (mktfhe_params.flags & MKTFHE_FLAG_TORUS32): 16-bit digit fields, two 16-bit limbs of the unshifted 32-bit keys, products added
// as R << 32.  Exact while 2 l N (Bg / 2) 2^15 <= 2^40; serves tfhe_parameters_80 (Bg = 2^10) and the gadget shapes of the CCS scheme.
template <int L, int GPC>
__global__ void __maxnreg__(MKF_MAXNREG) blind_rotate_fft_t32_kernel(mk::BlindRotateArgs p, const cpx* __restrict__ key_fft, const cpx* __restrict__ tw_g) {
    blind_rotate_body<L, GPC, LIMBS_T32, true>(p, key_fft, tw_g);
}

// parity hook: acc_out[g] = ExtProd(acc_in[g], key[elem[g]]).  T32: Torus32 mode (values in the top halves of the words); BODY: the mask
// operand is zero and neither read nor decomposed (the rounds of the CCS hybrid product)
template <int L, int GPC, bool T32 = false, bool BODY = false>
__global__ void __maxnreg__(MKF_MAXNREG) extprod_fft_kernel(int G, const cpx* __restrict__ key_fft, const cpx* __restrict__ tw_g, int bgbit,
                                                             const int32_t* __restrict__ elem, const int64_t* __restrict__ acc_in, int64_t* __restrict__ acc_out) {
    constexpr int NL = T32 ? LIMBS_T32 : LIMBS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* tw = reinterpret_cast<cpx*>(smem_raw);
    stage_tables(tw, tw_g);
    __syncthreads();
    const int slot = threadIdx.x / TPG, gtid = threadIdx.x - slot * TPG, bar_id = 1 + slot;
    const int g = blockIdx.x * GPC + slot;
    if (g >= G) return;
    const GateMem m = gate_mem<L, T32>(smem_raw, slot);
    for (int i = gtid + (BODY ? N : 0); i < 2 * N; i += TPG) m.acc[i] = (u64)acc_in[(size_t)g * 2 * N + i];
    mk::gate_barrier<WPG>(bar_id);
    extprod_step<L, false, NL, T32, BODY>(m, tw, key_fft + (size_t)elem[g] * bsk_elem_cpx(L, NL), 0, bgbit, bar_id, gtid);
    for (int i = gtid; i < 2 * N; i += TPG) acc_out[(size_t)g * 2 * N + i] = (int64_t)m.acc[i];
}

// One warp per (key polynomial, limb): raw int64 key -> spectrum of the limb, scaled by 1 / 512, in the FFT layout.
// raw: [n][4 parts][l][N] int64 of one party; task = ((j*4 + part)*l + q)*nl + limb; nl = 3 (Torus64 words) or 2 (Torus32 mode: 32-bit values).
constexpr int XF_WARPS = 4;
__global__ void __launch_bounds__(XF_WARPS * 32) bsk_transform_fft_kernel(const int64_t* __restrict__ raw, cpx* __restrict__ bsk, int n, int l, int nl, int party,
                                                                            const cpx* __restrict__ tw_g, int ntasks) {
    __shared__ cpx bufs[XF_WARPS * M];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * XF_WARPS + warp;
    if (task >= ntasks) return;
    const int limb = task % nl, pq = task / nl;
    const int q = pq % l, part = (pq / l) & 3, j = pq / (4 * l);
    // part_1: body<-body, part_2: body<-mask, part_3: mask<-mask, part_4: mask<-body  (tgsw_3gen.jl:109-110)
    const int out = part < 2 ? 1 : 0;
    const int src = (part == 0 || part == 3) ? 0 : 1;
    const int64_t* poly = raw + (size_t)pq * N;
    const int h = lane >> 4, l16 = lane & 15;
    cpx v[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int jj = 16 * r + l16;
        v[r] = fwd_stage0_real(key_limb(poly[jj], limb, nl), key_limb(poly[jj + 256], limb, nl), key_limb(poly[jj + 512], limb, nl), key_limb(poly[jj + 768], limb, nl), h);
    }
    fwd_passA(v, tw_g, h);
    rows_to_cols(v, bufs + warp * M, lane);
    fwd_passB(v, tw_g, lane);
    const size_t e = (size_t)party * n + j;
    cpx* dst = bsk + e * bsk_elem_cpx(l, nl) + (((size_t)(src * l + q) * 2 + out) * nl + limb) * M;
#pragma unroll
    for (int c = 0; c < 16; c++) dst[c * 32 + lane] = cpx{v[c].x * (1.0 / M), v[c].y * (1.0 / M)};
}

}  // namespace mkf
