// kernels2k.cuh -- the bootstrapped-gate path for the N = 2048 parameter sets (mktfhe_parameters_{16,32,64,128}party_3gen with l = 1,
// mktfhe_parameters_256party_3gen with l = 2; mk_api.jl:214-310): gadget digits of up to 27 bits, four 28-bit primes (ntt2048.cuh),
// one gate per CTA.
//
// First, functional version of this path: same step structure as kernels.cuh (decompose -> forward transforms -> multiply-accumulate
// -> inverse transforms -> CRT), eight warps per gate = (prime, polynomial), the accumulator resident in shared memory for all k*n
// steps.  Not tuned yet: per-lane twiddles come from global memory (their 127 KB do not fit beside the tiles), the key switch is the
// stand-alone kernel.  Reference semantics as in kernels.cuh (3gen_mk_internals.jl:59-116, tgsw_3gen.jl:102-113, tgsw.jl:112-138,
// rlwe.jl:70-74, keyswitch.jl:45-80).
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "ntt2048.cuh"

#ifndef MK2K_CRT_UNROLL
#define MK2K_CRT_UNROLL 4   // Garner chains of the CRT phase interleaved per thread (16 coefficients per thread and step); A/B 1 / 4 / 16 in profiles/ab_r1.txt
#endif

namespace mk2k {

using rns::uint2_;
constexpr int N = rns2k::N, NP = rns2k::NP;
constexpr int WARPS = 2 * NP, THREADS = 32 * WARPS;          // warp = (prime, polynomial)
constexpr int TILE_WORDS = rns2k::TILE_WORDS, TILE_STRIDE = rns2k::TILE_STRIDE;
constexpr int CRT_UNROLL = MK2K_CRT_UNROLL;

__constant__ rns2k::Consts c_k2;

// BSK layout (u32): [elem = party*n + j][prime][s = src*l + q][out][2048 key slots]; (out, src) <-> reference parts as in kernels.cuh.
// key slot of transformed position 32 (lane + 32 h) + c: a warp-wide 128-bit load is one contiguous 512-byte segment
// BSK rows per (element, prime): s = src * l + q (digit polynomial) x out
__host__ __device__ inline size_t bsk_elem_words(int l) { return (size_t)NP * 2 * l * 2 * N; }
__host__ __device__ inline int key_slot(int lane, int h, int c) { return ((h * 8 + (c >> 2)) * 32 + lane) * 4 + (c & 3); }
__host__ __device__ constexpr size_t smem_bytes(int l) { return (size_t)2 * N * 8 + (size_t)2 * l * N * 4 + (size_t)WARPS * TILE_WORDS * 4; }

struct Args {
    int G, n, k, bgbit;
    const u32* bsk;
    const uint2_* twB;
    const int32_t *xa, *xb, *ya, *yb, *za, *zb;
    mk::GateLinear lin;
    const int32_t* gate_ids;
    int64_t mu;
    int32_t* ext_out;   // [G][N+1], or nullptr when the key switch is fused and nobody asked for the extracted samples
    int64_t* acc_out;   // [G][2][N] or nullptr
    // fused key switch (epilogue of the same kernel): ksk [k][N][t][B-1][ks_row_stride(n)] followed by one all-zero row
    const int32_t* ksk; // nullptr: no key switch in this launch
    int ks_t, ks_basebit;
    int32_t *oa, *ob;   // [G][k][n], [G]
};

// the fused epilogue covers the key-switch shapes of the N = 2048 sets (t = 4, 5, 8) when a row fits the CTA four columns per thread
__host__ __device__ constexpr bool ks_fusable(int n, int t) { return mk::ks_row_stride(n) <= 4 * THREADS && (t == 4 || t == 5 || t == 8); }

// decode_message(x, 2N) for N = 2048: (x + 2^19) >> 20
__device__ __forceinline__ int mod_switch_2N(int32_t x) { return (int32_t)((uint32_t)x + (1u << 19)) >> 20; }

// forward transform of this warp's polynomial: x[r] = a[32 r + lane] in [0, 2p) -> y[h][c] = position 32 (lane + 32 h) + c, < 14p
__device__ __forceinline__ void warp_fwd(u32 (&x)[64], u32 (&y)[2][32], u32* tile, const uint2_* twB, int pi, u32 p, int lane) {
    rns2k::fwd_passA64(x, c_k2.twA[pi][0], p);
#pragma unroll
    for (int r = 0; r < 64; r++) tile[r * TILE_STRIDE + lane] = x[r];
    __syncwarp();
    const u32 p8 = rns2k::opaque_multiple(8 * p), p4 = rns2k::opaque_multiple(4 * p);
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
        for (int c = 0; c < 32; c++) y[h][c] = rns2k::reduce_to_4p(tile[(lane + 32 * h) * TILE_STRIDE + c], p8, p4);
        rns2k::fwd_passB32(y[h], twB + rns2k::twB_index(pi, 0, h, 0, lane), p);
    }
    __syncwarp();
}
// inverse: y[h][c] in [0, 4p) -> x[r] = N * a[32 r + lane] in [0, 4p)
__device__ __forceinline__ void warp_inv(u32 (&y)[2][32], u32 (&x)[64], u32* tile, const uint2_* twB, int pi, u32 p, int lane) {
    const u32 p4 = rns2k::opaque_multiple(4 * p);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        rns2k::inv_passB32(y[h], twB + rns2k::twB_index(pi, 1, h, 0, lane), p, p4);
#pragma unroll
        for (int c = 0; c < 32; c++) tile[(lane + 32 * h) * TILE_STRIDE + c] = y[h][c];
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 64; r++) x[r] = tile[r * TILE_STRIDE + lane];
    __syncwarp();
    rns2k::inv_passA64(x, c_k2.twA[pi][1], p, p4);
}

// Sample extraction + multi-key key switch of the gate by its CTA (rlwe.jl:70-74, mk_internals.jl:730-744, keyswitch.jl:45-80), same
// scheme as mk::fused_keyswitch: thread c owns the four output columns 4c .. 4c+3 (one LDG.128 per row), zero digits read the all-zero
// row so that the row reads of four consecutive coefficients are in flight together.
template <int T>
__device__ __forceinline__ void fused_keyswitch2k(const u64* __restrict__ acc, u32* __restrict__ s_a, const Args& p, int g, int tid) {
    const int n = p.n, stride = mk::ks_row_stride(n), bb = p.ks_basebit, B1 = (1 << bb) - 1;
    const uint32_t prec_offset = 1u << (32 - (1 + bb * T));   // keyswitch.jl:58
    for (int i = tid; i < N; i += THREADS) {
        const u64 v = i == 0 ? acc[0] : (u64)0 - acc[N - i];
        const int32_t ai = mk::t64tot32((int64_t)v);
        if (p.ext_out) p.ext_out[(size_t)g * (N + 1) + i] = ai;
        s_a[i] = (uint32_t)ai + prec_offset;
    }
    const int32_t eb = mk::t64tot32((int64_t)acc[N]);
    if (p.ext_out && tid == 0) p.ext_out[(size_t)g * (N + 1) + N] = eb;
    __syncthreads();
    const int col0 = 4 * tid;
    if (col0 >= stride) return;                               // no barrier below
    const size_t party_words = (size_t)N * T * B1 * stride;
    const uint4* zero_row = reinterpret_cast<const uint4*>(p.ksk + (size_t)p.k * party_words) + tid;
    uint32_t bsum = 0;
    for (int party = 0; party < p.k; party++) {
        uint4 out = make_uint4(0, 0, 0, 0);
        const int32_t* rows = p.ksk + (size_t)party * party_words;
#pragma unroll 1
        for (int i = 0; i < N; i += 4) {
#pragma unroll
            for (int ii = 0; ii < 4; ii++) {
                const uint32_t ai = s_a[i + ii];
#pragma unroll
                for (int j = 1; j <= T; j++) {
                    const uint32_t d = (ai >> (32 - j * bb)) & (uint32_t)B1;
                    const uint4* r = d ? reinterpret_cast<const uint4*>(rows + (((size_t)(i + ii) * T + (j - 1)) * B1 + (d - 1)) * stride) + tid : zero_row;
                    const uint4 v = __ldg(r);
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
    if (col0 <= n && n < col0 + 4) p.ob[g] = (int32_t)((uint32_t)eb + bsum);
}

// one gate per CTA; acc[0] = mask, acc[1] = body.  Warp (prime w, o): transforms the digit polynomials 2i + o, i < L, one after the
// other, and accumulates output polynomial o over all 2L of them (its own from registers, the partner's from the partner's tile)
template <int L>
__global__ void __launch_bounds__(THREADS, 1) blind_rotate2k_kernel(Args p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);
    int32_t* dig = reinterpret_cast<int32_t*>(smem_raw + (size_t)2 * N * 8);           // [s = src*L + q][N] signed digits; src 0 = body, 1 = mask
    u32* tiles = reinterpret_cast<u32*>(smem_raw + (size_t)2 * N * 8 + (size_t)2 * L * N * 4);
    const int tid = threadIdx.x, gw = tid >> 5, lane = tid & 31;
    const int w = gw >> 1, o = gw & 1;                     // prime; output polynomial accumulated = parity of the digit polynomials transformed
    const int g = blockIdx.x;
    const u32 pr = c_k2.p[w], pinv = c_k2.pinv_neg[w];
    u32* tile = tiles + gw * TILE_WORDS;
    const u32* ptile = tiles + (gw ^ 1) * TILE_WORDS;
    const int kn = p.k * p.n;
    const mk::GateLinear lin = p.gate_ids ? mk::gate_linear(__ldg(p.gate_ids + g)) : p.lin;
    auto rotation = [&](const int32_t* xs, const int32_t* ys, const int32_t* zs, size_t idx, uint32_t mu0) {
        uint32_t t = mu0 + (uint32_t)lin.cx * (uint32_t)__ldg(xs + idx);
        if (lin.cy) t += (uint32_t)lin.cy * (uint32_t)__ldg(ys + idx);
        if (lin.cz) t += (uint32_t)lin.cz * (uint32_t)__ldg(zs + idx);
        return mod_switch_2N((int32_t)t);
    };
    {   // acc = (0, X^{-barb} * testvect)  (3gen_mk_internals.jl:88-92)
        const int barb = rotation(p.xb, p.yb, p.zb, g, (uint32_t)lin.mu0);
        const int sh = (-barb) & (2 * N - 1);
        for (int i = tid; i < N; i += THREADS) {
            const int idx = (i - sh) & (2 * N - 1);
            acc[i] = 0;
            acc[N + i] = (idx & N) ? (u64)0 - (u64)p.mu : (u64)p.mu;
        }
    }
    __syncthreads();
    u64 off = 0;
#pragma unroll
    for (int q = 1; q <= L; q++) off += ((u64)1 << (64 - q * p.bgbit)) << (p.bgbit - 1);   // tgsw.jl:24-30
    const u64 dmask = ((u64)1 << p.bgbit) - 1;
    const int64_t half = (int64_t)1 << (p.bgbit - 1);
    const size_t abase = (size_t)g * kn;
    for (int it = 0; it < kn; it++) {
        const int a = rotation(p.xa, p.ya, p.za, abase + it, 0u);
        if (a == 0) continue;                                              // 3gen_mk_internals.jl:69
        // ---- decompose X^a * acc - acc (tgsw.jl:112-138, l = 1)
        for (int i = tid; i < 2 * N; i += THREADS) {
            const int c = i >> 11, ii = i & (N - 1);
            const u64* poly = acc + c * N;
            const int idx = (ii - a) & (2 * N - 1);
            u64 v = poly[idx & (N - 1)];
            if (idx & N) v = 0 - v;
            const u64 t = v - poly[ii] + off;
#pragma unroll
            for (int q = 0; q < L; q++) dig[((1 - c) * L + q) * N + ii] = (int32_t)((int64_t)((t >> (64 - (q + 1) * p.bgbit)) & dmask) - half);
        }
        __syncthreads();
        // ---- forward transforms of the digit polynomials 2i + o under prime w; exchange with the partner (w, 1 - o); multiply-accumulate
        u32 x[64], y[2][32], accv[2][32];
        const u32* kw = p.bsk + (size_t)it * bsk_elem_words(L) + (size_t)w * (2 * L * 2 * N);
#pragma unroll 1
        for (int i = 0; i < L; i++) {
            const int s_own = 2 * i + o, s_for = 2 * i + 1 - o;
#pragma unroll
            for (int r = 0; r < 64; r++) x[r] = (u32)(dig[s_own * N + 32 * r + lane] + (int32_t)pr);      // signed digit + p in [0, 2p)
            warp_fwd(x, y, tile, p.twB, w, pr, lane);
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int c = 0; c < 32; c++) tile[(h * 32 + c) * 32 + lane] = y[h][c];
            asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");
            // both digit polynomials of the pair times their key rows of output o, one Montgomery reduction per point
            const uint4* k_own = reinterpret_cast<const uint4*>(kw + (size_t)(s_own * 2 + o) * N) + lane;
            const uint4* k_for = reinterpret_cast<const uint4*>(kw + (size_t)(s_for * 2 + o) * N) + lane;
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int q4 = 0; q4 < 8; q4++) {
                    const uint4 ka = __ldg(k_own + (h * 8 + q4) * 32), kb = __ldg(k_for + (h * 8 + q4) * 32);
                    const u32* pt = ptile + (h * 32 + 4 * q4) * 32 + lane;
                    const u32 v0 = rns::mont_mul2(y[h][4 * q4 + 0], ka.x, pt[0], kb.x, pr, pinv);       // < 2.75p
                    const u32 v1 = rns::mont_mul2(y[h][4 * q4 + 1], ka.y, pt[32], kb.y, pr, pinv);
                    const u32 v2 = rns::mont_mul2(y[h][4 * q4 + 2], ka.z, pt[64], kb.z, pr, pinv);
                    const u32 v3 = rns::mont_mul2(y[h][4 * q4 + 3], ka.w, pt[96], kb.w, pr, pinv);
                    accv[h][4 * q4 + 0] = i == 0 ? v0 : accv[h][4 * q4 + 0] + v0;
                    accv[h][4 * q4 + 1] = i == 0 ? v1 : accv[h][4 * q4 + 1] + v1;
                    accv[h][4 * q4 + 2] = i == 0 ? v2 : accv[h][4 * q4 + 2] + v2;
                    accv[h][4 * q4 + 3] = i == 0 ? v3 : accv[h][4 * q4 + 3] + v3;
                }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + w) : "memory");          // the partner is done with this warp's tile
        }
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const u32 v = accv[h][c];                                      // < 2.75 L p
                y[h][c] = L > 1 ? rns::umin32(v, v - 4 * pr) : v;              // [0, 4p)
            }
        // ---- inverse transform, residues to the tile in coefficient order, CRT by the whole gate
        warp_inv(y, x, tile, p.twB, w, pr, lane);
#pragma unroll
        for (int r = 0; r < 64; r++) tile[32 * r + lane] = x[r];
        __syncthreads();
#pragma unroll CRT_UNROLL
        for (int i = tid; i < 2 * N; i += THREADS) {
            const int o = i >> 11, ii = i & (N - 1);
            const u32 r[NP] = {tiles[(0 * 2 + o) * TILE_WORDS + ii], tiles[(1 * 2 + o) * TILE_WORDS + ii], tiles[(2 * 2 + o) * TILE_WORDS + ii],
                               tiles[(3 * 2 + o) * TILE_WORDS + ii]};
            acc[i] += rns2k::lift4(r, c_k2);
        }
        __syncthreads();
    }
    if (p.acc_out) {
        int64_t* ao = p.acc_out + (size_t)g * 2 * N;
        for (int i = tid; i < 2 * N; i += THREADS) ao[i] = (int64_t)acc[i];
    }
    if (p.ksk) {   // fused extraction + key switch (the host only sets ksk when ks_fusable(n, t)); the tiles are free now
        if (p.ks_t == 4) fused_keyswitch2k<4>(acc, tiles, p, g, tid);
        else if (p.ks_t == 5) fused_keyswitch2k<5>(acc, tiles, p, g, tid);
        else fused_keyswitch2k<8>(acc, tiles, p, g, tid);
        return;
    }
    // rlwe_extract_sample_64 (rlwe.jl:70-74): a'_0 = mask_0, a'_i = -mask_{N-i}, b' = body_0
    int32_t* ext = p.ext_out + (size_t)g * (N + 1);
    for (int i = tid; i < N; i += THREADS) {
        const u64 v = i == 0 ? acc[0] : (u64)0 - acc[N - i];
        ext[i] = mk::t64tot32((int64_t)v);
    }
    if (tid == 0) ext[N] = mk::t64tot32((int64_t)acc[N]);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Second-generation kernel (round 2): SIXTEEN warps per gate, warp = (prime w, output o, half h), 32 coefficients per thread.
// The first kernel above keeps one polynomial in one warp (64 coefficients per thread, 217..255 registers, two warps per scheduler):
// ncu showed the fma pipe 59 % busy with `wait` and `math throttle` on top -- too few warps to hide the butterfly chains.  Here a
// 2048-point transform is split over the two warps (w, o, 0) and (w, o, 1) by the TOP bit of the coefficient row r (coefficient
// 32 r + lane, r = 32 h + r'):
//   forward: stage 0 pairs rows r' and r' + 32, i.e. the two halves: both warps read both digits from shared memory and each keeps its
//            own output (X + w0 Y for h = 0, X - w0 Y for h = 1; the product is computed by both -- 1024 of 11264 products twice); the
//            remaining ten stages are a 1024-point-shaped transform inside the warp (ct32 over r', transpose, ct32 over the lane bits);
//   inverse: the mirror image; the last Gentleman-Sande stage exchanges the halves through the tiles.
// Same butterflies, twiddles and key layout as the first kernel: bit-identical residues.  Four warps per scheduler at <= 128 registers,
// forward per-lane twiddles and the pass-A tables staged in shared memory (the inverse per-lane tables stay in L1/L2: 127 KB do not fit).
constexpr int WARPS16 = 4 * NP, THREADS16 = 32 * WARPS16;
constexpr int TILE16_WORDS = 32 * 33;
constexpr int TWB16_ENTRIES = NP * 2 * 31 * 32;            // forward pass-B tables [prime][half][31][32]
constexpr int TWA16_ENTRIES = NP * 2 * 2 * 32;             // pass-A half tables [prime][dir][half][32 (31 used)]
__host__ __device__ constexpr size_t smem_bytes16(int l) {
    return (size_t)2 * N * 8 + (size_t)2 * l * N * 4 + (size_t)WARPS16 * TILE16_WORDS * 4 + (size_t)(TWB16_ENTRIES + TWA16_ENTRIES) * 8;
}
__host__ __device__ constexpr bool ks_fusable16(int n, int t) { return mk::ks_row_stride(n) <= 4 * THREADS16 && (t == 4 || t == 5 || t == 8); }

// extraction + key switch by the 512 threads of the gate: a row needs stride / 4 threads, so up to three groups of threads take the
// coefficients 4q, 4q + 1, .. (mod 4 * groups) and their partial sums meet in shared memory (as mk::fused_keyswitch does for its
// latency launch)
template <int T>
__device__ __forceinline__ void fused_keyswitch2k16(const u64* __restrict__ acc, u32* __restrict__ s_a, const Args& p, int g, int tid) {
    const int n = p.n, stride = mk::ks_row_stride(n), bb = p.ks_basebit, B1 = (1 << bb) - 1;
    const uint32_t prec_offset = 1u << (32 - (1 + bb * T));   // keyswitch.jl:58
    for (int i = tid; i < N; i += THREADS16) {
        const u64 v = i == 0 ? acc[0] : (u64)0 - acc[N - i];
        const int32_t ai = mk::t64tot32((int64_t)v);
        if (p.ext_out) p.ext_out[(size_t)g * (N + 1) + i] = ai;
        s_a[i] = (uint32_t)ai + prec_offset;
    }
    const int32_t eb = mk::t64tot32((int64_t)acc[N]);
    if (p.ext_out && tid == 0) p.ext_out[(size_t)g * (N + 1) + N] = eb;
    __syncthreads();
    const int tpr = stride / 4;
    constexpr int MAXG = 4;
    const int groups = min(MAXG, THREADS16 / tpr);
    const int grp = tid / tpr, t = tid - grp * tpr;
    const bool active = grp < groups;
    const int col0 = 4 * t;
    const size_t party_words = (size_t)N * T * B1 * stride;
    const uint4* zero_row = reinterpret_cast<const uint4*>(p.ksk + (size_t)p.k * party_words) + t;
    uint4* part = reinterpret_cast<uint4*>(s_a + N);          // [2 (party parity)][MAXG - 1][tpr] partial sums
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
                        const uint32_t d = (ai >> (32 - j * bb)) & (uint32_t)B1;
                        const uint4* r = d ? reinterpret_cast<const uint4*>(rows + (((size_t)(i + ii) * T + (j - 1)) * B1 + (d - 1)) * stride) + t : zero_row;
                        const uint4 v = __ldg(r);
                        out.x -= v.x; out.y -= v.y; out.z -= v.z; out.w -= v.w;
                    }
                }
            }
        }
        if (groups > 1) {
            uint4* pp = part + (size_t)(party & 1) * (MAXG - 1) * tpr;
            if (active && grp > 0) pp[(grp - 1) * tpr + t] = out;
            __syncthreads();
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
    if (grp == 0 && col0 <= n && n < col0 + 4) p.ob[g] = (int32_t)((uint32_t)eb + bsum);
}

template <int L>
__global__ void __launch_bounds__(THREADS16, 1) blind_rotate2k16_kernel(Args p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);
    int32_t* dig = reinterpret_cast<int32_t*>(smem_raw + (size_t)2 * N * 8);           // [s = src*L + q][N] signed digits; src 0 = body, 1 = mask
    u32* tiles = reinterpret_cast<u32*>(smem_raw + (size_t)2 * N * 8 + (size_t)2 * L * N * 4);
    uint2_* twB_s = reinterpret_cast<uint2_*>(tiles + WARPS16 * TILE16_WORDS);        // forward pass-B tables [prime][half][31][32]
    uint2_* twA_s = twB_s + TWB16_ENTRIES;                                            // pass-A half tables [prime][dir][half][32]
    const int tid = threadIdx.x, gw = tid >> 5, lane = tid & 31;
    const int w = gw >> 2, o = (gw >> 1) & 1, h = gw & 1;  // prime; output polynomial = parity of the digit polynomials transformed; half
    const int g = blockIdx.x;
    // ---- tables: forward per-lane twiddles of all four primes; the 31-entry network of pass-A stages 1..5 restricted to half h
    for (int i = tid; i < TWB16_ENTRIES; i += THREADS16) {
        const int pi = i / (2 * 31 * 32), rest = i - pi * (2 * 31 * 32);
        twB_s[i] = p.twB[rns2k::twB_index(pi, 0, 0, 0, 0) + rest];
    }
    for (int i = tid; i < TWA16_ENTRIES; i += THREADS16) {
        const int e = i & 31, hh = (i >> 5) & 1, dir = (i >> 6) & 1, pi = i >> 7;
        if (e < 31) {
            const int k = 31 - __clz(e + 1), b = e + 1 - (1 << k);                    // stage k' and block b' of the 32-point network
            twA_s[i] = c_k2.twA[pi][dir][(2 << k) - 1 + (hh << k) + b];               // = stage k' + 1, block hh 2^k' + b' of the 64-point one
        }
    }
    const u32 pr = c_k2.p[w], pinv = c_k2.pinv_neg[w];
    u32* tile = tiles + gw * TILE16_WORDS;
    const u32* otile = tiles + (gw ^ 2) * TILE16_WORDS;    // same prime and half, other output: the partner of the digit exchange
    const u32* htile = tiles + (gw ^ 1) * TILE16_WORDS;    // same prime and output, other half: the partner of the last inverse stage
    const uint2_* twBf = twB_s + ((size_t)(w * 2 + h) * 31) * 32 + lane;
    const uint2_* twBi = p.twB + rns2k::twB_index(w, 1, h, 0, lane);
    const uint2_* twAf = twA_s + ((w * 2 + 0) * 2 + h) * 32;
    const uint2_* twAi = twA_s + ((w * 2 + 1) * 2 + h) * 32;
    const int pbar = 1 + w;                                 // one named barrier per prime: its four warps
    auto prime_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(pbar) : "memory"); };
    const int kn = p.k * p.n;
    const mk::GateLinear lin = p.gate_ids ? mk::gate_linear(__ldg(p.gate_ids + g)) : p.lin;
    auto rotation = [&](const int32_t* xs, const int32_t* ys, const int32_t* zs, size_t idx, uint32_t mu0) {
        uint32_t t = mu0 + (uint32_t)lin.cx * (uint32_t)__ldg(xs + idx);
        if (lin.cy) t += (uint32_t)lin.cy * (uint32_t)__ldg(ys + idx);
        if (lin.cz) t += (uint32_t)lin.cz * (uint32_t)__ldg(zs + idx);
        return mod_switch_2N((int32_t)t);
    };
    {   // acc = (0, X^{-barb} * testvect)  (3gen_mk_internals.jl:88-92)
        const int barb = rotation(p.xb, p.yb, p.zb, g, (uint32_t)lin.mu0);
        const int sh = (-barb) & (2 * N - 1);
        for (int i = tid; i < N; i += THREADS16) {
            const int idx = (i - sh) & (2 * N - 1);
            acc[i] = 0;
            acc[N + i] = (idx & N) ? (u64)0 - (u64)p.mu : (u64)p.mu;
        }
    }
    __syncthreads();
    u64 off = 0;
#pragma unroll
    for (int q = 1; q <= L; q++) off += ((u64)1 << (64 - q * p.bgbit)) << (p.bgbit - 1);   // tgsw.jl:24-30
    const u64 dmask = ((u64)1 << p.bgbit) - 1;
    const int64_t half = (int64_t)1 << (p.bgbit - 1);
    const size_t abase = (size_t)g * kn;
    const uint2_ w0f = c_k2.twA[w][0][0], w0i = c_k2.twA[w][1][0];
    const u32 p2 = rns2k::opaque_multiple(2 * pr), p4 = rns2k::opaque_multiple(4 * pr), p8 = rns2k::opaque_multiple(8 * pr);
    int a_next = rotation(p.xa, p.ya, p.za, abase, 0u);
    for (int it = 0; it < kn; it++) {
        const int a = a_next;
        if (it + 1 < kn) a_next = rotation(p.xa, p.ya, p.za, abase + it + 1, 0u);
        if (a == 0) continue;                                              // 3gen_mk_internals.jl:69
        // ---- decompose X^a * acc - acc (tgsw.jl:112-138)
#pragma unroll 4
        for (int i = tid; i < 2 * N; i += THREADS16) {
            const int c = i >> 11, ii = i & (N - 1);
            const u64* poly = acc + c * N;
            const int idx = (ii - a) & (2 * N - 1);
            u64 v = poly[idx & (N - 1)];
            if (idx & N) v = 0 - v;
            const u64 t = v - poly[ii] + off;
#pragma unroll
            for (int q = 0; q < L; q++) dig[((1 - c) * L + q) * N + ii] = (int32_t)((int64_t)((t >> (64 - (q + 1) * p.bgbit)) & dmask) - half);
        }
        __syncthreads();
        u32 x[32], accv[32];
        const u32* kw = p.bsk + (size_t)it * bsk_elem_words(L) + (size_t)w * (2 * L * 2 * N);
#pragma unroll 1
        for (int i = 0; i < L; i++) {
            const int s_own = 2 * i + o, s_for = 2 * i + 1 - o;
            // ---- forward: stage 0 across the halves, straight from the digit buffer
            const int32_t* d0 = dig + s_own * N + lane;
            if (h == 0) {
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    const u32 X = (u32)(d0[32 * r] + (int32_t)pr), Y = (u32)(d0[1024 + 32 * r] + (int32_t)pr);     // signed digit + p in [0, 2p)
                    x[r] = rns::alu_add(X, rns::shoup_mul(Y, w0f.x, w0f.y, pr));                                    // X + w0 Y in [0, 4p)
                }
            } else {
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    const u32 X = (u32)(d0[32 * r] + (int32_t)pr), Y = (u32)(d0[1024 + 32 * r] + (int32_t)pr);
                    x[r] = X - rns::shoup_mul(Y, w0f.x, w0f.y, pr) + p2;                                            // X - w0 Y in [0, 4p)
                }
            }
            rns::ct32(x, rns2k::TwUniform{twAf}, pr);                                                           // [0, 14p)
#pragma unroll
            for (int r = 0; r < 32; r++) tile[r * 33 + lane] = x[r];
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 32; c++) x[c] = rns2k::reduce_to_4p(tile[lane * 33 + c], p8, p4);
            __syncwarp();
            rns::ct32(x, rns2k::TwLane{twBf}, pr);                                                              // positions 32 (lane + 32 h) + c, [0, 14p)
#pragma unroll
            for (int c = 0; c < 32; c++) tile[c * 32 + lane] = x[c];
            prime_barrier();
            // ---- both digit polynomials of the pair times their key rows of output o, one Montgomery reduction per point
            const uint4* k_own = reinterpret_cast<const uint4*>(kw + (size_t)(s_own * 2 + o) * N) + lane + h * 256;
            const uint4* k_for = reinterpret_cast<const uint4*>(kw + (size_t)(s_for * 2 + o) * N) + lane + h * 256;
#pragma unroll
            for (int q4 = 0; q4 < 8; q4++) {
                const uint4 ka = __ldg(k_own + q4 * 32), kb = __ldg(k_for + q4 * 32);
                const u32* pt = otile + (4 * q4) * 32 + lane;
                const u32 v0 = rns::mont_mul2(x[4 * q4 + 0], ka.x, pt[0], kb.x, pr, pinv);       // < 2.75p
                const u32 v1 = rns::mont_mul2(x[4 * q4 + 1], ka.y, pt[32], kb.y, pr, pinv);
                const u32 v2 = rns::mont_mul2(x[4 * q4 + 2], ka.z, pt[64], kb.z, pr, pinv);
                const u32 v3 = rns::mont_mul2(x[4 * q4 + 3], ka.w, pt[96], kb.w, pr, pinv);
                accv[4 * q4 + 0] = i == 0 ? v0 : accv[4 * q4 + 0] + v0;
                accv[4 * q4 + 1] = i == 0 ? v1 : accv[4 * q4 + 1] + v1;
                accv[4 * q4 + 2] = i == 0 ? v2 : accv[4 * q4 + 2] + v2;
                accv[4 * q4 + 3] = i == 0 ? v3 : accv[4 * q4 + 3] + v3;
            }
            prime_barrier();                                               // the partner is done with this warp's tile
        }
#pragma unroll
        for (int c = 0; c < 32; c++) {
            const u32 v = accv[c];                                         // < 2.75 L p
            x[c] = L > 1 ? rns::umin32(v, v - p4) : v;                     // [0, 4p)
        }
        // ---- inverse: ten stages inside the warp, then the last stage across the halves
        rns2k::gs_net<5>(x, rns2k::TwLane{twBi}, pr, p4);
#pragma unroll
        for (int c = 0; c < 32; c++) tile[lane * 33 + c] = x[c];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; r++) x[r] = tile[r * 33 + lane];
        __syncwarp();
        rns2k::gs_net<5>(x, rns2k::TwUniform{twAi}, pr, p4);               // [0, 4p)
#pragma unroll
        for (int r = 0; r < 32; r++) tile[r * 32 + lane] = x[r];
        prime_barrier();
        if (h == 0) {                                                      // X + Y, Y from the other half
#pragma unroll
            for (int r = 0; r < 32; r++) {
                const u32 sgm = rns::alu_add(x[r], htile[r * 32 + lane]);
                x[r] = rns::umin32(sgm, sgm - p4);
            }
        } else {                                                           // (X - Y) w0^-1, X from the other half
#pragma unroll
            for (int r = 0; r < 32; r++) x[r] = rns::shoup_mul(htile[r * 32 + lane] - x[r] + p4, w0i.x, w0i.y, pr);
        }
        prime_barrier();                                                   // both halves have read each other's tile
#pragma unroll
        for (int r = 0; r < 32; r++) tile[r * 32 + lane] = x[r];           // residues of coefficients 1024 h + 32 r + lane
        __syncthreads();
        // ---- CRT by the whole gate: residues of (prime w', output oo, half hh) are in tile (w' * 2 + oo) * 2 + hh
#pragma unroll CRT_UNROLL
        for (int i = tid; i < 2 * N; i += THREADS16) {
            const int oo = i >> 11, hh = (i >> 10) & 1, ii = i & 1023;
            const u32* rt = tiles + (oo * 2 + hh) * TILE16_WORDS + ii;
            const u32 r[NP] = {rt[0], rt[4 * TILE16_WORDS], rt[8 * TILE16_WORDS], rt[12 * TILE16_WORDS]};
            acc[i] += rns2k::lift4(r, c_k2);
        }
        __syncthreads();
    }
    if (p.acc_out) {
        int64_t* ao = p.acc_out + (size_t)g * 2 * N;
        for (int i = tid; i < 2 * N; i += THREADS16) ao[i] = (int64_t)acc[i];
    }
    if (p.ksk) {   // fused extraction + key switch (the host only sets ksk when ks_fusable16(n, t)); the tiles are free now
        if (p.ks_t == 4) fused_keyswitch2k16<4>(acc, tiles, p, g, tid);
        else if (p.ks_t == 5) fused_keyswitch2k16<5>(acc, tiles, p, g, tid);
        else fused_keyswitch2k16<8>(acc, tiles, p, g, tid);
        return;
    }
    int32_t* ext = p.ext_out + (size_t)g * (N + 1);
    for (int i = tid; i < N; i += THREADS16) {
        const u64 v = i == 0 ? acc[0] : (u64)0 - acc[N - i];
        ext[i] = mk::t64tot32((int64_t)v);
    }
    if (tid == 0) ext[N] = mk::t64tot32((int64_t)acc[N]);
}

// One warp per (polynomial, prime): raw int64 key of one party, [n][4 parts][l][N] -> transformed residues in the streaming layout
constexpr int XF_WARPS = 2;
__global__ void __launch_bounds__(XF_WARPS * 32) bsk_transform2k_kernel(const int64_t* __restrict__ raw, u32* __restrict__ bsk, int n, int l, int party,
                                                                         const uint2_* __restrict__ twB, int ntasks) {
    __shared__ u32 tiles[XF_WARPS * TILE_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * XF_WARPS + warp;
    if (task >= ntasks) return;
    const int pi = task % NP, pq = task / NP;
    const int q = pq % l, part = (pq / l) & 3, j = pq / (4 * l);
    const int out = part < 2 ? 1 : 0;                       // part_1: body<-body, part_2: body<-mask, part_3: mask<-mask, part_4: mask<-body
    const int src = (part == 0 || part == 3) ? 0 : 1;
    const u32 p = c_k2.p[pi];
    const int64_t* poly = raw + (size_t)pq * N;
    u32 x[64], y[2][32];
#pragma unroll
    for (int r = 0; r < 64; r++) x[r] = rns::residue_i64(poly[32 * r + lane], p);
    warp_fwd(x, y, tiles + warp * TILE_WORDS, twB, pi, p, lane);
    u32* dst = bsk + ((size_t)party * n + j) * bsk_elem_words(l) + ((size_t)(pi * 2 * l + (src * l + q)) * 2 + out) * N;
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int c = 0; c < 32; c++) dst[key_slot(lane, h, c)] = rns::mulmod(y[h][c] % p, rns2k::stored_key_scale(c_k2, pi), p);
}

// exact c = a * b mod (X^2048 + 1, 2^64) for |a_i| <= 2^25 (key generation primitive / parity hook); NP warps = NP primes
__global__ void __launch_bounds__(32 * NP) negacyclic_mul2k_kernel(const int64_t* __restrict__ a, const int64_t* __restrict__ b, int64_t* __restrict__ c,
                                                                   const uint2_* __restrict__ twB) {
    __shared__ u32 tiles[NP * TILE_WORDS];
    const int pi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t g = blockIdx.x;
    const u32 p = c_k2.p[pi], pinv = c_k2.pinv_neg[pi];
    u32* tile = tiles + pi * TILE_WORDS;
    u32 x[64], ya[2][32], yb[2][32];
#pragma unroll
    for (int r = 0; r < 64; r++) x[r] = rns::residue_i64(a[g * N + 32 * r + lane], p);
    warp_fwd(x, ya, tile, twB, pi, p, lane);
#pragma unroll
    for (int r = 0; r < 64; r++) x[r] = rns::residue_i64(b[g * N + 32 * r + lane], p);
    warp_fwd(x, yb, tile, twB, pi, p, lane);
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int cc = 0; cc < 32; cc++) {
            const u32 ks = rns::mulmod(yb[h][cc] % p, rns2k::stored_key_scale(c_k2, pi), p);
            ya[h][cc] = rns::mont_mul(ya[h][cc], ks, p, pinv);
        }
    warp_inv(ya, x, tile, twB, pi, p, lane);
#pragma unroll
    for (int r = 0; r < 64; r++) tile[32 * r + lane] = x[r];
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += 32 * NP) {
        const u32 r[NP] = {tiles[i], tiles[TILE_WORDS + i], tiles[2 * TILE_WORDS + i], tiles[3 * TILE_WORDS + i]};
        c[g * N + i] = (int64_t)rns2k::lift4(r, c_k2);
    }
}

// stand-alone multi-key key switch for N = 2048 (same algorithm as mk::keyswitch_kernel)
__global__ void __launch_bounds__(mk::KS_THREADS) keyswitch2k_kernel(int n, int k, int t, int basebit, const int32_t* __restrict__ ksk,
                                                                      const int32_t* __restrict__ ext, int32_t* __restrict__ oa, int32_t* __restrict__ ob) {
    __shared__ uint32_t s_a[N];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int B1 = (1 << basebit) - 1, row = n + 1, stride = mk::ks_row_stride(n);
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));   // keyswitch.jl:58
    const int32_t* e = ext + (size_t)g * (N + 1);
    for (int i = tid; i < N; i += mk::KS_THREADS) s_a[i] = (uint32_t)e[i] + prec_offset;
    __syncthreads();
    const int ncols = (row + mk::KS_THREADS - 1) / mk::KS_THREADS;
    uint32_t bsum = 0;
    for (int p = 0; p < k; p++) {
        uint32_t acc[mk::KS_MAXCOLS];
#pragma unroll
        for (int c = 0; c < mk::KS_MAXCOLS; c++) acc[c] = 0;
        const int32_t* rows = ksk + (size_t)p * N * t * B1 * stride;
        for (int i = 0; i < N; i++) {
            const uint32_t ai = s_a[i];
            for (int j = 1; j <= t; j++) {
                const uint32_t d = (ai >> (32 - j * basebit)) & (uint32_t)B1;
                if (d != 0) {
                    const int32_t* r = rows + (((size_t)i * t + (j - 1)) * B1 + (d - 1)) * stride;
#pragma unroll
                    for (int c = 0; c < mk::KS_MAXCOLS; c++) {
                        const int col = tid + c * mk::KS_THREADS;
                        if (c < ncols && col < row) acc[c] -= (uint32_t)__ldg(r + col);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < mk::KS_MAXCOLS; c++) {
            const int col = tid + c * mk::KS_THREADS;
            if (c < ncols && col < n) oa[((size_t)g * k + p) * n + col] = (int32_t)acc[c];
            if (c < ncols && col == n) bsum += acc[c];
        }
    }
    if (tid == n % mk::KS_THREADS) ob[g] = (int32_t)((uint32_t)e[N] + bsum);
}

}  // namespace mk2k
