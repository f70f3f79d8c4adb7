// kernels.cuh -- sm_100a kernels of the 3gen MK-TFHE bootstrapped-gate path.
//
//   bsk_transform_kernel   one-time: int64 key polys -> two NTT-domain limbs, streaming layout
//   blind_rotate_kernel    gate prologue + mod-switch + k*n mux-rotate steps + sample extraction
//   keyswitch_kernel       multi-key LWE key switch (gather-accumulate over ksk rows)
//   extprod_kernel / negacyclic_mul_kernel   parity hooks built from the same device code
//
// Reference semantics (3-gen-mk-tfhe/src/): 3gen_mk_internals.jl:59-116,
// tgsw_3gen.jl:102-113, tgsw.jl:112-138, rlwe.jl:70-74, keyswitch.jl:45-80,
// mk_internals.jl:730-744, numeric-functions.jl:70-73,109-111.
#pragma once
#include <cuda_runtime.h>
#include "ntt1024.cuh"

namespace mk {

constexpr int N = ntt::N;
constexpr int BR_THREADS = 128;  // 4 warps: one polynomial transform each
constexpr int BR_WARPS = BR_THREADS / 32;

// BSK streaming layout (u64): [elem = party*n + j][outlimb = out*2 + limb][r 32][lane 32][s = src*l + q]
//   out : 0 = mask', 1 = body'   (accumulator polynomial written)
//   src : 0 = body digits (c0), 1 = mask digits (c1)
//   (out, src) -> reference part: body<-body part_1, body<-mask part_2, mask<-mask part_3, mask<-body part_4
//   (r, lane) -> NTT index brev5(lane) + 32*brev5(r)  (ntt1024.cuh NTT-domain layout)
__host__ __device__ inline size_t bsk_elem_u64(int l) { return (size_t)4 * N * 2 * l; }

struct GateLinear {   // temp = mu0 + cx*x + cy*y + cz*z   (3gen_mk_gates.jl:8-74)
    int32_t mu0, cx, cy, cz;
};

struct BlindRotateArgs {
    int n, k, bgbit;
    const u64* bsk;
    const u64* tw_fwd;
    const u64* tw_inv;
    const int32_t *xa, *xb, *ya, *yb, *za, *zb;
    GateLinear lin;
    int64_t mu;
    int32_t* ext_out;   // [G][N+1]
    int64_t* acc_out;   // [G][2][N] or nullptr
};

__device__ __forceinline__ int tiles_needed(int l) { return 2 * l > BR_WARPS ? 2 * l : BR_WARPS; }

// decode_message(x, 2N), numeric-functions.jl:70-73 (wrapping add, arithmetic shift)
__device__ __forceinline__ int mod_switch_2N(int32_t x) {
    return (int32_t)((uint32_t)x + (1u << 20)) >> 21;   // N = 1024: 32 - log2(2N) = 21
}
// t64tot32, numeric-functions.jl:109-111: Int64 -> Float64 (RN), /2^32 (exact), trunc toward zero.
// The reference throws when the quotient is 2^31 (p ~ 2^-54); __double2int_rz saturates.
__device__ __forceinline__ int32_t t64tot32(int64_t v) {
    return __double2int_rz(__ll2double_rn(v) * (1.0 / 4294967296.0));
}

// One external product (tgsw_extern_mul_3gen, tgsw_3gen.jl:102-113) on the
// accumulator held in shared memory, by the 4 warps of the CTA.
//   MUX = true : acc += ExtProd(X^a * acc - acc, key)   (mk_mux_rotate_3gen, 3gen_mk_internals.jl:59-62)
//   MUX = false: acc  = ExtProd(acc, key)
// acc: [2][N] u64, [0] = mask, [1] = body.  tiles: tiles_needed(L) padded 32x33 u64 tiles.
template <int L, bool MUX>
__device__ __forceinline__ void extprod_step(u64* __restrict__ acc, u64* __restrict__ tiles, const u64* __restrict__ key,
                                             int a, int bgbit, const u64* __restrict__ tw_fwd,
                                             const u64* __restrict__ tw_inv) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ---- phase A: rotate-subtract, gadget-decompose (tgsw.jl:112-138), forward NTT of the 2L digit polynomials
    u64 off = 0;
#pragma unroll
    for (int q = 1; q <= L; q++) off += ((u64)1 << (64 - q * bgbit)) << (bgbit - 1);   // tgsw.jl:24-30
    const int64_t dmask = ((int64_t)1 << bgbit) - 1, dhalf = (int64_t)1 << (bgbit - 1);
    for (int s = warp; s < 2 * L; s += BR_WARPS) {
        const int src = s / L, q = s - src * L;
        const u64* poly = acc + (1 - src) * N;   // src 0 = body = acc[1]
        const int sh = 64 - (q + 1) * bgbit;
        u64 x[32];
#pragma unroll
        for (int i1 = 0; i1 < 32; i1++) {
            const int i = 32 * i1 + lane;
            u64 t;
            if (MUX) {
                const int idx = (i - a) & (2 * N - 1);        // (X^a * p)[i] = +-p[(i - a) mod 2N]
                u64 v = poly[idx & (N - 1)];
                if (idx & N) v = 0 - v;
                t = v - poly[i];
            } else {
                t = poly[i];
            }
            const int64_t d = (((int64_t)(t + off) >> sh) & dmask) - dhalf;
            x[i1] = gl::from_i64(d);
        }
        ntt::fwd_pass1(x, tw_fwd, lane);
        u64* tile = tiles + s * ntt::TILE_ELEMS;
#pragma unroll
        for (int r = 0; r < 32; r++) tile[r * ntt::TILE_STRIDE + lane] = x[r];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j++) x[j] = tile[lane * ntt::TILE_STRIDE + j];
        __syncwarp();
        ntt::fwd_pass2(x);
#pragma unroll
        for (int r = 0; r < 32; r++) tile[r * ntt::TILE_STRIDE + lane] = x[r];
    }
    __syncthreads();
    // ---- phase B: warp w accumulates output polynomial (out, limb) = (w >> 1, w & 1) in the NTT domain
    u64 y[32];
    {
        const u64* kp = key + (size_t)warp * (N * 2 * L) + lane * (2 * L);
#pragma unroll
        for (int r = 0; r < 32; r++) {
            u64 kv[2 * L];
            const ulonglong2* kp2 = reinterpret_cast<const ulonglong2*>(kp + (size_t)r * 32 * 2 * L);
#pragma unroll
            for (int s2 = 0; s2 < L; s2++) {
                ulonglong2 v = __ldg(kp2 + s2);
                kv[2 * s2] = v.x;
                kv[2 * s2 + 1] = v.y;
            }
            u64 sum = 0;
#pragma unroll
            for (int s = 0; s < 2 * L; s++)
                sum = gl::add(sum, gl::mul(tiles[s * ntt::TILE_ELEMS + r * ntt::TILE_STRIDE + lane], kv[s]));
            y[r] = sum;
        }
    }
    __syncthreads();   // all digit tiles consumed; tiles are scratch again
    // ---- inverse NTT of the 4 output polynomials
    ntt::inv_pass1(y, tw_inv, lane);
    u64* tile = tiles + warp * ntt::TILE_ELEMS;
#pragma unroll
    for (int j = 0; j < 32; j++) tile[lane * ntt::TILE_STRIDE + j] = y[j];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; r++) y[r] = tile[r * ntt::TILE_STRIDE + lane];
    __syncwarp();
    ntt::inv_pass2(y);
    // ---- phase C: recombine limbs (exact mod 2^64) and update the accumulator
    const int out = warp >> 1, limb = warp & 1;
    if (limb == 1) {
        u32* hb = reinterpret_cast<u32*>(tile);
#pragma unroll
        for (int i1 = 0; i1 < 32; i1++) hb[32 * i1 + lane] = (u32)gl::lift(y[i1]);
    }
    __syncthreads();
    if (limb == 0) {
        const u32* hb = reinterpret_cast<const u32*>(tiles + (warp + 1) * ntt::TILE_ELEMS);
        u64* ap = acc + out * N;
#pragma unroll
        for (int i1 = 0; i1 < 32; i1++) {
            const int i = 32 * i1 + lane;
            const u64 v = gl::lift(y[i1]) + ((u64)hb[i] << 32);
            ap[i] = MUX ? ap[i] + v : v;
        }
    }
    __syncthreads();
}

// One CTA per gate.  Accumulator resident in shared memory for all k*n steps.
template <int L>
__global__ void __launch_bounds__(BR_THREADS, 4) blind_rotate_kernel(BlindRotateArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);
    u64* tiles = acc + 2 * N;
    const int NT = 2 * L > BR_WARPS ? 2 * L : BR_WARPS;
    int16_t* bara = reinterpret_cast<int16_t*>(tiles + NT * ntt::TILE_ELEMS);
    __shared__ int s_barb;
    const int g = blockIdx.x, tid = threadIdx.x, kn = p.k * p.n;
    // gate prologue (3gen_mk_gates.jl) + mod switch (3gen_mk_internals.jl:102-103)
    for (int i = tid; i < kn; i += BR_THREADS) {
        uint32_t t = (uint32_t)p.lin.cx * (uint32_t)p.xa[(size_t)g * kn + i];
        if (p.lin.cy) t += (uint32_t)p.lin.cy * (uint32_t)p.ya[(size_t)g * kn + i];
        if (p.lin.cz) t += (uint32_t)p.lin.cz * (uint32_t)p.za[(size_t)g * kn + i];
        bara[i] = (int16_t)mod_switch_2N((int32_t)t);
    }
    if (tid == 0) {
        uint32_t t = (uint32_t)p.lin.mu0 + (uint32_t)p.lin.cx * (uint32_t)p.xb[g];
        if (p.lin.cy) t += (uint32_t)p.lin.cy * (uint32_t)p.yb[g];
        if (p.lin.cz) t += (uint32_t)p.lin.cz * (uint32_t)p.zb[g];
        s_barb = mod_switch_2N((int32_t)t);
    }
    __syncthreads();
    // acc = (0, X^{-barb} * testvect), testvect = mu * (1 + X + ... + X^{N-1})  (:88-92, rlwe.jl:113-119)
    {
        const int s = (-s_barb) & (2 * N - 1);
        for (int i = tid; i < N; i += BR_THREADS) {
            const int idx = (i - s) & (2 * N - 1);
            acc[i] = 0;
            acc[N + i] = (idx & N) ? (u64)0 - (u64)p.mu : (u64)p.mu;
        }
    }
    __syncthreads();
    // mk_blind_rotate_3gen: parties outer, coefficients inner (:66-84); element index = party*n + j
    const size_t estride = bsk_elem_u64(L);
    for (int it = 0; it < kn; it++) {
        const int a = bara[it];
        if (a == 0) continue;   // :69 (uniform across the CTA)
        extprod_step<L, true>(acc, tiles, p.bsk + (size_t)it * estride, a, p.bgbit, p.tw_fwd, p.tw_inv);
    }
    // rlwe_extract_sample_64 (rlwe.jl:70-74): a'_0 = mask_0, a'_i = -mask_{N-i}, b' = body_0
    int32_t* ext = p.ext_out + (size_t)g * (N + 1);
    for (int i = tid; i < N; i += BR_THREADS) {
        const u64 v = i == 0 ? acc[0] : (u64)0 - acc[N - i];
        ext[i] = t64tot32((int64_t)v);
    }
    if (tid == 0) ext[N] = t64tot32((int64_t)acc[N]);
    if (p.acc_out) {
        int64_t* ao = p.acc_out + (size_t)g * 2 * N;
        for (int i = tid; i < 2 * N; i += BR_THREADS) ao[i] = (int64_t)acc[i];
    }
}

// parity hook: acc_out[g] = ExtProd(acc_in[g], bsk[elem[g]])
template <int L>
__global__ void __launch_bounds__(BR_THREADS, 4) extprod_kernel(const u64* bsk, const u64* tw_fwd, const u64* tw_inv, int bgbit,
                                                                 const int32_t* elem, const int64_t* acc_in, int64_t* acc_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);
    u64* tiles = acc + 2 * N;
    const int g = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 2 * N; i += BR_THREADS) acc[i] = (u64)acc_in[(size_t)g * 2 * N + i];
    __syncthreads();
    extprod_step<L, false>(acc, tiles, bsk + (size_t)elem[g] * bsk_elem_u64(L), 0, bgbit, tw_fwd, tw_inv);
    for (int i = tid; i < 2 * N; i += BR_THREADS) acc_out[(size_t)g * 2 * N + i] = (int64_t)acc[i];
}

// One warp per (polynomial, limb): raw int64 key -> NTT-domain streaming layout.
// raw: [n][4 parts][l][N] int64 of one party; task = ((j*4 + part)*l + q)*2 + limb.
__global__ void __launch_bounds__(BR_THREADS) bsk_transform_kernel(const int64_t* __restrict__ raw, u64* __restrict__ bsk, int n, int l,
                                                                     int party, const u64* __restrict__ tw_fwd, int ntasks) {
    __shared__ u64 tiles[BR_WARPS * ntt::TILE_ELEMS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * BR_WARPS + warp;
    if (task >= ntasks) return;
    const int limb = task & 1, pq = task >> 1;
    const int q = pq % l, part = (pq / l) & 3, j = pq / (4 * l);
    // part_1: body<-body, part_2: body<-mask, part_3: mask<-mask, part_4: mask<-body  (tgsw_3gen.jl:109-110)
    const int out = part < 2 ? 1 : 0;
    const int src = (part == 0 || part == 3) ? 0 : 1;
    const int64_t* poly = raw + (size_t)pq * N;
    u64 x[32];
#pragma unroll
    for (int i1 = 0; i1 < 32; i1++) {
        const u64 v = (u64)poly[32 * i1 + lane];
        x[i1] = limb ? (v >> 32) : (v & gl::EPS);
    }
    ntt::fwd_pass1(x, tw_fwd, lane);
    u64* tile = tiles + warp * ntt::TILE_ELEMS;
#pragma unroll
    for (int r = 0; r < 32; r++) tile[r * ntt::TILE_STRIDE + lane] = x[r];
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 32; jj++) x[jj] = tile[lane * ntt::TILE_STRIDE + jj];
    ntt::fwd_pass2(x);
    const size_t e = (size_t)party * n + j;
    u64* dst = bsk + e * bsk_elem_u64(l) + (size_t)(out * 2 + limb) * (N * 2 * l) + (size_t)lane * 2 * l + (src * l + q);
#pragma unroll
    for (int r = 0; r < 32; r++) dst[(size_t)r * 32 * 2 * l] = x[r];
}

// parity hook: exact c = a * b mod (X^N + 1, 2^64); 3 warps: a, b_lo, b_hi
__global__ void __launch_bounds__(96) negacyclic_mul_kernel(const int64_t* __restrict__ a, const int64_t* __restrict__ b, int64_t* __restrict__ c,
                                                             const u64* __restrict__ tw_fwd, const u64* __restrict__ tw_inv) {
    __shared__ u64 tiles[3 * ntt::TILE_ELEMS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t g = blockIdx.x;
    u64 x[32];
#pragma unroll
    for (int i1 = 0; i1 < 32; i1++) {
        const int i = 32 * i1 + lane;
        if (warp == 0) x[i1] = gl::from_i64(a[g * N + i]);
        else { const u64 v = (u64)b[g * N + i]; x[i1] = warp == 2 ? (v >> 32) : (v & gl::EPS); }
    }
    ntt::fwd_pass1(x, tw_fwd, lane);
    u64* tile = tiles + warp * ntt::TILE_ELEMS;
#pragma unroll
    for (int r = 0; r < 32; r++) tile[r * ntt::TILE_STRIDE + lane] = x[r];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; j++) x[j] = tile[lane * ntt::TILE_STRIDE + j];
    __syncwarp();
    ntt::fwd_pass2(x);
    if (warp == 0) {
#pragma unroll
        for (int r = 0; r < 32; r++) tile[r * ntt::TILE_STRIDE + lane] = x[r];
    }
    __syncthreads();
    if (warp > 0) {
#pragma unroll
        for (int r = 0; r < 32; r++) x[r] = gl::mul(x[r], tiles[r * ntt::TILE_STRIDE + lane]);
        ntt::inv_pass1(x, tw_inv, lane);
#pragma unroll
        for (int j = 0; j < 32; j++) tile[lane * ntt::TILE_STRIDE + j] = x[j];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; r++) x[r] = tile[r * ntt::TILE_STRIDE + lane];
        __syncwarp();
        ntt::inv_pass2(x);
        if (warp == 2) {
            u32* hb = reinterpret_cast<u32*>(tile);
#pragma unroll
            for (int i1 = 0; i1 < 32; i1++) hb[32 * i1 + lane] = (u32)gl::lift(x[i1]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        const u32* hb = reinterpret_cast<const u32*>(tiles + 2 * ntt::TILE_ELEMS);
#pragma unroll
        for (int i1 = 0; i1 < 32; i1++) {
            const int i = 32 * i1 + lane;
            c[g * N + i] = (int64_t)(gl::lift(x[i1]) + ((u64)hb[i] << 32));
        }
    }
}

// Multi-key LWE key switch (mk_keyswitch_3gen mk_internals.jl:730-744, keyswitch keyswitch.jl:45-80).
// One CTA per sample; thread c owns output columns c, c+KS_THREADS, ... of the (n+1)-wide rows.
// ksk: int32 [k][N][t][B-1][n+1].
constexpr int KS_THREADS = 256;
constexpr int KS_MAXCOLS = 4;   // n + 1 <= 1024
__global__ void __launch_bounds__(KS_THREADS) keyswitch_kernel(int n, int k, int t, int basebit, const int32_t* __restrict__ ksk,
                                                                const int32_t* __restrict__ ext, int32_t* __restrict__ oa, int32_t* __restrict__ ob) {
    __shared__ uint32_t s_a[N];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int B1 = (1 << basebit) - 1, row = n + 1;
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
        const int32_t* rows = ksk + (size_t)p * N * t * B1 * row;
        for (int i = 0; i < N; i++) {
            const uint32_t ai = s_a[i];
            for (int j = 1; j <= t; j++) {
                const uint32_t d = (ai >> (32 - j * basebit)) & (uint32_t)B1;   // :65-67
                if (d != 0) {                                                 // :74-76
                    const int32_t* r = rows + (((size_t)i * t + (j - 1)) * B1 + (d - 1)) * row;
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

}  // namespace mk
