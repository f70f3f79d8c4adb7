// ntt2048_gpu.cu -- the N = 2048 arithmetic core (torus-fhe_b200/csrc/ntt2048.cuh) on the GPU: one warp per (product, prime) computes
// c = a * b mod (X^2048 + 1, p) with the warp-level transform (64 coefficients per thread, padded 64 x 33 tile), the host lifts the
// four residue polynomials with crt4_lift and compares with an exact schoolbook product mod 2^64.  a: 26-bit signed gadget digits,
// b: 64-bit keys -- the operand shapes of an external product at the 16-party parameters.  Stand-alone check of the core under
// kernels2k.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I torus-fhe_b200/csrc tests/host_emu/ntt2048_gpu.cu -o ntt2048_gpu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "tables2048.h"

using namespace rns2k;

__constant__ Consts c_k;

// forward transform of this warp's polynomial: x[r] = a[32 r + lane] in [0, 2p) -> y[h][c] = position 32 (lane + 32 h) + c, < 14p
__device__ __forceinline__ void warp_fwd(u32 (&x)[64], u32 (&y)[2][32], u32* tile, const uint2_* twB, int pi, u32 p, int lane) {
    fwd_passA64(x, c_k.twA[pi][0], p);
#pragma unroll
    for (int r = 0; r < 64; r++) tile[r * TILE_STRIDE + lane] = x[r];
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
        for (int c = 0; c < 32; c++) y[h][c] = rns2k::reduce_to_4p(tile[(lane + 32 * h) * TILE_STRIDE + c], rns2k::opaque_multiple(8 * p), rns2k::opaque_multiple(4 * p));
        fwd_passB32(y[h], twB + twB_index(pi, 0, h, 0, lane), p);
    }
    __syncwarp();
}
__device__ __forceinline__ void warp_inv(u32 (&y)[2][32], u32 (&x)[64], u32* tile, const uint2_* twB, int pi, u32 p, int lane) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
        inv_passB32(y[h], twB + twB_index(pi, 1, h, 0, lane), p, rns2k::opaque_multiple(4 * p));
#pragma unroll
        for (int c = 0; c < 32; c++) tile[(lane + 32 * h) * TILE_STRIDE + c] = y[h][c];
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 64; r++) x[r] = tile[r * TILE_STRIDE + lane];
    __syncwarp();
    inv_passA64(x, c_k.twA[pi][1], p, rns2k::opaque_multiple(4 * p));
}

// grid = products, block = NP warps; res[g][prime][N] residues in [0, 4p) in coefficient order
__global__ void __launch_bounds__(32 * NP) negacyclic_mul2048_kernel(const int64_t* __restrict__ a, const int64_t* __restrict__ b, u32* __restrict__ res,
                                                                     const uint2_* __restrict__ twB) {
    __shared__ u32 tiles[NP * TILE_WORDS];
    const int pi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t g = blockIdx.x;
    const u32 p = c_k.p[pi], pinv = c_k.pinv_neg[pi];
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
        for (int c = 0; c < 32; c++) {
            const u32 ks = rns::mulmod(yb[h][c] % p, c_k.key_scale[pi], p);      // what the stored key would hold
            ya[h][c] = rns::mont_mul(ya[h][c], ks, p, pinv);                     // [0, 2p)
        }
    warp_inv(ya, x, tile, twB, pi, p, lane);
#pragma unroll
    for (int r = 0; r < 64; r++) res[(g * NP + pi) * N + 32 * r + lane] = x[r];
}

static u64 rnd_state = 0x1234567ull;
static u64 rnd() { rnd_state ^= rnd_state << 13; rnd_state ^= rnd_state >> 7; rnd_state ^= rnd_state << 17; return rnd_state; }

int main() {
    HostTables T;
    const int G = 6;
    std::vector<int64_t> a((size_t)G * N), b((size_t)G * N);
    for (size_t i = 0; i < a.size(); i++) { a[i] = (int64_t)(rnd() % (1u << 26)) - (1 << 25); b[i] = (int64_t)rnd(); }
    for (int i = 0; i < N; i++) { a[1 * N + i] = -(1 << 25); b[1 * N + i] = INT64_MIN; }             // extreme magnitudes
    for (int i = 0; i < N; i++) { a[2 * N + i] = (1 << 25) - 1; b[2 * N + i] = INT64_MAX; }
    for (int i = 0; i < N; i++) { a[3 * N + i] = rnd() % 3 - 1; }                                     // ternary (key generation)
    for (int i = 0; i < N; i++) { a[4 * N + i] = 0; }
    int64_t *da, *db; u32* dres; uint2_* dtw;
    if (cudaMalloc(&da, a.size() * 8) != cudaSuccess) { printf("ntt2048_gpu: no CUDA device\n"); return 2; }
    cudaMalloc(&db, b.size() * 8); cudaMalloc(&dres, (size_t)G * NP * N * 4); cudaMalloc(&dtw, T.twB.size() * sizeof(uint2_));
    cudaMemcpy(da, a.data(), a.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dtw, T.twB.data(), T.twB.size() * sizeof(uint2_), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_k, &T.c, sizeof(Consts));
    negacyclic_mul2048_kernel<<<G, 32 * NP>>>(da, db, dres, dtw);
    std::vector<u32> res((size_t)G * NP * N);
    cudaError_t e = cudaMemcpy(res.data(), dres, res.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("ntt2048_gpu: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    int fails = 0;
    for (int g = 0; g < G && !fails; g++) {
        std::vector<u64> ref(N, 0);
        for (int i = 0; i < N; i++) {
            if (!a[(size_t)g * N + i]) continue;
            for (int j = 0; j < N; j++) {
                const u64 t = (u64)a[(size_t)g * N + i] * (u64)b[(size_t)g * N + j];
                if (i + j < N) ref[i + j] += t; else ref[i + j - N] -= t;
            }
        }
        for (int i = 0; i < N; i++) {
            const u32 r[NP] = {res[((size_t)g * NP + 0) * N + i], res[((size_t)g * NP + 1) * N + i], res[((size_t)g * NP + 2) * N + i], res[((size_t)g * NP + 3) * N + i]};
            for (int q = 0; q < NP; q++) if (r[q] >= 4ull * T.c.p[q]) { fails++; printf("FAIL range g=%d i=%d\n", g, i); break; }
            if (crt4_lift(r, T.c) != ref[i]) { fails++; printf("FAIL product g=%d i=%d\n", g, i); break; }
        }
    }
    printf(fails ? "ntt2048_gpu: %d FAILURES\n" : "ntt2048_gpu: OK (%d exact negacyclic products of degree 2048 through four primes)\n", fails ? fails : G);
    return fails ? 1 : 0;
}
