// ntt_pass_ubench.cu -- ceiling probe: the 32-point in-register pass of the production NTT (rns::ct32, per-lane twiddles from shared
// memory) in a loop, at different occupancies / register caps / elements per thread.  Prints achieved fma-pipe slot rate vs the
// measured IMAD peak.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../torus-fhe_b200/csrc -o ntt_pass_ubench ntt_pass_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "rns.cuh"
using namespace rns;
struct TwL { const uint2_* t; __device__ uint2_ operator()(int e) const { return t[e * 32]; } };

template <int E>   // E elements per thread: 32 (5 stages) or 16 (4 stages)
__device__ __forceinline__ void pass(u32 (&x)[E], TwL tw, u32 p) {
    const u32 p2 = keep_in_register(2 * p);
    constexpr int ST = E == 32 ? 5 : 4;
#pragma unroll
    for (int k = 0; k < ST; k++) {
        const int g = (E / 2) >> k;
#pragma unroll
        for (int b = 0; b < (1 << k); b++) {
            const uint2_ w = tw((1 << k) - 1 + b);
#pragma unroll
            for (int j = 0; j < g; j++) ct_bfly<false>(x[2 * g * b + j], x[2 * g * b + j + g], w.x, w.y, p, p2);
        }
    }
}
// variant: explicit phase-split order (all quotients, then all products, then the adds) as a scheduling hint
template <int E>
__device__ __forceinline__ void pass_split(u32 (&x)[E], TwL tw, u32 p) {
    const u32 p2 = keep_in_register(2 * p);
    constexpr int ST = E == 32 ? 5 : 4;
#pragma unroll
    for (int k = 0; k < ST; k++) {
        const int g = (E / 2) >> k;
        u32 q[E / 2], t[E / 2];
        uint2_ w[E / 2];
#pragma unroll
        for (int b = 0; b < (1 << k); b++) {
            const uint2_ ww = tw((1 << k) - 1 + b);
#pragma unroll
            for (int j = 0; j < g; j++) w[b * g + j] = ww;
        }
#pragma unroll
        for (int b = 0; b < (1 << k); b++)
#pragma unroll
            for (int j = 0; j < g; j++) q[b * g + j] = mulhi32(x[2 * g * b + j + g], w[b * g + j].y);
#pragma unroll
        for (int b = 0; b < (1 << k); b++)
#pragma unroll
            for (int j = 0; j < g; j++) t[b * g + j] = x[2 * g * b + j + g] * w[b * g + j].x;
#pragma unroll
        for (int b = 0; b < (1 << k); b++)
#pragma unroll
            for (int j = 0; j < g; j++) t[b * g + j] -= q[b * g + j] * p;
#pragma unroll
        for (int b = 0; b < (1 << k); b++)
#pragma unroll
            for (int j = 0; j < g; j++) {
                const u32 X = x[2 * g * b + j];
                x[2 * g * b + j] = alu_add(X, t[b * g + j]);
                x[2 * g * b + j + g] = X - t[b * g + j] + p2;
            }
    }
}
template <int E, int REGS, bool SPLIT = false>
__global__ void __maxnreg__(REGS) k(u32* out, const uint2_* twg, int iters, u32 p) {
    extern __shared__ uint2_ tw[];
    for (int i = threadIdx.x; i < 31 * 32; i += blockDim.x) tw[i] = twg[i];
    __syncthreads();
    u32 x[E];
#pragma unroll
    for (int i = 0; i < E; i++) x[i] = threadIdx.x * 977 + i * 131 + blockIdx.x;
    const u32 p4 = keep_in_register(4 * p);
    for (int it = 0; it < iters; it++) {
        if (SPLIT) pass_split<E>(x, TwL{tw + (threadIdx.x & 31)}, p); else pass<E>(x, TwL{tw + (threadIdx.x & 31)}, p);
#pragma unroll
        for (int i = 0; i < E; i++) x[i] = reduce_to_4p(x[i], p4);
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < E; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int E, int REGS, bool SPLIT = false>
void run(int warps_per_sm) {
    uint2_* tw; u32* out;
    cudaMalloc(&tw, 31 * 32 * 8); cudaMemset(tw, 0x5a, 31 * 32 * 8);
    const int threads = warps_per_sm * 32, iters = 2000;
    cudaMalloc(&out, 148 * threads * 4);
    cudaFuncSetAttribute(k<E, REGS, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<E, REGS, SPLIT><<<148, threads, 200 * 1024>>>(out, tw, 10, 268369921u);    // 200 KB smem: exactly one CTA per SM
    cudaEventRecord(e0);
    k<E, REGS, SPLIT><<<148, threads, 200 * 1024>>>(out, tw, iters, 268369921u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bf = (double)148 * threads * iters * (E == 32 ? 80 : 32);
    const double slots = bf * 4;
    printf("%sE=%2d regs<=%3d warps/SM=%2d: %7.2f ms  %6.2f T fma-slots/s = %4.1f %% of 18.26 T  (err %s)\n", SPLIT ? "split " : "      ", E, REGS, warps_per_sm, ms, slots / ms / 1e9,
           100 * slots / ms / 1e9 / 18.26, cudaGetErrorString(cudaGetLastError()));
    cudaFree(tw); cudaFree(out);
}
int main() {
    run<32, 168>(4); run<32, 168>(8); run<32, 168>(12); run<32, 128>(16); run<32, 96>(20); run<32, 80>(24);
    run<32, 168, true>(12); run<32, 128, true>(16);
    run<16, 168>(12); run<16, 128>(16); run<16, 96>(20); run<16, 80>(24); run<16, 64>(32);
    return 0;
}
