// pipe_ubench.cu -- integer-pipe micro-benchmark for the roofline denominators of the NTT kernels
// ("IMAD peak ... measure with a micro-benchmark and record").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_ubench pipe_ubench.cu ; run on the B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;
#define ITER 4096
#define CHAINS 8
template <int OP> __global__ void bench(u32* out, u32 a, u32 b, u32 c) {
    u32 x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x[i] = threadIdx.x * 7 + i + a;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) x[i] = x[i] * b + c;                       // IMAD
            if (OP == 1) x[i] = __umulhi(x[i], b) + c;              // IMAD.HI.U32
            if (OP == 2) { uint64_t w = (uint64_t)x[i] * b + (((uint64_t)c << 32) | x[i]); x[i] = (u32)(w >> 32) ^ (u32)w; }  // IMAD.WIDE.U32 + LOP3
            if (OP == 3) x[i] = min(x[i] + b, x[i]);                // VIADDMNMX.U32
            if (OP == 4) x[i] = x[i] + b - c + (x[i] >> 31);        // IADD3 + SHF mix (alu)
            if (OP == 5) {                                          // Harvey CT butterfly pair (x[i], x[i^1])
                u32 X = x[i], Y = x[(i + 1) % CHAINS];
                u32 xr = min(X, X - 2 * c);
                u32 q = __umulhi(Y, b);
                u32 t = Y * a - q * c;
                x[i] = xr + t;
            }
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, double ops_per_iter_chain) {
    u32* out; cudaMalloc(&out, 148 * 16 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<OP><<<148 * 8, 256>>>(out, 3, 0x9E3779B1u, 12345);
    cudaEventRecord(e0);
    bench<OP><<<148 * 8, 256>>>(out, 3, 0x9E3779B1u, 12345);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = 148.0 * 8 * 256 * ITER * CHAINS * ops_per_iter_chain;
    printf("%-28s %8.3f ms  %8.2f T thread-ops/s  (%.1f lanes/clk/SM at 1.965 GHz)\n", name, ms, n / ms / 1e9, n / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    run<0>("IMAD (lo)", 1); run<1>("IMAD.HI.U32 (+IADD)", 1); run<2>("IMAD.WIDE.U32 (+LOP3)", 1);
    run<3>("VIADDMNMX.U32", 1); run<4>("IADD3+SHF (2 alu)", 2); run<5>("CT butterfly half (3 fma+2 alu)", 1);
    return 0;
}
