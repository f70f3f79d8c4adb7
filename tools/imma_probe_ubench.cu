// imma_probe_ubench.cu -- round-2 go/no-go probe for the only idea that lowers the kernels' INSTRUCTION count:
// pass A of the forward transforms of the gadget digits as a matrix product on the tensor cores.  The digits are 7-bit, pass A applies
// the same 32 x 32 twiddle matrix in every lane, so per warp and digit polynomial it is T[32 x 32] * D[32 x 32 lanes] with T split in four
// 7-bit planes: 32 mma.sync.m16n8k32.s8 (legacy IMMA path; whether sm_100 still runs it at a useful rate is exactly what is unknown)
// plus a recombination (two shift-adds, one Shoup product, one add per point) against 2.5 butterflies = 11 instructions per point today.
// Estimated gain if IMMA is fast and co-issues with IMAD: about a quarter of pass A's fmaheavy slots, i.e. ~8 % of the butterfly work.
// This probe measures
//   1. IMMA.16832.S8 alone: MMA instructions per clock and SM (each is 16 x 8 x 32 = 4096 MACs);
//   2. IMMA warps next to IMAD warps on the same SM (does the tensor pipe co-issue with the binding integer pipe?);
//   3. a self-check of the fragment layouts the formulation relies on (C = A * B against a scalar loop), so a wrong reading of the
//      PTX layouts shows up here and not inside the kernel.
// Decision rule (written before measuring): go only if (1) sustains >= 1 MMA per 16 clocks per SM sub-partition (32 MMAs of a digit
// polynomial then cost <= 512 clocks, below the ~700 clocks pass A's stage-1..4 butterflies hold the fmaheavy pipe) AND (2) leaves the
// IMAD rate within 10 % of its stand-alone value.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_probe_ubench imma_probe_ubench.cu ; run on the B200.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef uint32_t u32;
#define ITER 2048
#define CHAINS 8

__device__ __forceinline__ void imma_16832(int (&c)[4], const u32 (&a)[4], const u32 (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MODE 0: IMAD in all warps; 1: IMMA in all warps; 2: even warps IMAD, odd warps IMMA
template <int MODE> __global__ void bench(u32* out, u32 m, u32 c0) {
    const int warp = threadIdx.x >> 5;
    const bool mma_warp = MODE == 1 || (MODE == 2 && (warp & 1));
    u32 x[CHAINS];
    int acc[CHAINS][4];
    u32 a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = 0x01020304u * (threadIdx.x + i + 1);
    b[0] = 0x04030201u + threadIdx.x; b[1] = 0x01010101u * (threadIdx.x & 7);
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { x[i] = threadIdx.x * 7 + i; acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = i; }
    if (mma_warp) {
        for (int it = 0; it < ITER; it++) {
#pragma unroll
            for (int i = 0; i < CHAINS; i++) imma_16832(acc[i], a, b);
        }
    } else {
        for (int it = 0; it < ITER; it++) {
#pragma unroll
            for (int i = 0; i < CHAINS; i++) x[i] = x[i] * m + c0;
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s ^= x[i] ^ (u32)(acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, double imad_share, double mma_share) {
    const int threads = 384, blocks = 148 * 4;
    u32* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<blocks, threads>>>(out, 0x9E3779B1u, 12345);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, threads>>>(out, 0x9E3779B1u, 12345);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 * 1.965e9;
    const double thread_ops = (double)blocks * threads * ITER * CHAINS;
    const double imad_lanes = thread_ops * imad_share / clk / 148;
    const double mma_per_clk_sm = thread_ops * mma_share / 32 / clk / 148;      // warp-level MMA instructions
    printf("%-34s %8.3f ms   IMAD %6.1f lanes/clk/SM   IMMA.16832 %6.3f /clk/SM (= %7.0f int8 MAC/clk/SM)\n", name, ms, imad_lanes, mma_per_clk_sm,
           mma_per_clk_sm * 4096);
    cudaFree(out);
}

// fragment-layout self-check: C[16 x 8] = A[16 x 32] (row-major) * B[32 x 8] (column-major), signed 8-bit
__global__ void layout_check(const int8_t* A, const int8_t* B, int* C) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    u32 a[4], b[2];
    int c[4] = {0, 0, 0, 0};
    auto pack = [](const int8_t* p) { return (u32)(uint8_t)p[0] | ((u32)(uint8_t)p[1] << 8) | ((u32)(uint8_t)p[2] << 16) | ((u32)(uint8_t)p[3] << 24); };
    a[0] = pack(A + g * 32 + 4 * t);             // row g,     k = 4t .. 4t+3
    a[1] = pack(A + (g + 8) * 32 + 4 * t);       // row g + 8, k = 4t .. 4t+3
    a[2] = pack(A + g * 32 + 16 + 4 * t);        // row g,     k = 16 + 4t ..
    a[3] = pack(A + (g + 8) * 32 + 16 + 4 * t);  // row g + 8, k = 16 + 4t ..
    b[0] = pack(B + g * 32 + 4 * t);             // column g (stored k-contiguous), k = 4t ..
    b[1] = pack(B + g * 32 + 16 + 4 * t);        // column g, k = 16 + 4t ..
    imma_16832(c, a, b);
    C[g * 8 + 2 * t] = c[0];                     // (row g,     col 2t)
    C[g * 8 + 2 * t + 1] = c[1];                 // (row g,     col 2t + 1)
    C[(g + 8) * 8 + 2 * t] = c[2];               // (row g + 8, col 2t)
    C[(g + 8) * 8 + 2 * t + 1] = c[3];           // (row g + 8, col 2t + 1)
}

int main() {
    run<0>("IMAD, all warps", 1, 0);
    run<1>("IMMA.16832.S8, all warps", 0, 1);
    run<2>("IMAD warps | IMMA warps (1:1)", 0.5, 0.5);
    int8_t hA[16 * 32], hB[8 * 32];
    int hC[16 * 8];
    srand(7);
    for (auto& v : hA) v = (int8_t)(rand() % 128);            // 7-bit twiddle planes
    for (auto& v : hB) v = (int8_t)(rand() % 128 - 64);       // signed gadget digits
    int8_t *dA, *dB; int* dC;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    layout_check<<<1, 32>>>(dA, dB, dC);
    cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 16; r++)
        for (int n = 0; n < 8; n++) {
            int s = 0;
            for (int k = 0; k < 32; k++) s += (int)hA[r * 32 + k] * (int)hB[n * 32 + k];
            bad += s != hC[r * 8 + n];
        }
    printf("fragment layout self-check: %s (%d mismatches of 128; cuda status %s)\n", bad ? "FAIL" : "OK", bad, cudaGetErrorString(cudaGetLastError()));
    return bad != 0;
}
