// fp64_pipe_ubench.cu -- round-2 go/no-go probe: is the FP64 pipe of the B200 a second, idle multiplier next to the IMAD pipe?
// The blind-rotate kernels are bound by the fmaheavy (IMAD) pipe (66 % busy, 0.61 of the algorithmic bound);
// sm_100 quotes ~40 TFLOP/s of vector FP64 = 64 DFMA lanes/clk/SM, the same width as IMAD, on a pipe of its own.  If DFMA and IMAD
// issue side by side, the residues of one of the three primes could be carried in doubles (exact: h = x*w, l = fma(x, w, -h),
// q = rint(h/p), r = fma(-q, p, h) + l) by the warps of that prime, taking a third of the butterflies off the binding pipe.
// This probe measures, per SM and clock:
//   1. DFMA alone, IMAD alone (same harness as pipe_ubench.cu);
//   2. both in every warp, interleaved;
//   3. warp-specialised: IMAD warps and DFMA warps resident on the same SM (the shape the kernel would use: warp = (prime, output));
//   4. a Harvey butterfly in u32 next to the same butterfly in doubles, alone and side by side.
// Ceiling known before measuring: the scheduler issues one instruction per clock; a u32 butterfly is 5 (forward) or 6 (inverse) instructions for 8 fmaheavy
// cycles, the double butterfly 8 instructions for 16 FP64 cycles, so 2 integer warps + 1 FP64 warp per scheduler are issue-bound at 18-20 cycles
// where 3 integer warps are pipe-bound at 24: at most 1.2-1.33x in the butterfly phases (mode 6 below measures exactly this mix).
// Decision rule (written before measuring): go if (3) sustains >= 1.6x the thread-ops of IMAD alone AND the double butterfly costs
// <= 2.2x the u32 butterfly in isolation (then a 4 + 2 warp split per gate balances the two pipes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe_ubench fp64_pipe_ubench.cu ; run on the B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;
#define ITER 4096
#define CHAINS 8
#define P28 268369921u                       // rns::PRIME0 of the N = 1024 kernels

__device__ __forceinline__ void imad_chain(u32 (&x)[CHAINS], u32 b, u32 c) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x[i] = x[i] * b + c;
}
__device__ __forceinline__ void dfma_chain(double (&y)[CHAINS], double b, double c) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) y[i] = fma(y[i], b, c);
}
// u32 Harvey butterfly half, as in pipe_ubench.cu OP 5: 1 IMAD.HI + 2 IMAD + 2 alu
__device__ __forceinline__ void bfly_u32(u32 (&x)[CHAINS], u32 w, u32 wq, u32 p) {
#pragma unroll
    for (int i = 0; i < CHAINS; i += 2) {
        u32 X = x[i], Y = x[i + 1];
        u32 xr = min(X, X - 2 * p);
        u32 q = __umulhi(Y, wq);
        u32 t = Y * w - q * p;
        x[i] = xr + t;
        x[i + 1] = xr - t + 2 * p;
    }
}
// the same butterfly on doubles holding exact integers: the product is split exactly by the fma error term, the quotient is rounded
// with the 1.5 * 2^52 trick, and no range correction is needed at all -- a 28-bit residue may grow by p per stage for far more than
// the ten stages of a transform before it leaves the 53-bit significand.  8 FP64-pipe instructions: DMUL, 3 DFMA, 4 DADD.
__device__ __forceinline__ void bfly_f64(double (&y)[CHAINS], double w, double pinv, double p) {
    const double magic = 6755399441055744.0;          // 1.5 * 2^52
#pragma unroll
    for (int i = 0; i < CHAINS; i += 2) {
        double X = y[i], Y = y[i + 1];
        double h = Y * w;
        double l = fma(Y, w, -h);
        double q = fma(h, pinv, magic) - magic;
        double t = fma(-q, p, h) + l;                 // Y * w mod p in (-p, p), exact
        y[i] = X + t;
        y[i + 1] = X - t;
    }
}

// MODE 0 IMAD all warps; 1 DFMA all warps; 2 both interleaved in every warp; 3 warp-specialised (odd warps DFMA, even warps IMAD);
// 4 u32 butterfly all warps; 5 f64 butterfly all warps; 6 butterflies warp-specialised 2:1 (warp % 3 == 2 runs doubles)
template <int MODE> __global__ void bench(u32* out, u32 a, u32 b, u32 c, double fb, double fc) {
    u32 x[CHAINS]; double y[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { x[i] = threadIdx.x * 7 + i + a; y[i] = (double)(threadIdx.x + i) * 1e-3; }
    const int warp = threadIdx.x >> 5;
    const bool f64_warp = MODE == 1 || MODE == 5 || (MODE == 3 && (warp & 1)) || (MODE == 6 && warp % 3 == 2);
    const bool both = MODE == 2;
    if (MODE <= 3) {
        if (both) for (int it = 0; it < ITER; it++) { imad_chain(x, b, c); dfma_chain(y, fb, fc); }
        else if (f64_warp) for (int it = 0; it < ITER; it++) dfma_chain(y, fb, fc);
        else for (int it = 0; it < ITER; it++) imad_chain(x, b, c);
    } else {
        const double p = (double)P28, pinv = 1.0 / (double)P28;
        if (f64_warp) { for (int i = 0; i < CHAINS; i++) y[i] = (double)((threadIdx.x * 977 + i) % P28);
                        for (int it = 0; it < ITER; it++) bfly_f64(y, fb, pinv, p); }
        else for (int it = 0; it < ITER; it++) bfly_u32(x, a, b, P28);
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s ^= x[i] ^ (u32)__double2ll_rn(y[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double int_share, double f64_share, double ops_int, double ops_f64) {
    const int threads = 384;                                    // the throughput kernel's CTA: 12 warps, one CTA per SM -> 2 CTAs here
    u32* out; cudaMalloc(&out, 148 * 8 * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<148 * 4, threads>>>(out, 3, 0x9E3779B1u, 12345, 123456789.0, 0.5);
    cudaEventRecord(e0);
    bench<MODE><<<148 * 4, threads>>>(out, 3, 0x9E3779B1u, 12345, 123456789.0, 0.5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double thr = 148.0 * 4 * threads * ITER * CHAINS;
    double ni = thr * int_share * ops_int, nf = thr * f64_share * ops_f64;
    printf("%-44s %8.3f ms  int %6.1f  f64 %6.1f  lanes/clk/SM (sum %6.1f)\n", name, ms, ni / (ms * 1e-3) / 148 / 1.965e9,
           nf / (ms * 1e-3) / 148 / 1.965e9, (ni + nf) / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    run<0>("IMAD, all warps", 1, 0, 1, 0);
    run<1>("DFMA, all warps", 0, 1, 0, 1);
    run<2>("IMAD + DFMA interleaved in every warp", 1, 1, 1, 1);
    run<3>("IMAD warps | DFMA warps (1:1)", 0.5, 0.5, 1, 1);
    run<4>("u32 butterfly (per output), all warps", 1, 0, 1, 0);
    run<5>("f64 butterfly (per output), all warps", 0, 1, 0, 1);
    run<6>("u32 butterfly warps | f64 butterfly warps 2:1", 2.0 / 3, 1.0 / 3, 1, 1);
    return 0;
}
