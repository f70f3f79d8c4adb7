// Host emulation of the FP64 FFT channel (torus-fhe_b200/csrc/fft64_core.cuh, tables_fft.h): the per-thread passes the CUDA kernels run,
// executed for 32 emulated lanes with the kernels' own buffer slots, against (1) the definition of the transform -- evaluation at the
// roots of X^512 - i -- and (2) the wrap-around integer external product sum_s digit_s * key_s mod (X^N + 1, 2^64), and mod 2^32 in the Torus32
// form (16-bit digit fields, two key limbs), including operands at the extremes of their ranges.  Prints the worst distance of a limb result to an integer (the margin of the rounding).
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "fft64_core.cuh"
#include "tables_fft.h"

using namespace mkf;
typedef std::vector<cpx> Buf;

static std::vector<cpx> TW;

// one warp: forward transform from the digit words (dig: 256 words) or from real coefficients; result in the spectrum layout [c 32 + lane]
static Buf warp_forward(const uint32_t* dig, const double* real, int half_bg, bool wide = false) {
    cpx v[32][16];
    Buf buf(M), out(M);
    for (int lane = 0; lane < 32; lane++) {
        const int h = lane >> 4, l16 = lane & 15;
        if (dig && wide) fwd_stage0_digits_wide(v[lane], dig, dig + 256, lane, half_bg);
        else if (dig) fwd_stage0_digits(v[lane], dig, lane, half_bg);
        else
            for (int r = 0; r < 16; r++) {
                const int j = 16 * r + l16;
                v[lane][r] = fwd_stage0_real(real[j], real[j + 256], real[j + 512], real[j + 768], h);
            }
        fwd_passA(v[lane], TW.data(), h);
        for (int r = 0; r < 16; r++) buf[transpose_slot(h, r, l16)] = v[lane][r];
    }
    for (int lane = 0; lane < 32; lane++) {
        const int h = lane >> 4, l16 = lane & 15;
        for (int c = 0; c < 16; c++) v[lane][c] = buf[transpose_slot(h, l16, c)];
        fwd_passB(v[lane], TW.data(), lane);
        for (int c = 0; c < 16; c++) out[c * 32 + lane] = v[lane][c];
    }
    return out;
}
// one warp: inverse transform of a spectrum up to the last stage; result in natural position order
static Buf warp_inverse(const Buf& spec) {
    cpx v[32][16];
    Buf buf(M), out(M);
    for (int lane = 0; lane < 32; lane++) {
        const int h = lane >> 4, l16 = lane & 15;
        for (int c = 0; c < 16; c++) v[lane][c] = spec[c * 32 + lane];
        inv_passB(v[lane]);
        for (int c = 0; c < 16; c++) buf[transpose_slot(h, l16, c)] = v[lane][c];
    }
    for (int lane = 0; lane < 32; lane++) {
        const int h = lane >> 4, l16 = lane & 15;
        for (int r = 0; r < 16; r++) v[lane][r] = buf[transpose_slot(h, r, l16)];
        inv_passA(v[lane], TW.data(), l16);
        for (int r = 0; r < 16; r++) out[h * 256 + r * 16 + l16] = v[lane][r];
    }
    return out;
}

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); fails++; } } while (0)

int main() {
    HostTablesFFT T;
    TW.resize(T_ENTRIES);
    for (int i = 0; i < T_ENTRIES; i++) TW[i] = cpx{T.tw[2 * i], T.tw[2 * i + 1]};
    std::mt19937_64 rng(7);
    const int bgbit = 7, half = 1 << (bgbit - 1);
    // ---- 1. forward transform vs the definition: spectrum point at position pos = lane 16 + c is a~(zeta omega^k), k = bitrev9(pos)
    {
        std::vector<uint32_t> dig(256);
        std::vector<double> a(N);
        for (int j = 0; j < 256; j++) {
            uint32_t w = 0;
            for (int b = 0; b < 4; b++) { const uint32_t d = rng() % (2 * half); w |= d << (8 * b); a[j + 256 * b] = (double)d - half; }
            dig[j] = w;
        }
        const Buf A = warp_forward(dig.data(), nullptr, half), A2 = warp_forward(nullptr, a.data(), 0);
        double worst = 0, worst2 = 0;
        for (int pos = 0; pos < M; pos += 37) {
            int k = 0;
            for (int b = 0; b < 9; b++) k |= ((pos >> b) & 1) << (8 - b);
            const long double PI = 3.14159265358979323846264338327950288L;
            std::complex<long double> x = std::polar(1.0L, PI / N + 2 * PI * k / M), xp = 1, sum = 0;
            for (int j = 0; j < M; j++) { sum += std::complex<long double>(a[j], a[j + M]) * xp; xp *= x; }
            const int lane = pos >> 4, c = pos & 15;
            worst = std::fmax(worst, (double)std::abs(sum - std::complex<long double>(A[c * 32 + lane].x, A[c * 32 + lane].y)));
            worst2 = std::fmax(worst2, std::hypot(A[c * 32 + lane].x - A2[c * 32 + lane].x, A[c * 32 + lane].y - A2[c * 32 + lane].y));
        }
        CHECK(worst < 1e-8, "forward transform differs from the definition by %g", worst);
        CHECK(worst2 == 0.0, "digit-byte and real-coefficient first stages differ by %g", worst2);
    }
    // ---- 2. external products: 2l digit polynomials x Torus64 key polynomials, three limbs, rounding, recombination
    double worst_frac = 0;
    for (int trial = 0; trial < 12; trial++) {
        const int L2 = 4;
        std::vector<std::vector<uint32_t>> dig(L2, std::vector<uint32_t>(256));
        std::vector<std::vector<int64_t>> d(L2, std::vector<int64_t>(N)), key(L2, std::vector<int64_t>(N));
        for (int s = 0; s < L2; s++)
            for (int i = 0; i < N; i++) {
                uint32_t byte;
                int64_t kv;
                if (trial == 0) { byte = 0; kv = s & 1 ? INT64_MIN : INT64_MAX; }                                       // every digit -Bg/2, keys at +-2^63
                else if (trial == 1) { byte = rng() & 1 ? 2 * half - 1 : 0; kv = (rng() & 1 ? 1 : -1) * (int64_t)0x1FFFFF1FFFFF1FFFll; }   // limb maxima
                else if (trial == 2) { byte = 0; kv = (int64_t)0x7FFFFBFFFFF00000ll * ((i & 1) ? 1 : -1); }
                else { byte = rng() % (2 * half); kv = (int64_t)rng(); }
                d[s][i] = (int64_t)byte - half;
                key[s][i] = kv;
                dig[s][i & 255] |= byte << (8 * (i >> 8));
            }
        // exact: wrap-around schoolbook
        std::vector<uint64_t> want(N, 0);
        for (int s = 0; s < L2; s++)
            for (int i = 0; i < N; i++) {
                if (!d[s][i]) continue;
                for (int j = 0; j < N; j++) {
                    const uint64_t p = (uint64_t)d[s][i] * (uint64_t)key[s][j];
                    if (i + j < N) want[i + j] += p; else want[i + j - N] -= p;
                }
            }
        // the kernel's way
        std::vector<Buf> spec(L2);
        for (int s = 0; s < L2; s++) spec[s] = warp_forward(dig[s].data(), nullptr, half);
        std::vector<Buf> Y(LIMBS);
        for (int limb = 0; limb < LIMBS; limb++) {
            Buf acc(M, cpx{0, 0});
            for (int s = 0; s < L2; s++) {
                std::vector<double> kl(N);
                for (int i = 0; i < N; i++) kl[i] = key_limb(key[s][i], limb);
                Buf K = warp_forward(nullptr, kl.data(), 0);
                for (int p = 0; p < M; p++) {
                    const cpx k = {K[p].x * (1.0 / M), K[p].y * (1.0 / M)}, x = spec[s][p];
                    acc[p].x = fma(x.x, k.x, fma(-x.y, k.y, acc[p].x));
                    acc[p].y = fma(x.x, k.y, fma(x.y, k.x, acc[p].y));
                }
            }
            Y[limb] = warp_inverse(acc);
        }
        for (int j = 0; j < 256; j++) {
            uint64_t R[4] = {0, 0, 0, 0};
            for (int limb = 0; limb < LIMBS; limb++) {
                recombine_limb(R, Y[limb][j], Y[limb][j + 256], TW[T_WJ + j], TW[T_UT + j], limb_shift(LIMBS, limb));
                // rounding margin of this limb: redo the last stage in plain arithmetic
                cpx lo = Y[limb][j], hi = Y[limb][j + 256];
                ct(lo, hi, TW[T_WJ + j]);
                const cpx e = cmul(lo, TW[T_UT + j]);
                worst_frac = std::fmax(worst_frac, std::fmax(std::fabs(e.x - std::nearbyint(e.x)), std::fabs(e.y - std::nearbyint(e.y))));
            }
            for (int b = 0; b < 4; b++)
                CHECK(R[b] - round_k(LIMBS) == want[j + 256 * b], "trial %d coefficient %d: got %llx want %llx", trial, j + 256 * b,
                      (unsigned long long)(R[b] - round_k(LIMBS)), (unsigned long long)want[j + 256 * b]);
            if (fails > 5) break;
        }
    }
    // ---- 3. Torus32 form: 16-bit digit fields (Bg = 2^10, l = 2), two 16-bit limbs of 32-bit key words, R mod 2^32
    for (int trial = 0; trial < 8; trial++) {
        const int L2 = 4, bg = 10, hb = 1 << (bg - 1);
        std::vector<std::vector<uint32_t>> dig(L2, std::vector<uint32_t>(512));
        std::vector<std::vector<int64_t>> d(L2, std::vector<int64_t>(N)), key(L2, std::vector<int64_t>(N));
        for (int s = 0; s < L2; s++)
            for (int i = 0; i < N; i++) {
                uint32_t f;
                int32_t kv;
                if (trial == 0) { f = 0; kv = s & 1 ? INT32_MIN : INT32_MAX; }
                else if (trial == 1) { f = rng() & 1 ? 2 * hb - 1 : 0; kv = (rng() & 1 ? 1 : -1) * 0x7FFF7FFF; }
                else { f = rng() % (2 * hb); kv = (int32_t)rng(); }
                d[s][i] = (int64_t)f - hb;
                key[s][i] = kv;
                const int b = i >> 8, j = i & 255;          // word (b & 1, j): field b >> 1
                dig[s][(b & 1) * 256 + j] |= f << (16 * (b >> 1));
            }
        std::vector<uint32_t> want(N, 0);
        for (int s = 0; s < L2; s++)
            for (int i = 0; i < N; i++) {
                if (!d[s][i]) continue;
                for (int j = 0; j < N; j++) {
                    const uint32_t p = (uint32_t)d[s][i] * (uint32_t)key[s][j];
                    if (i + j < N) want[i + j] += p; else want[i + j - N] -= p;
                }
            }
        std::vector<Buf> spec(L2);
        for (int s = 0; s < L2; s++) spec[s] = warp_forward(dig[s].data(), nullptr, hb, true);
        std::vector<Buf> Y(LIMBS_T32);
        for (int limb = 0; limb < LIMBS_T32; limb++) {
            Buf acc(M, cpx{0, 0});
            for (int s = 0; s < L2; s++) {
                std::vector<double> kl(N);
                for (int i = 0; i < N; i++) kl[i] = key_limb(key[s][i], limb, LIMBS_T32);
                Buf K = warp_forward(nullptr, kl.data(), 0);
                for (int p = 0; p < M; p++) {
                    const cpx k = {K[p].x * (1.0 / M), K[p].y * (1.0 / M)}, x = spec[s][p];
                    acc[p].x = fma(x.x, k.x, fma(-x.y, k.y, acc[p].x));
                    acc[p].y = fma(x.x, k.y, fma(x.y, k.x, acc[p].y));
                }
            }
            Y[limb] = warp_inverse(acc);
        }
        for (int j = 0; j < 256 && fails <= 5; j++) {
            uint64_t R[4] = {0, 0, 0, 0};
            for (int limb = 0; limb < LIMBS_T32; limb++)
                recombine_limb(R, Y[limb][j], Y[limb][j + 256], TW[T_WJ + j], TW[T_UT + j], limb_shift(LIMBS_T32, limb));
            for (int b = 0; b < 4; b++) {
                const uint64_t top = (R[b] - round_k(LIMBS_T32)) << 32;           // what the kernel adds to the accumulator
                CHECK(top == ((uint64_t)want[j + 256 * b] << 32), "Torus32 trial %d coefficient %d: got %llx want %llx", trial, j + 256 * b,
                      (unsigned long long)top, (unsigned long long)want[j + 256 * b] << 32);
            }
        }
    }
    CHECK(worst_frac < 1.0 / 64, "a limb result is %g away from an integer", worst_frac);
    printf("worst distance of a limb result to an integer: %.3g (2^%.1f)\n", worst_frac, std::log2(worst_frac));
    if (!fails) printf("fft64_emu: OK\n");
    return fails != 0;
}
