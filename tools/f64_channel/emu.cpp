// Host emulation of the FP64 RNS channel prototype (ntt_f64.cuh) for 32 emulated lanes:
//   1. mulmod on doubles == 64-bit integer arithmetic over the ranges the transforms produce;
//   2. forward and inverse 1024-point transforms in doubles give, position by position, the residues of the u32 channel of the
//      same prime (torus-fhe_b200/csrc/ntt_rns.cuh);
//   3. the three-prime exact product with prime 2 carried in doubles (primes 0 and 1 in u32) and the unchanged Garner lift equals
//      the schoolbook negacyclic product mod 2^64, including the largest-magnitude operands;
//   4. every intermediate stays an exact integer far below 2^53.
// Build: g++ -O2 -std=c++17 -ffp-contract=off -I torus-fhe_b200/csrc -I tools/f64_channel tools/f64_channel/emu.cpp -o f64_emu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tables.h"
#include "ntt_f64.cuh"

using namespace rns;
static HostTables T;
static const int PI = 2;                       // the channel that would move to the FP64 pipe

static u64 rnd_state = 0x9E3779B97F4A7C15ull;
static u64 rnd() { rnd_state ^= rnd_state << 13; rnd_state ^= rnd_state >> 7; rnd_state ^= rnd_state << 17; return rnd_state; }

static double twA_f[2][31], twB_f[2][31][32];
static rnsf::Mod M;
static double max_abs = 0;

static void track(const double* x, int n) {
    for (int i = 0; i < n; i++) {
        if (x[i] != std::nearbyint(x[i])) { printf("FAIL non-integer intermediate\n"); exit(1); }
        if (std::fabs(x[i]) > max_abs) max_abs = std::fabs(x[i]);
    }
}

// ---- u32 channel (reference for the residues), as in tests/host_emu/ntt_emu.cpp
static void warp_fwd_u32(int pi, const u32* a, u32 out[32][32]) {
    static u32 tile[TILE_WORDS];
    const u32 p = T.c.p[pi];
    u32 x[32];
    for (int lane = 0; lane < 32; lane++) {
        for (int r = 0; r < 32; r++) x[r] = a[32 * r + lane];
        fwd_passA(x, T.c.twA[pi][0], p);
        for (int r = 0; r < 32; r++) tile[r * TILE_STRIDE + lane] = x[r];
    }
    for (int lane = 0; lane < 32; lane++) {
        for (int c = 0; c < 32; c++) x[c] = reduce_to_4p(tile[lane * TILE_STRIDE + c], 4 * p);
        fwd_passB(x, T.twB.data() + ((size_t)pi * 2 + 0) * 31 * 32 + lane, p);
        for (int c = 0; c < 32; c++) out[lane][c] = x[c];
    }
}
static void warp_inv_u32(int pi, u32 in[32][32], u32* a) {
    static u32 tile[TILE_WORDS];
    const u32 p = T.c.p[pi];
    u32 x[32];
    for (int lane = 0; lane < 32; lane++) {
        for (int c = 0; c < 32; c++) x[c] = in[lane][c];
        inv_passB(x, T.twB.data() + ((size_t)pi * 2 + 1) * 31 * 32 + lane, p);
        for (int c = 0; c < 32; c++) tile[lane * TILE_STRIDE + c] = x[c];
    }
    for (int lane = 0; lane < 32; lane++) {
        for (int r = 0; r < 32; r++) x[r] = tile[r * TILE_STRIDE + lane];
        inv_passA(x, T.c.twA[pi][1], p);
        for (int r = 0; r < 32; r++) a[32 * r + lane] = x[r];
    }
}

// ---- the same two transforms in doubles: no range correction anywhere
static void warp_fwd_f64(const double* a, double out[32][32]) {
    static double tile[32 * 33];
    double x[32];
    for (int lane = 0; lane < 32; lane++) {
        for (int r = 0; r < 32; r++) x[r] = a[32 * r + lane];
        rnsf::ct32(x, rnsf::TwUniform{twA_f[0]}, M);
        track(x, 32);
        for (int r = 0; r < 32; r++) tile[r * 33 + lane] = x[r];
    }
    for (int lane = 0; lane < 32; lane++) {
        for (int c = 0; c < 32; c++) x[c] = tile[lane * 33 + c];
        rnsf::ct32(x, rnsf::TwLane{&twB_f[0][0][lane]}, M);
        track(x, 32);
        for (int c = 0; c < 32; c++) out[lane][c] = x[c];
    }
}
static void warp_inv_f64(double in[32][32], double* a) {
    static double tile[32 * 33];
    double x[32];
    for (int lane = 0; lane < 32; lane++) {
        for (int c = 0; c < 32; c++) x[c] = in[lane][c];
        rnsf::gs32(x, rnsf::TwLane{&twB_f[1][0][lane]}, M);
        track(x, 32);
        for (int c = 0; c < 32; c++) tile[lane * 33 + c] = x[c];
    }
    for (int lane = 0; lane < 32; lane++) {
        for (int r = 0; r < 32; r++) x[r] = tile[r * 33 + lane];
        rnsf::gs32(x, rnsf::TwUniform{twA_f[1]}, M);
        track(x, 32);
        for (int r = 0; r < 32; r++) a[32 * r + lane] = x[r];
    }
}

static u32 res_of(double v, u32 p) {          // exact integer double -> [0, p) through 64-bit integers (checker side)
    const int64_t i = (int64_t)v;
    const int64_t r = i % (int64_t)p;
    return (u32)(r < 0 ? r + p : r);
}

int main() {
    const u32 p = T.c.p[PI];
    M.p = (double)p;
    M.pinv = 1.0 / (double)p;
    for (int dir = 0; dir < 2; dir++) {
        for (int e = 0; e < 31; e++) {
            twA_f[dir][e] = (double)T.c.twA[PI][dir][e].x;
            for (int lane = 0; lane < 32; lane++) twB_f[dir][e][lane] = (double)T.twB[(((size_t)PI * 2 + dir) * 31 + e) * 32 + lane].x;
        }
    }
    int fails = 0;
    // 1. scalar primitive over the ranges in use (|y| up to 2^41: the inverse pass sums; w any residue), and the integer bridges
    for (int it = 0; it < 400000 && !fails; it++) {
        const int64_t y = (int64_t)(rnd() % ((u64)1 << 42)) - ((int64_t)1 << 41);
        const u32 w = rnd() % p;
        const double t = rnsf::mulmod((double)y, (double)w, M);
        const int64_t ym = ((y % (int64_t)p) + p) % p;
        const u32 want = (u32)((u64)ym * w % p);
        if (t != std::nearbyint(t) || std::fabs(t) > 0.5 * p + std::fabs((double)y) * 6e-8 + 1 || res_of(t, p) != want) { fails++; printf("FAIL mulmod y=%lld w=%u t=%.1f\n", (long long)y, w, t); }
        if (rnsf::to_residue(t, M) != want) { fails++; printf("FAIL to_residue\n"); }
        const int32_t v = (int32_t)rnd();
        if (rnsf::from_i32(v) != (double)v) { fails++; printf("FAIL from_i32\n"); }
    }
    // 2. transforms against the u32 channel, position by position
    for (int trial = 0; trial < 4 && !fails; trial++) {
        std::vector<u32> au(N), backu(N);
        std::vector<double> af(N), backf(N);
        for (int i = 0; i < N; i++) {
            const int d = trial == 0 ? (int)(rnd() % 128) - 64 : trial == 1 ? -64 : trial == 2 ? 63 : (int)(rnd() % (2u * p)) - (int)p;
            af[i] = (double)d;
            au[i] = d >= 0 ? (u32)d % p : p - (u32)(-d) % p;
        }
        static u32 AU[32][32];
        static double AF[32][32];
        warp_fwd_u32(PI, au.data(), AU);
        warp_fwd_f64(af.data(), AF);
        for (int l = 0; l < 32; l++)
            for (int c = 0; c < 32; c++)
                if (rnsf::to_residue(AF[l][c], M) != AU[l][c] % p) { fails++; printf("FAIL fwd trial=%d lane=%d c=%d\n", trial, l, c); l = 32; break; }
        for (int l = 0; l < 32; l++) for (int c = 0; c < 32; c++) AU[l][c] = reduce_to_4p(AU[l][c], 4 * p);
        warp_inv_u32(PI, AU, backu.data());
        warp_inv_f64(AF, backf.data());
        for (int i = 0; i < N; i++)
            if (rnsf::to_residue(backf[i], M) != backu[i] % p) { fails++; printf("FAIL inv trial=%d i=%d\n", trial, i); break; }
    }
    // 3. exact product: primes 0, 1 in u32, prime 2 in doubles, unchanged Garner lift
    for (int trial = 0; trial < 5 && !fails; trial++) {
        const int L2 = 8;   // 2l = 8 digit polynomials (l = 4, the widest N = 1024 set) accumulated before the inverse
        std::vector<int64_t> d(L2 * N), key(L2 * N);
        std::vector<u64> ref(N, 0);
        const int half = trial >= 1 && trial <= 3 ? 8 : 64;      // Bg/2: 2l N Bg/2 2^63 must stay below M/4 -> l = 4 goes with Bg = 2^4
        for (int i = 0; i < L2 * N; i++) { d[i] = (int64_t)(rnd() % (2 * 8)) - 8; key[i] = (int64_t)rnd(); }
        if (trial == 1) for (int i = 0; i < L2 * N; i++) { d[i] = -half; key[i] = INT64_MIN; }
        if (trial == 2) for (int i = 0; i < L2 * N; i++) { d[i] = half - 1; key[i] = INT64_MAX; }
        if (trial == 3) for (int i = 0; i < L2 * N; i++) { d[i] = (i & 1) ? half - 1 : -half; key[i] = (i & 2) ? INT64_MIN : INT64_MAX; }
        if (trial == 4) for (int i = 0; i < L2 * N; i++) { d[i] = i < 4 * N ? (int64_t)(rnd() % 128) - 64 : 0; }   // l = 2, 7-bit digits
        for (int s = 0; s < L2; s++)
            for (int i = 0; i < N; i++) {
                if (!d[s * N + i]) continue;
                for (int j = 0; j < N; j++) {
                    const u64 t = (u64)d[s * N + i] * (u64)key[s * N + j];
                    if (i + j < N) ref[i + j] += t; else ref[i + j - N] -= t;
                }
            }
        static u32 res[NP][1024];
        for (int pi = 0; pi < 2; pi++) {                         // u32 channels, as the kernels run them
            const u32 q = T.c.p[pi];
            static u32 D[32][32], K[32][32];
            static u64 ACC[32][32];
            for (int l = 0; l < 32; l++) for (int c = 0; c < 32; c++) ACC[l][c] = 0;
            std::vector<u32> tmp(N);
            for (int s = 0; s < L2; s++) {
                for (int i = 0; i < N; i++) tmp[i] = (u32)(d[s * N + i] + 64) + (q - 64);
                warp_fwd_u32(pi, tmp.data(), D);
                for (int i = 0; i < N; i++) tmp[i] = residue_i64(key[s * N + i], q);
                warp_fwd_u32(pi, tmp.data(), K);
                for (int l = 0; l < 32; l++)
                    for (int c = 0; c < 32; c++) ACC[l][c] += (u64)D[l][c] * mulmod(K[l][c] % q, T.c.key_scale[pi], q);
            }
            static u32 R[32][32];
            for (int l = 0; l < 32; l++)
                for (int c = 0; c < 32; c++) {
                    const u32 m = (u32)ACC[l][c] * T.c.pinv_neg[pi];
                    u32 v = (u32)((ACC[l][c] + (u64)m * q) >> 32);
                    v = umin32(v, v - 8 * q);
                    R[l][c] = umin32(v, v - 4 * q);
                }
            warp_inv_u32(pi, R, res[pi]);
        }
        {                                                        // the FP64 channel
            static double D[32][32], K[32][32], ACC[32][32];
            for (int l = 0; l < 32; l++) for (int c = 0; c < 32; c++) ACC[l][c] = 0;
            std::vector<double> tmp(N);
            const u32 ninv = invmod(N % p, p);
            for (int s = 0; s < L2; s++) {
                for (int i = 0; i < N; i++) tmp[i] = (double)d[s * N + i];                        // signed digits as they are
                warp_fwd_f64(tmp.data(), D);
                for (int i = 0; i < N; i++) tmp[i] = (double)residue_i64(key[s * N + i], p);
                warp_fwd_f64(tmp.data(), K);
                for (int l = 0; l < 32; l++)
                    for (int c = 0; c < 32; c++) {
                        const double kk = (double)mulmod(rnsf::to_residue(K[l][c], M), ninv, p);   // stored key: NTT(K) N^-1 mod p, as a double
                        ACC[l][c] += rnsf::mulmod(D[l][c], kk, M);
                    }
            }
            track(&ACC[0][0], 1024);
            std::vector<double> back(N);
            warp_inv_f64(ACC, back.data());
            for (int i = 0; i < N; i++) res[2][i] = rnsf::to_residue(back[i], M);
        }
        for (int i = 0; i < N; i++) {
            const u64 got = crt_lift(res[0][i], res[1][i], res[2][i], T.c.crt);
            if (got != ref[i]) { fails++; printf("FAIL product trial=%d i=%d got=%llx ref=%llx\n", trial, i, (unsigned long long)got, (unsigned long long)ref[i]); break; }
        }
    }
    // 5. feasibility of a 47-bit prime in the same arithmetic (the N = 2048 sets: one such channel would replace two u32 channels):
    //    operands larger than an 11-stage forward pass lets them grow (|y| < 12 p; growth is < 0.54 p per stage), product checked against 128-bit
    //    integers.  The hard limit is the quotient: |y w / p| must stay below 2^51 for the 1.5 * 2^52 rounding trick, i.e. |y| < 16 p at 47 bits.
    {
        auto mulmod128 = [](u64 a, u64 b, u64 m) { return (u64)((unsigned __int128)a * b % m); };
        auto powmod128 = [&](u64 a, u64 e, u64 m) { u64 r = 1; while (e) { if (e & 1) r = mulmod128(r, a, m); a = mulmod128(a, a, m); e >>= 1; } return r; };
        auto is_prime = [&](u64 n) {
            for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
                if (n % a == 0) return n == a;
                u64 d = n - 1; int r = 0;
                while (!(d & 1)) { d >>= 1; r++; }
                u64 x = powmod128(a, d, n);
                if (x == 1 || x == n - 1) continue;
                bool comp = true;
                for (int i = 1; i < r && comp; i++) { x = mulmod128(x, x, n); if (x == n - 1) comp = false; }
                if (comp) return false;
            }
            return true;
        };
        u64 p47 = 0;
        for (u64 c = ((u64)1 << 47) / 4096; c > 0 && !p47; c--) if (is_prime(c * 4096 + 1)) p47 = c * 4096 + 1;   // largest p = 1 (mod 4096) below 2^47
        rnsf::Mod M47{(double)p47, 1.0 / (double)p47};
        double worst = 0;
        for (int it = 0; it < 400000 && !fails; it++) {
            const int64_t y = (int64_t)(rnd() % (24 * p47)) - (int64_t)(12 * p47);
            const u64 w = rnd() % p47;
            const double t = rnsf::mulmod((double)y, (double)w, M47);
            const __int128 diff = (__int128)y * (__int128)w - (__int128)(int64_t)t;
            if (t != std::nearbyint(t) || diff % (__int128)p47 != 0) { fails++; printf("FAIL 47-bit mulmod y=%lld w=%llu\n", (long long)y, (unsigned long long)w); }
            if (std::fabs(t) / (double)p47 > worst) worst = std::fabs(t) / (double)p47;
        }
        printf("47-bit prime %llu: mulmod exact for |y| < 12 p, |result| <= %.3f p\n", (unsigned long long)p47, worst);
        if (worst > 1.6) { fails++; printf("FAIL 47-bit range\n"); }
    }
    printf("largest intermediate: 2^%.1f (exact integers need < 2^53)\n", std::log2(max_abs));
    if (max_abs >= 9007199254740992.0 / 256) { fails++; printf("FAIL headroom\n"); }
    printf(fails ? "f64_emu: %d FAILURES\n" : "f64_emu: OK\n", fails);
    return fails ? 1 : 0;
}
