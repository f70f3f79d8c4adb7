// Host emulation of the N = 2048 arithmetic core (torus-fhe_b200/csrc/ntt2048.cuh, groundwork for the 16..256-party parameter sets):
// the per-thread passes run for 32 emulated lanes and are checked against the O(N^2) definition of the negacyclic transform, the
// round trip, and the whole four-prime pipeline (26-bit gadget digits x 64-bit key, Montgomery pointwise products, four-prime Garner
// lift) against an exact schoolbook negacyclic product mod 2^64, extreme magnitudes included.
// Build: g++ -O2 -std=c++17 -I torus-fhe_b200/csrc tests/host_emu/ntt2048_emu.cpp -o ntt2048_emu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tables2048.h"

using namespace rns2k;
static HostTables T;

static u64 rnd_state = 0x9e3779b97f4a7c15ull;
static u64 rnd() { rnd_state ^= rnd_state << 13; rnd_state ^= rnd_state >> 7; rnd_state ^= rnd_state << 17; return rnd_state; }

// position (in the transformed, bit-reversed-order polynomial) of element c of half h of thread `lane`
static int ntt_pos(int lane, int h, int c) { return 32 * (lane + 32 * h) + c; }

// coefficient order in (values in [0, 2p)) -> transformed out[lane][h][c], values in [0, 14p)
static void warp_fwd(int pi, const u32* a, u32 out[32][2][32]) {
    static u32 tile[TILE_WORDS];
    const u32 p = T.c.p[pi];
    for (int lane = 0; lane < 32; lane++) {
        u32 x[64];
        for (int r = 0; r < 64; r++) x[r] = a[32 * r + lane];
        fwd_passA64(x, T.c.twA[pi][0], p);
        for (int r = 0; r < 64; r++) {
            if (x[r] >= 14ull * p) { printf("FAIL passA range\n"); exit(1); }
            tile[r * TILE_STRIDE + lane] = x[r];
        }
    }
    for (int lane = 0; lane < 32; lane++)
        for (int h = 0; h < 2; h++) {
            u32 y[32];
            for (int c = 0; c < 32; c++) y[c] = rns2k::reduce_to_4p(tile[(lane + 32 * h) * TILE_STRIDE + c], 8 * p, 4 * p);
            fwd_passB32(y, T.twB.data() + twB_index(pi, 0, h, 0, lane), p);
            for (int c = 0; c < 32; c++) out[lane][h][c] = y[c];
        }
}
// transformed in[lane][h][c] (values in [0, 4p)) -> coefficient order, scaled by N, values in [0, 4p)
static void warp_inv(int pi, u32 in[32][2][32], u32* a) {
    static u32 tile[TILE_WORDS];
    const u32 p = T.c.p[pi];
    for (int lane = 0; lane < 32; lane++)
        for (int h = 0; h < 2; h++) {
            u32 y[32];
            for (int c = 0; c < 32; c++) y[c] = in[lane][h][c];
            inv_passB32(y, T.twB.data() + twB_index(pi, 1, h, 0, lane), p, opaque_multiple(4 * p));
            for (int c = 0; c < 32; c++) tile[(lane + 32 * h) * TILE_STRIDE + c] = y[c];
        }
    for (int lane = 0; lane < 32; lane++) {
        u32 x[64];
        for (int r = 0; r < 64; r++) x[r] = tile[r * TILE_STRIDE + lane];
        inv_passA64(x, T.c.twA[pi][1], p, opaque_multiple(4 * p));
        for (int r = 0; r < 64; r++) a[32 * r + lane] = x[r];
    }
}

int main() {
    int fails = 0;
    for (int pi = 0; pi < NP; pi++) {
        const u32 p = T.c.p[pi];
        if (p >= (1u << 28) || (p - 1) % (2 * N) || rns::powmod(T.psi[pi], N, p) != p - 1) { printf("FAIL prime %d\n", pi); return 1; }
        if ((u32)(p * (0u - T.c.pinv_neg[pi])) != 1u) { printf("FAIL pinv %d\n", pi); return 1; }
        std::vector<u32> a(N), back(N);
        for (auto& v : a) v = rnd() % (2 * p);
        static u32 A[32][2][32];
        warp_fwd(pi, a.data(), A);
        std::vector<u32> psipow(2 * N);
        psipow[0] = 1;
        for (int i = 1; i < 2 * N; i++) psipow[i] = rns::mulmod(psipow[i - 1], T.psi[pi], p);
        for (int lane = 0; lane < 32 && !fails; lane += 5)
            for (int h = 0; h < 2; h++)
                for (int c = 0; c < 32; c += 7) {
                    const u32 e = 2 * rns::brev(ntt_pos(lane, h, c), LOGN) + 1;      // out[pos] = A(psi^(2 brev(pos) + 1))
                    u64 s = 0;
                    for (int i = 0; i < N; i++) s = (s + (u64)(a[i] % p) * psipow[(u64)i * e % (2 * N)]) % p;
                    if (A[lane][h][c] >= 14ull * p || A[lane][h][c] % p != s) { fails++; printf("FAIL fwd prime=%d lane=%d h=%d c=%d\n", pi, lane, h, c); break; }
                }
        for (int l = 0; l < 32; l++) for (int h = 0; h < 2; h++) for (int c = 0; c < 32; c++) A[l][h][c] %= p;
        warp_inv(pi, A, back.data());
        for (int i = 0; i < N; i++)
            if (back[i] >= 4ull * p || back[i] % p != (u64)(a[i] % p) * N % p) { fails++; printf("FAIL roundtrip prime=%d i=%d\n", pi, i); break; }
    }
    // Garner lift on random signed integers of up to 109 bits (|R| < M/4), from their residues in the lazy range [0, 4p)
    for (int it = 0; it < 200000 && !fails; it++) {
        const __int128 mag = ((__int128)(rnd() & ((1ull << 45) - 1)) << 64) | rnd();                 // < 2^109
        const __int128 R = (rnd() & 1) ? -mag : mag;
        u32 r[NP];
        for (int i = 0; i < NP; i++) {
            const __int128 m = R % (__int128)T.c.p[i];
            r[i] = (u32)(m < 0 ? m + T.c.p[i] : m) + (u32)(rnd() % 4) * T.c.p[i];
        }
        if (crt4_lift(r, T.c) != (u64)R) { fails++; printf("FAIL crt4\n"); }
        {   // the quotient-corrected lift on y_i = R (M / p_i)^-1 mod p_i, in a lazy representative (+ 0..3 p_i)
            u32 y[NP];
            for (int i = 0; i < NP; i++) y[i] = rns::mulmod(r[i] % PRIMES[i], T.c.yscale[i], PRIMES[i]) + (u32)((it + i) & 3) * PRIMES[i];
            if (crt4_lift_kappa(y, T.c) != (u64)R) { fails++; printf("FAIL crt4 kappa\n"); }
        }
    }
    // exact negacyclic product: 2l = 2 polynomials of 26-bit signed digits against 64-bit keys (l = 1, Bg = 2^26), through the four primes
    for (int trial = 0; trial < 4 && !fails; trial++) {
        const int L2 = 2;
        std::vector<int64_t> d(L2 * N), key(L2 * N);
        std::vector<u64> ref(N, 0);
        for (int i = 0; i < L2 * N; i++) { d[i] = (int64_t)(rnd() % (1u << 26)) - (1 << 25); key[i] = (int64_t)rnd(); }
        if (trial == 1) for (int i = 0; i < L2 * N; i++) { d[i] = -(1 << 25); key[i] = INT64_MIN; }    // largest magnitude: 2^100
        if (trial == 2) for (int i = 0; i < L2 * N; i++) { d[i] = (1 << 25) - 1; key[i] = INT64_MAX; }
        if (trial == 3) for (int i = 0; i < L2 * N; i++) { d[i] = (i & 1) ? (1 << 25) - 1 : -(1 << 25); key[i] = (i & 2) ? INT64_MIN : INT64_MAX; }
        for (int s = 0; s < L2; s++)
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) {
                    const u64 t = (u64)d[s * N + i] * (u64)key[s * N + j];
                    if (i + j < N) ref[i + j] += t; else ref[i + j - N] -= t;
                }
        static u32 res[NP][N];
        for (int pi = 0; pi < NP; pi++) {
            const u32 p = T.c.p[pi];
            static u32 D[2][32][2][32], K[2][32][2][32], ACC[32][2][32];
            std::vector<u32> tmp(N);
            for (int s = 0; s < L2; s++) {
                for (int i = 0; i < N; i++) tmp[i] = (u32)(d[s * N + i] + (int64_t)p);               // signed digit + p in [0, 2p)
                warp_fwd(pi, tmp.data(), D[s]);
                for (int i = 0; i < N; i++) tmp[i] = rns::residue_i64(key[s * N + i], p);
                warp_fwd(pi, tmp.data(), K[s]);
            }
            for (int l = 0; l < 32; l++)
                for (int h = 0; h < 2; h++)
                    for (int c = 0; c < 32; c++) {
                        const u32 k0 = rns::mulmod(K[0][l][h][c] % p, T.c.key_scale[pi], p), k1 = rns::mulmod(K[1][l][h][c] % p, T.c.key_scale[pi], p);
                        const u32 v = rns::mont_mul2(D[0][l][h][c], k0, D[1][l][h][c], k1, p, T.c.pinv_neg[pi]);   // both digit polynomials, one reduction
                        if (v >= 3ull * p) { fails++; printf("FAIL pointwise range\n"); }
                        ACC[l][h][c] = v;                                                                // < 2.75p: already inside the inverse's [0, 4p)
                    }
            warp_inv(pi, ACC, res[pi]);
        }
        for (int i = 0; i < N; i++) {
            const u32 r[NP] = {res[0][i], res[1][i], res[2][i], res[3][i]};
            const u64 got = crt4_lift(r, T.c);
            if (got != ref[i]) { fails++; printf("FAIL product trial=%d i=%d got=%llx ref=%llx\n", trial, i, (unsigned long long)got, (unsigned long long)ref[i]); break; }
        }
    }
    printf(fails ? "ntt2048_emu: %d FAILURES\n" : "ntt2048_emu: OK\n", fails);
    return fails ? 1 : 0;
}
