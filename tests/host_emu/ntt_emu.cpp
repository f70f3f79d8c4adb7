// Host emulation of the warp-level 1024-point negacyclic RNS NTT used by the CUDA kernels: runs the very
// same per-thread passes (torus-fhe_b200/csrc/ntt_rns.cuh) for 32 emulated lanes and checks them against
// the O(N^2) definition, and the whole three-prime product (Montgomery pointwise products, folded key
// scaling, Garner CRT) against an exact schoolbook negacyclic product mod 2^64.
// Build: g++ -O2 -std=c++17 -I torus-fhe_b200/csrc tests/host_emu/ntt_emu.cpp -o ntt_emu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tables.h"

using namespace rns;
static HostTables T;

static u64 rnd_state = 0x1234567ull;
static u64 rnd() { rnd_state ^= rnd_state << 13; rnd_state ^= rnd_state >> 7; rnd_state ^= rnd_state << 17; return rnd_state; }

// coefficient order in (values in [0, 2p)) -> transformed out[lane][c] (position 32 lane + c, values in [0, 14p))
static void warp_fwd(int pi, const u32* a, u32 out[32][32]) {
    static u32 tile[TILE_WORDS];
    const u32 p = T.c.p[pi];
    u32 x[32];
    for (int lane = 0; lane < 32; lane++) {
        for (int r = 0; r < 32; r++) x[r] = a[32 * r + lane];
        fwd_passA(x, T.c.twA[pi][0], p);
        for (int r = 0; r < 32; r++) tile[r * TILE_STRIDE + lane] = x[r];
    }
    for (int lane = 0; lane < 32; lane++) {
        for (int c = 0; c < 32; c++) {
            if (tile[lane * TILE_STRIDE + c] >= 12 * p) { printf("FAIL passA range\n"); exit(1); }
            x[c] = reduce_to_4p(tile[lane * TILE_STRIDE + c], 4 * p);
        }
        fwd_passB(x, T.twB.data() + ((size_t)pi * 2 + 0) * 31 * 32 + lane, p);
        for (int c = 0; c < 32; c++) out[lane][c] = x[c];
    }
}
// transformed in[lane][c] (values in [0, 4p)) -> coefficient order, scaled by N
static void warp_inv(int pi, u32 in[32][32], u32* a) {
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

int main() {
    int fails = 0;
    for (int pi = 0; pi < NP; pi++) {
        const u32 p = T.c.p[pi];
        if (p >= (1u << 28) || (p - 1) % (2 * N) || powmod(T.psi[pi], N, p) != p - 1) { printf("FAIL prime %d\n", pi); return 1; }
        if ((u32)(p * (0u - T.c.pinv_neg[pi])) != 1u) { printf("FAIL pinv %d\n", pi); return 1; }
        // scalar primitives vs 64-bit arithmetic, including the lazy ranges
        for (int it = 0; it < 200000; it++) {
            const u32 w = rnd() % p, ws = shoup_of(w, p), y = (u32)rnd();
            const u32 r = shoup_mul(y, w, ws, p);
            if (r >= 2 * p || r % p != (u64)y * w % p) { fails++; printf("FAIL shoup\n"); break; }
            const u32 d = (u32)rnd(), k = rnd() % p;
            const u32 m = mont_mul(d, k, p, T.c.pinv_neg[pi]);
            if (m >= 2 * p || (u64)m * (((u64)1 << 32) % p) % p != (u64)d * k % p) { fails++; printf("FAIL mont\n"); break; }
            const u32 d2 = (u32)(rnd() % (14ull * p)), k2 = rnd() % p, d1 = d % (14 * p);
            const u32 m2 = mont_mul2(d1, k, d2, k2, p, T.c.pinv_neg[pi]);
            if (m2 >= 3 * p || (u64)m2 * (((u64)1 << 32) % p) % p != ((u64)d1 * k % p + (u64)d2 * k2 % p) % p) { fails++; printf("FAIL mont2\n"); break; }
            {   // eight products of the kernel's ranges accumulate in 64 bits without overflow and reduce to < 8p
                u64 acc = 0, ref = 0;
                for (int e = 0; e < 8; e++) {
                    const u32 dd = it == 0 ? 14 * p - 1 : (u32)(rnd() % (14ull * p)), kk = it == 0 ? p - 1 : rnd() % p;
                    acc += (u64)dd * kk;
                    ref = (ref + (u64)dd * kk % p) % p;
                }
                const u32 mm = (u32)acc * T.c.pinv_neg[pi];
                const u32 v = (u32)((acc + (u64)mm * p) >> 32);
                if (acc + (u64)mm * p < acc || v >= 8 * p || (u64)v * (((u64)1 << 32) % p) % p != ref) { fails++; printf("FAIL acc64\n"); break; }
            }
            u32 X = rnd() % (4 * p), Y = rnd() % (4 * p), X0 = X, Y0 = Y;
            ct_bfly<true>(X, Y, w, ws, p, 2 * p);
            const u64 wy = (u64)Y0 % p * w % p;
            if (X >= 4 * p || Y >= 4 * p || X % p != (X0 % p + wy) % p || Y % p != (X0 % p + p - wy) % p) { fails++; printf("FAIL ct\n"); break; }
            X = X0 + 8 * p; Y = (u32)rnd();                      // lazy form: X < 12p, any 32-bit Y
            const u32 Xl = X, Yl = Y;
            ct_bfly<false>(X, Y, w, ws, p, 2 * p);
            const u64 wyl = (u64)Yl % p * w % p;
            if (X >= 14 * p || Y >= 14 * p || X % p != (Xl % p + wyl) % p || Y % p != (Xl % p + p - wyl) % p) { fails++; printf("FAIL ct lazy\n"); break; }
            X = X0; Y = Y0;
            gs_bfly<true>(X, Y, w, ws, p, 4 * p, 4 * p);
            if (X >= 4 * p || Y >= 2 * p || X % p != ((u64)X0 + Y0) % p || Y % p != ((u64)X0 % p + p - Y0 % p) % p * w % p) { fails++; printf("FAIL gs\n"); break; }
        }
        // pass A with the first stage's products taken from a table (gadget digits) == plain pass A, modulo p
        {
            u32 xa[32], xb[32];
            const u32 w0 = T.c.twA[pi][0][0].x;
            for (int r = 0; r < 32; r++) {
                const int d = (int)(rnd() % 128) - 64;
                xa[r] = (u32)(d + 64) + (p - 64);
                xb[r] = r < 16 ? xa[r] : mulmod(d >= 0 ? (u32)d : p - (u32)(-d), w0, p);
            }
            fwd_passA(xa, T.c.twA[pi][0], p);
            fwd_passA_pre(xb, T.c.twA[pi][0], p);
            for (int r = 0; r < 32; r++)
                if (xb[r] >= 12 * p || xa[r] % p != xb[r] % p) { fails++; printf("FAIL passA_pre prime=%d r=%d\n", pi, r); break; }
        }
        // forward transform vs the definition: out[pos] = A(psi^(2 brev(pos) + 1))
        std::vector<u32> a(N), back(N);
        for (auto& v : a) v = rnd() % (2 * p);
        static u32 A[32][32];
        warp_fwd(pi, a.data(), A);
        std::vector<u32> psipow(2 * N);
        psipow[0] = 1;
        for (int i = 1; i < 2 * N; i++) psipow[i] = mulmod(psipow[i - 1], T.psi[pi], p);
        for (int lane = 0; lane < 32 && !fails; lane += 3)
            for (int c = 0; c < 32; c += 5) {
                const u32 e = 2 * brev(ntt_pos(lane, c), LOGN) + 1;
                u64 s = 0;
                for (int i = 0; i < N; i++) s = (s + (u64)(a[i] % p) * psipow[(u64)i * e % (2 * N)]) % p;
                if (A[lane][c] >= 14 * p || A[lane][c] % p != s) { fails++; printf("FAIL fwd prime=%d lane=%d c=%d\n", pi, lane, c); break; }
            }
        for (int l = 0; l < 32; l++) for (int c = 0; c < 32; c++) A[l][c] = reduce_to_4p(A[l][c], 4 * p);
        warp_inv(pi, A, back.data());
        for (int i = 0; i < N; i++)
            if (back[i] >= 4 * p || back[i] % p != (u64)(a[i] % p) * N % p) { fails++; printf("FAIL roundtrip prime=%d i=%d\n", pi, i); break; }
    }
    // exact negacyclic product of signed digits with a 64-bit key through the three primes
    for (int trial = 0; trial < 5; trial++) {
        const int L2 = 4;   // 2l polynomials accumulated before the inverse, as in the external product
        std::vector<int64_t> d(L2 * N), key(L2 * N);
        std::vector<u64> ref(N, 0);
        for (int i = 0; i < L2 * N; i++) { d[i] = (int64_t)(rnd() % 128) - 64; key[i] = (int64_t)rnd(); }
        if (trial == 1) for (int i = 0; i < L2 * N; i++) { d[i] = -64; key[i] = INT64_MIN; }          // largest magnitude: 2^81
        if (trial == 2) for (int i = 0; i < L2 * N; i++) { d[i] = 63; key[i] = INT64_MAX; }
        if (trial == 3) for (int i = 0; i < L2 * N; i++) { d[i] = (i & 1) ? 63 : -64; key[i] = (i & 2) ? INT64_MIN : INT64_MAX; }
        for (int s = 0; s < L2; s++)
            for (int i = 0; i < N; i++) {
                if (!d[s * N + i]) continue;
                for (int j = 0; j < N; j++) {
                    const u64 t = (u64)d[s * N + i] * (u64)key[s * N + j];
                    if (i + j < N) ref[i + j] += t; else ref[i + j - N] -= t;
                }
            }
        static u32 res[NP][1024];
        for (int pi = 0; pi < NP; pi++) {
            const u32 p = T.c.p[pi];
            static u32 D[32][32], K[32][32], ACC[32][32];
            for (int l = 0; l < 32; l++) for (int c = 0; c < 32; c++) ACC[l][c] = 0;
            std::vector<u32> tmp(N);
            for (int s = 0; s < L2; s++) {
                for (int i = 0; i < N; i++) tmp[i] = (u32)(d[s * N + i] + 64) + (p - 64);              // biased byte + (p - Bg/2)
                warp_fwd(pi, tmp.data(), D);
                for (int i = 0; i < N; i++) tmp[i] = residue_i64(key[s * N + i], p);
                warp_fwd(pi, tmp.data(), K);
                for (int l = 0; l < 32; l++)
                    for (int c = 0; c < 32; c++) {
                        const u32 kk = mulmod(K[l][c] % p, T.c.key_scale[pi], p);                      // stored key: NTT(K) N^-1 2^32
                        ACC[l][c] += mont_mul(D[l][c], kk, p, T.c.pinv_neg[pi]);
                    }
            }
            for (int l = 0; l < 32; l++)
                for (int c = 0; c < 32; c++) {
                    u32 v = ACC[l][c];
                    if (v >= 8 * p) { fails++; printf("FAIL acc range\n"); }
                    ACC[l][c] = umin32(v, v - 4 * p);
                }
            warp_inv(pi, ACC, res[pi]);
        }
        for (int i = 0; i < N; i++) {
            const u64 got = crt_lift(res[0][i], res[1][i], res[2][i], T.c.crt);
            if (got != ref[i]) { fails++; printf("FAIL product trial=%d i=%d got=%llx ref=%llx\n", trial, i, (unsigned long long)got, (unsigned long long)ref[i]); break; }
        }
    }
    printf(fails ? "ntt_emu: %d FAILURES\n" : "ntt_emu: OK\n", fails);
    return fails ? 1 : 0;
}
