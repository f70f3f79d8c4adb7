// Host emulation of the warp-level 1024-point negacyclic NTT used by the CUDA
// kernels: runs the very same per-thread passes (ntt1024.cuh) for 32 emulated
// lanes and checks them against the O(N^2) definition and against an exact
// schoolbook negacyclic product with the two-limb key split.
// Build: g++ -O2 -std=c++17 -I torus-fhe_b200/csrc tests/host_emu/ntt_emu.cpp -o ntt_emu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ntt1024.cuh"
#include "tables.h"

using namespace ntt;
static Tables T;

static u64 rnd_state = 0x1234567ull;
static u64 rnd() { rnd_state ^= rnd_state << 13; rnd_state ^= rnd_state >> 7; rnd_state ^= rnd_state << 17; return rnd_state; }

// coefficient layout in -> NTT-domain layout out, reg[lane][r]
static void warp_fwd(const u64* a, u64 out[32][32]) {
    static u64 tile[TILE_ELEMS];
    u64 x[32];
    for (int lane = 0; lane < 32; lane++) {
        for (int i1 = 0; i1 < 32; i1++) x[i1] = a[32 * i1 + lane];
        fwd_pass1(x, T.tw_fwd.data(), lane);
        for (int r = 0; r < 32; r++) tile[r * TILE_STRIDE + lane] = x[r];
    }
    for (int lane = 0; lane < 32; lane++) {
        for (int j = 0; j < 32; j++) x[j] = tile[lane * TILE_STRIDE + j];
        fwd_pass2(x);
        for (int r = 0; r < 32; r++) out[lane][r] = x[r];
    }
}
static void warp_inv(u64 in[32][32], u64* a) {
    static u64 tile[TILE_ELEMS];
    u64 x[32];
    for (int lane = 0; lane < 32; lane++) {
        for (int r = 0; r < 32; r++) x[r] = in[lane][r];
        inv_pass1(x, T.tw_inv.data(), lane);
        for (int j = 0; j < 32; j++) tile[lane * TILE_STRIDE + j] = x[j];
    }
    for (int lane = 0; lane < 32; lane++) {
        for (int r = 0; r < 32; r++) x[r] = tile[r * TILE_STRIDE + lane];
        inv_pass2(x);
        for (int i1 = 0; i1 < 32; i1++) a[32 * i1 + lane] = x[i1];
    }
}

int main() {
    int fails = 0;
    if (T.psi == 0 || gl::pow(T.psi, 32) != 8 || gl::pow(T.psi, 1024) != gl::P - 1) { printf("FAIL psi\n"); return 1; }
    // field ops vs __int128
    for (int it = 0; it < 200000; it++) {
        u64 a = rnd() % gl::P, b = rnd() % gl::P;
        if (it < 64) { a = gl::P - 1 - (it & 7); b = gl::P - 1 - (it >> 3); }
        unsigned __int128 m = (unsigned __int128)a * b;
        if (gl::mul(a, b) != (u64)(m % gl::P)) { fails++; printf("FAIL mul\n"); break; }
        if (gl::add(a, b) != (u64)(((unsigned __int128)a + b) % gl::P)) { fails++; printf("FAIL add\n"); break; }
        if (gl::sub(a, b) != (u64)(((unsigned __int128)a + gl::P - b) % gl::P)) { fails++; printf("FAIL sub\n"); break; }
        int s = it % 192;
        if (gl::mul_pow2(a, s) != gl::mul(a, gl::pow(2, s))) { fails++; printf("FAIL mul_pow2 s=%d\n", s); break; }
    }
    // size-32 transforms vs definition
    {
        u64 x[32], y[32], ref[32];
        for (int i = 0; i < 32; i++) x[i] = y[i] = rnd() % gl::P;
        for (int k = 0; k < 32; k++) { u64 s = 0; for (int i = 0; i < 32; i++) s = gl::add(s, gl::mul(x[i], gl::pow(64, (u64)(i * k) % 32))); ref[k] = s; }
        dif32(y);
        for (int r = 0; r < 32; r++) if (y[r] != ref[brev5(r)]) { fails++; printf("FAIL dif32 r=%d\n", r); break; }
        dit32_inv(y);
        for (int i = 0; i < 32; i++) if (y[i] != gl::mul(x[i], 32)) { fails++; printf("FAIL dit32_inv i=%d\n", i); break; }
        for (int i = 0; i < 32; i++) y[i] = x[i];
        twist32(y); untwist32(y);
        for (int i = 0; i < 32; i++) if (y[i] != x[i]) { fails++; printf("FAIL twist i=%d\n", i); break; }
    }
    // full 1024-point NTT vs definition A[k] = sum a_i psi^(i(2k+1)), layout check
    {
        std::vector<u64> a(N), back(N);
        for (auto& v : a) v = rnd() % gl::P;
        static u64 A[32][32];
        warp_fwd(a.data(), A);
        std::vector<u64> psipow(2048);
        psipow[0] = 1; for (int i = 1; i < 2048; i++) psipow[i] = gl::mul(psipow[i - 1], T.psi);
        for (int lane = 0; lane < 32 && !fails; lane += 5)
            for (int r = 0; r < 32; r += 3) {
                int k = brev5(lane) + 32 * brev5(r);
                u64 s = 0;
                for (int i = 0; i < N; i++) s = gl::add(s, gl::mul(a[i], psipow[(size_t)i * (2 * k + 1) % 2048]));
                if (s != A[lane][r]) { fails++; printf("FAIL ntt1024 lane=%d r=%d\n", lane, r); break; }
            }
        warp_inv(A, back.data());
        for (int i = 0; i < N; i++) if (back[i] != a[i]) { fails++; printf("FAIL roundtrip i=%d\n", i); break; }
    }
    // exact negacyclic product, signed 7-bit digits x 64-bit key, two 32-bit key limbs
    for (int trial = 0; trial < 4; trial++) {
        std::vector<int64_t> d(N); std::vector<u64> key(N), ref(N, 0), got(N);
        for (int i = 0; i < N; i++) { d[i] = (int64_t)(rnd() % 128) - 64; key[i] = rnd(); }
        if (trial == 1) for (int i = 0; i < N; i++) { d[i] = -64; key[i] = ~0ull; }
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
            u64 t = (u64)d[i] * key[j];
            if (i + j < N) ref[i + j] += t; else ref[i + j - N] -= t;
        }
        std::vector<u64> df(N), lo(N), hi(N), rlo(N), rhi(N);
        for (int i = 0; i < N; i++) { df[i] = gl::from_i64(d[i]); lo[i] = key[i] & gl::EPS; hi[i] = key[i] >> 32; }
        static u64 D[32][32], L[32][32], H[32][32];
        warp_fwd(df.data(), D); warp_fwd(lo.data(), L); warp_fwd(hi.data(), H);
        for (int l = 0; l < 32; l++) for (int r = 0; r < 32; r++) { L[l][r] = gl::mul(L[l][r], D[l][r]); H[l][r] = gl::mul(H[l][r], D[l][r]); }
        warp_inv(L, rlo.data()); warp_inv(H, rhi.data());
        for (int i = 0; i < N; i++) {
            got[i] = gl::lift(rlo[i]) + (gl::lift(rhi[i]) << 32);
            if (got[i] != ref[i]) { fails++; printf("FAIL product trial=%d i=%d\n", trial, i); break; }
        }
    }
    printf(fails ? "ntt_emu: %d FAILURES\n" : "ntt_emu: OK\n", fails);
    return fails ? 1 : 0;
}
