// tables2048.h -- host-side generation of the twiddle tables and Garner constants of ntt2048.cuh (groundwork for N = 2048).
#pragma once
#include <vector>
#include "ntt2048.cuh"
#include "tables.h"   // brev, find_psi-style helpers (powmod / invmod / shoup_of come from rns.cuh)

namespace rns2k {

inline u32 find_psi(u32 p) {   // primitive 2N-th root of unity mod p
    for (u32 g = 2;; g++) {
        const u32 psi = rns::powmod(g, (p - 1) / (2 * N), p);
        if (rns::powmod(psi, N, p) == p - 1) return psi;
    }
}

struct HostTables {
    Consts c;
    std::vector<uint2_> twB;   // [prime][dir][half][31][32], see twB_index
    u32 psi[NP];

    HostTables() : twB((size_t)NP * 2 * 2 * 31 * 32) {
        for (int i = 0; i < NP; i++) {
            const u32 p = PRIMES[i];
            psi[i] = find_psi(p);
            c.p[i] = p;
            u32 inv = p;
            for (int it = 0; it < 5; it++) inv *= 2 - p * inv;
            c.pinv_neg[i] = 0u - inv;
            c.key_scale[i] = rns::mulmod(rns::invmod(N % p, p), (u32)(((u64)1 << 32) % p), p);
            std::vector<u32> psi_br(N), psi_br_inv(N);
            for (u32 k = 0; k < (u32)N; k++) {
                psi_br[k] = rns::powmod(psi[i], rns::brev(k, LOGN), p);
                psi_br_inv[k] = rns::invmod(psi_br[k], p);
            }
            for (int e = 0; e < 63; e++) {            // stages with 1 .. 32 blocks: block b of stage k is psi_br[2^k + b] = entry 2^k - 1 + b
                c.twA[i][0][e] = {psi_br[e + 1], rns::shoup_of(psi_br[e + 1], p)};
                c.twA[i][1][e] = {psi_br_inv[e + 1], rns::shoup_of(psi_br_inv[e + 1], p)};
            }
            for (int half = 0; half < 2; half++)
                for (int k = 0; k < 5; k++)           // stages with 64 * 2^k blocks inside the run of block q = lane + 32 half
                    for (int b = 0; b < (1 << k); b++)
                        for (int lane = 0; lane < 32; lane++) {
                            const int e = (1 << k) - 1 + b, q = lane + 32 * half, idx = (1 << k) * (64 + q) + b;
                            twB[twB_index(i, 0, half, e, lane)] = {psi_br[idx], rns::shoup_of(psi_br[idx], p)};
                            twB[twB_index(i, 1, half, e, lane)] = {psi_br_inv[idx], rns::shoup_of(psi_br_inv[idx], p)};
                        }
        }
        for (int i = 0; i < NP; i++)
            for (int j = 0; j < NP; j++) {
                c.ginv[i][j] = i < j ? rns::invmod(PRIMES[i] % PRIMES[j], PRIMES[j]) : 0;
                c.ginv_s[i][j] = i < j ? rns::shoup_of(c.ginv[i][j], PRIMES[j]) : 0;
            }
        c.pp[0] = 1;
        for (int j = 1; j < NP; j++) c.pp[j] = c.pp[j - 1] * (u64)PRIMES[j - 1];   // wraps mod 2^64 from j = 3
        c.m_mod64 = c.pp[NP - 1] * (u64)PRIMES[NP - 1];
        for (int i = 0; i < NP; i++) {
            const u32 p = PRIMES[i];
            u64 cm = 1;                    // M / p_i mod 2^64
            u32 cmp = 1;                   // M / p_i mod p_i
            for (int j = 0; j < NP; j++)
                if (j != i) { cm *= (u64)PRIMES[j]; cmp = rns::mulmod(cmp, PRIMES[j] % p, p); }
            c.cm[i] = cm;
            c.yscale[i] = rns::invmod(cmp, p);
            c.key_scale_k[i] = rns::mulmod(c.key_scale[i], c.yscale[i], p);
            c.kf[i] = (u32)((((u64)1) << 59) / p);
        }
    }
};

// log2 of M / 4: the magnitude bound below which crt4_lift's sign decision is unambiguous
inline double log2_crt_bound() { return 111.997 - 2.0; }

}  // namespace rns2k
