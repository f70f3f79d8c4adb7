// tables.h -- host-side generation of the NTT twiddle tables and CRT constants (rns.cuh / ntt_rns.cuh).
#pragma once
#include <vector>
#include "ntt_rns.cuh"

namespace rns {

inline u32 brev(u32 x, int bits) {
    u32 r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

// smallest generator-derived primitive 2N-th root of unity mod p
inline u32 find_psi(u32 p) {
    for (u32 g = 2;; g++) {
        const u32 psi = powmod(g, (p - 1) / (2 * N), p);
        if (powmod(psi, N, p) == p - 1) return psi;   // psi^N = -1  <=>  order exactly 2N
    }
}

inline u32 invmod(u32 a, u32 p) { return powmod(a, p - 2, p); }

struct HostTables {
    Consts c;                       // goes to __constant__ memory
    std::vector<uint2_> twB;        // [prime][dir][31][32 lanes], goes to global memory (staged in shared memory by the kernels)
    u32 psi[NP];

    HostTables() : twB((size_t)NP * 2 * 31 * 32) {
        const u32 primes[NP] = {PRIME0, PRIME1, PRIME2};
        for (int i = 0; i < NP; i++) {
            const u32 p = primes[i];
            psi[i] = find_psi(p);
            c.p[i] = p;
            // -p^-1 mod 2^32 by Newton iteration
            u32 inv = p;
            for (int it = 0; it < 5; it++) inv *= 2 - p * inv;
            c.pinv_neg[i] = 0u - inv;
            // key scaling: N^-1 (inverse-NTT normalisation) * 2^32 (Montgomery factor of the pointwise products)
            c.key_scale[i] = mulmod(invmod(N % p, p), (u32)(((u64)1 << 32) % p), p);
            // psi_br[k] = psi^brev(k): twiddle of stage m, block i is psi_br[m + i] (merged negacyclic CT NTT)
            std::vector<u32> psi_br(N), psi_br_inv(N);
            for (u32 k = 0; k < (u32)N; k++) {
                psi_br[k] = powmod(psi[i], brev(k, LOGN), p);
                psi_br_inv[k] = invmod(psi_br[k], p);
            }
            for (int e = 0; e < 31; e++) {          // pass over the high index bits: index e + 1, uniform across lanes
                c.twA[i][0][e] = {psi_br[e + 1], shoup_of(psi_br[e + 1], p)};
                c.twA[i][1][e] = {psi_br_inv[e + 1], shoup_of(psi_br_inv[e + 1], p)};
            }
            for (int k = 0; k < 5; k++)             // pass over the low index bits: index 2^k (32 + lane) + b
                for (int b = 0; b < (1 << k); b++)
                    for (int lane = 0; lane < 32; lane++) {
                        const int e = (1 << k) - 1 + b, idx = (1 << k) * (32 + lane) + b;
                        twB[(((size_t)i * 2 + 0) * 31 + e) * 32 + lane] = {psi_br[idx], shoup_of(psi_br[idx], p)};
                        twB[(((size_t)i * 2 + 1) * 31 + e) * 32 + lane] = {psi_br_inv[idx], shoup_of(psi_br_inv[idx], p)};
                    }
        }
        Crt& t = c.crt;
        for (int i = 0; i < NP; i++) t.p[i] = primes[i];
        t.c01 = invmod(primes[0] % primes[1], primes[1]); t.c01s = shoup_of(t.c01, primes[1]);
        t.c02 = invmod(primes[0] % primes[2], primes[2]); t.c02s = shoup_of(t.c02, primes[2]);
        t.c12 = invmod(primes[1] % primes[2], primes[2]); t.c12s = shoup_of(t.c12, primes[2]);
        t.p01 = (u64)primes[0] * primes[1];
        t.m_mod64 = t.p01 * (u64)primes[2];          // wraps mod 2^64
        for (int i = 0; i < NP; i++) {
            const u32 pa = primes[(i + 1) % NP], pb = primes[(i + 2) % NP], pi = primes[i];
            t.C[i] = (u64)pa * pb;
            t.inv_p[i] = 1.0f / (float)pi;
            t.yscale[i] = invmod((u32)(t.C[i] % pi), pi);
#if MK_CRT_FLOAT
            c.key_scale[i] = mulmod(c.key_scale[i], t.yscale[i], pi);     // the lift's (M / p_i)^-1 rides in the stored key
#endif
        }
    }
};

// log2 of M / 4 (the magnitude bound below which crt_lift's sign decision is unambiguous), for parameter checks
inline double log2_crt_bound() { return 3 * 27.9996 - 2.0; }

}  // namespace rns
