/*
 * mk_oracle.h -- CPU ORACLE for the 3gen multi-key TFHE bootstrapped-gate path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (torus-fhe_b200/) may
 * include, link or call this; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * PARITY UNPINNED: the reference (Animesh005/Torus-FHE, 3-gen-mk-tfhe/, Julia)
 * holds no golden vector, known-answer test or seeded fixture for the 3gen
 * path (3-gen-mk-tfhe/test/runtests.jl covers single-key and CCS gates only)
 * and Julia is not available in this image, so this restatement is pinned
 * only functionally (gate truth tables / integer circuits after decryption)
 * and by self-generated fixtures under tests/golden/.
 *
 * Every function cites the reference file:line it restates; paths are
 * relative to /root/reference/3-gen-mk-tfhe/src/.
 */
#ifndef MK_ORACLE_H
#define MK_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SchemeParameters_3gen, api.jl:50-67 (rlwe_mask_size is always 1 and
 * rlwe_is32 always false in the 3gen sets, mk_api.jl:32-322). */
typedef struct {
    int32_t n;        /* lwe_size */
    int32_t N;        /* rlwe_polynomial_degree (power of two) */
    int32_t k;        /* parties */
    int32_t l;        /* gsw_decomp_length */
    int32_t bgbit;    /* gsw_log2_base */
    int32_t t;        /* ks_decomp_length */
    int32_t basebit;  /* ks_log2_base */
    int32_t _pad;
    double sigma_lwe; /* lwe_noise_stddev */
    double sigma_gsw; /* gsw_noise_stddev */
    double sigma_ks;  /* ks_noise_stddev  */
} mko_params;

/* multiplication back-ends for the external product */
enum { MKO_EXACT_SCHOOLBOOK = 0, MKO_EXACT_NTT = 1, MKO_FFT = 2 };

/* gate ids (3gen_mk_gates.jl:8-74) -- same numbering as include/mktfhe_b200.h */
enum { MKO_GATE_NAND = 0, MKO_GATE_OR = 1, MKO_GATE_AND = 2, MKO_GATE_XOR = 3, MKO_GATE_AND3 = 4 };

typedef struct mko_keyset mko_keyset;

/* ---- scalar helpers (numeric-functions.jl) ---- */
int32_t mko_encode_message32(int64_t mu, int space);           /* :86-89  */
int64_t mko_encode_message64(int64_t mu, int space);           /* :92-95  */
int32_t mko_decode_message32(int32_t phase, int space);        /* :70-73  */
int32_t mko_dtot32(double d);                                  /* :101-103 */
int64_t mko_dtot64(double d);                                  /* :105-107 */
int32_t mko_t64tot32(int64_t d);                               /* :109-111 */

/* ---- polynomial / gadget helpers ---- */
int64_t mko_gadget_offset(int l, int bgbit);                                   /* tgsw.jl:24-30 */
void mko_decompose(const int64_t *poly, int N, int l, int bgbit, int64_t *digits /*[l][N]*/); /* tgsw.jl:112-138 */
void mko_mul_by_monomial(const int64_t *p, int N, int64_t shift, int64_t *out); /* DarkIntegers mul_by_monomial */
void mko_negacyclic_mul_schoolbook(const int64_t *a, const int64_t *b, int N, int64_t *out); /* exact mod 2^64 */
void mko_negacyclic_mul_ntt(const int64_t *small, const int64_t *big, int N, int64_t *out);  /* exact if |small|*N*2^32 < 2^62 */
void mko_negacyclic_mul_fft(const int64_t *a, const int64_t *b, int N, int64_t *out);        /* polynomials.jl:245-247 */

/* ---- keys ---- */
/* Generates the full 3gen key material as multikey_3gen.jl:15-30 does (own
 * seeded xoshiro256** RNG; Julia's MersenneTwister streams are not
 * reproducible here). */
mko_keyset *mko_keygen(const mko_params *p, uint64_t seed, int nthreads);
/* Key set with caller-supplied raw bootstrapping / key-switching keys and no
 * secrets (used by the benchmarks on synthetic random keys).
 * bsk: int64 [k][n][4][l][N]; ksk: int32 [k][N][t][B-1][n+1]. */
mko_keyset *mko_keyset_from_raw(const mko_params *p, const int64_t *bsk, const int32_t *ksk);
void mko_keyset_free(mko_keyset *ks);
const mko_params *mko_keyset_params(const mko_keyset *ks);
const int64_t *mko_bsk(const mko_keyset *ks);      /* [k][n][4 (part_1..part_4)][l][N] */
const int32_t *mko_ksk(const mko_keyset *ks);      /* [k][N][t][B-1][n+1] (a[0..n-1], b) */
const int32_t *mko_lwe_keys(const mko_keyset *ks); /* [k][n] in {0,1} */
const int64_t *mko_rlwe_keys(const mko_keyset *ks);/* [k][N] in {-1,0,1} */
size_t mko_bsk_len(const mko_keyset *ks);
size_t mko_ksk_len(const mko_keyset *ks);
/* prepare the Complex{Float64} transformed key (TransformedBootstrapKeyPart_3gen,
 * 3gen_mk_internals.jl:45-55); called lazily by the FFT back-end. */
void mko_prepare_fft_key(mko_keyset *ks);
void mko_prepare_ntt_key(mko_keyset *ks);

/* ---- encrypt / decrypt (mk_api.jl:519-536, 607-610; mk_internals.jl:85-91) ---- */
void mko_encrypt(const mko_keyset *ks, uint64_t seed, int count, const uint8_t *bits,
                 int32_t *a /*[count][k][n]*/, int32_t *b /*[count]*/);
void mko_phase(const mko_keyset *ks, int count, const int32_t *a, const int32_t *b, int32_t *phase);

/* ---- hot path ---- */
/* tgsw_extern_mul_3gen, tgsw_3gen.jl:102-113.  acc = [mask, body] each N int64;
 * elem = party*n + j. out may alias acc_in. */
void mko_extprod(mko_keyset *ks, int backend, int party, int j,
                 const int64_t *acc_in /*[2][N]*/, int64_t *acc_out /*[2][N]*/);
/* one blind-rotate step mk_mux_rotate_3gen, 3gen_mk_internals.jl:59-62 */
void mko_mux_rotate(mko_keyset *ks, int backend, int party, int j, int32_t bara, int64_t *acc /*[2][N]*/);
/* mk_bootstrap_wo_keyswitch_3gen (:99-109): returns the extracted LWE sample
 * (dim N, Torus32) and, if acc_out != NULL, the final accumulator.
 * If digit_log != NULL it receives a 64-bit FNV hash per iteration of the
 * digit stream (for the lock-step exact/FFT comparison). */
void mko_bootstrap_wo_keyswitch(mko_keyset *ks, int backend, int64_t mu, const int32_t *a /*[k][n]*/, int32_t b,
                                int32_t *ext_a /*[N]*/, int32_t *ext_b, int64_t *acc_out /*[2][N] or NULL*/,
                                uint64_t *digit_log /*[k*n] or NULL*/);
/* mk_keyswitch_3gen, mk_internals.jl:730-744 + keyswitch.jl:45-80 */
void mko_keyswitch(const mko_keyset *ks, const int32_t *ext_a /*[N]*/, int32_t ext_b,
                   int32_t *out_a /*[k][n]*/, int32_t *out_b);
/* mk_bootstrap_3gen (:112-116), batch over `count` samples with pthreads */
void mko_bootstrap_batch(mko_keyset *ks, int backend, int64_t mu, int count, const int32_t *a, const int32_t *b,
                         int32_t *out_a, int32_t *out_b, int nthreads);
/* linear prologue of a gate (3gen_mk_gates.jl:8-74): returns temp = mu0 + cx*x + cy*y (+ cz*z) */
void mko_gate_prologue(const mko_params *p, int gate, const int32_t *xa, int32_t xb, const int32_t *ya, int32_t yb,
                       const int32_t *za, int32_t zb, int32_t *ta, int32_t *tb);
/* mk_gate_*_3gen, batch */
void mko_gate_batch(mko_keyset *ks, int backend, int gate, int count,
                    const int32_t *xa, const int32_t *xb, const int32_t *ya, const int32_t *yb,
                    const int32_t *za, const int32_t *zb, int32_t *oa, int32_t *ob, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
