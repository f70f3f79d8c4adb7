/*
 * mk_oracle.c -- CPU ORACLE (test infrastructure, NOT product code) for the
 * 3gen multi-key TFHE bootstrapped-gate path of Animesh005/Torus-FHE.
 *
 * PARITY UNPINNED (see mk_oracle.h): no reference KAT exists for this path and
 * Julia cannot run here.  Citations `file:line` are relative to
 * /root/reference/3-gen-mk-tfhe/src/.
 *
 * Three multiplication back-ends for the external product:
 *   MKO_EXACT_SCHOOLBOOK  exact negacyclic product mod 2^64 (ground truth)
 *   MKO_EXACT_NTT         exact, Goldilocks radix-2 NTT, the key in 2 limbs of 32 bits
 *                         (N = 1024 sets) or 3 of 22 bits (N = 2048 sets, 26-bit digits),
 *                         see ntt_limb_plan (bulk tests; cross-validated against
 *                         schoolbook in tests/)
 *   MKO_FFT               Float64 folded negacyclic FFT exactly as
 *                         polynomials.jl:208-242 (the reference's arithmetic;
 *                         also the timed CPU baseline)
 */
#define _GNU_SOURCE
#include "mk_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define MKO_CLONES __attribute__((target_clones("default", "avx2")))
#else
#define MKO_CLONES
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------------ */
/* RNG: splitmix64-seeded xoshiro256**, Box-Muller gaussians.                */
/* (numeric-functions.jl:7-62 uses Julia's MersenneTwister/randn/StatsBase;   */
/*  those streams cannot be reproduced, only the distributions are restated.) */
/* ------------------------------------------------------------------------ */
typedef struct { u64 s[4]; int have_spare; double spare; } mko_rng;

static u64 splitmix64(u64 *x) {
    u64 z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_seed(mko_rng *r, u64 seed, u64 stream) {
    u64 x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
    r->have_spare = 0; r->spare = 0.0;
}
static inline u64 rotl64(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
static inline u64 rng_u64(mko_rng *r) {
    u64 *s = r->s;
    u64 result = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
static inline double rng_unit(mko_rng *r) { return (double)(rng_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
static double rng_normal(mko_rng *r) {
    if (r->have_spare) { r->have_spare = 0; return r->spare; }
    double u1, u2;
    do { u1 = rng_unit(r); } while (u1 <= 0.0);
    u2 = rng_unit(r);
    double m = sqrt(-2.0 * log(u1)), a = 6.283185307179586476925286766559 * u2;
    r->spare = m * sin(a); r->have_spare = 1;
    return m * cos(a);
}
/* rand_negative_binary64, numeric-functions.jl:26-28: P(-1)=P(+1)=0.113546097609674 */
static inline int64_t rng_ternary(mko_rng *r) {
    const double w = 0.113546097609674;
    double u = rng_unit(r);
    return u < w ? -1 : (u < 1.0 - w ? 0 : 1);
}

/* ------------------------------------------------------------------------ */
/* scalar helpers, numeric-functions.jl                                       */
/* ------------------------------------------------------------------------ */
static inline int ilog2(int x) { return __builtin_ctz((unsigned)x); }

int32_t mko_encode_message32(int64_t mu, int space) {   /* :86-89 */
    return (int32_t)((uint32_t)(int32_t)mu << (32 - ilog2(space)));
}
int64_t mko_encode_message64(int64_t mu, int space) {   /* :92-95 */
    return (int64_t)((u64)mu << (64 - ilog2(space)));
}
int32_t mko_decode_message32(int32_t phase, int space) { /* :70-73: wrap add, arithmetic shift */
    int lg = ilog2(space);
    int32_t s = (int32_t)((uint32_t)phase + (1u << (32 - lg - 1)));
    return s >> (32 - lg);
}
int32_t mko_dtot32(double d) { return (int32_t)trunc(d * 4294967296.0); }          /* :101-103 */
int64_t mko_dtot64(double d) { return (int64_t)trunc(d * 18446744073709551616.0); } /* :105-107 */
/* :109-111  t64tot32(d) = trunc(Int32, d / 2^32): Int64 -> Float64 (round to nearest
 * even), exact division, truncation toward zero.  Julia throws InexactError when
 * the quotient is 2^31 (d >= 2^63-2^9, probability ~2^-54); we saturate there. */
int32_t mko_t64tot32(int64_t d) {
    double x = (double)d / 4294967296.0;
    x = trunc(x);
    if (x >= 2147483647.0) return INT32_MAX;
    if (x <= -2147483648.0) return INT32_MIN;
    return (int32_t)x;
}

/* ------------------------------------------------------------------------ */
/* gadget / decomposition, tgsw.jl                                            */
/* ------------------------------------------------------------------------ */
/* TGswParams ctor, tgsw.jl:24-30: offset = wrap64( sum_q 2^(64-q*bg) * 2^(bg-1) ) */
int64_t mko_gadget_offset(int l, int bgbit) {
    u64 off = 0;
    for (int q = 1; q <= l; q++) off += ((u64)1 << (64 - q * bgbit)) << (bgbit - 1);
    return (int64_t)off;
}
/* decompose, tgsw.jl:112-138 */
void mko_decompose(const int64_t *poly, int N, int l, int bgbit, int64_t *digits) {
    const int64_t mask = ((int64_t)1 << bgbit) - 1, half = (int64_t)1 << (bgbit - 1);
    const u64 off = (u64)mko_gadget_offset(l, bgbit);
    for (int q = 1; q <= l; q++) {
        int sh = 64 - q * bgbit;
        for (int i = 0; i < N; i++) {
            int64_t v = (int64_t)((u64)poly[i] + off);
            digits[(size_t)(q - 1) * N + i] = ((v >> sh) & mask) - half;
        }
    }
}
/* DarkIntegers mul_by_monomial on a negacyclic polynomial: p(X)*X^shift mod X^N+1 */
void mko_mul_by_monomial(const int64_t *p, int N, int64_t shift, int64_t *out) {
    int64_t s = shift % (2 * N);
    if (s < 0) s += 2 * N;
    for (int i = 0; i < N; i++) {
        int64_t j = i + s;
        int neg = 0;
        if (j >= 2 * N) j -= 2 * N;
        if (j >= N) { j -= N; neg = 1; }
        out[j] = neg ? (int64_t)(0 - (u64)p[i]) : p[i];
    }
}

/* exact negacyclic product mod 2^64; `a` is scanned for zeros (digits / ternary) */
MKO_CLONES
static void negacyclic_mac_schoolbook(const int64_t *a, const int64_t *b, int N, u64 *acc) {
    for (int i = 0; i < N; i++) {
        u64 d = (u64)a[i];
        if (!d) continue;
        u64 *hi = acc + i;
        const u64 *bb = (const u64 *)b;
        int m = N - i;
        for (int j = 0; j < m; j++) hi[j] += d * bb[j];
        u64 *lo = acc - m;
        for (int j = m; j < N; j++) lo[j] -= d * bb[j];
    }
}
void mko_negacyclic_mul_schoolbook(const int64_t *a, const int64_t *b, int N, int64_t *out) {
    u64 *acc = (u64 *)calloc((size_t)N, sizeof(u64));
    negacyclic_mac_schoolbook(a, b, N, acc);
    memcpy(out, acc, (size_t)N * sizeof(u64));
    free(acc);
}

/* ------------------------------------------------------------------------ */
/* Goldilocks field p = 2^64 - 2^32 + 1, plain radix-2 NTT                     */
/* ------------------------------------------------------------------------ */
#define GL_P 0xFFFFFFFF00000001ull
#define GL_EPS 0xFFFFFFFFull
static inline u64 gl_add(u64 a, u64 b) { u64 s = a + b; if (s < a || s >= GL_P) s -= GL_P; return s; }
static inline u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : a + (GL_P - b); }
static inline u64 gl_reduce128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 h0 = hi & GL_EPS, h1 = hi >> 32;
    u64 t = lo - h1; if (lo < h1) t -= GL_EPS;          /* 2^96 = -1 */
    u64 m = (h0 << 32) - h0;                             /* 2^64 = 2^32-1 */
    u64 r = t + m; if (r < t) r += GL_EPS;
    if (r >= GL_P) r -= GL_P;
    return r;
}
static inline u64 gl_mul(u64 a, u64 b) { return gl_reduce128((u128)a * b); }
static u64 gl_pow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; } return r; }
static inline u64 gl_from_i64(int64_t v) { return v >= 0 ? (u64)v : GL_P - (u64)(-v); } /* |v| < p */
static inline int64_t gl_lift(u64 v) { return v > GL_P / 2 ? (int64_t)(v - GL_P) : (int64_t)v; } /* wraps mod 2^64 */

typedef struct {
    int N, logN;
    u64 *psi_pow;      /* psi^i          */
    u64 *psi_inv_pow;  /* psi^-i / N     */
    u64 *w_pow;        /* omega^i, i<N/2 */
    u64 *w_inv_pow;
    /* FFT tables (polynomials.jl:81-160) */
    double *twist;     /* [N/2][2]: exp(-i*pi*j/N) */
    double *fft_w;     /* [N/4][2]: exp(-2*pi*i*j/(N/2)) */
} mko_tables;

static mko_tables *tables_new(int N) {
    mko_tables *t = (mko_tables *)calloc(1, sizeof(*t));
    t->N = N; t->logN = ilog2(N);
    t->psi_pow = malloc(sizeof(u64) * N); t->psi_inv_pow = malloc(sizeof(u64) * N);
    t->w_pow = malloc(sizeof(u64) * (N / 2)); t->w_inv_pow = malloc(sizeof(u64) * (N / 2));
    u64 psi = gl_pow(7, (GL_P - 1) / (u64)(2 * N));   /* 7 generates F_p^* */
    u64 psi_inv = gl_pow(psi, GL_P - 2), n_inv = gl_pow((u64)N, GL_P - 2);
    u64 w = gl_mul(psi, psi), w_inv = gl_mul(psi_inv, psi_inv);
    u64 a = 1, b = n_inv;
    for (int i = 0; i < N; i++) { t->psi_pow[i] = a; t->psi_inv_pow[i] = b; a = gl_mul(a, psi); b = gl_mul(b, psi_inv); }
    a = 1; b = 1;
    for (int i = 0; i < N / 2; i++) { t->w_pow[i] = a; t->w_inv_pow[i] = b; a = gl_mul(a, w); b = gl_mul(b, w_inv); }
    int M = N / 2;
    t->twist = malloc(sizeof(double) * 2 * M);
    t->fft_w = malloc(sizeof(double) * 2 * (M / 2 > 0 ? M / 2 : 1));
    const double PI = 3.14159265358979323846264338327950288;
    for (int j = 0; j < M; j++) { t->twist[2 * j] = cos(-PI * j / N); t->twist[2 * j + 1] = sin(-PI * j / N); }
    for (int j = 0; j < M / 2; j++) { t->fft_w[2 * j] = cos(-2.0 * PI * j / M); t->fft_w[2 * j + 1] = sin(-2.0 * PI * j / M); }
    return t;
}
static void tables_free(mko_tables *t) {
    if (!t) return;
    free(t->psi_pow); free(t->psi_inv_pow); free(t->w_pow); free(t->w_inv_pow); free(t->twist); free(t->fft_w); free(t);
}
static void bitrev_u64(u64 *x, int N, int logN) {
    for (int i = 0; i < N; i++) {
        int j = 0;
        for (int b = 0; b < logN; b++) j |= ((i >> b) & 1) << (logN - 1 - b);
        if (j > i) { u64 t = x[i]; x[i] = x[j]; x[j] = t; }
    }
}
static void gl_ntt_cyclic(u64 *x, int N, int logN, const u64 *wp) {
    bitrev_u64(x, N, logN);
    for (int len = 1; len < N; len <<= 1) {
        int step = N / (2 * len);
        for (int i = 0; i < N; i += 2 * len)
            for (int j = 0; j < len; j++) {
                u64 u = x[i + j], v = gl_mul(x[i + j + len], wp[j * step]);
                x[i + j] = gl_add(u, v); x[i + j + len] = gl_sub(u, v);
            }
    }
}
static void gl_negacyclic_fwd(const mko_tables *t, const u64 *in, u64 *out) {
    for (int i = 0; i < t->N; i++) out[i] = gl_mul(in[i], t->psi_pow[i]);
    gl_ntt_cyclic(out, t->N, t->logN, t->w_pow);
}
static void gl_negacyclic_inv(const mko_tables *t, u64 *x) {
    gl_ntt_cyclic(x, t->N, t->logN, t->w_inv_pow);
    for (int i = 0; i < t->N; i++) x[i] = gl_mul(x[i], t->psi_inv_pow[i]);
}
/* exact small*big product: big split into 32-bit limbs so each limb product is < 2^62 */
void mko_negacyclic_mul_ntt(const int64_t *small, const int64_t *big, int N, int64_t *out) {
    mko_tables *t = tables_new(N);
    u64 *s = malloc(sizeof(u64) * N), *lo = malloc(sizeof(u64) * N), *hi = malloc(sizeof(u64) * N), *tmp = malloc(sizeof(u64) * N);
    for (int i = 0; i < N; i++) { tmp[i] = gl_from_i64(small[i]); }
    gl_negacyclic_fwd(t, tmp, s);
    for (int i = 0; i < N; i++) tmp[i] = (u64)big[i] & GL_EPS;
    gl_negacyclic_fwd(t, tmp, lo);
    for (int i = 0; i < N; i++) tmp[i] = (u64)big[i] >> 32;
    gl_negacyclic_fwd(t, tmp, hi);
    for (int i = 0; i < N; i++) { lo[i] = gl_mul(lo[i], s[i]); hi[i] = gl_mul(hi[i], s[i]); }
    gl_negacyclic_inv(t, lo); gl_negacyclic_inv(t, hi);
    for (int i = 0; i < N; i++) out[i] = (int64_t)((u64)gl_lift(lo[i]) + ((u64)gl_lift(hi[i]) << 32));
    free(s); free(lo); free(hi); free(tmp); tables_free(t);
}

/* ------------------------------------------------------------------------ */
/* Float64 folded negacyclic FFT, polynomials.jl:81-247                        */
/* ------------------------------------------------------------------------ */
/* in-place complex radix-2 FFT of size M, sign = -1 forward / +1 inverse (unscaled) */
static void fft_complex(double *x, int M, const double *w, int sign) {
    int logM = ilog2(M);
    for (int i = 0; i < M; i++) {
        int j = 0;
        for (int b = 0; b < logM; b++) j |= ((i >> b) & 1) << (logM - 1 - b);
        if (j > i) { double tr = x[2 * i], ti = x[2 * i + 1]; x[2 * i] = x[2 * j]; x[2 * i + 1] = x[2 * j + 1]; x[2 * j] = tr; x[2 * j + 1] = ti; }
    }
    for (int len = 1; len < M; len <<= 1) {
        int step = M / (2 * len);
        for (int i = 0; i < M; i += 2 * len)
            for (int j = 0; j < len; j++) {
                double wr = w[2 * j * step], wi = sign < 0 ? w[2 * j * step + 1] : -w[2 * j * step + 1];
                double *a = x + 2 * (i + j), *b = x + 2 * (i + j + len);
                double vr = b[0] * wr - b[1] * wi, vi = b[0] * wi + b[1] * wr;
                b[0] = a[0] - vr; b[1] = a[1] - vi; a[0] += vr; a[1] += vi;
            }
    }
}
/* forward_transform, polynomials.jl:208-214: buf = (c[lo] - i*c[hi]) .* exp(-i*pi*j/N); fft(buf) */
static void fft_forward(const mko_tables *t, const int64_t *c, double *out /*[N/2][2]*/) {
    int M = t->N / 2;
    for (int j = 0; j < M; j++) {
        double re = (double)c[j], im = -(double)c[j + M];
        double wr = t->twist[2 * j], wi = t->twist[2 * j + 1];
        out[2 * j] = re * wr - im * wi; out[2 * j + 1] = re * wi + im * wr;
    }
    fft_complex(out, M, t->fft_w, -1);
}
/* to_int64(x::Float64) = wrap(round(Int128, x)), polynomials.jl:217-220 (ties to even) */
static inline int64_t fft_to_int64(double x) { return (int64_t)(u64)(u128)(__int128)rint(x); }
/* inverse_transform, polynomials.jl:224-242: ifft; conj .* coeffs; real -> low half, imag -> high half */
static void fft_inverse(const mko_tables *t, double *x /*[N/2][2], destroyed*/, int64_t *out) {
    int M = t->N / 2;
    fft_complex(x, M, t->fft_w, +1);
    double sc = 1.0 / M;
    for (int j = 0; j < M; j++) {
        double re = x[2 * j] * sc, im = -x[2 * j + 1] * sc;           /* conj */
        double wr = t->twist[2 * j], wi = t->twist[2 * j + 1];
        out[j] = fft_to_int64(re * wr - im * wi);
        out[j + M] = fft_to_int64(re * wi + im * wr);
    }
}
void mko_negacyclic_mul_fft(const int64_t *a, const int64_t *b, int N, int64_t *out) { /* transformed_mul :245-247 */
    mko_tables *t = tables_new(N);
    double *fa = malloc(sizeof(double) * N), *fb = malloc(sizeof(double) * N);
    fft_forward(t, a, fa); fft_forward(t, b, fb);
    for (int j = 0; j < N / 2; j++) {
        double r = fa[2 * j] * fb[2 * j] - fa[2 * j + 1] * fb[2 * j + 1], i = fa[2 * j] * fb[2 * j + 1] + fa[2 * j + 1] * fb[2 * j];
        fa[2 * j] = r; fa[2 * j + 1] = i;
    }
    fft_inverse(t, fa, out);
    free(fa); free(fb); tables_free(t);
}

/* ------------------------------------------------------------------------ */
/* key set                                                                    */
/* ------------------------------------------------------------------------ */
struct mko_keyset {
    mko_params p;
    int32_t *lwe_keys;   /* [k][n]            SecretKey_3gen, api.jl:196-204 / lwe.jl:11 */
    int64_t *rlwe_keys;  /* [k][N]            RLweKey(.., true), rlwe.jl:13-31 */
    int64_t *bsk;        /* [k][n][4][l][N]   BootstrapKeyPart_3gen, 3gen_mk_internals.jl:10-43 */
    int32_t *ksk;        /* [k][N][t][B-1][n+1] KeyswitchKey, keyswitch.jl:7-42 */
    double *bsk_fft;     /* [k][n][4][l][N/2][2] TransformedBootstrapKeyPart_3gen :45-55 */
    u64 *bsk_ntt;        /* [k][n][4][l][ntt_limbs][N]: the key in limbs of ntt_limb_bits bits, each transformed */
    int ntt_limbs, ntt_limb_bits;
    mko_tables *tab;
    pthread_mutex_t lock;
};

static size_t bsk_len(const mko_params *p) { return (size_t)p->k * p->n * 4 * p->l * p->N; }
static size_t ksk_len(const mko_params *p) { return (size_t)p->k * p->N * p->t * ((1 << p->basebit) - 1) * (p->n + 1); }
size_t mko_bsk_len(const mko_keyset *ks) { return bsk_len(&ks->p); }
size_t mko_ksk_len(const mko_keyset *ks) { return ksk_len(&ks->p); }
const mko_params *mko_keyset_params(const mko_keyset *ks) { return &ks->p; }
const int64_t *mko_bsk(const mko_keyset *ks) { return ks->bsk; }
const int32_t *mko_ksk(const mko_keyset *ks) { return ks->ksk; }
const int32_t *mko_lwe_keys(const mko_keyset *ks) { return ks->lwe_keys; }
const int64_t *mko_rlwe_keys(const mko_keyset *ks) { return ks->rlwe_keys; }

static mko_keyset *keyset_alloc(const mko_params *p) {
    mko_keyset *ks = (mko_keyset *)calloc(1, sizeof(*ks));
    ks->p = *p;
    ks->bsk = (int64_t *)malloc(bsk_len(p) * sizeof(int64_t));
    ks->ksk = (int32_t *)malloc(ksk_len(p) * sizeof(int32_t));
    ks->tab = tables_new(p->N);
    pthread_mutex_init(&ks->lock, NULL);
    return ks;
}
void mko_keyset_free(mko_keyset *ks) {
    if (!ks) return;
    free(ks->lwe_keys); free(ks->rlwe_keys); free(ks->bsk); free(ks->ksk); free(ks->bsk_fft); free(ks->bsk_ntt);
    tables_free(ks->tab); pthread_mutex_destroy(&ks->lock); free(ks);
}
mko_keyset *mko_keyset_from_raw(const mko_params *p, const int64_t *bsk, const int32_t *ksk) {
    mko_keyset *ks = keyset_alloc(p);
    memcpy(ks->bsk, bsk, bsk_len(p) * sizeof(int64_t));
    memcpy(ks->ksk, ksk, ksk_len(p) * sizeof(int32_t));
    return ks;
}

typedef struct { int tid, nthreads; void (*fn)(void *, int, int); void *arg; int count; } par_job;
static void *par_trampoline(void *v) {
    par_job *j = (par_job *)v;
    for (int i = j->tid; i < j->count; i += j->nthreads) j->fn(j->arg, i, j->tid);
    return NULL;
}
static void parallel_for(int count, int nthreads, void (*fn)(void *, int, int), void *arg) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > count) nthreads = count > 0 ? count : 1;
    if (nthreads == 1) { for (int i = 0; i < count; i++) fn(arg, i, 0); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    par_job *jobs = malloc(sizeof(par_job) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (par_job){t, nthreads, fn, arg, count};
        pthread_create(&th[t], NULL, par_trampoline, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

typedef struct {
    mko_keyset *ks; u64 seed;
    const int64_t *crp;    /* [N]   CRP_3gen(a_same=true), mk_internals.jl:183-195 */
    const int64_t *common; /* [l][N] CommonPubKey_3gen, mk_internals.jl:325-345 */
} keygen_ctx;

/* tgsw_encrypt_3gen (wo_FFT=1, negative_random=true), tgsw_3gen.jl:41-95, for
 * LWE key bit s_p[j]; element index e = p*n + j. */
static void keygen_bsk_elem(void *v, int e, int tid) {
    (void)tid;
    keygen_ctx *c = (keygen_ctx *)v;
    const mko_params *P = &c->ks->p;
    int N = P->N, l = P->l;
    mko_rng r; rng_seed(&r, c->seed, 0x1000000ull + (u64)e);
    int64_t msg = c->ks->lwe_keys[e];
    int64_t *out = c->ks->bsk + (size_t)e * 4 * l * N;
    int64_t *r1 = malloc(sizeof(int64_t) * N * l), *r2 = malloc(sizeof(int64_t) * N * l);
    for (int i = 0; i < N * l; i++) r1[i] = rng_ternary(&r);       /* :61 */
    for (int i = 0; i < N * l; i++) r2[i] = rng_ternary(&r);       /* :62 */
    /* error_part_1..4, :64-67 */
    for (int part = 0; part < 4; part++)
        for (int i = 0; i < l * N; i++) out[(size_t)part * l * N + i] = mko_dtot64(rng_normal(&r) * P->sigma_gsw);
    for (int q = 0; q < l; q++) {
        u64 g = (u64)1 << (64 - (q + 1) * P->bgbit);               /* gadget_values, tgsw.jl:26 */
        u64 *p1 = (u64 *)out + ((size_t)0 * l + q) * N, *p2 = (u64 *)out + ((size_t)1 * l + q) * N;
        u64 *p3 = (u64 *)out + ((size_t)2 * l + q) * N, *p4 = (u64 *)out + ((size_t)3 * l + q) * N;
        negacyclic_mac_schoolbook(r1 + q * N, c->common + (size_t)q * N, N, p1);  /* part_1 = r1*B + m*g + e1  :88 */
        negacyclic_mac_schoolbook(r2 + q * N, c->common + (size_t)q * N, N, p2);  /* part_2 = r2*B + e2        :89 */
        negacyclic_mac_schoolbook(r2 + q * N, c->crp, N, p3);                     /* part_3 = r2*a + m*g + e3  :90 */
        negacyclic_mac_schoolbook(r1 + q * N, c->crp, N, p4);                     /* part_4 = r1*a + e4        :91 */
        p1[0] += (u64)msg * g;  /* Polynomial .+ scalar adds to the constant coefficient (DarkIntegers) */
        p3[0] += (u64)msg * g;
    }
    free(r1); free(r2);
}

/* KeyswitchKey ctor, keyswitch.jl:14-41, party p = idx */
static void keygen_ksk_party(void *v, int p, int tid) {
    (void)tid;
    keygen_ctx *c = (keygen_ctx *)v;
    const mko_params *P = &c->ks->p;
    int N = P->N, n = P->n, t = P->t, B1 = (1 << P->basebit) - 1;
    mko_rng r; rng_seed(&r, c->seed, 0x2000000ull + (u64)p);
    size_t cnt = (size_t)N * t * B1;
    double *noise = malloc(sizeof(double) * cnt), sum = 0.0;
    for (size_t i = 0; i < cnt; i++) { noise[i] = rng_normal(&r) * P->sigma_ks; sum += noise[i]; }  /* :27-29 */
    double mean = sum / (double)cnt;
    const int32_t *s = c->ks->lwe_keys + (size_t)p * n;     /* out_key */
    const int64_t *z = c->ks->rlwe_keys + (size_t)p * N;    /* in_key = extract_lwe_key(rlwe_key), rlwe.jl:34-40 */
    int32_t *rows = c->ks->ksk + (size_t)p * cnt * (n + 1);
    for (int i = 0; i < N; i++)
        for (int j = 1; j <= t; j++)
            for (int h = 1; h <= B1; h++) {
                size_t idx = ((size_t)i * t + (j - 1)) * B1 + (h - 1);
                int32_t *row = rows + idx * (n + 1);
                /* message(i,j,h) = (in_key[i]*h) << (32 - j*log2_base)  :35 */
                uint32_t msg = (uint32_t)((int32_t)z[i] * h) << (32 - j * P->basebit);
                uint32_t dot = 0;
                for (int c2 = 0; c2 < n; c2++) { uint32_t a = (uint32_t)rng_u64(&r); row[c2] = (int32_t)a; dot += a * (uint32_t)s[c2]; }
                /* lwe_encrypt with given noise, lwe.jl:47-53 */
                row[n] = (int32_t)(msg + (uint32_t)mko_dtot32(noise[idx] - mean) + dot);
            }
    free(noise);
}

mko_keyset *mko_keygen(const mko_params *P, uint64_t seed, int nthreads) {
    mko_keyset *ks = keyset_alloc(P);
    int N = P->N, n = P->n, k = P->k, l = P->l;
    ks->lwe_keys = malloc(sizeof(int32_t) * k * n);
    ks->rlwe_keys = malloc(sizeof(int64_t) * k * N);
    mko_rng r; rng_seed(&r, seed, 1);
    for (int i = 0; i < k * n; i++) ks->lwe_keys[i] = (int32_t)(rng_u64(&r) & 1);   /* rand_uniform_bool, lwe.jl:11 */
    for (int i = 0; i < k * N; i++) ks->rlwe_keys[i] = rng_ternary(&r);            /* rlwe.jl:24-27 */
    int64_t *crp = malloc(sizeof(int64_t) * N);
    for (int i = 0; i < N; i++) crp[i] = (int64_t)rng_u64(&r);                    /* CRP_3gen a_same, mk_internals.jl:189-191 */
    /* PublicKey(rng, rlwe_key, alpha, crp, tgsw_params, wo_FFT=1): b_p[q] = z_p*a + e, mk_internals.jl:266-298;
     * CommonPubKey_3gen: b[q] = sum_p b_p[q], :331-343 */
    u64 *common = calloc((size_t)l * N, sizeof(u64));
    for (int p = 0; p < k; p++)
        for (int q = 0; q < l; q++) {
            negacyclic_mac_schoolbook(ks->rlwe_keys + (size_t)p * N, crp, N, common + (size_t)q * N);
            for (int i = 0; i < N; i++) common[(size_t)q * N + i] += (u64)mko_dtot64(rng_normal(&r) * P->sigma_gsw);
        }
    memset(ks->bsk, 0, bsk_len(P) * sizeof(int64_t));
    keygen_ctx c = {ks, seed, crp, (const int64_t *)common};
    parallel_for(k * n, nthreads, keygen_bsk_elem, &c);
    parallel_for(k, nthreads, keygen_ksk_party, &c);
    free(crp); free(common);
    return ks;
}

static void prep_fft_elem(void *v, int e, int tid) {
    (void)tid;
    mko_keyset *ks = (mko_keyset *)v;
    int N = ks->p.N, l = ks->p.l;
    for (int q = 0; q < 4 * l; q++)
        fft_forward(ks->tab, ks->bsk + ((size_t)e * 4 * l + q) * N, ks->bsk_fft + ((size_t)e * 4 * l + q) * N);
}
void mko_prepare_fft_key(mko_keyset *ks) {
    pthread_mutex_lock(&ks->lock);
    if (!ks->bsk_fft) {
        double *buf = malloc(bsk_len(&ks->p) * sizeof(double));
        ks->bsk_fft = buf;
        parallel_for(ks->p.k * ks->p.n, 8, prep_fft_elem, ks);
    }
    pthread_mutex_unlock(&ks->lock);
}
/* Limb split of the exact Goldilocks back-end: the sum over the 2l digit polynomials of digit * limb products must stay below
 * 2^62 (half the field, so that the signed lift is unambiguous): 2l * N * 2^(bgbit-1) * 2^limb_bits < 2^62.  Two 32-bit limbs cover
 * the N = 1024 sets (7-bit digits); the 24..26-bit digits of the N = 2048 sets need three limbs of 22 bits. */
static void ntt_limb_plan(const mko_params *p, int *limbs, int *bits) {
    int room = 62 - (ilog2(2 * p->l * p->N) + (p->bgbit - 1));   /* ilog2 of a non power of two rounds down: add one below */
    if ((2 * p->l * p->N) & (2 * p->l * p->N - 1)) room -= 1;
    if (room >= 32) { *limbs = 2; *bits = 32; }
    else if (room >= 22) { *limbs = 3; *bits = 22; }
    else { *limbs = 4; *bits = 16; }
}
static void prep_ntt_elem(void *v, int e, int tid) {
    (void)tid;
    mko_keyset *ks = (mko_keyset *)v;
    int N = ks->p.N, l = ks->p.l, NL = ks->ntt_limbs, LB = ks->ntt_limb_bits;
    const u64 lmask = LB == 64 ? ~0ull : (((u64)1 << LB) - 1);
    u64 *tmp = malloc(sizeof(u64) * N);
    for (int q = 0; q < 4 * l; q++) {
        const int64_t *src = ks->bsk + ((size_t)e * 4 * l + q) * N;
        u64 *dst = ks->bsk_ntt + ((size_t)e * 4 * l + q) * NL * N;
        for (int limb = 0; limb < NL; limb++) {
            /* the top limb keeps every remaining bit (it is the only one that may exceed limb_bits when NL * LB < 64: never here) */
            for (int i = 0; i < N; i++) tmp[i] = limb == NL - 1 ? (u64)src[i] >> (limb * LB) : ((u64)src[i] >> (limb * LB)) & lmask;
            gl_negacyclic_fwd(ks->tab, tmp, dst + (size_t)limb * N);
        }
    }
    free(tmp);
}
void mko_prepare_ntt_key(mko_keyset *ks) {
    pthread_mutex_lock(&ks->lock);
    if (!ks->bsk_ntt) {
        ntt_limb_plan(&ks->p, &ks->ntt_limbs, &ks->ntt_limb_bits);
        ks->bsk_ntt = malloc(bsk_len(&ks->p) * ks->ntt_limbs * sizeof(u64));
        parallel_for(ks->p.k * ks->p.n, 8, prep_ntt_elem, ks);
    }
    pthread_mutex_unlock(&ks->lock);
}

/* ------------------------------------------------------------------------ */
/* encrypt / phase                                                            */
/* ------------------------------------------------------------------------ */
/* mk_encrypt_3gen, mk_api.jl:519-536: mu = +-2^29, b = mu + e + sum_p <a_p, s_p> */
void mko_encrypt(const mko_keyset *ks, uint64_t seed, int count, const uint8_t *bits, int32_t *a, int32_t *b) {
    const mko_params *P = &ks->p;
    int kn = P->k * P->n;
    for (int g = 0; g < count; g++) {
        mko_rng r; rng_seed(&r, seed, 0x3000000ull + (u64)g);
        uint32_t dot = 0;
        for (int i = 0; i < kn; i++) { uint32_t x = (uint32_t)rng_u64(&r); a[(size_t)g * kn + i] = (int32_t)x; dot += x * (uint32_t)ks->lwe_keys[i]; }
        uint32_t mu = (uint32_t)mko_encode_message32(bits[g] ? 1 : -1, 8);
        b[g] = (int32_t)(mu + (uint32_t)mko_dtot32(rng_normal(&r) * P->sigma_lwe) + dot);
    }
}
/* mk_lwe_phase, mk_internals.jl:85-91 (as used by mk_decrypt_3gen, mk_api.jl:607-610:
 * phase = b - sum_p <a_p, s_p>; the reference's lwe_phase of a (a, b=0) sample is -<a,s>) */
void mko_phase(const mko_keyset *ks, int count, const int32_t *a, const int32_t *b, int32_t *phase) {
    int kn = ks->p.k * ks->p.n;
    for (int g = 0; g < count; g++) {
        uint32_t dot = 0;
        for (int i = 0; i < kn; i++) dot += (uint32_t)a[(size_t)g * kn + i] * (uint32_t)ks->lwe_keys[i];
        phase[g] = (int32_t)((uint32_t)b[g] - dot);
    }
}

/* ------------------------------------------------------------------------ */
/* external product, tgsw_3gen.jl:102-113                                     */
/* c0 = accum.a[2] (body), c1 = accum.a[1] (mask)                             */
/* body' = sum_q dec(c0)_q*part_1[q] + sum_q dec(c1)_q*part_2[q]              */
/* mask' = sum_q dec(c0)_q*part_4[q] + sum_q dec(c1)_q*part_3[q]              */
/* ------------------------------------------------------------------------ */
static uint64_t fnv_digits(const int64_t *d, size_t cnt) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < cnt; i++) { h ^= (uint64_t)d[i] & 0xFFFF; h *= 0x100000001b3ull; }
    return h;
}

static void extprod_impl(mko_keyset *ks, int backend, int party, int j, const int64_t *acc_in, int64_t *acc_out,
                         uint64_t *digit_hash) {
    const mko_params *P = &ks->p;
    int N = P->N, l = P->l;
    size_t e = (size_t)party * P->n + j;
    int64_t *dig = malloc(sizeof(int64_t) * 2 * l * N);      /* [src: 0=body(c0), 1=mask(c1)][q][N] */
    mko_decompose(acc_in + N, N, l, P->bgbit, dig);           /* g_c0 = decompose(body) :106 */
    mko_decompose(acc_in, N, l, P->bgbit, dig + (size_t)l * N); /* g_c1 = decompose(mask) :107 */
    if (digit_hash) *digit_hash = fnv_digits(dig, (size_t)2 * l * N);
    /* key part index for (out, src): body<-body part_1(0), body<-mask part_2(1), mask<-body part_4(3), mask<-mask part_3(2) */
    static const int part_of[2][2] = {{3, 2}, {0, 1}};       /* [out: 0=mask,1=body][src: 0=body,1=mask] */
    if (backend == MKO_EXACT_SCHOOLBOOK) {
        const int64_t *key = ks->bsk + e * 4 * l * N;
        u64 *res = calloc((size_t)2 * N, sizeof(u64));
        for (int out = 0; out < 2; out++)
            for (int src = 0; src < 2; src++)
                for (int q = 0; q < l; q++)
                    negacyclic_mac_schoolbook(dig + ((size_t)src * l + q) * N, key + ((size_t)part_of[out][src] * l + q) * N, N, res + (size_t)out * N);
        memcpy(acc_out, res, sizeof(u64) * 2 * N);
        free(res);
    } else if (backend == MKO_EXACT_NTT) {
        mko_prepare_ntt_key(ks);
        const int NL = ks->ntt_limbs, LB = ks->ntt_limb_bits;
        const u64 *key = ks->bsk_ntt + e * 4 * l * NL * N;
        u64 *dh = malloc(sizeof(u64) * 2 * l * N), *tmp = malloc(sizeof(u64) * N), *r = malloc(sizeof(u64) * NL * N);
        for (int s = 0; s < 2 * l; s++) {
            for (int i = 0; i < N; i++) tmp[i] = gl_from_i64(dig[(size_t)s * N + i]);
            gl_negacyclic_fwd(ks->tab, tmp, dh + (size_t)s * N);
        }
        for (int out = 0; out < 2; out++) {
            for (int limb = 0; limb < NL; limb++) {
                u64 *rr = r + (size_t)limb * N;
                memset(rr, 0, sizeof(u64) * N);
                for (int src = 0; src < 2; src++)
                    for (int q = 0; q < l; q++) {
                        const u64 *kp = key + (((size_t)part_of[out][src] * l + q) * NL + limb) * N;
                        const u64 *dp = dh + ((size_t)src * l + q) * N;
                        for (int i = 0; i < N; i++) rr[i] = gl_add(rr[i], gl_mul(dp[i], kp[i]));
                    }
                gl_negacyclic_inv(ks->tab, rr);
            }
            for (int i = 0; i < N; i++) {
                u64 v = 0;
                for (int limb = 0; limb < NL; limb++) v += (u64)gl_lift(r[(size_t)limb * N + i]) << (limb * LB);
                acc_out[(size_t)out * N + i] = (int64_t)v;
            }
        }
        free(dh); free(tmp); free(r);
    } else { /* MKO_FFT */
        mko_prepare_fft_key(ks);
        int M = N / 2;
        const double *key = ks->bsk_fft + e * 4 * l * N;
        double *dh = malloc(sizeof(double) * 2 * l * N), *sum0 = malloc(sizeof(double) * N), *sum1 = malloc(sizeof(double) * N);
        for (int s = 0; s < 2 * l; s++) fft_forward(ks->tab, dig + (size_t)s * N, dh + (size_t)s * N);
        for (int out = 0; out < 2; out++) {
            /* sum(ft(g_c0) .* part_a) + sum(ft(g_c1) .* part_b): each sum over q first, then added  :109-110 */
            for (int src = 0; src < 2; src++) {
                double *sum = src == 0 ? sum0 : sum1;
                for (int q = 0; q < l; q++) {
                    const double *kp = key + ((size_t)part_of[out][src] * l + q) * N;
                    const double *dp = dh + ((size_t)src * l + q) * N;
                    for (int i = 0; i < M; i++) {
                        double re = dp[2 * i] * kp[2 * i] - dp[2 * i + 1] * kp[2 * i + 1];
                        double im = dp[2 * i] * kp[2 * i + 1] + dp[2 * i + 1] * kp[2 * i];
                        if (q == 0) { sum[2 * i] = re; sum[2 * i + 1] = im; } else { sum[2 * i] += re; sum[2 * i + 1] += im; }
                    }
                }
            }
            for (int i = 0; i < N; i++) sum0[i] += sum1[i];
            fft_inverse(ks->tab, sum0, acc_out + (size_t)out * N);
        }
        free(dh); free(sum0); free(sum1);
    }
    free(dig);
}
void mko_extprod(mko_keyset *ks, int backend, int party, int j, const int64_t *acc_in, int64_t *acc_out) {
    int N = ks->p.N;
    int64_t *tmp = malloc(sizeof(int64_t) * 2 * N);
    extprod_impl(ks, backend, party, j, acc_in, tmp, NULL);
    memcpy(acc_out, tmp, sizeof(int64_t) * 2 * N);
    free(tmp);
}

/* mk_mux_rotate_3gen, 3gen_mk_internals.jl:59-62: acc + ExtProd(X^bara*acc - acc, bk) */
static void mux_rotate_impl(mko_keyset *ks, int backend, int party, int j, int32_t bara, int64_t *acc, uint64_t *dh) {
    int N = ks->p.N;
    int64_t *tmp = malloc(sizeof(int64_t) * 2 * N), *prod = malloc(sizeof(int64_t) * 2 * N);
    for (int c = 0; c < 2; c++) {
        mko_mul_by_monomial(acc + (size_t)c * N, N, bara, tmp + (size_t)c * N);
        for (int i = 0; i < N; i++) tmp[(size_t)c * N + i] = (int64_t)((u64)tmp[(size_t)c * N + i] - (u64)acc[(size_t)c * N + i]);
    }
    extprod_impl(ks, backend, party, j, tmp, prod, dh);
    for (int i = 0; i < 2 * N; i++) acc[i] = (int64_t)((u64)acc[i] + (u64)prod[i]);
    free(tmp); free(prod);
}
void mko_mux_rotate(mko_keyset *ks, int backend, int party, int j, int32_t bara, int64_t *acc) {
    mux_rotate_impl(ks, backend, party, j, bara, acc, NULL);
}

/* mk_bootstrap_wo_keyswitch_3gen :99-109 -> mk_blind_rotate_and_extract_3gen :88-95
 * -> mk_blind_rotate_3gen :78-84 -> mk_ith_blind_rotate_3gen :66-74 -> rlwe_extract_sample_64 rlwe.jl:70-74 */
void mko_bootstrap_wo_keyswitch(mko_keyset *ks, int backend, int64_t mu, const int32_t *a, int32_t b,
                                int32_t *ext_a, int32_t *ext_b, int64_t *acc_out, uint64_t *digit_log) {
    const mko_params *P = &ks->p;
    int N = P->N, n = P->n, k = P->k;
    int32_t barb = mko_decode_message32(b, 2 * N);                   /* :102 */
    int64_t *acc = calloc((size_t)2 * N, sizeof(int64_t));           /* [mask, body]; rlwe_noiseless_trivial rlwe.jl:113-119 */
    int64_t *tv = malloc(sizeof(int64_t) * N);
    for (int i = 0; i < N; i++) tv[i] = mu;                          /* testvect :106 */
    mko_mul_by_monomial(tv, N, -(int64_t)barb, acc + N);             /* testvectbis :91 */
    for (int p = 0; p < k; p++)                                      /* parties outer :80 */
        for (int j = 0; j < n; j++) {                                /* coefficients inner :67 */
            int32_t bara = mko_decode_message32(a[(size_t)p * n + j], 2 * N);   /* :103 */
            uint64_t h = 0;
            if (bara != 0) mux_rotate_impl(ks, backend, p, j, bara, acc, digit_log ? &h : NULL);  /* :69-71 */
            if (digit_log) digit_log[(size_t)p * n + j] = h;
        }
    /* rlwe_extract_sample_64: a = t64tot32.(reverse_polynomial(mask)), b = t64tot32(body[1]) */
    ext_a[0] = mko_t64tot32(acc[0]);
    for (int i = 1; i < N; i++) ext_a[i] = mko_t64tot32((int64_t)(0 - (u64)acc[N - i]));   /* polynomials.jl:69-72 */
    *ext_b = mko_t64tot32(acc[N]);
    if (acc_out) memcpy(acc_out, acc, sizeof(int64_t) * 2 * N);
    free(acc); free(tv);
}

/* mk_keyswitch_3gen mk_internals.jl:730-744; keyswitch keyswitch.jl:45-80 */
void mko_keyswitch(const mko_keyset *ks, const int32_t *ext_a, int32_t ext_b, int32_t *out_a, int32_t *out_b) {
    const mko_params *P = &ks->p;
    int N = P->N, n = P->n, k = P->k, t = P->t, bb = P->basebit, B1 = (1 << bb) - 1;
    uint32_t prec_offset = 1u << (32 - (1 + bb * t));               /* :58 */
    uint32_t bsum = (uint32_t)ext_b;
    for (int p = 0; p < k; p++) {
        uint32_t *res = calloc((size_t)n + 1, sizeof(uint32_t));     /* lwe_noiseless_trivial(0) */
        const int32_t *rows = ks->ksk + (size_t)p * N * t * B1 * (n + 1);
        for (int i = 0; i < N; i++) {
            uint32_t aibar = (uint32_t)ext_a[i] + prec_offset;       /* :59 */
            for (int j = 1; j <= t; j++) {
                uint32_t d = (uint32_t)((int32_t)aibar >> (32 - j * bb)) & (uint32_t)B1;   /* :65-67 */
                if (d != 0) {                                        /* :74-76 */
                    const int32_t *row = rows + (((size_t)i * t + (j - 1)) * B1 + (d - 1)) * (n + 1);
                    for (int c = 0; c <= n; c++) res[c] -= (uint32_t)row[c];
                }
            }
        }
        for (int c = 0; c < n; c++) out_a[(size_t)p * n + c] = (int32_t)res[c];
        bsum += res[n];
        free(res);
    }
    *out_b = (int32_t)bsum;
}

typedef struct {
    mko_keyset *ks; int backend; int64_t mu; const int32_t *a, *b; int32_t *oa, *ob;
} boot_ctx;
static void boot_one(void *v, int g, int tid) {
    (void)tid;
    boot_ctx *c = (boot_ctx *)v;
    const mko_params *P = &c->ks->p;
    size_t kn = (size_t)P->k * P->n;
    int32_t *ea = malloc(sizeof(int32_t) * P->N), eb;
    mko_bootstrap_wo_keyswitch(c->ks, c->backend, c->mu, c->a + g * kn, c->b[g], ea, &eb, NULL, NULL);
    mko_keyswitch(c->ks, ea, eb, c->oa + g * kn, &c->ob[g]);
    free(ea);
}
void mko_bootstrap_batch(mko_keyset *ks, int backend, int64_t mu, int count, const int32_t *a, const int32_t *b,
                         int32_t *out_a, int32_t *out_b, int nthreads) {
    if (backend == MKO_FFT) mko_prepare_fft_key(ks);
    if (backend == MKO_EXACT_NTT) mko_prepare_ntt_key(ks);
    boot_ctx c = {ks, backend, mu, a, b, out_a, out_b};
    parallel_for(count, nthreads, boot_one, &c);
}

/* gates, 3gen_mk_gates.jl:8-14 (NAND), 24-30 (OR), 40-46 (AND), 55-64 (3AND), 68-74 (XOR) */
void mko_gate_prologue(const mko_params *P, int gate, const int32_t *xa, int32_t xb, const int32_t *ya, int32_t yb,
                       const int32_t *za, int32_t zb, int32_t *ta, int32_t *tb) {
    int kn = P->k * P->n;
    uint32_t mu0, cx, cy, cz = 0;
    switch (gate) {
    case MKO_GATE_NAND: mu0 = (uint32_t)mko_encode_message32(1, 8); cx = (uint32_t)-1; cy = (uint32_t)-1; break;
    case MKO_GATE_OR:   mu0 = (uint32_t)mko_encode_message32(1, 8); cx = 1; cy = 1; break;
    case MKO_GATE_AND:  mu0 = (uint32_t)mko_encode_message32(-1, 8); cx = 1; cy = 1; break;
    case MKO_GATE_XOR:  mu0 = (uint32_t)mko_encode_message32(1, 4); cx = 2; cy = 2; break;
    case MKO_GATE_AND3: mu0 = (uint32_t)mko_encode_message32(-1, 4); cx = 1; cy = 1; cz = 1; break;
    default: mu0 = 0; cx = cy = 0; break;
    }
    for (int i = 0; i < kn; i++)
        ta[i] = (int32_t)(cx * (uint32_t)xa[i] + cy * (uint32_t)ya[i] + (cz ? cz * (uint32_t)za[i] : 0u));
    *tb = (int32_t)(mu0 + cx * (uint32_t)xb + cy * (uint32_t)yb + (cz ? cz * (uint32_t)zb : 0u));
}
void mko_gate_batch(mko_keyset *ks, int backend, int gate, int count,
                    const int32_t *xa, const int32_t *xb, const int32_t *ya, const int32_t *yb,
                    const int32_t *za, const int32_t *zb, int32_t *oa, int32_t *ob, int nthreads) {
    const mko_params *P = &ks->p;
    size_t kn = (size_t)P->k * P->n;
    int32_t *ta = malloc(sizeof(int32_t) * kn * (size_t)count), *tb = malloc(sizeof(int32_t) * (size_t)count);
    for (int g = 0; g < count; g++)
        mko_gate_prologue(P, gate, xa + g * kn, xb[g], ya + g * kn, yb[g], za ? za + g * kn : NULL, zb ? zb[g] : 0,
                          ta + g * kn, &tb[g]);
    /* output message mu = encode_message64(1, 8) = 2^61 since rlwe_is32 == false (3gen_mk_gates.jl:12) */
    mko_bootstrap_batch(ks, backend, mko_encode_message64(1, 8), count, ta, tb, oa, ob, nthreads);
    free(ta); free(tb);
}
