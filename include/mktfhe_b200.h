/*
 * mktfhe_b200.h -- C ABI of libmktfhe_b200.so, the B200 (sm_100a) engine for the
 * bootstrapped-gate path of the 3rd-generation multi-key TFHE scheme of
 * Animesh005/Torus-FHE.
 *
 * The reference has no FFI layer: its seam is the set of plain Julia functions
 * exported from module TFHE (3-gen-mk-tfhe/src/TFHE.jl:113-119, 177-196).  Each
 * entry point below names the reference function (file:line, relative to
 * 3-gen-mk-tfhe/src/) it replaces; INTEGRATION.md shows the Julia `ccall`
 * binding and the Python `ctypes` binding.
 *
 * Conventions
 *  - every function returns 0 on success or a negative MKTFHE_E* code; the
 *    message is available from mktfhe_last_error().  No exceptions or aborts
 *    cross this boundary.
 *  - a context is used by one host thread at a time.  mktfhe_create binds it to
 *    one GPU; mktfhe_create_multi spans several GPUs of the box behind the same
 *    handle: keys are loaded once and broadcast GPU to GPU inside
 *    mktfhe_finalize_keys, and every host-pointer batch call shards [G]
 *    contiguously over the GPUs (one internal host thread and stream per GPU, no
 *    exchange on the data path).  The one-context-per-rank form (torchrun,
 *    mktfhe_key_buffers + NCCL) remains available.
 *  - plain (non-_dev) entry points take HOST pointers; the caller owns them and
 *    nothing is retained after return except uploaded keys (copied).
 *  - *_dev entry points take DEVICE pointers valid on the context's GPU and a
 *    cudaStream_t (as void*, NULL = the context's own stream); they are
 *    asynchronous with respect to the host.  The context's own stream is
 *    non-blocking: with NULL the caller must make sure the operands are
 *    complete (they are not ordered after work on the legacy default stream).
 *    All calls on one context share its scratch buffers and timing events, so
 *    they must be stream-ordered with each other (same stream, or ordered by
 *    events): two *_dev calls in flight on unordered streams race.
 *  - arithmetic: every external product is exact mod (X^N + 1, 2^64).  Contexts with
 *    N = 1024 compute it on the FP64 pipe (an exact folded complex FFT with the
 *    key word split in three limbs, two in Torus32 mode: csrc/fft64.cuh); N = 2048
 *    on four 28-bit NTT primes with a CRT lift.  Results are bit-identical either way;
 *    MKTFHE_B200_FFT=0 in the environment of mktfhe_create selects the NTT kernels
 *    (A/B runs), mktfhe_describe reports which engine a context runs.
 *  - ciphertext layout: MKLweSample (mk_internals.jl:23-37) `a::Array{Int32,2}`
 *    of shape (n, k), column-major == int32 [k][n] per sample; batches are
 *    int32 a[G][k][n], int32 b[G].
 */
#ifndef MKTFHE_B200_H
#define MKTFHE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MKTFHE_OK 0
#define MKTFHE_EINVAL (-1)   /* bad argument / unsupported parameter set */
#define MKTFHE_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define MKTFHE_ESTATE (-3)   /* keys not loaded / not finalized */
#define MKTFHE_ENOMEM (-4)

/* gate ids; the linear prologue constants are those of 3gen_mk_gates.jl */
#define MKTFHE_GATE_NAND 0   /* mk_gate_nand_3gen  3gen_mk_gates.jl:8-14   +1/8 - x - y   */
#define MKTFHE_GATE_OR 1     /* mk_gate_or_3gen    3gen_mk_gates.jl:24-30  +1/8 + x + y   */
#define MKTFHE_GATE_AND 2    /* mk_gate_and_3gen   3gen_mk_gates.jl:40-46  -1/8 + x + y   */
#define MKTFHE_GATE_XOR 3    /* mk_gate_xor_3gen   3gen_mk_gates.jl:68-74  +1/4 + 2x + 2y */
#define MKTFHE_GATE_AND3 4   /* mk_gate_3and_3gen  3gen_mk_gates.jl:55-64  -1/4 + x+y+z   */

/* Integer part of SchemeParameters_3gen (api.jl:50-67); rlwe_mask_size = 1 and
 * rlwe_is32 = false as in every 3gen set (mk_api.jl:32-322).  Supported:
 *  - N = 1024 (2..8 parties, mk_api.jl:32-146): 1 <= l <= 4, bgbit <= 8, 2l*N*2^(bgbit-1)*2^63 < M/4 with M ~ 2^84 the
 *    product of the three NTT primes (so the CRT result is exact), n*k <= 8192; with MKTFHE_FLAG_TORUS32: l = 2..4,
 *    bgbit <= 16, l*bgbit <= 32, n*k <= 32768;
 *  - N = 2048 (16..256 parties, mk_api.jl:214-310): l = 1 or 2, bgbit <= 27, four NTT primes (M ~ 2^112), n*k <= 2^18;
 *  - t*basebit <= 31, basebit <= 16.
 * Rejected with MKTFHE_EINVAL: N = 4096 (512 parties). */
typedef struct {
    int32_t n;        /* lwe_size */
    int32_t N;        /* rlwe_polynomial_degree */
    int32_t k;        /* parties */
    int32_t l;        /* gsw_decomp_length */
    int32_t bgbit;    /* gsw_log2_base */
    int32_t t;        /* ks_decomp_length */
    int32_t basebit;  /* ks_log2_base */
    int32_t reserved; /* flags: 0, or MKTFHE_FLAG_TORUS32 */
} mktfhe_params;

/* Torus32 mode (N = 1024 only): for schemes whose ring elements are Torus32 (rlwe_is32 = true: the single-key sets of api.jl:76-113
 * and the CCS multi-key sets) with a gadget base above 2^8.  The bootstrapping key is loaded UNSHIFTED -- int64 words holding the
 * 32-bit signed Torus32 values -- gadget digits may have up to 16 bits (l * bgbit <= 32), accumulators, test-vector message and
 * the parity hooks carry Torus32 values in the top half of their 64-bit words (v << 32), and each external product is added as
 * R << 32.  Without the flag a Torus32 key can still be served through the default kernels as K << 32 when bgbit <= 8
 * (torus-fhe_b200/tfhe1.py does that for tfhe_parameters_128: their table-driven first transform stage is faster). */
#define MKTFHE_FLAG_TORUS32 1

typedef struct mktfhe_ctx mktfhe_ctx;

/* -- lifetime ----------------------------------------------------------- */
int mktfhe_create(const mktfhe_params *params, int device, mktfhe_ctx **out);
/* One context spanning n_devices GPUs of this box (the gate API of 3gen_mk_gates.jl has no notion of
 * devices, so this is what lets `mk_gate_nand_3gen(bk, ks, xs, ys)` on a vector use the whole box).  devices = NULL means
 * GPUs 0 .. n_devices-1, n_devices = 0 means every visible GPU; a device may be listed more than once (replicas sharing a
 * GPU: a test aid).  Keys are loaded into the first device and broadcast in mktfhe_finalize_keys -- a binomial tree of
 * cudaMemcpyPeerAsync over NVLink, or one grouped ncclBroadcast per key buffer when MKTFHE_B200_BCAST=nccl (libnccl.so.2
 * is bound with dlopen at that point; no link-time dependency).  With n_devices = 1 the result is a plain context. */
int mktfhe_create_multi(const mktfhe_params *params, int n_devices, const int *devices, mktfhe_ctx **out);
void mktfhe_destroy(mktfhe_ctx *ctx);
/* GPUs a context spans (1 for mktfhe_create) */
int mktfhe_device_count(const mktfhe_ctx *ctx);
/* replica i of a multi-device context as a single-device context (i = 0: ctx itself) and its CUDA device ordinal, for the
 * *_dev entry points, which take pointers into ONE GPU: called on the spanning handle they address its first GPU (replica 0),
 * the other GPUs are reached through their replicas.  A replica is owned by ctx (do not destroy it). */
int mktfhe_device_ctx(mktfhe_ctx *ctx, int i, mktfhe_ctx **replica, int *device);
/* the contiguous slice [lo, hi) of a batch of G that replica i processes in the host-pointer calls */
int mktfhe_shard_bounds(const mktfhe_ctx *ctx, size_t G, int i, size_t *lo, size_t *hi);
/* page-lock / release a caller buffer (cudaHostRegister, portable): with pinned ciphertext buffers the per-GPU copies of a
 * multi-device batch call are true DMA and overlap; pageable memory works but is staged by the driver. */
int mktfhe_pin_host(void *p, size_t bytes);
int mktfhe_unpin_host(void *p);
/* message of the last failing call on ctx (ctx == NULL: last mktfhe_create failure) */
const char *mktfhe_last_error(const mktfhe_ctx *ctx);

/* -- keys ---------------------------------------------------------------- */
/* Replaces TransformedBootstrapKeyPart_3gen(bk::BootstrapKeyPart_3gen)
 * (3gen_mk_internals.jl:45-55): takes party `party`'s INTEGER key
 * bk.gsw_key[j].part_{1..4}[q].coeffs as int64 [n][4][l][N] (host), splits every
 * polynomial modulo the three 28-bit NTT primes, transforms the residues on the
 * GPU and stores them in the streaming layout of the blind-rotate kernel. */
int mktfhe_load_bsk(mktfhe_ctx *ctx, int party, const int64_t *polys);
/* KeyswitchKey.key::Array{LweSample,3} of dims (base-1, t, N) (keyswitch.jl:7-42),
 * flattened as int32 [N][t][base-1][n+1] (a[0..n-1] then b), host. */
int mktfhe_load_ksk(mktfhe_ctx *ctx, int party, const int32_t *rows);
/* KeyswitchKey(rng, alpha, params, out_key, rlwe_key) (keyswitch.jl:14-41) generated ON the GPU, straight into the context's
 * key-switching key of `party` (no 44 MB .. 140 MB per party built on the host and copied): lwe_key = the party's LWE secret,
 * int32 [n]; rlwe_key = its RLWE secret polynomial, int64 [N] (extract_lwe_key, rlwe.jl:35-41); sigma = ks_noise_stddev; seed
 * keys the counter-based generator (Philox4x32-10): row (i, j, h) = (a uniform, b = (z_i h) << (32 - j basebit) + e + <a, s>),
 * e Gaussian and re-centred to zero mean over the key as :28-29.  The secrets are erased from device scratch before return. */
int mktfhe_generate_ksk(mktfhe_ctx *ctx, int party, const int32_t *lwe_key, const int64_t *rlwe_key, double sigma, uint64_t seed);
/* all parties loaded (or received through mktfhe_key_buffers) -> ready */
int mktfhe_finalize_keys(mktfhe_ctx *ctx);
/* device-resident key buffers, for a one-time NCCL broadcast from rank 0:
 * non-root ranks receive into these pointers, then call mktfhe_finalize_keys
 * with no loads. */
int mktfhe_key_buffers(mktfhe_ctx *ctx, void **bsk_dev, size_t *bsk_bytes, void **ksk_dev, size_t *ksk_bytes);
int mktfhe_mark_keys_received(mktfhe_ctx *ctx);

/* -- the hot path ---------------------------------------------------------- */
/* mk_bootstrap_3gen(bk, ks, mu, x) (3gen_mk_internals.jl:112-116) on a batch:
 * mod-switch, blind rotate over the k*n key elements, sample extraction
 * (rlwe_extract_sample_64, rlwe.jl:70-74) and mk_keyswitch_3gen
 * (mk_internals.jl:730-744).  mu is the Torus64 test-vector message. */
int mktfhe_bootstrap_batch(mktfhe_ctx *ctx, int64_t mu, size_t G, const int32_t *a_in, const int32_t *b_in,
                           int32_t *a_out, int32_t *b_out);
/* mk_gate_{nand,or,and,xor,3and}_3gen(bk, ks, x, y[, z]) (3gen_mk_gates.jl:8-74)
 * on a batch: linear prologue fused into the bootstrap; output message
 * encode_message64(1, 8) = 2^61.  za/zb are only read for MKTFHE_GATE_AND3. */
int mktfhe_gate_batch(mktfhe_ctx *ctx, int gate, size_t G, const int32_t *xa, const int32_t *xb,
                      const int32_t *ya, const int32_t *yb, const int32_t *za, const int32_t *zb,
                      int32_t *oa, int32_t *ob);
/* device-pointer variants (inputs/outputs resident in HBM) */
int mktfhe_bootstrap_batch_dev(mktfhe_ctx *ctx, int64_t mu, size_t G, const int32_t *a_in, const int32_t *b_in,
                               int32_t *a_out, int32_t *b_out, void *stream);
int mktfhe_gate_batch_dev(mktfhe_ctx *ctx, int gate, size_t G, const int32_t *xa, const int32_t *xb,
                          const int32_t *ya, const int32_t *yb, const int32_t *za, const int32_t *zb,
                          int32_t *oa, int32_t *ob, void *stream);

/* General form of the gates above and of the single-key gates (gates.jl:16-142: NAND, OR, AND, XOR, XNOR, NOR, ANDNY, ANDYN,
 * ORNY, ORYN): temp = mu0 + cx*x + cy*y + cz*z (Int32 wrap-around, lwe.jl:62-76), then bootstrap(temp) with test-vector
 * message mu (bootstrap.jl:97-100 / 3gen_mk_internals.jl:112-116).  y/z are read only when cy/cz != 0 (may be NULL). */
int mktfhe_affine_bootstrap_batch(mktfhe_ctx *ctx, int32_t mu0, int32_t cx, int32_t cy, int32_t cz, int64_t mu, size_t G,
                                  const int32_t *xa, const int32_t *xb, const int32_t *ya, const int32_t *yb,
                                  const int32_t *za, const int32_t *zb, int32_t *oa, int32_t *ob);
int mktfhe_affine_bootstrap_batch_dev(mktfhe_ctx *ctx, int32_t mu0, int32_t cx, int32_t cy, int32_t cz, int64_t mu, size_t G,
                                      const int32_t *xa, const int32_t *xb, const int32_t *ya, const int32_t *yb,
                                      const int32_t *za, const int32_t *zb, int32_t *oa, int32_t *ob, void *stream);

/* One launch for a batch whose gates differ: gate_ids[g] in MKTFHE_GATE_* selects the prologue of gate g (a dependency
 * level of a circuit such as mk_add_3gen, 3gen_mk_gates.jl:183-220, mixes XOR/AND/OR gates).  za/zb may be NULL when
 * no gate is MKTFHE_GATE_AND3.  The _dev variant takes device pointers (gate_ids included) and does not validate ids:
 * an id outside 0..4 bootstraps the zero sample. */
int mktfhe_gate_batch_mixed(mktfhe_ctx *ctx, size_t G, const int32_t *gate_ids, const int32_t *xa, const int32_t *xb,
                            const int32_t *ya, const int32_t *yb, const int32_t *za, const int32_t *zb,
                            int32_t *oa, int32_t *ob);
int mktfhe_gate_batch_mixed_dev(mktfhe_ctx *ctx, size_t G, const int32_t *gate_ids, const int32_t *xa, const int32_t *xb,
                                const int32_t *ya, const int32_t *yb, const int32_t *za, const int32_t *zb,
                                int32_t *oa, int32_t *ob, void *stream);

/* -- parity hooks (host pointers; one call per stage of the path) ---------- */
/* tgsw_extern_mul_3gen(accum, bk[party].gsw_key[j]) (tgsw_3gen.jl:102-113) for
 * G accumulators: acc = int64 [G][2][N] with [0] = mask (accum.a[1]) and
 * [1] = body (accum.a[2]); elem[g] = party*n + j. */
int mktfhe_extprod_batch(mktfhe_ctx *ctx, size_t G, const int32_t *elem, const int64_t *acc_in, int64_t *acc_out);
/* the same on device pointers (elem included), asynchronous on `stream`: the building block of the CCS scheme's hybrid product
 * (UniProduct, mk_internals.jl:471-535), which torus-fhe_b200/tfhe_ccs.py composes from 2 (k+1) of these per blind-rotate step */
int mktfhe_extprod_batch_dev(mktfhe_ctx *ctx, size_t G, const int32_t *elem, const int64_t *acc_in, int64_t *acc_out, void *stream);
/* mk_bootstrap_wo_keyswitch of the CCS scheme (mk_internals.jl:793-850) on a batch, as a composition run entirely by the library: per
 * blind-rotate step two launches of G (parties + 1) external products with the rotation / accumulation kernels between them;
 * accumulators stay in HBM.  ctx: a Torus32-mode context whose parties * (parties + 2) "parties" hold the key elements of the hybrid
 * product (UniProduct, :471-535) -- pseudo-party p (parties + 1) + i (i = 0: the b polynomial, i >= 1: a_i): part_1 = d[p][j],
 * part_4 = -a or b_i; pseudo-party parties (parties + 1) + p: part_1 = f0[p][j], part_4 = f1[p][j]; part_2 = part_3 = 0
 * (torus-fhe_b200/tfhe_ccs.py, build_elements).  mu: the Torus32 test-vector message.  a_in int32 [G][parties][n], b_in [G] ->
 * the extracted sample with one mask per party, ext_a int32 [G][parties][N], ext_b [G]. */
int mktfhe_ccs_blind_rotate_batch(mktfhe_ctx *ctx, int parties, int32_t mu, size_t G, const int32_t *a_in, const int32_t *b_in,
                                  int32_t *ext_a, int32_t *ext_b);
/* mk_keyswitch(ks, sample::MKLweSample) of the CCS scheme (mk_internals.jl:703-719): the extracted sample has one mask per party,
 * ext_a = int32 [G][k][N], ext_b = int32 [G]; party p's mask goes through party p's key. */
int mktfhe_mk_keyswitch_batch(mktfhe_ctx *ctx, size_t G, const int32_t *ext_a, const int32_t *ext_b, int32_t *a_out, int32_t *b_out);
/* mk_bootstrap_wo_keyswitch_3gen (3gen_mk_internals.jl:99-109): returns the
 * extracted sample ext = int32 [G][N+1] (a'[0..N-1], b') and, if acc_out is not
 * NULL, the final accumulator int64 [G][2][N]. */
int mktfhe_blind_rotate_batch(mktfhe_ctx *ctx, int64_t mu, size_t G, const int32_t *a_in, const int32_t *b_in,
                              int32_t *ext_out, int64_t *acc_out);
/* mk_keyswitch_3gen (mk_internals.jl:730-744) on ext = int32 [G][N+1] */
int mktfhe_keyswitch_batch(mktfhe_ctx *ctx, size_t G, const int32_t *ext, int32_t *a_out, int32_t *b_out);
/* exact negacyclic products c = a * b mod (X^N+1, 2^64) through the same three-prime NTT + CRT:
 * a = int64 [G][N] with |a_i| <= 2^8 at N = 1024, <= 2^25 at N = 2048 ("digit" / ternary operand: N * |a| * 2^63 stays inside
 * the CRT range; larger entries are rejected with MKTFHE_EINVAL),
 * b = int64 [G][N].  Also the primitive of key generation (tgsw_3gen.jl:85-88 uses DarkIntegers' exact `*`). */
int mktfhe_negacyclic_mul_batch(mktfhe_ctx *ctx, size_t G, const int64_t *a, const int64_t *b, int64_t *c);

/* -- introspection ---------------------------------------------------------- */
/* hash of the kernel sources this library was built from (the ncu captures under profiles/ record the id they apply to) */
const char *mktfhe_build_id(void);
/* one-line JSON description of the context: build id, devices, how the keys were broadcast, whether the most recent
 * bootstrap ran the key switch as the blind-rotate kernel's epilogue, the external-product engine ("fft64" / "ntt_rns"),
 * gates per CTA, key bytes */
int mktfhe_describe(const mktfhe_ctx *ctx, char *buf, size_t cap);
/* kernels launched by this context so far (all devices) */
uint64_t mktfhe_launch_count(const mktfhe_ctx *ctx);
/* device time (ms, CUDA events on the launching stream) of the blind-rotate and
 * key-switch kernels of the most recent batch call (multi-device: the slowest GPU); blocks until they finished */
int mktfhe_last_kernel_ms(mktfhe_ctx *ctx, float *blind_rotate_ms, float *keyswitch_ms);
/* bytes of the streamed bootstrapping key read per gate (three u32 residues per coefficient) and of ksk rows gathered per gate */
int mktfhe_algorithmic_bytes(const mktfhe_ctx *ctx, double *bsk_bytes_per_gate, double *ksk_bytes_per_gate);

#ifdef __cplusplus
}
#endif
#endif
