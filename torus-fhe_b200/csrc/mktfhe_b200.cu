// mktfhe_b200.cu -- C ABI (include/mktfhe_b200.h) over the sm_100a kernels in kernels.cuh.
//
// One context owns one GPU: the NTT-domain bootstrapping key (three 28-bit-prime residues per
// reference polynomial), the key-switching key, the twiddle tables, one stream and
// growable device staging buffers.  There is no CPU fallback: every entry point
// either launches the CUDA kernels or returns an error code.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <dlfcn.h>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mktfhe_b200.h"
#include "kernels.cuh"
#include "kernels2k.cuh"
#include "fft64.cuh"
#include "tables_fft.h"
#include "tables2048.h"
#include "tables.h"

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct mktfhe_ctx {
    mktfhe_params prm{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // br start/stop, ks start/stop
    bool ev_valid = false;
    bool last_fused = false;     // the most recent bootstrap-like call ran the key switch as the blind-rotate kernel's epilogue
    u32* d_bsk = nullptr;
    size_t bsk_bytes = 0;
    int32_t* d_ksk = nullptr;
    size_t ksk_bytes = 0;
    rns::uint2_* d_twB = nullptr;
    // FP64 FFT channel (fft64.cuh; N = 1024, Torus64, byte digits; MKTFHE_B200_FFT=0 selects the three-prime NTT kernels instead: A/B runs):
    // d_bsk then holds the bootstrapping key as limb spectra (d_bsk_fft == d_bsk) and every launch shape runs the FFT kernels
    bool fft = false;
    bool fft_small = true;       // batches (and tails) of at most one gate per SM: one six-warp FFT gate per CTA (MKTFHE_B200_FFT_SMALL=0: the
                                 // NTT latency kernel -- needs the NTT key layout, so it then also keeps that copy, stored in front of the spectra)
    size_t bsk_ntt_bytes = 0;    // size of the NTT-layout part of d_bsk (0 when every launch shape runs the FFT kernels)
    mkf::cpx* d_bsk_fft = nullptr;
    mkf::cpx* d_twF = nullptr;
    int gpc = 1;                 // gates per CTA of the blind-rotate / external-product kernels
    int num_sms = 0;
    bool split_tail_all = false; // MKTFHE_B200_SPLIT_TAIL=2: the tail launch at every l, not only l = 2
    bool split_tail = true;      // MKTFHE_B200_SPLIT_TAIL=0: no separate one-gate-per-CTA launch for the tail of a large batch (A/B)
    bool latency_kernel = true;  // MKTFHE_B200_LATENCY=0 turns the 12-warp small-batch launch off (A/B)
    bool t32 = false;            // Torus32 mode (MKTFHE_FLAG_TORUS32): blind_rotate_t32_kernel / extprod_t32_kernel
    bool two_k16 = true;         // N = 2048: the sixteen-warp kernel (MKTFHE_B200_2K=8 selects the first, eight-warp kernel: A/B runs)
    bool fuse_ks = true;         // key switch as the epilogue of the blind-rotate kernel (MKTFHE_B200_FUSE_KS=0 disables: A/B runs)
    std::vector<char> bsk_loaded, ksk_loaded;
    bool ready = false;
    DevBuf in[6], ext, oa, ob, accin, accout, elem, raw, gids;
    uint64_t launches = 0;
    std::string err;
    // multi-device context (mktfhe_create_multi): this context is replica 0 and owns one further single-device context per
    // additional GPU; host-pointer batch calls shard [G] contiguously over all replicas, one host thread per replica
    std::vector<mktfhe_ctx*> kids;
    bool is_kid = false;
    mktfhe_ctx* parent = nullptr;   // replicas 1.. point at the spanning context
    uint64_t seq = 0;               // spanning context: number of bootstrap-like calls so far; ev_seq: the call that last recorded this replica's events
    uint64_t ev_seq = 0;
    bool in_shard_call = false;     // a sharded host-pointer call is in flight: its replicas all belong to call `seq`
    std::string bcast_how;       // how the keys reached the replicas ("p2p" / "nccl"), for mktfhe_describe
};

namespace {

int fail(mktfhe_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU_TRY(ctx, call)                                                                     \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? MKTFHE_ENOMEM : MKTFHE_ECUDA,  \
                        "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int reserve(mktfhe_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return MKTFHE_OK;
    if (b.p) CU_TRY(c, cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    CU_TRY(c, cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return MKTFHE_OK;
}

int check_params(const mktfhe_params* p) {
    if (!p) return fail(nullptr, MKTFHE_EINVAL, "params is NULL");
    // shifts by these fields follow: bound them before anything computes 1 << bgbit or 1 << basebit
    if (p->bgbit < 1 || p->bgbit > 30) return fail(nullptr, MKTFHE_EINVAL, "unsupported gsw_log2_base=%d", p->bgbit);
    if (p->basebit < 1 || p->basebit > 16) return fail(nullptr, MKTFHE_EINVAL, "unsupported ks_log2_base=%d (1..16)", p->basebit);
    if (p->N == mk2k::N) {
        if (p->reserved) return fail(nullptr, MKTFHE_EINVAL, "flags are not defined for N=2048");
        // N = 2048 sets (mk_api.jl:214-310): l = 1 or 2 with a wide gadget base, four-prime exact product (kernels2k.cuh)
        if (p->l < 1 || p->l > 2) return fail(nullptr, MKTFHE_EINVAL, "N=2048 is supported with gsw_decomp_length l=1 or 2 (got %d)", p->l);
        if (p->bgbit < 1 || p->bgbit > 27) return fail(nullptr, MKTFHE_EINVAL, "N=2048: unsupported gsw_log2_base=%d (1..27)", p->bgbit);
        if (std::log2((double)(2 * p->l) * p->N) + (p->bgbit - 1) + 63.0 >= rns2k::log2_crt_bound())
            return fail(nullptr, MKTFHE_EINVAL, "2l*N*Bg/2*2^63 exceeds the CRT range of the four-prime exact product");
        if (p->k < 1 || p->n < 1 || (long long)p->n * p->k > (1 << 18)) return fail(nullptr, MKTFHE_EINVAL, "need 1 <= n*k <= 2^18");
        if (p->n + 1 > mk::KS_THREADS * mk::KS_MAXCOLS) return fail(nullptr, MKTFHE_EINVAL, "n too large for the key-switch kernel");
        if (p->t < 1 || p->basebit < 1 || p->t * p->basebit > 31) return fail(nullptr, MKTFHE_EINVAL, "need t*basebit <= 31");
        return MKTFHE_OK;
    }
    if (p->N != mk::N) return fail(nullptr, MKTFHE_EINVAL, "unsupported N=%d (1024 or 2048)", p->N);
    if (p->reserved & ~MKTFHE_FLAG_TORUS32) return fail(nullptr, MKTFHE_EINVAL, "unknown flags 0x%x", p->reserved);
    if (p->reserved & MKTFHE_FLAG_TORUS32) {
        // Torus32 mode: unshifted 32-bit keys, 16-bit digit fields, products added as R << 32: |R| <= 2l * N * 2^(bgbit-1) * 2^31 < 2^59
        if (p->l < 2 || p->l > 4) return fail(nullptr, MKTFHE_EINVAL, "Torus32 mode is built for gsw_decomp_length l=2..4 (got %d)", p->l);
        if (p->bgbit < 1 || p->bgbit > 16 || p->l * p->bgbit > 32) return fail(nullptr, MKTFHE_EINVAL, "Torus32 mode: need bgbit <= 16 and l*bgbit <= 32");
        // k may count the pseudo-parties of a CCS context (parties * (parties + 2) sets of n key elements): up to 2^15 elements
        if (p->k < 1 || p->n < 1 || (long long)p->n * p->k > 32768) return fail(nullptr, MKTFHE_EINVAL, "need 1 <= n*k <= 32768");
        if (p->n + 1 > mk::KS_THREADS * mk::KS_MAXCOLS) return fail(nullptr, MKTFHE_EINVAL, "n too large for the key-switch kernel");
        if (p->t < 1 || p->t * p->basebit > 31) return fail(nullptr, MKTFHE_EINVAL, "need t*basebit <= 31");
        return MKTFHE_OK;
    }
    if (p->l < 1 || p->l > 4) return fail(nullptr, MKTFHE_EINVAL, "unsupported gsw_decomp_length l=%d (1..4)", p->l);
    if (p->bgbit < 1 || p->l * p->bgbit > 32) return fail(nullptr, MKTFHE_EINVAL, "unsupported l*bgbit=%d (<=32)", p->l * p->bgbit);
    if ((1 << p->bgbit) > mk::LUT_BYTES_MAX)
        return fail(nullptr, MKTFHE_EINVAL, "unsupported gsw_log2_base=%d (digit lookup table holds %d values)", p->bgbit, mk::LUT_BYTES_MAX);
    // exactness of the three-prime CRT: |sum| <= 2l * N * 2^(bgbit-1) * 2^63 must stay below M/4
    if (std::log2((double)(2 * p->l) * p->N) + (p->bgbit - 1) + 63.0 >= rns::log2_crt_bound())
        return fail(nullptr, MKTFHE_EINVAL, "2l*N*Bg/2*2^63 exceeds the CRT range of the three-prime exact product");
    if (p->k < 1 || p->n < 1 || (long long)p->n * p->k > 8192) return fail(nullptr, MKTFHE_EINVAL, "need 1 <= n*k <= 8192");
    if (p->n + 1 > mk::KS_THREADS * mk::KS_MAXCOLS) return fail(nullptr, MKTFHE_EINVAL, "n too large for the key-switch kernel");
    if (p->t < 1 || p->basebit < 1 || p->t * p->basebit > 31) return fail(nullptr, MKTFHE_EINVAL, "need t*basebit <= 31");
    return MKTFHE_OK;
}

#ifndef MK_EXTRA_SMEM
#define MK_EXTRA_SMEM 0     // diagnostic: unused shared memory to move the L1 carve-out
#endif
size_t br_smem_bytes(const mktfhe_ctx* c) { return mk::cta_smem_bytes(c->prm.l, c->t32) + MK_EXTRA_SMEM; }

// (L, GPC) instantiations: mk::gpc_for(l) gates per CTA
#define MK_DISPATCH_L(c, KERNEL, ...)                                      \
    switch ((c)->prm.l) {                                                  \
    case 1: KERNEL(1, mk::gpc_for(1), __VA_ARGS__); break;                 \
    case 2: KERNEL(2, mk::gpc_for(2), __VA_ARGS__); break;                 \
    case 3: KERNEL(3, mk::gpc_for(3), __VA_ARGS__); break;                 \
    default: KERNEL(4, mk::gpc_for(4), __VA_ARGS__); break;                \
    }

// Torus32-mode instantiations: l = 2, 3, 4
#define MK_DISPATCH_T32(c, KERNEL, ...)                                          \
    switch ((c)->prm.l) {                                                        \
    case 2: KERNEL(2, mk::gpc_for(2, true), __VA_ARGS__); break;                 \
    case 3: KERNEL(3, mk::gpc_for(3, true), __VA_ARGS__); break;                 \
    default: KERNEL(4, mk::gpc_for(4, true), __VA_ARGS__); break;                \
    }

// FFT-channel instantiations: mkf::gpc_for(l) gates per CTA
#define MK_DISPATCH_FFT(c, KERNEL, ...)                                     \
    switch ((c)->prm.l) {                                                   \
    case 1: KERNEL(1, mkf::gpc_for(1), __VA_ARGS__); break;                 \
    case 2: KERNEL(2, mkf::gpc_for(2), __VA_ARGS__); break;                 \
    case 3: KERNEL(3, mkf::gpc_for(3), __VA_ARGS__); break;                 \
    default: KERNEL(4, mkf::gpc_for(4), __VA_ARGS__); break;                \
    }
// Torus32-mode FFT instantiations: l = 2, 3, 4
#define MK_DISPATCH_FFT_T32(c, KERNEL, ...)                                       \
    switch ((c)->prm.l) {                                                         \
    case 2: KERNEL(2, mkf::gpc_for(2, true), __VA_ARGS__); break;                 \
    case 3: KERNEL(3, mkf::gpc_for(3, true), __VA_ARGS__); break;                 \
    default: KERNEL(4, mkf::gpc_for(4, true), __VA_ARGS__); break;                \
    }
size_t fft_smem_bytes(const mktfhe_ctx* c) { return mkf::cta_bytes(c->prm.l, mkf::gpc_for(c->prm.l, c->t32), c->t32); }
int fft_nl(const mktfhe_ctx* c) { return c->t32 ? mkf::LIMBS_T32 : mkf::LIMBS; }

int set_attrs(mktfhe_ctx* c) {
    if (c->prm.N == mk2k::N) {
        if (c->prm.l == 1) CU_TRY(c, cudaFuncSetAttribute(mk2k::blind_rotate2k_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mk2k::smem_bytes(1)));
        else CU_TRY(c, cudaFuncSetAttribute(mk2k::blind_rotate2k_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mk2k::smem_bytes(2)));
        if (c->prm.l == 1) CU_TRY(c, cudaFuncSetAttribute(mk2k::blind_rotate2k16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mk2k::smem_bytes16(1)));
        else CU_TRY(c, cudaFuncSetAttribute(mk2k::blind_rotate2k16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mk2k::smem_bytes16(2)));
        return MKTFHE_OK;
    }
    if (c->fft && c->t32) {
        const int smf = (int)fft_smem_bytes(c), smf1 = (int)mkf::cta_bytes(c->prm.l, 1, true);
#define SET_ATTR_FFT_T32(L, GPC, dummy)                                                                                            \
    CU_TRY(c, cudaFuncSetAttribute(mkf::blind_rotate_fft_t32_kernel<L, GPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf));   \
    CU_TRY(c, cudaFuncSetAttribute(mkf::blind_rotate_fft_t32_kernel<L, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf1));    \
    CU_TRY(c, cudaFuncSetAttribute(mkf::extprod_fft_kernel<L, GPC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf));      \
    CU_TRY(c, cudaFuncSetAttribute(mkf::extprod_fft_kernel<L, GPC, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf));
        MK_DISPATCH_FFT_T32(c, SET_ATTR_FFT_T32, 0)
#undef SET_ATTR_FFT_T32
        return MKTFHE_OK;
    }
    const int sm = (int)br_smem_bytes(c);
    if (c->t32) {
#define SET_ATTR_T32(L, GPC, dummy)                                                                                                \
    CU_TRY(c, cudaFuncSetAttribute(mk::blind_rotate_t32_kernel<L, GPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));         \
    CU_TRY(c, cudaFuncSetAttribute(mk::extprod_t32_kernel<L, GPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));         \
    CU_TRY(c, cudaFuncSetAttribute(mk::extprod_t32_kernel<L, GPC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        MK_DISPATCH_T32(c, SET_ATTR_T32, 0)
#undef SET_ATTR_T32
        return MKTFHE_OK;
    }
    if (c->fft) {
        const int smf = (int)fft_smem_bytes(c);
#define SET_ATTR_FFT(L, GPC, dummy)                                                                                                \
    CU_TRY(c, cudaFuncSetAttribute(mkf::blind_rotate_fft_kernel<L, GPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf));       \
    CU_TRY(c, cudaFuncSetAttribute(mkf::extprod_fft_kernel<L, GPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf));
        MK_DISPATCH_FFT(c, SET_ATTR_FFT, 0)
#undef SET_ATTR_FFT
        const int smf1 = (int)mkf::cta_bytes(c->prm.l, 1);
#define SET_ATTR_FFT1(L, GPC, dummy) CU_TRY(c, cudaFuncSetAttribute(mkf::blind_rotate_fft_kernel<L, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smf1));
        MK_DISPATCH_FFT(c, SET_ATTR_FFT1, 0)
#undef SET_ATTR_FFT1
    }
#define SET_ATTR(L, GPC, dummy)                                                                                                    \
    CU_TRY(c, cudaFuncSetAttribute(mk::blind_rotate_kernel<L, GPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));             \
    if (getenv("MKTFHE_B200_CARVEOUT"))                                                                                            \
        CU_TRY(c, cudaFuncSetAttribute(mk::blind_rotate_kernel<L, GPC>, cudaFuncAttributePreferredSharedMemoryCarveout,            \
                                       atoi(getenv("MKTFHE_B200_CARVEOUT"))));                                                     \
    CU_TRY(c, cudaFuncSetAttribute(mk::extprod_kernel<L, GPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    MK_DISPATCH_L(c, SET_ATTR, 0)
#undef SET_ATTR
    // small batches and tails: one gate per CTA (see launch_blind_rotate)
    const int l = c->prm.l;
    const int sm1 = (int)(mk::TW_SMEM_BYTES + mk::gate_smem_bytes(l));        // six warps, one gate per CTA (l = 1, or MKTFHE_B200_LATENCY=0)
#define SET_ATTR1(L, GPC, dummy) CU_TRY(c, cudaFuncSetAttribute(mk::blind_rotate_kernel<L, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm1));
    MK_DISPATCH_L(c, SET_ATTR1, 0)
#undef SET_ATTR1
    const int sml = (int)(mk::TW_SMEM_BYTES + mk::gate_smem_bytes(l, mk::lat_wpg(l)));
    if (l == 2) CU_TRY(c, cudaFuncSetAttribute(mk::blind_rotate_lat_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sml));
    if (l == 3) CU_TRY(c, cudaFuncSetAttribute(mk::blind_rotate_lat_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, sml));
    if (l == 4) CU_TRY(c, cudaFuncSetAttribute(mk::blind_rotate_lat_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sml));
    return MKTFHE_OK;
}

// one gate per CTA for gates [g0, g1): the latency kernel of 6 l warps per gate (l >= 2), else the six-warp kernel alone on its SM
void launch_one_gate_per_cta(mktfhe_ctx* c, mk::BlindRotateArgs a, size_t g0, size_t g1, cudaStream_t st) {
    a.g0 = (int)g0; a.G = (int)g1;
    const unsigned grid = (unsigned)(g1 - g0);
    const int l = c->prm.l;
    const size_t sml = mk::TW_SMEM_BYTES + mk::gate_smem_bytes(l, mk::lat_wpg(l));
    if (c->fft && c->fft_small) {
        const size_t smf1 = mkf::cta_bytes(l, 1);
#define LAUNCH_BR_FFT1(L, GPC, dummy) mkf::blind_rotate_fft_kernel<L, 1><<<grid, mkf::TPG, smf1, st>>>(a, c->d_bsk_fft, c->d_twF)
        MK_DISPATCH_FFT(c, LAUNCH_BR_FFT1, 0)
#undef LAUNCH_BR_FFT1
    } else if (l == 2 && c->latency_kernel) mk::blind_rotate_lat_kernel<2><<<grid, 32 * mk::lat_wpg(2), sml, st>>>(a);
    else if (l == 3 && c->latency_kernel) mk::blind_rotate_lat_kernel<3><<<grid, 32 * mk::lat_wpg(3), sml, st>>>(a);
    else if (l == 4 && c->latency_kernel) mk::blind_rotate_lat_kernel<4><<<grid, 32 * mk::lat_wpg(4), sml, st>>>(a);
    else {
        const size_t sm1 = mk::TW_SMEM_BYTES + mk::gate_smem_bytes(l);
#define LAUNCH_BR1(L, GPC, dummy) mk::blind_rotate_kernel<L, 1><<<grid, mk::TPG, sm1, st>>>(a)
        MK_DISPATCH_L(c, LAUNCH_BR1, 0)
#undef LAUNCH_BR1
    }
    c->launches++;
}

// Launch shapes.  Throughput: gpc gates per CTA, one CTA per SM, i.e. waves of gpc * num_sms gates.  A batch -- or the tail a batch
// leaves after its full waves -- of at most one gate per SM runs one gate per CTA instead, on the latency kernel (6 l warps per gate): at
// l = 2 a gate alone on an SM finishes in 7.3 ms against 12.3 ms for two gates sharing it.  Bit-identical results.
void launch_blind_rotate(mktfhe_ctx* c, const mk::BlindRotateArgs& a, size_t G, cudaStream_t st) {
    if (c->t32 && c->fft) {   // Torus32 mode on the FFT channel: two gates per CTA, or one for batches of at most one gate per SM
        mk::BlindRotateArgs h = a;
        h.g0 = 0; h.G = (int)G;
        if (G <= (size_t)c->num_sms) {
            const size_t smf1 = mkf::cta_bytes(c->prm.l, 1, true);
#define LAUNCH_BR_FFT_T32_1(L, GPC, dummy) mkf::blind_rotate_fft_t32_kernel<L, 1><<<(unsigned)G, mkf::TPG, smf1, st>>>(h, c->d_bsk_fft, c->d_twF)
            MK_DISPATCH_FFT_T32(c, LAUNCH_BR_FFT_T32_1, 0)
#undef LAUNCH_BR_FFT_T32_1
        } else {
            const size_t smf = fft_smem_bytes(c);
            const int gf = mkf::gpc_for(c->prm.l, true);
            const unsigned gridf = (unsigned)((G + gf - 1) / gf);
#define LAUNCH_BR_FFT_T32(L, GPC, dummy) mkf::blind_rotate_fft_t32_kernel<L, GPC><<<gridf, GPC * mkf::TPG, smf, st>>>(h, c->d_bsk_fft, c->d_twF)
            MK_DISPATCH_FFT_T32(c, LAUNCH_BR_FFT_T32, 0)
#undef LAUNCH_BR_FFT_T32
        }
        c->launches++;
        return;
    }
    if (c->t32) {   // Torus32 mode, RNS kernels: one launch shape (two six-warp gates per CTA)
        mk::BlindRotateArgs h = a;
        h.g0 = 0; h.G = (int)G;
        const size_t sm = br_smem_bytes(c);
        const unsigned grid = (unsigned)((G + c->gpc - 1) / c->gpc);
#define LAUNCH_BR_T32(L, GPC, dummy) mk::blind_rotate_t32_kernel<L, GPC><<<grid, GPC * mk::TPG, sm, st>>>(h)
        MK_DISPATCH_T32(c, LAUNCH_BR_T32, 0)
#undef LAUNCH_BR_T32
        c->launches++;
        return;
    }
    const int gpc = c->fft ? mkf::gpc_for(c->prm.l) : c->gpc;
    const size_t wave = (size_t)gpc * (size_t)c->num_sms;
    size_t tail = gpc > 1 ? G % wave : 0;
    // the tail of a larger batch is split off only at l = 2, where it was measured to pay on two workloads; at l = 3 the split lost 1.2 %
    // (the partial last wave of the throughput grid cost 3 ms there, not a full wave), profiles/ab_r1.txt
    // (RNS kernels).  MKTFHE_B200_SPLIT_TAIL=2 splits at every l (A/B knob for the FFT kernels)
    if (tail > (size_t)c->num_sms || ((!c->split_tail || (c->prm.l != 2 && !c->split_tail_all)) && tail != G)) tail = 0;
    const size_t head = G - tail;
    if (head) {
        mk::BlindRotateArgs h = a;
        h.g0 = 0; h.G = (int)head;
        const unsigned grid = (unsigned)((head + gpc - 1) / gpc);
        if (c->fft) {   // throughput launch on the FP64 FFT channel
            const size_t smf = fft_smem_bytes(c);
#define LAUNCH_BR_FFT(L, GPC, dummy) mkf::blind_rotate_fft_kernel<L, GPC><<<grid, GPC * mkf::TPG, smf, st>>>(h, c->d_bsk_fft, c->d_twF)
            MK_DISPATCH_FFT(c, LAUNCH_BR_FFT, 0)
#undef LAUNCH_BR_FFT
        } else {
            const size_t sm = br_smem_bytes(c);
#define LAUNCH_BR(L, GPC, dummy) mk::blind_rotate_kernel<L, GPC><<<grid, GPC * mk::TPG, sm, st>>>(h)
            MK_DISPATCH_L(c, LAUNCH_BR, 0)
#undef LAUNCH_BR
        }
        c->launches++;
    }
    if (tail) launch_one_gate_per_cta(c, a, head, G, st);
}

mk::GateLinear gate_linear(int gate, bool* ok) {
    *ok = gate >= MKTFHE_GATE_NAND && gate <= MKTFHE_GATE_AND3;
    return mk::gate_linear(gate);
}

// device-pointer core shared by every bootstrap-like entry point.  `st` = the caller's stream or nullptr for the context's own.
// Per-context scratch (c->ext when the key switch is a separate launch, the timing events) is shared by all calls: calls on one
// context must be stream-ordered with each other (include/mktfhe_b200.h, conventions).
int run_bootstrap_dev(mktfhe_ctx* c, mk::GateLinear lin, int64_t mu, size_t G, const int32_t* xa, const int32_t* xb,
                      const int32_t* ya, const int32_t* yb, const int32_t* za, const int32_t* zb, int32_t* oa, int32_t* ob,
                      int32_t* ext_out, int64_t* acc_out, bool do_keyswitch, cudaStream_t st, const int32_t* gate_ids = nullptr) {
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (G == 0) return MKTFHE_OK;
    if (G > 0x7fffffffu) return fail(c, MKTFHE_EINVAL, "batch too large");
    if (!st) st = c->stream;
    const bool big = c->prm.N == mk2k::N;
    const bool fuse = do_keyswitch && c->fuse_ks && (big ? mk2k::ks_fusable(c->prm.n, c->prm.t) : mk::ks_fusable(c->prm.n, c->prm.t));
    c->last_fused = fuse;
    int32_t* ext = ext_out;
    if (!ext && !fuse) {
        int rc = reserve(c, c->ext, G * ((size_t)c->prm.N + 1) * sizeof(int32_t));
        if (rc) return rc;
        ext = (int32_t*)c->ext.p;
    }
    c->ev_valid = false;
    {
        mktfhe_ctx* root = c->parent ? c->parent : c;
        if (!root->in_shard_call) root->seq++;
        c->ev_seq = root->seq;
    }
    CU_TRY(c, cudaEventRecord(c->ev[0], st));
    if (big) {
        // N = 2048: one gate per CTA (kernels2k.cuh); the key switch is the kernel's epilogue when its shape is covered
        mk2k::Args a{};
        a.G = (int)G; a.n = c->prm.n; a.k = c->prm.k; a.bgbit = c->prm.bgbit;
        a.bsk = c->d_bsk; a.twB = c->d_twB;
        a.xa = xa; a.xb = xb; a.ya = ya; a.yb = yb; a.za = za; a.zb = zb;
        a.lin = lin; a.gate_ids = gate_ids; a.mu = mu; a.ext_out = ext; a.acc_out = acc_out;
        if (fuse) { a.ksk = c->d_ksk; a.ks_t = c->prm.t; a.ks_basebit = c->prm.basebit; a.oa = oa; a.ob = ob; }
        if (c->two_k16) {
            if (c->prm.l == 1) mk2k::blind_rotate2k16_kernel<1><<<(unsigned)G, mk2k::THREADS16, mk2k::smem_bytes16(1), st>>>(a);
            else mk2k::blind_rotate2k16_kernel<2><<<(unsigned)G, mk2k::THREADS16, mk2k::smem_bytes16(2), st>>>(a);
        } else {
            if (c->prm.l == 1) mk2k::blind_rotate2k_kernel<1><<<(unsigned)G, mk2k::THREADS, mk2k::smem_bytes(1), st>>>(a);
            else mk2k::blind_rotate2k_kernel<2><<<(unsigned)G, mk2k::THREADS, mk2k::smem_bytes(2), st>>>(a);
        }
        c->launches++;
    } else {
        mk::BlindRotateArgs a{};
        a.G = (int)G; a.n = c->prm.n; a.k = c->prm.k; a.bgbit = c->prm.bgbit;
        a.bsk = c->d_bsk; a.twB = c->d_twB;
        a.xa = xa; a.xb = xb; a.ya = ya; a.yb = yb; a.za = za; a.zb = zb;
        a.lin = lin; a.gate_ids = gate_ids; a.mu = mu; a.ext_out = ext; a.acc_out = acc_out;
        if (fuse) {
            a.ksk = c->d_ksk; a.ks_t = c->prm.t; a.ks_basebit = c->prm.basebit; a.oa = oa; a.ob = ob;
            if (!ext_out) a.ext_out = nullptr;   // nobody asked for the extracted samples
        }
        launch_blind_rotate(c, a, G, st);
    }
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaEventRecord(c->ev[1], st));
    CU_TRY(c, cudaEventRecord(c->ev[2], st));
    if (do_keyswitch && !fuse) {
        if (big) mk2k::keyswitch2k_kernel<<<(unsigned)G, mk::KS_THREADS, 0, st>>>(c->prm.n, c->prm.k, c->prm.t, c->prm.basebit, c->d_ksk, ext, oa, ob);
        else mk::keyswitch_kernel<<<(unsigned)G, mk::KS_THREADS, 0, st>>>(c->prm.n, c->prm.k, c->prm.t, c->prm.basebit, c->d_ksk, ext, oa, ob);
        c->launches++;
        CU_TRY(c, cudaGetLastError());
    }
    CU_TRY(c, cudaEventRecord(c->ev[3], st));
    c->ev_valid = true;
    return MKTFHE_OK;
}

// ---- multi-device contexts (mktfhe_create_multi) ---------------------------------------------------------------------------
// gates shard across GPUs, every GPU holds a full key replica, the one-time key broadcast is the only
// exchange.  The context the caller holds is replica 0; c->kids are single-device contexts on the further GPUs.
inline size_t n_replicas(const mktfhe_ctx* c) { return 1 + c->kids.size(); }
inline mktfhe_ctx* replica(mktfhe_ctx* c, size_t i) { return i == 0 ? c : c->kids[i - 1]; }
// contiguous slice [lo, hi) of G units owned by replica r of R; sizes differ by at most one
inline void shard_bounds(size_t G, size_t R, size_t r, size_t* lo, size_t* hi) {
    const size_t base = G / R, rem = G % R;
    *lo = r * base + (r < rem ? r : rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}
// f(replica, lo, hi) for every replica with a non-empty slice, one host thread per replica so that their host<->device copies
// (pinned or pageable caller memory alike), launches and waits overlap.  First failing replica's code and message win.
template <class F>
int for_each_shard(mktfhe_ctx* c, size_t G, F f) {
    const size_t R = n_replicas(c);
    std::vector<int> rc(R, MKTFHE_OK);
    std::vector<std::thread> th;
    th.reserve(R);
    c->seq++;
    c->in_shard_call = true;
    for (size_t r = 1; r < R; r++) {
        size_t lo, hi;
        shard_bounds(G, R, r, &lo, &hi);
        if (hi == lo) continue;
        try {
            th.emplace_back([&rc, &f, c, r, lo, hi]() { rc[r] = f(replica(c, r), lo, hi); });
        } catch (...) {
            rc[r] = f(replica(c, r), lo, hi);   // no thread to be had: run the slice here
        }
    }
    size_t lo, hi;
    shard_bounds(G, R, 0, &lo, &hi);
    if (hi > lo) rc[0] = f(c, lo, hi);
    for (auto& t : th) t.join();
    c->in_shard_call = false;
    for (size_t r = 0; r < R; r++)
        if (rc[r]) {
            if (r) c->err = "device " + std::to_string(replica(c, r)->device) + ": " + replica(c, r)->err;
            return rc[r];
        }
    return MKTFHE_OK;
}

// NCCL, bound at run time (no link-time dependency; torch ships its own libnccl.so.2 and a process must not hold two)
struct NcclApi {
    void* h = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (h) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (h) break;
        }
        if (!h) return false;
#define NCCL_SYM(field, sym) *(void**)(&field) = dlsym(h, sym)
        NCCL_SYM(CommInitAll, "ncclCommInitAll"); NCCL_SYM(CommDestroy, "ncclCommDestroy"); NCCL_SYM(GroupStart, "ncclGroupStart");
        NCCL_SYM(GroupEnd, "ncclGroupEnd"); NCCL_SYM(Broadcast, "ncclBroadcast"); NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast && GetErrorString;
    }
};

// one grouped ncclBroadcast per key buffer from replica 0 (single process, one communicator per device)
int broadcast_keys_nccl(mktfhe_ctx* c) {
    static NcclApi api;
    if (!api.load()) return fail(c, MKTFHE_ECUDA, "MKTFHE_B200_BCAST=nccl: libnccl.so.2 not loadable (%s)", dlerror());
    const size_t R = n_replicas(c);
    std::vector<int> devs(R);
    for (size_t r = 0; r < R; r++) devs[r] = replica(c, r)->device;
    std::vector<void*> comms(R, nullptr);
    int e = api.CommInitAll(comms.data(), (int)R, devs.data());
    if (e) return fail(c, MKTFHE_ECUDA, "ncclCommInitAll: %s", api.GetErrorString(e));
    const size_t kbytes = c->ksk_bytes;
    int bad = 0;
    for (int which = 0; which < 2 && !bad; which++) {
        api.GroupStart();
        for (size_t r = 0; r < R; r++) {
            mktfhe_ctx* d = replica(c, r);
            void* buf = which == 0 ? (void*)d->d_bsk : (void*)d->d_ksk;
            e = api.Broadcast(buf, buf, which == 0 ? c->bsk_bytes : kbytes, /* ncclInt8 */ 0, 0, comms[r], d->stream);
            if (e) bad = e;
        }
        e = api.GroupEnd();
        if (e) bad = e;
    }
    for (size_t r = 0; r < R; r++) {
        cudaSetDevice(replica(c, r)->device);
        cudaStreamSynchronize(replica(c, r)->stream);
    }
    for (void* cm : comms) if (cm) api.CommDestroy(cm);
    if (bad) return fail(c, MKTFHE_ECUDA, "ncclBroadcast: %s", api.GetErrorString(bad));
    c->bcast_how = "nccl";
    return MKTFHE_OK;
}

// binomial tree of peer copies: in the round of span s the replicas [0, s) each send both key buffers to replica + s
// (cudaMemcpyPeerAsync: NVLink when peer access is on, staged through the host otherwise)
int broadcast_keys_p2p(mktfhe_ctx* c) {
    const size_t R = n_replicas(c);
    for (size_t span = 1; span < R; span <<= 1) {
        for (size_t src = 0; src < span && src + span < R; src++) {
            mktfhe_ctx *s = replica(c, src), *d = replica(c, src + span);
            CU_TRY(c, cudaSetDevice(d->device));
            CU_TRY(c, cudaMemcpyPeerAsync(d->d_bsk, d->device, s->d_bsk, s->device, c->bsk_bytes, d->stream));
            CU_TRY(c, cudaMemcpyPeerAsync(d->d_ksk, d->device, s->d_ksk, s->device, c->ksk_bytes, d->stream));
        }
        for (size_t src = 0; src < span && src + span < R; src++) {
            mktfhe_ctx* d = replica(c, src + span);
            CU_TRY(c, cudaSetDevice(d->device));
            CU_TRY(c, cudaStreamSynchronize(d->stream));
        }
    }
    c->bcast_how = "p2p";
    return MKTFHE_OK;
}

int broadcast_keys(mktfhe_ctx* c) {
    const char* how = getenv("MKTFHE_B200_BCAST");
    bool distinct = true;
    for (size_t i = 0; i < n_replicas(c); i++)
        for (size_t j = 0; j < i; j++) distinct &= replica(c, i)->device != replica(c, j)->device;
    int rc = (how && !strcmp(how, "nccl") && distinct) ? broadcast_keys_nccl(c) : broadcast_keys_p2p(c);
    if (rc) return rc;
    for (mktfhe_ctx* k : c->kids) {
        k->bsk_loaded.assign(k->prm.k, 1);
        k->ksk_loaded.assign(k->prm.k, 1);
        k->ready = true;
    }
    CU_TRY(c, cudaSetDevice(c->device));
    return MKTFHE_OK;
}

}  // namespace

extern "C" {

int mktfhe_create(const mktfhe_params* params, int device, mktfhe_ctx** out) {
    if (!out) return fail(nullptr, MKTFHE_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = check_params(params);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MKTFHE_ECUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, MKTFHE_EINVAL, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) return fail(nullptr, MKTFHE_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    mktfhe_ctx* c = new (std::nothrow) mktfhe_ctx();
    if (!c) return fail(nullptr, MKTFHE_ENOMEM, "out of host memory");
    c->prm = *params;
    c->device = device;
    c->bsk_loaded.assign(params->k, 0);
    c->ksk_loaded.assign(params->k, 0);
#define CREATE_TRY(call)                                                                                  \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            fail(nullptr, e_ == cudaErrorMemoryAllocation ? MKTFHE_ENOMEM : MKTFHE_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
            mktfhe_destroy(c);                                                                            \
            return e_ == cudaErrorMemoryAllocation ? MKTFHE_ENOMEM : MKTFHE_ECUDA;                        \
        }                                                                                                 \
    } while (0)
    CREATE_TRY(cudaSetDevice(device));
    CREATE_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto& ev : c->ev) CREATE_TRY(cudaEventCreate(&ev));
    const int B1 = (1 << params->basebit) - 1;
    const bool big = params->N == mk2k::N;
    c->bsk_bytes = (size_t)params->k * params->n * (big ? mk2k::bsk_elem_words(params->l) : mk::bsk_elem_words(params->l)) * sizeof(u32);
    c->t32 = !big && (params->reserved & MKTFHE_FLAG_TORUS32);
    c->bsk_ntt_bytes = c->bsk_bytes;
    // FP64 FFT channel: N = 1024, Torus64 keys, byte digits; exact while 2 l N (Bg / 2) 2^21 <= 2^40 (limb products recovered by rounding)
    //                   Torus32 mode: two 16-bit limbs of the unshifted 32-bit keys, 2 l N (Bg / 2) 2^15 <= 2^40
    c->fft = !big && std::log2((double)(2 * params->l) * params->N) + (params->bgbit - 1) + (c->t32 ? 15.0 : 21.0) <= 40.0;
    if (const char* e = getenv("MKTFHE_B200_FFT")) c->fft = c->fft && atoi(e) != 0;
    if (const char* e = getenv("MKTFHE_B200_FFT_SMALL")) c->fft_small = atoi(e) != 0;
    if (c->fft && (c->fft_small || c->t32)) c->bsk_ntt_bytes = c->bsk_bytes = 0;
    if (c->fft) c->bsk_bytes += (size_t)params->k * params->n * mkf::bsk_elem_cpx(params->l, fft_nl(c)) * sizeof(mkf::cpx);
    c->gpc = big ? 1 : mk::gpc_for(params->l, c->t32);
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device);
    if (const char* e = getenv("MKTFHE_B200_LATENCY")) c->latency_kernel = atoi(e) != 0;
    if (const char* e = getenv("MKTFHE_B200_SPLIT_TAIL")) { c->split_tail = atoi(e) != 0; c->split_tail_all = atoi(e) == 2; }
    if (const char* e = getenv("MKTFHE_B200_FUSE_KS")) c->fuse_ks = atoi(e) != 0;
    if (const char* e = getenv("MKTFHE_B200_2K")) c->two_k16 = atoi(e) != 8;
    const size_t ks_stride = mk::ks_row_stride(params->n);   // rows padded to 16 bytes on the device
    c->ksk_bytes = (size_t)params->k * params->N * params->t * B1 * ks_stride * sizeof(int32_t);
    CREATE_TRY(cudaMalloc(&c->d_bsk, c->bsk_bytes));
    if (c->fft) {
        c->d_bsk_fft = reinterpret_cast<mkf::cpx*>(reinterpret_cast<char*>(c->d_bsk) + c->bsk_ntt_bytes);
        mkf::HostTablesFFT TF;
        CREATE_TRY(cudaMalloc(&c->d_twF, (size_t)mkf::TW_BYTES));
        CREATE_TRY(cudaMemcpy(c->d_twF, TF.tw.data(), (size_t)mkf::TW_BYTES, cudaMemcpyHostToDevice));
    }
    // the key-switching key is followed by one all-zero row (read by the fused key switch for zero digits)
    CREATE_TRY(cudaMalloc(&c->d_ksk, c->ksk_bytes + ks_stride * sizeof(int32_t)));
    CREATE_TRY(cudaMemset(c->d_ksk, 0, c->ksk_bytes + ks_stride * sizeof(int32_t)));
    if (big) {
        rns2k::HostTables T2;
        CREATE_TRY(cudaMalloc(&c->d_twB, T2.twB.size() * sizeof(rns::uint2_)));
        CREATE_TRY(cudaMemcpyToSymbol(mk2k::c_k2, &T2.c, sizeof(rns2k::Consts)));
        CREATE_TRY(cudaMemcpy(c->d_twB, T2.twB.data(), T2.twB.size() * sizeof(rns::uint2_), cudaMemcpyHostToDevice));
    } else {
        CREATE_TRY(cudaMalloc(&c->d_twB, (size_t)mk::TWB_WORDS * 4));
        rns::HostTables T;
        CREATE_TRY(cudaMemcpyToSymbol(mk::c_rns, &T.c, sizeof(rns::Consts)));
        CREATE_TRY(cudaMemcpy(c->d_twB, T.twB.data(), (size_t)mk::TWB_WORDS * 4, cudaMemcpyHostToDevice));
    }
#undef CREATE_TRY
    rc = set_attrs(c);
    if (rc) { g_create_error = c->err; mktfhe_destroy(c); return rc; }
    *out = c;
    return MKTFHE_OK;
}

void mktfhe_destroy(mktfhe_ctx* c) {
    if (!c) return;
    for (mktfhe_ctx* k : c->kids) mktfhe_destroy(k);
    c->kids.clear();
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->in[0], &c->in[1], &c->in[2], &c->in[3], &c->in[4], &c->in[5], &c->ext, &c->oa, &c->ob, &c->accin, &c->accout, &c->elem, &c->raw, &c->gids};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    if (c->d_bsk) cudaFree(c->d_bsk);
    if (c->d_ksk) cudaFree(c->d_ksk);
    if (c->d_twB) cudaFree(c->d_twB);
    if (c->d_twF) cudaFree(c->d_twF);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* mktfhe_last_error(const mktfhe_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int mktfhe_load_bsk(mktfhe_ctx* c, int party, const int64_t* polys) {
    if (!c) return MKTFHE_EINVAL;
    if (party < 0 || party >= c->prm.k || !polys) return fail(c, MKTFHE_EINVAL, "load_bsk: bad party %d or NULL key", party);
    CU_TRY(c, cudaSetDevice(c->device));
    const int n = c->prm.n, l = c->prm.l;
    const size_t raw_bytes = (size_t)n * 4 * l * c->prm.N * sizeof(int64_t);
    int rc = reserve(c, c->raw, raw_bytes);
    if (rc) return rc;
    CU_TRY(c, cudaMemcpyAsync(c->raw.p, polys, raw_bytes, cudaMemcpyHostToDevice, c->stream));
    if (c->prm.N == mk2k::N) {
        const int ntasks = n * 4 * l * rns2k::NP;
        mk2k::bsk_transform2k_kernel<<<(ntasks + mk2k::XF_WARPS - 1) / mk2k::XF_WARPS, mk2k::XF_WARPS * 32, 0, c->stream>>>(
            (const int64_t*)c->raw.p, c->d_bsk, n, l, party, c->d_twB, ntasks);
    } else {
        const int ntasks = n * 4 * l * rns::NP;
        if (c->bsk_ntt_bytes)
            mk::bsk_transform_kernel<<<(ntasks + mk::XF_WARPS - 1) / mk::XF_WARPS, mk::XF_WARPS * 32, 0, c->stream>>>(
                (const int64_t*)c->raw.p, c->d_bsk, n, l, party, c->d_twB, ntasks);
        if (c->fft) {
            const int nt = n * 4 * l * fft_nl(c);
            mkf::bsk_transform_fft_kernel<<<(nt + mkf::XF_WARPS - 1) / mkf::XF_WARPS, mkf::XF_WARPS * 32, 0, c->stream>>>(
                (const int64_t*)c->raw.p, c->d_bsk_fft, n, l, fft_nl(c), party, c->d_twF, nt);
            c->launches++;
        }
    }
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->bsk_loaded[party] = 1;
    c->ready = false;
    return MKTFHE_OK;
}

int mktfhe_load_ksk(mktfhe_ctx* c, int party, const int32_t* rows) {
    if (!c) return MKTFHE_EINVAL;
    if (party < 0 || party >= c->prm.k || !rows) return fail(c, MKTFHE_EINVAL, "load_ksk: bad party %d or NULL key", party);
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t per = c->ksk_bytes / c->prm.k, nrows = per / (mk::ks_row_stride(c->prm.n) * sizeof(int32_t));
    const size_t src_pitch = (size_t)(c->prm.n + 1) * sizeof(int32_t), dst_pitch = (size_t)mk::ks_row_stride(c->prm.n) * sizeof(int32_t);
    CU_TRY(c, cudaMemcpy2DAsync((char*)c->d_ksk + per * party, dst_pitch, rows, src_pitch, src_pitch, nrows, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->ksk_loaded[party] = 1;
    c->ready = false;
    return MKTFHE_OK;
}

int mktfhe_generate_ksk(mktfhe_ctx* c, int party, const int32_t* lwe_key, const int64_t* rlwe_key, double sigma, uint64_t seed) {
    if (!c) return MKTFHE_EINVAL;
    if (party < 0 || party >= c->prm.k || !lwe_key || !rlwe_key) return fail(c, MKTFHE_EINVAL, "generate_ksk: bad party %d or NULL key", party);
    if (!(sigma >= 0.0) || sigma > 0.25) return fail(c, MKTFHE_EINVAL, "generate_ksk: noise standard deviation %g out of range", sigma);
    CU_TRY(c, cudaSetDevice(c->device));
    const int n = c->prm.n, N = c->prm.N, t = c->prm.t, bb = c->prm.basebit, B1 = (1 << bb) - 1;
    const size_t rows = (size_t)N * t * B1, per = c->ksk_bytes / c->prm.k;
    // scratch: [rows] noise doubles, the sum, then the two secret vectors (erased below)
    const size_t off_sum = rows * 8, off_s = off_sum + 8, off_z = (off_s + (size_t)n * 4 + 7) & ~(size_t)7, total = off_z + (size_t)N * 8;
    int rc = reserve(c, c->raw, total);
    if (rc) return rc;
    char* base = (char*)c->raw.p;
    CU_TRY(c, cudaMemsetAsync(base + off_sum, 0, 8, c->stream));
    CU_TRY(c, cudaMemcpyAsync(base + off_s, lwe_key, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(base + off_z, rlwe_key, (size_t)N * 8, cudaMemcpyHostToDevice, c->stream));
    const uint2 key = make_uint2((uint32_t)seed ^ (uint32_t)(0x9E3779B9u * (uint32_t)(party + 1)), (uint32_t)(seed >> 32));
    mk::ksk_noise_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, c->stream>>>((double*)base, (double*)(base + off_sum), rows, sigma, key);
    CU_TRY(c, cudaGetLastError());
    mk::ksk_generate_kernel<<<(unsigned)((rows + mk::KG_WARPS - 1) / mk::KG_WARPS), 32 * mk::KG_WARPS, 0, c->stream>>>(
        (int32_t*)((char*)c->d_ksk + per * party), (const double*)base, (const double*)(base + off_sum), (const int32_t*)(base + off_s),
        (const int64_t*)(base + off_z), n, t, bb, rows, key);
    CU_TRY(c, cudaGetLastError());
    c->launches += 2;
    CU_TRY(c, cudaMemsetAsync(base + off_s, 0, total - off_s, c->stream));      // the secrets do not stay in device scratch
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->ksk_loaded[party] = 1;
    c->ready = false;
    return MKTFHE_OK;
}

int mktfhe_finalize_keys(mktfhe_ctx* c) {
    if (!c) return MKTFHE_EINVAL;
    for (int p = 0; p < c->prm.k; p++)
        if (!c->bsk_loaded[p] || !c->ksk_loaded[p]) return fail(c, MKTFHE_ESTATE, "party %d: bootstrapping or key-switching key not loaded", p);
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (!c->kids.empty()) {   // multi-device context: the one-time key broadcast from replica 0
        int rc = broadcast_keys(c);
        if (rc) return rc;
    }
    c->ready = true;
    return MKTFHE_OK;
}

int mktfhe_key_buffers(mktfhe_ctx* c, void** bsk_dev, size_t* bsk_bytes, void** ksk_dev, size_t* ksk_bytes) {
    if (!c) return MKTFHE_EINVAL;
    if (bsk_dev) *bsk_dev = c->d_bsk;
    if (bsk_bytes) *bsk_bytes = c->bsk_bytes;
    if (ksk_dev) *ksk_dev = c->d_ksk;
    if (ksk_bytes) *ksk_bytes = c->ksk_bytes;
    return MKTFHE_OK;
}

int mktfhe_mark_keys_received(mktfhe_ctx* c) {
    if (!c) return MKTFHE_EINVAL;
    c->bsk_loaded.assign(c->prm.k, 1);
    c->ksk_loaded.assign(c->prm.k, 1);
    return MKTFHE_OK;
}

int mktfhe_bootstrap_batch_dev(mktfhe_ctx* c, int64_t mu, size_t G, const int32_t* a_in, const int32_t* b_in, int32_t* a_out,
                               int32_t* b_out, void* stream) {
    if (!c) return MKTFHE_EINVAL;
    if (G && (!a_in || !b_in || !a_out || !b_out)) return fail(c, MKTFHE_EINVAL, "bootstrap_batch: NULL buffer");
    CU_TRY(c, cudaSetDevice(c->device));
    return run_bootstrap_dev(c, {0, 1, 0, 0}, mu, G, a_in, b_in, nullptr, nullptr, nullptr, nullptr, a_out, b_out, nullptr, nullptr, true,
                             (cudaStream_t)stream);
}

int mktfhe_gate_batch_dev(mktfhe_ctx* c, int gate, size_t G, const int32_t* xa, const int32_t* xb, const int32_t* ya,
                          const int32_t* yb, const int32_t* za, const int32_t* zb, int32_t* oa, int32_t* ob, void* stream) {
    if (!c) return MKTFHE_EINVAL;
    bool ok;
    mk::GateLinear lin = gate_linear(gate, &ok);
    if (!ok) return fail(c, MKTFHE_EINVAL, "unknown gate id %d", gate);
    if (G && (!xa || !xb || !ya || !yb || !oa || !ob || (lin.cz && (!za || !zb)))) return fail(c, MKTFHE_EINVAL, "gate_batch: NULL buffer");
    CU_TRY(c, cudaSetDevice(c->device));
    // output message encode_message64(1, 8) = 2^61 (rlwe_is32 == false, 3gen_mk_gates.jl:12)
    return run_bootstrap_dev(c, lin, (int64_t)1 << 61, G, xa, xb, ya, yb, za, zb, oa, ob, nullptr, nullptr, true, (cudaStream_t)stream);
}

static int stage_in(mktfhe_ctx* c, DevBuf& b, const void* host, size_t bytes) {
    int rc = reserve(c, b, bytes);
    if (rc) return rc;
    CU_TRY(c, cudaMemcpyAsync(b.p, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return MKTFHE_OK;
}

// host-pointer core of every "linear prologue + bootstrap" call: temp = mu0 + cx x + cy y + cz z, then mk_bootstrap_3gen(temp) with
// test-vector message mu.  y / z are staged only when their coefficient is non-zero.
static int affine_batch_1(mktfhe_ctx* c, mk::GateLinear lin, int64_t mu, size_t G, const int32_t* xa, const int32_t* xb, const int32_t* ya,
                          const int32_t* yb, const int32_t* za, const int32_t* zb, int32_t* oa, int32_t* ob, const char* what) {
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (G == 0) return MKTFHE_OK;
    if (!xa || !xb || !oa || !ob || (lin.cy && (!ya || !yb)) || (lin.cz && (!za || !zb))) return fail(c, MKTFHE_EINVAL, "%s: NULL buffer", what);
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t kn = (size_t)c->prm.n * c->prm.k, abytes = G * kn * 4, bbytes = G * 4;
    int rc;
    if ((rc = stage_in(c, c->in[0], xa, abytes)) || (rc = stage_in(c, c->in[1], xb, bbytes))) return rc;
    if (lin.cy && ((rc = stage_in(c, c->in[2], ya, abytes)) || (rc = stage_in(c, c->in[3], yb, bbytes)))) return rc;
    if (lin.cz && ((rc = stage_in(c, c->in[4], za, abytes)) || (rc = stage_in(c, c->in[5], zb, bbytes)))) return rc;
    if ((rc = reserve(c, c->oa, abytes)) || (rc = reserve(c, c->ob, bbytes))) return rc;
    rc = run_bootstrap_dev(c, lin, mu, G, (int32_t*)c->in[0].p, (int32_t*)c->in[1].p, (int32_t*)c->in[2].p, (int32_t*)c->in[3].p,
                           (int32_t*)c->in[4].p, (int32_t*)c->in[5].p, (int32_t*)c->oa.p, (int32_t*)c->ob.p, nullptr, nullptr, true, nullptr);
    if (rc) return rc;
    CU_TRY(c, cudaMemcpyAsync(oa, c->oa.p, abytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(ob, c->ob.p, bbytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return MKTFHE_OK;
}

static int mktfhe_gate_batch_1(mktfhe_ctx* c, int gate, size_t G, const int32_t* xa, const int32_t* xb, const int32_t* ya, const int32_t* yb,
                      const int32_t* za, const int32_t* zb, int32_t* oa, int32_t* ob) {
    if (!c) return MKTFHE_EINVAL;
    bool ok;
    mk::GateLinear lin = gate_linear(gate, &ok);
    if (!ok) return fail(c, MKTFHE_EINVAL, "unknown gate id %d", gate);
    // output message encode_message64(1, 8) = 2^61 (rlwe_is32 == false, 3gen_mk_gates.jl:12)
    return affine_batch_1(c, lin, (int64_t)1 << 61, G, xa, xb, ya, yb, za, zb, oa, ob, "gate_batch");
}

int mktfhe_gate_batch_mixed_dev(mktfhe_ctx* c, size_t G, const int32_t* gate_ids, const int32_t* xa, const int32_t* xb, const int32_t* ya,
                                const int32_t* yb, const int32_t* za, const int32_t* zb, int32_t* oa, int32_t* ob, void* stream) {
    if (!c) return MKTFHE_EINVAL;
    if (G && (!gate_ids || !xa || !xb || !ya || !yb || !oa || !ob)) return fail(c, MKTFHE_EINVAL, "gate_batch_mixed: NULL buffer");
    CU_TRY(c, cudaSetDevice(c->device));
    // za/zb are read only by MKTFHE_GATE_AND3 gates; point them at x when absent so no gate dereferences NULL
    return run_bootstrap_dev(c, {0, 0, 0, 0}, (int64_t)1 << 61, G, xa, xb, ya, yb, za ? za : xa, zb ? zb : xb, oa, ob, nullptr, nullptr, true,
                             (cudaStream_t)stream, gate_ids);
}

static int mktfhe_gate_batch_mixed_1(mktfhe_ctx* c, size_t G, const int32_t* gate_ids, const int32_t* xa, const int32_t* xb, const int32_t* ya,
                            const int32_t* yb, const int32_t* za, const int32_t* zb, int32_t* oa, int32_t* ob) {
    if (!c) return MKTFHE_EINVAL;
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (G == 0) return MKTFHE_OK;
    if (!gate_ids || !xa || !xb || !ya || !yb || !oa || !ob) return fail(c, MKTFHE_EINVAL, "gate_batch_mixed: NULL buffer");
    bool need_z = false;
    for (size_t g = 0; g < G; g++) {
        if (gate_ids[g] < MKTFHE_GATE_NAND || gate_ids[g] > MKTFHE_GATE_AND3) return fail(c, MKTFHE_EINVAL, "gate_ids[%zu] = %d is not a gate id", g, gate_ids[g]);
        need_z |= gate_ids[g] == MKTFHE_GATE_AND3;
    }
    if (need_z && (!za || !zb)) return fail(c, MKTFHE_EINVAL, "gate_batch_mixed: a 3AND gate needs z");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t kn = (size_t)c->prm.n * c->prm.k, abytes = G * kn * 4, bbytes = G * 4;
    int rc;
    if ((rc = stage_in(c, c->gids, gate_ids, bbytes)) || (rc = stage_in(c, c->in[0], xa, abytes)) || (rc = stage_in(c, c->in[1], xb, bbytes)) ||
        (rc = stage_in(c, c->in[2], ya, abytes)) || (rc = stage_in(c, c->in[3], yb, bbytes)))
        return rc;
    if (need_z && ((rc = stage_in(c, c->in[4], za, abytes)) || (rc = stage_in(c, c->in[5], zb, bbytes)))) return rc;
    if ((rc = reserve(c, c->oa, abytes)) || (rc = reserve(c, c->ob, bbytes))) return rc;
    rc = mktfhe_gate_batch_mixed_dev(c, G, (int32_t*)c->gids.p, (int32_t*)c->in[0].p, (int32_t*)c->in[1].p, (int32_t*)c->in[2].p, (int32_t*)c->in[3].p,
                                     need_z ? (int32_t*)c->in[4].p : nullptr, need_z ? (int32_t*)c->in[5].p : nullptr, (int32_t*)c->oa.p,
                                     (int32_t*)c->ob.p, nullptr);
    if (rc) return rc;
    CU_TRY(c, cudaMemcpyAsync(oa, c->oa.p, abytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(ob, c->ob.p, bbytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return MKTFHE_OK;
}

static int mktfhe_bootstrap_batch_1(mktfhe_ctx* c, int64_t mu, size_t G, const int32_t* a_in, const int32_t* b_in, int32_t* a_out, int32_t* b_out) {
    if (!c) return MKTFHE_EINVAL;
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (G == 0) return MKTFHE_OK;
    if (!a_in || !b_in || !a_out || !b_out) return fail(c, MKTFHE_EINVAL, "bootstrap_batch: NULL buffer");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t kn = (size_t)c->prm.n * c->prm.k, abytes = G * kn * 4, bbytes = G * 4;
    int rc;
    if ((rc = stage_in(c, c->in[0], a_in, abytes)) || (rc = stage_in(c, c->in[1], b_in, bbytes))) return rc;
    if ((rc = reserve(c, c->oa, abytes)) || (rc = reserve(c, c->ob, bbytes))) return rc;
    rc = run_bootstrap_dev(c, {0, 1, 0, 0}, mu, G, (int32_t*)c->in[0].p, (int32_t*)c->in[1].p, nullptr, nullptr, nullptr, nullptr,
                           (int32_t*)c->oa.p, (int32_t*)c->ob.p, nullptr, nullptr, true, nullptr);
    if (rc) return rc;
    CU_TRY(c, cudaMemcpyAsync(a_out, c->oa.p, abytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(b_out, c->ob.p, bbytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return MKTFHE_OK;
}

static int mktfhe_blind_rotate_batch_1(mktfhe_ctx* c, int64_t mu, size_t G, const int32_t* a_in, const int32_t* b_in, int32_t* ext_out, int64_t* acc_out) {
    if (!c) return MKTFHE_EINVAL;
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (G == 0) return MKTFHE_OK;
    if (!a_in || !b_in || !ext_out) return fail(c, MKTFHE_EINVAL, "blind_rotate_batch: NULL buffer");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t kn = (size_t)c->prm.n * c->prm.k, abytes = G * kn * 4, bbytes = G * 4;
    const size_t ebytes = G * ((size_t)c->prm.N + 1) * 4, accbytes = G * 2 * (size_t)c->prm.N * 8;
    int rc;
    if ((rc = stage_in(c, c->in[0], a_in, abytes)) || (rc = stage_in(c, c->in[1], b_in, bbytes))) return rc;
    if ((rc = reserve(c, c->ext, ebytes))) return rc;
    if (acc_out && (rc = reserve(c, c->accout, accbytes))) return rc;
    rc = run_bootstrap_dev(c, {0, 1, 0, 0}, mu, G, (int32_t*)c->in[0].p, (int32_t*)c->in[1].p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                           (int32_t*)c->ext.p, acc_out ? (int64_t*)c->accout.p : nullptr, false, nullptr);
    if (rc) return rc;
    CU_TRY(c, cudaMemcpyAsync(ext_out, c->ext.p, ebytes, cudaMemcpyDeviceToHost, c->stream));
    if (acc_out) CU_TRY(c, cudaMemcpyAsync(acc_out, c->accout.p, accbytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return MKTFHE_OK;
}

static int mktfhe_keyswitch_batch_1(mktfhe_ctx* c, size_t G, const int32_t* ext, int32_t* a_out, int32_t* b_out) {
    if (!c) return MKTFHE_EINVAL;
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (G == 0) return MKTFHE_OK;
    if (!ext || !a_out || !b_out) return fail(c, MKTFHE_EINVAL, "keyswitch_batch: NULL buffer");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t kn = (size_t)c->prm.n * c->prm.k, abytes = G * kn * 4, bbytes = G * 4, ebytes = G * ((size_t)c->prm.N + 1) * 4;
    int rc;
    if ((rc = stage_in(c, c->ext, ext, ebytes)) || (rc = reserve(c, c->oa, abytes)) || (rc = reserve(c, c->ob, bbytes))) return rc;
    if (c->prm.N == mk2k::N)
        mk2k::keyswitch2k_kernel<<<(unsigned)G, mk::KS_THREADS, 0, c->stream>>>(c->prm.n, c->prm.k, c->prm.t, c->prm.basebit, c->d_ksk,
                                                                              (const int32_t*)c->ext.p, (int32_t*)c->oa.p, (int32_t*)c->ob.p);
    else
        mk::keyswitch_kernel<<<(unsigned)G, mk::KS_THREADS, 0, c->stream>>>(c->prm.n, c->prm.k, c->prm.t, c->prm.basebit, c->d_ksk,
                                                                            (const int32_t*)c->ext.p, (int32_t*)c->oa.p, (int32_t*)c->ob.p);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(a_out, c->oa.p, abytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(b_out, c->ob.p, bbytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return MKTFHE_OK;
}

static int mktfhe_extprod_batch_1(mktfhe_ctx* c, size_t G, const int32_t* elem, const int64_t* acc_in, int64_t* acc_out) {
    if (!c) return MKTFHE_EINVAL;
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (G == 0) return MKTFHE_OK;
    if (!elem || !acc_in || !acc_out) return fail(c, MKTFHE_EINVAL, "extprod_batch: NULL buffer");
    if (c->prm.N == mk2k::N) return fail(c, MKTFHE_EINVAL, "extprod_batch: the single-external-product hook exists for N=1024 only");
    for (size_t g = 0; g < G; g++)
        if (elem[g] < 0 || elem[g] >= c->prm.n * c->prm.k) return fail(c, MKTFHE_EINVAL, "extprod_batch: elem[%zu]=%d out of range", g, elem[g]);
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t accbytes = G * 2 * mk::N * 8;
    int rc;
    if ((rc = stage_in(c, c->elem, elem, G * 4)) || (rc = stage_in(c, c->accin, acc_in, accbytes)) || (rc = reserve(c, c->accout, accbytes))) return rc;
    const size_t sm = br_smem_bytes(c);
    const unsigned grid = (unsigned)((G + c->gpc - 1) / c->gpc);
    if (c->t32 && c->fft) {
        const size_t smf = fft_smem_bytes(c);
        const int gf = mkf::gpc_for(c->prm.l, true);
        const unsigned gridf = (unsigned)((G + gf - 1) / gf);
#define LAUNCH_EP_FFT_T32(L, GPC, dummy)                                                                                                  \
    mkf::extprod_fft_kernel<L, GPC, true><<<gridf, GPC * mkf::TPG, smf, c->stream>>>((int)G, c->d_bsk_fft, c->d_twF, c->prm.bgbit, (const int32_t*)c->elem.p, \
                                                                                      (const int64_t*)c->accin.p, (int64_t*)c->accout.p)
        MK_DISPATCH_FFT_T32(c, LAUNCH_EP_FFT_T32, 0)
#undef LAUNCH_EP_FFT_T32
    } else if (c->t32) {
#define LAUNCH_EP_T32(L, GPC, dummy)                                                                                                   \
    mk::extprod_t32_kernel<L, GPC><<<grid, GPC * mk::TPG, sm, c->stream>>>((int)G, c->d_bsk, c->d_twB, c->prm.bgbit, (const int32_t*)c->elem.p, \
                                                                           (const int64_t*)c->accin.p, (int64_t*)c->accout.p)
        MK_DISPATCH_T32(c, LAUNCH_EP_T32, 0)
#undef LAUNCH_EP_T32
    } else if (c->fft) {
        const size_t smf = fft_smem_bytes(c);
        const int gf = mkf::gpc_for(c->prm.l);
        const unsigned gridf = (unsigned)((G + gf - 1) / gf);
#define LAUNCH_EP_FFT(L, GPC, dummy)                                                                                                      \
    mkf::extprod_fft_kernel<L, GPC><<<gridf, GPC * mkf::TPG, smf, c->stream>>>((int)G, c->d_bsk_fft, c->d_twF, c->prm.bgbit, (const int32_t*)c->elem.p, \
                                                                                (const int64_t*)c->accin.p, (int64_t*)c->accout.p)
        MK_DISPATCH_FFT(c, LAUNCH_EP_FFT, 0)
#undef LAUNCH_EP_FFT
    } else {
#define LAUNCH_EP(L, GPC, dummy)                                                                                          \
    mk::extprod_kernel<L, GPC><<<grid, GPC * mk::TPG, sm, c->stream>>>((int)G, c->d_bsk, c->d_twB, c->prm.bgbit, (const int32_t*)c->elem.p, \
                                                                       (const int64_t*)c->accin.p, (int64_t*)c->accout.p)
    MK_DISPATCH_L(c, LAUNCH_EP, 0)
#undef LAUNCH_EP
    }
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(acc_out, c->accout.p, accbytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return MKTFHE_OK;
}

static int mktfhe_negacyclic_mul_batch_1(mktfhe_ctx* c, size_t G, const int64_t* a, const int64_t* b, int64_t* out) {
    if (!c) return MKTFHE_EINVAL;
    if (G == 0) return MKTFHE_OK;
    if (!a || !b || !out) return fail(c, MKTFHE_EINVAL, "negacyclic_mul_batch: NULL buffer");
    {   // exactness of the CRT lift: N * max|a_i| * 2^63 must stay below M/4 (M ~ 2^84 at N = 1024, 2^112 at N = 2048)
        const int64_t lim = c->prm.N == mk2k::N ? ((int64_t)1 << 25) : ((int64_t)1 << 8);
        const size_t cnt = G * (size_t)c->prm.N;
        for (size_t i = 0; i < cnt; i++)
            if (a[i] > lim || a[i] < -lim)
                return fail(c, MKTFHE_EINVAL, "negacyclic_mul_batch: |a[%zu]| = %lld exceeds 2^%d, the exact range of the small operand at N = %d",
                            i, (long long)(a[i] < 0 ? -a[i] : a[i]), c->prm.N == mk2k::N ? 25 : 8, c->prm.N);
    }
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t bytes = G * (size_t)c->prm.N * 8;
    int rc;
    if ((rc = stage_in(c, c->accin, a, bytes)) || (rc = stage_in(c, c->raw, b, bytes)) || (rc = reserve(c, c->accout, bytes))) return rc;
    if (c->prm.N == mk2k::N)
        mk2k::negacyclic_mul2k_kernel<<<(unsigned)G, 32 * mk2k::NP, 0, c->stream>>>((const int64_t*)c->accin.p, (const int64_t*)c->raw.p,
                                                                                 (int64_t*)c->accout.p, c->d_twB);
    else
        mk::negacyclic_mul_kernel<<<(unsigned)G, mk::NM_THREADS, 0, c->stream>>>((const int64_t*)c->accin.p, (const int64_t*)c->raw.p,
                                                                          (int64_t*)c->accout.p, c->d_twB);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out, c->accout.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return MKTFHE_OK;
}

// ---- public host-pointer entry points: single-device contexts run the body above, multi-device contexts shard [G] ------------
#define MK_NOT_READY(c)                                                                       \
    if (!(c)) return MKTFHE_EINVAL;                                                           \
    if (!(c)->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)")

int mktfhe_gate_batch(mktfhe_ctx* c, int gate, size_t G, const int32_t* xa, const int32_t* xb, const int32_t* ya, const int32_t* yb,
                      const int32_t* za, const int32_t* zb, int32_t* oa, int32_t* ob) {
    if (!c) return MKTFHE_EINVAL;
    if (c->kids.empty()) return mktfhe_gate_batch_1(c, gate, G, xa, xb, ya, yb, za, zb, oa, ob);
    MK_NOT_READY(c);
    if (G && (!xa || !xb || !ya || !yb || !oa || !ob)) return fail(c, MKTFHE_EINVAL, "gate_batch: NULL buffer");
    const size_t kn = (size_t)c->prm.n * c->prm.k;
    return for_each_shard(c, G, [=](mktfhe_ctx* r, size_t lo, size_t hi) {
        return mktfhe_gate_batch_1(r, gate, hi - lo, xa + lo * kn, xb + lo, ya + lo * kn, yb + lo, za ? za + lo * kn : nullptr, zb ? zb + lo : nullptr,
                                   oa + lo * kn, ob + lo);
    });
}

int mktfhe_affine_bootstrap_batch(mktfhe_ctx* c, int32_t mu0, int32_t cx, int32_t cy, int32_t cz, int64_t mu, size_t G, const int32_t* xa,
                                  const int32_t* xb, const int32_t* ya, const int32_t* yb, const int32_t* za, const int32_t* zb, int32_t* oa,
                                  int32_t* ob) {
    if (!c) return MKTFHE_EINVAL;
    const mk::GateLinear lin{mu0, cx, cy, cz};
    if (c->kids.empty()) return affine_batch_1(c, lin, mu, G, xa, xb, ya, yb, za, zb, oa, ob, "affine_bootstrap_batch");
    MK_NOT_READY(c);
    if (G && (!xa || !xb || !oa || !ob)) return fail(c, MKTFHE_EINVAL, "affine_bootstrap_batch: NULL buffer");
    const size_t kn = (size_t)c->prm.n * c->prm.k;
    return for_each_shard(c, G, [=](mktfhe_ctx* r, size_t lo, size_t hi) {
        return affine_batch_1(r, lin, mu, hi - lo, xa + lo * kn, xb + lo, ya ? ya + lo * kn : nullptr, yb ? yb + lo : nullptr,
                              za ? za + lo * kn : nullptr, zb ? zb + lo : nullptr, oa + lo * kn, ob + lo, "affine_bootstrap_batch");
    });
}

int mktfhe_affine_bootstrap_batch_dev(mktfhe_ctx* c, int32_t mu0, int32_t cx, int32_t cy, int32_t cz, int64_t mu, size_t G, const int32_t* xa,
                                      const int32_t* xb, const int32_t* ya, const int32_t* yb, const int32_t* za, const int32_t* zb,
                                      int32_t* oa, int32_t* ob, void* stream) {
    if (!c) return MKTFHE_EINVAL;
    if (G && (!xa || !xb || !oa || !ob || (cy && (!ya || !yb)) || (cz && (!za || !zb)))) return fail(c, MKTFHE_EINVAL, "affine_bootstrap_batch: NULL buffer");
    CU_TRY(c, cudaSetDevice(c->device));
    return run_bootstrap_dev(c, {mu0, cx, cy, cz}, mu, G, xa, xb, ya, yb, za, zb, oa, ob, nullptr, nullptr, true, (cudaStream_t)stream);
}

int mktfhe_gate_batch_mixed(mktfhe_ctx* c, size_t G, const int32_t* gate_ids, const int32_t* xa, const int32_t* xb, const int32_t* ya,
                            const int32_t* yb, const int32_t* za, const int32_t* zb, int32_t* oa, int32_t* ob) {
    if (!c) return MKTFHE_EINVAL;
    if (c->kids.empty()) return mktfhe_gate_batch_mixed_1(c, G, gate_ids, xa, xb, ya, yb, za, zb, oa, ob);
    MK_NOT_READY(c);
    if (G && (!gate_ids || !xa || !xb || !ya || !yb || !oa || !ob)) return fail(c, MKTFHE_EINVAL, "gate_batch_mixed: NULL buffer");
    const size_t kn = (size_t)c->prm.n * c->prm.k;
    return for_each_shard(c, G, [=](mktfhe_ctx* r, size_t lo, size_t hi) {
        return mktfhe_gate_batch_mixed_1(r, hi - lo, gate_ids + lo, xa + lo * kn, xb + lo, ya + lo * kn, yb + lo, za ? za + lo * kn : nullptr,
                                         zb ? zb + lo : nullptr, oa + lo * kn, ob + lo);
    });
}

int mktfhe_bootstrap_batch(mktfhe_ctx* c, int64_t mu, size_t G, const int32_t* a_in, const int32_t* b_in, int32_t* a_out, int32_t* b_out) {
    if (!c) return MKTFHE_EINVAL;
    if (c->kids.empty()) return mktfhe_bootstrap_batch_1(c, mu, G, a_in, b_in, a_out, b_out);
    MK_NOT_READY(c);
    if (G && (!a_in || !b_in || !a_out || !b_out)) return fail(c, MKTFHE_EINVAL, "bootstrap_batch: NULL buffer");
    const size_t kn = (size_t)c->prm.n * c->prm.k;
    return for_each_shard(c, G, [=](mktfhe_ctx* r, size_t lo, size_t hi) {
        return mktfhe_bootstrap_batch_1(r, mu, hi - lo, a_in + lo * kn, b_in + lo, a_out + lo * kn, b_out + lo);
    });
}

int mktfhe_blind_rotate_batch(mktfhe_ctx* c, int64_t mu, size_t G, const int32_t* a_in, const int32_t* b_in, int32_t* ext_out, int64_t* acc_out) {
    if (!c) return MKTFHE_EINVAL;
    if (c->kids.empty()) return mktfhe_blind_rotate_batch_1(c, mu, G, a_in, b_in, ext_out, acc_out);
    MK_NOT_READY(c);
    if (G && (!a_in || !b_in || !ext_out)) return fail(c, MKTFHE_EINVAL, "blind_rotate_batch: NULL buffer");
    const size_t kn = (size_t)c->prm.n * c->prm.k, N = (size_t)c->prm.N;
    return for_each_shard(c, G, [=](mktfhe_ctx* r, size_t lo, size_t hi) {
        return mktfhe_blind_rotate_batch_1(r, mu, hi - lo, a_in + lo * kn, b_in + lo, ext_out + lo * (N + 1), acc_out ? acc_out + lo * 2 * N : nullptr);
    });
}

int mktfhe_keyswitch_batch(mktfhe_ctx* c, size_t G, const int32_t* ext, int32_t* a_out, int32_t* b_out) {
    if (!c) return MKTFHE_EINVAL;
    if (c->kids.empty()) return mktfhe_keyswitch_batch_1(c, G, ext, a_out, b_out);
    MK_NOT_READY(c);
    if (G && (!ext || !a_out || !b_out)) return fail(c, MKTFHE_EINVAL, "keyswitch_batch: NULL buffer");
    const size_t kn = (size_t)c->prm.n * c->prm.k, N = (size_t)c->prm.N;
    return for_each_shard(c, G, [=](mktfhe_ctx* r, size_t lo, size_t hi) {
        return mktfhe_keyswitch_batch_1(r, hi - lo, ext + lo * (N + 1), a_out + lo * kn, b_out + lo);
    });
}

int mktfhe_extprod_batch(mktfhe_ctx* c, size_t G, const int32_t* elem, const int64_t* acc_in, int64_t* acc_out) {
    if (!c) return MKTFHE_EINVAL;
    if (c->kids.empty()) return mktfhe_extprod_batch_1(c, G, elem, acc_in, acc_out);
    MK_NOT_READY(c);
    if (G && (!elem || !acc_in || !acc_out)) return fail(c, MKTFHE_EINVAL, "extprod_batch: NULL buffer");
    const size_t N = (size_t)c->prm.N;
    return for_each_shard(c, G, [=](mktfhe_ctx* r, size_t lo, size_t hi) {
        return mktfhe_extprod_batch_1(r, hi - lo, elem + lo, acc_in + lo * 2 * N, acc_out + lo * 2 * N);
    });
}

int mktfhe_extprod_batch_dev(mktfhe_ctx* c, size_t G, const int32_t* elem, const int64_t* acc_in, int64_t* acc_out, void* stream) {
    if (!c) return MKTFHE_EINVAL;
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (G == 0) return MKTFHE_OK;
    if (!elem || !acc_in || !acc_out) return fail(c, MKTFHE_EINVAL, "extprod_batch_dev: NULL buffer");
    if (c->prm.N == mk2k::N) return fail(c, MKTFHE_EINVAL, "extprod_batch: the single-external-product hook exists for N=1024 only");
    if (G > 0x7fffffffu) return fail(c, MKTFHE_EINVAL, "batch too large");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const size_t sm = br_smem_bytes(c);
    const unsigned grid = (unsigned)((G + c->gpc - 1) / c->gpc);
    if (c->t32 && c->fft) {
        const size_t smf = fft_smem_bytes(c);
        const int gf = mkf::gpc_for(c->prm.l, true);
        const unsigned gridf = (unsigned)((G + gf - 1) / gf);
#define LAUNCH_EPD_FFT_T32(L, GPC, dummy) \
    mkf::extprod_fft_kernel<L, GPC, true><<<gridf, GPC * mkf::TPG, smf, st>>>((int)G, c->d_bsk_fft, c->d_twF, c->prm.bgbit, elem, acc_in, acc_out)
        MK_DISPATCH_FFT_T32(c, LAUNCH_EPD_FFT_T32, 0)
#undef LAUNCH_EPD_FFT_T32
    } else if (c->t32) {
#define LAUNCH_EPD_T32(L, GPC, dummy) mk::extprod_t32_kernel<L, GPC><<<grid, GPC * mk::TPG, sm, st>>>((int)G, c->d_bsk, c->d_twB, c->prm.bgbit, elem, acc_in, acc_out)
        MK_DISPATCH_T32(c, LAUNCH_EPD_T32, 0)
#undef LAUNCH_EPD_T32
    } else if (c->fft) {
        const size_t smf = fft_smem_bytes(c);
        const int gf = mkf::gpc_for(c->prm.l);
        const unsigned gridf = (unsigned)((G + gf - 1) / gf);
#define LAUNCH_EPD_FFT(L, GPC, dummy) mkf::extprod_fft_kernel<L, GPC><<<gridf, GPC * mkf::TPG, smf, st>>>((int)G, c->d_bsk_fft, c->d_twF, c->prm.bgbit, elem, acc_in, acc_out)
        MK_DISPATCH_FFT(c, LAUNCH_EPD_FFT, 0)
#undef LAUNCH_EPD_FFT
    } else {
#define LAUNCH_EPD(L, GPC, dummy) mk::extprod_kernel<L, GPC><<<grid, GPC * mk::TPG, sm, st>>>((int)G, c->d_bsk, c->d_twB, c->prm.bgbit, elem, acc_in, acc_out)
        MK_DISPATCH_L(c, LAUNCH_EPD, 0)
#undef LAUNCH_EPD
    }
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return MKTFHE_OK;
}

int mktfhe_ccs_blind_rotate_batch(mktfhe_ctx* c, int parties, int32_t mu, size_t G, const int32_t* a_in, const int32_t* b_in, int32_t* ext_a,
                                  int32_t* ext_b) {
    if (!c) return MKTFHE_EINVAL;
    if (!c->kids.empty()) return fail(c, MKTFHE_EINVAL, "ccs_blind_rotate_batch runs on a single-device context");
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (!c->t32) return fail(c, MKTFHE_EINVAL, "the CCS composition needs a Torus32-mode context (MKTFHE_FLAG_TORUS32)");
    const int k = parties, n = c->prm.n, N = mk::N;
    if (k < 1 || k * (k + 2) != c->prm.k)
        return fail(c, MKTFHE_EINVAL, "a CCS context for %d parties holds %d pseudo-parties of key elements, this one has %d", k, k * (k + 2), c->prm.k);
    if (G == 0) return MKTFHE_OK;
    if (!a_in || !b_in || !ext_a || !ext_b) return fail(c, MKTFHE_EINVAL, "ccs_blind_rotate_batch: NULL buffer");
    if (G * (size_t)(k + 1) > 0x7fffffffu) return fail(c, MKTFHE_EINVAL, "batch too large");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t P = (size_t)G * (k + 1), abytes = G * (size_t)k * n * 4, xbytes = P * 2 * N * 8;
    int rc;
    if ((rc = stage_in(c, c->in[0], a_in, abytes)) || (rc = stage_in(c, c->in[1], b_in, G * 4)) || (rc = reserve(c, c->raw, P * N * 8)) ||
        (rc = reserve(c, c->accin, xbytes)) || (rc = reserve(c, c->accout, xbytes)) || (rc = reserve(c, c->elem, P * 4)) ||
        (rc = reserve(c, c->ext, G * (size_t)k * N * 4)) || (rc = reserve(c, c->ob, G * 4)))
        return rc;
    u64* acc = (u64*)c->raw.p;
    u64 *xin = (u64*)c->accin.p, *xout = (u64*)c->accout.p;
    int32_t* elem = (int32_t*)c->elem.p;
    const int32_t* d_a = (const int32_t*)c->in[0].p;
    cudaStream_t st = c->stream;
    CU_TRY(c, cudaMemsetAsync(xin, 0, xbytes, st));                                   // the mask operands stay zero
    mk::ccs_init_kernel<<<(unsigned)G, mk::CCS_THREADS, 0, st>>>(acc, (const int32_t*)c->in[1].p, k, (int64_t)mu << 32);
    const size_t sm = br_smem_bytes(c);
    const unsigned grid = (unsigned)((P + c->gpc - 1) / c->gpc);
    const dim3 gi((unsigned)G, (unsigned)(k + 1));
    const size_t smf = c->fft ? fft_smem_bytes(c) : 0;
    const int gf = mkf::gpc_for(c->prm.l, true);
    const unsigned gridf = (unsigned)((P + gf - 1) / gf);
    auto products = [&]() {
        if (c->fft) {   // body-only Torus32 products on the FFT channel
#define LAUNCH_CCS_FFT(L, GPC, dummy) \
    mkf::extprod_fft_kernel<L, GPC, true, true><<<gridf, GPC * mkf::TPG, smf, st>>>((int)P, c->d_bsk_fft, c->d_twF, c->prm.bgbit, elem, (const int64_t*)xin, (int64_t*)xout)
            MK_DISPATCH_FFT_T32(c, LAUNCH_CCS_FFT, 0)
#undef LAUNCH_CCS_FFT
            return;
        }
#define LAUNCH_CCS_T32(L, GPC, dummy) \
    mk::extprod_t32_kernel<L, GPC, true><<<grid, GPC * mk::TPG, sm, st>>>((int)P, c->d_bsk, c->d_twB, c->prm.bgbit, elem, (const int64_t*)xin, (int64_t*)xout)
        MK_DISPATCH_T32(c, LAUNCH_CCS_T32, 0)
#undef LAUNCH_CCS_T32
    };
    for (int party = 0; party < k; party++)                                           // mk_blind_rotate (mk_internals.jl:804-816): parties outer
        for (int j = 0; j < n; j++) {
            mk::ccs_round1_kernel<<<gi, mk::CCS_THREADS, 0, st>>>(acc, d_a, xin, elem, k, n, party, j);
            products();
            mk::ccs_round2_kernel<<<gi, mk::CCS_THREADS, 0, st>>>(acc, xout, xin, elem, k, n, party, j);
            products();
            mk::ccs_accumulate_kernel<<<(unsigned)G, mk::CCS_THREADS, 0, st>>>(acc, xout, k, party);
            c->launches += 5;
        }
    CU_TRY(c, cudaGetLastError());
    mk::ccs_extract_kernel<<<(unsigned)G, mk::CCS_THREADS, 0, st>>>(acc, (int32_t*)c->ext.p, (int32_t*)c->ob.p, k);
    c->launches += 2;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(ext_a, c->ext.p, G * (size_t)k * N * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(c, cudaMemcpyAsync(ext_b, c->ob.p, G * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(c, cudaStreamSynchronize(st));
    return MKTFHE_OK;
}

int mktfhe_mk_keyswitch_batch(mktfhe_ctx* c, size_t G, const int32_t* ext_a, const int32_t* ext_b, int32_t* a_out, int32_t* b_out) {
    if (!c) return MKTFHE_EINVAL;
    if (!c->kids.empty()) return fail(c, MKTFHE_EINVAL, "mk_keyswitch_batch runs on a single-device context");
    if (!c->ready) return fail(c, MKTFHE_ESTATE, "keys not finalized (mktfhe_finalize_keys)");
    if (G == 0) return MKTFHE_OK;
    if (!ext_a || !ext_b || !a_out || !b_out) return fail(c, MKTFHE_EINVAL, "mk_keyswitch_batch: NULL buffer");
    if (c->prm.N != mk::N) return fail(c, MKTFHE_EINVAL, "mk_keyswitch_batch exists for N=1024 only");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t kn = (size_t)c->prm.n * c->prm.k, abytes = G * kn * 4, bbytes = G * 4, eabytes = G * (size_t)c->prm.k * c->prm.N * 4;
    int rc;
    if ((rc = stage_in(c, c->ext, ext_a, eabytes)) || (rc = stage_in(c, c->in[1], ext_b, bbytes)) || (rc = reserve(c, c->oa, abytes)) || (rc = reserve(c, c->ob, bbytes)))
        return rc;
    mk::mk_keyswitch_kernel<<<(unsigned)G, mk::KS_THREADS, 0, c->stream>>>(c->prm.n, c->prm.k, c->prm.t, c->prm.basebit, c->d_ksk, (const int32_t*)c->ext.p,
                                                                           (const int32_t*)c->in[1].p, (int32_t*)c->oa.p, (int32_t*)c->ob.p);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(a_out, c->oa.p, abytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(b_out, c->ob.p, bbytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return MKTFHE_OK;
}

int mktfhe_negacyclic_mul_batch(mktfhe_ctx* c, size_t G, const int64_t* a, const int64_t* b, int64_t* out) {
    if (!c) return MKTFHE_EINVAL;
    if (c->kids.empty()) return mktfhe_negacyclic_mul_batch_1(c, G, a, b, out);
    if (G && (!a || !b || !out)) return fail(c, MKTFHE_EINVAL, "negacyclic_mul_batch: NULL buffer");
    const size_t N = (size_t)c->prm.N;
    return for_each_shard(c, G, [=](mktfhe_ctx* r, size_t lo, size_t hi) {
        return mktfhe_negacyclic_mul_batch_1(r, hi - lo, a + lo * N, b + lo * N, out + lo * N);
    });
}

// ---- multi-device lifetime -----------------------------------------------------------------------------------------------------
int mktfhe_create_multi(const mktfhe_params* params, int n_devices, const int* devices, mktfhe_ctx** out) {
    if (!out) return fail(nullptr, MKTFHE_EINVAL, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MKTFHE_ECUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (n_devices == 0) n_devices = ndev;                       // 0 = every visible GPU
    if (n_devices < 1 || n_devices > 64) return fail(nullptr, MKTFHE_EINVAL, "n_devices=%d out of range (1..64, or 0 for all)", n_devices);
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; i++) devs[i] = devices ? devices[i] : i;
    mktfhe_ctx* c = nullptr;
    int rc = mktfhe_create(params, devs[0], &c);
    if (rc) return rc;
    for (int i = 1; i < n_devices; i++) {
        mktfhe_ctx* k = nullptr;
        rc = mktfhe_create(params, devs[i], &k);
        if (rc) { mktfhe_destroy(c); return rc; }               // g_create_error holds the message
        k->is_kid = true;
        k->parent = c;
        c->kids.push_back(k);
    }
    // peer access in both directions between every pair of distinct devices, where the hardware offers it (NVLink / NVSwitch):
    // the key broadcast then moves GPU to GPU; without it cudaMemcpyPeerAsync stages through the host
    for (int i = 0; i < n_devices; i++)
        for (int j = 0; j < n_devices; j++) {
            if (devs[i] == devs[j]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) == cudaSuccess && can) {
                cudaSetDevice(devs[i]);
                cudaError_t pe = cudaDeviceEnablePeerAccess(devs[j], 0);
                if (pe != cudaSuccess) cudaGetLastError();      // already enabled (by us or by the host framework): fine
            }
        }
    cudaSetDevice(devs[0]);
    *out = c;
    return MKTFHE_OK;
}

int mktfhe_device_count(const mktfhe_ctx* c) { return c ? (int)n_replicas(c) : 0; }

int mktfhe_device_ctx(mktfhe_ctx* c, int i, mktfhe_ctx** out, int* device) {
    if (!c) return MKTFHE_EINVAL;
    if (i < 0 || (size_t)i >= n_replicas(c)) return fail(c, MKTFHE_EINVAL, "device_ctx: replica %d out of range (%zu devices)", i, n_replicas(c));
    if (out) *out = replica(c, i);
    if (device) *device = replica(c, i)->device;
    return MKTFHE_OK;
}

int mktfhe_shard_bounds(const mktfhe_ctx* c, size_t G, int i, size_t* lo, size_t* hi) {
    if (!c || !lo || !hi || i < 0 || (size_t)i >= n_replicas(c)) return MKTFHE_EINVAL;
    shard_bounds(G, n_replicas(c), (size_t)i, lo, hi);
    return MKTFHE_OK;
}

int mktfhe_pin_host(void* p, size_t bytes) {
    if (!p || !bytes) return MKTFHE_EINVAL;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, MKTFHE_ECUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
    return MKTFHE_OK;
}
int mktfhe_unpin_host(void* p) {
    if (!p) return MKTFHE_EINVAL;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, MKTFHE_ECUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
    return MKTFHE_OK;
}

#ifndef MK_BUILD_ID
#define MK_BUILD_ID "unknown"
#endif
const char* mktfhe_build_id(void) { return MK_BUILD_ID; }

int mktfhe_describe(const mktfhe_ctx* c, char* buf, size_t cap) {
    if (!c || !buf || !cap) return MKTFHE_EINVAL;
    std::string d = "{\"build_id\": \"" MK_BUILD_ID "\", \"devices\": [";
    for (size_t r = 0; r < n_replicas(c); r++) d += (r ? ", " : "") + std::to_string(replica(const_cast<mktfhe_ctx*>(c), r)->device);
    d += "], \"key_broadcast\": \"" + (c->kids.empty() ? std::string("none") : (c->bcast_how.empty() ? std::string("pending") : c->bcast_how)) + "\"";
    d += ", \"keyswitch_fused\": " + std::string(c->last_fused ? "true" : "false");
    d += ", \"external_product\": \"" + std::string(c->fft ? "fft64" : "ntt_rns") + "\"";
    d += ", \"gates_per_cta\": " + std::to_string(c->gpc) + ", \"sms\": " + std::to_string(c->num_sms);
    d += ", \"bsk_bytes\": " + std::to_string(c->bsk_bytes) + ", \"ksk_bytes\": " + std::to_string(c->ksk_bytes) + "}";
    snprintf(buf, cap, "%s", d.c_str());
    return d.size() < cap ? MKTFHE_OK : MKTFHE_EINVAL;
}

uint64_t mktfhe_launch_count(const mktfhe_ctx* c) {
    if (!c) return 0;
    uint64_t n = c->launches;
    for (const mktfhe_ctx* k : c->kids) n += k->launches;
    return n;
}

int mktfhe_last_kernel_ms(mktfhe_ctx* c, float* blind_rotate_ms, float* keyswitch_ms) {
    if (!c) return MKTFHE_EINVAL;
    float br = 0.f, ks = 0.f;
    bool any = false;
    for (size_t r = 0; r < n_replicas(c); r++) {   // multi-device: the slowest of the replicas that took part in the most recent call
        mktfhe_ctx* d = replica(c, r);
        if (!d->ev_valid || (!c->kids.empty() && d->ev_seq != c->seq)) continue;
        any = true;
        CU_TRY(c, cudaSetDevice(d->device));
        CU_TRY(c, cudaEventSynchronize(d->ev[3]));
        float b1 = 0.f, k1 = 0.f;
        CU_TRY(c, cudaEventElapsedTime(&b1, d->ev[0], d->ev[1]));
        CU_TRY(c, cudaEventElapsedTime(&k1, d->ev[2], d->ev[3]));
        br = b1 > br ? b1 : br;
        ks = k1 > ks ? k1 : ks;
    }
    if (!any) return fail(c, MKTFHE_ESTATE, "no batch has run yet");
    CU_TRY(c, cudaSetDevice(c->device));
    if (blind_rotate_ms) *blind_rotate_ms = br;
    if (keyswitch_ms) *keyswitch_ms = ks;
    return MKTFHE_OK;
}

int mktfhe_algorithmic_bytes(const mktfhe_ctx* c, double* bsk_bytes_per_gate, double* ksk_bytes_per_gate) {
    if (!c) return MKTFHE_EINVAL;
    const mktfhe_params& p = c->prm;
    const double B = (double)(1 << p.basebit);
    if (bsk_bytes_per_gate) *bsk_bytes_per_gate = (double)c->bsk_bytes;
    if (ksk_bytes_per_gate) *ksk_bytes_per_gate = (double)p.k * p.N * p.t * (1.0 - 1.0 / B) * (p.n + 1) * 4.0;
    return MKTFHE_OK;
}

}  // extern "C"
