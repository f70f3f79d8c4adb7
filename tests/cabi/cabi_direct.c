/* cabi_direct.c -- the C ABI (include/mktfhe_b200.h) driven from plain C, the way a cgo / ccall / JNI host would bind it: no Python,
 * no torch.  Keys and ciphertexts come from the CPU oracle (TEST INFRASTRUCTURE, linked here only as the checker); the program runs
 * one batch of every bootstrapped gate through libmktfhe_b200.so and compares the bytes with the oracle's exact back-end, then checks
 * the error behaviour (negative codes + mktfhe_last_error, no aborts).
 * Build (tests/test_cabi_direct.py does this): gcc -O2 -I include -I oracle tests/cabi/cabi_direct.c -o cabi_direct \
 *        torus-fhe_b200/libmktfhe_b200.so oracle/liboracle_mk3gen.so -Wl,-rpath,... -lm */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mktfhe_b200.h"
#include "mk_oracle.h"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

int main(int argc, char **argv) {
    /* reduced LWE dimension so that CPU key generation takes a second; ring and gadget as the 2-party default (mk_api.jl:32-38) */
    mko_params op = {64, 1024, 2, 2, 7, 3, 3, 0, 1.0 / 11000.0, 5.7e-10, 1.0 / 11000.0};
    mko_keyset *ks = mko_keygen(&op, 7, 8);
    CHECK(ks, "oracle keygen");
    mktfhe_params gp;
    memset(&gp, 0, sizeof gp);
    gp.n = op.n; gp.N = op.N; gp.k = op.k; gp.l = op.l; gp.bgbit = op.bgbit; gp.t = op.t; gp.basebit = op.basebit;
    mktfhe_ctx *ctx = NULL;
    int rc = mktfhe_create(&gp, 0, &ctx);
    CHECK(rc == MKTFHE_OK && ctx, "mktfhe_create rc=%d: %s", rc, mktfhe_last_error(NULL));

    /* gates before the keys are finalized must fail with a state error, not crash */
    int32_t dummy_a[2 * 64] = {0}, dummy_b = 0, out_a[2 * 64], out_b;
    rc = mktfhe_gate_batch(ctx, MKTFHE_GATE_NAND, 1, dummy_a, &dummy_b, dummy_a, &dummy_b, NULL, NULL, out_a, &out_b);
    CHECK(rc == MKTFHE_ESTATE && strlen(mktfhe_last_error(ctx)) > 0, "expected MKTFHE_ESTATE before finalize, got %d", rc);

    const size_t bsk_party = (size_t)op.n * 4 * op.l * op.N, ksk_party = (size_t)op.N * op.t * ((1 << op.basebit) - 1) * (op.n + 1);
    for (int p = 0; p < op.k; p++) {
        CHECK(mktfhe_load_bsk(ctx, p, mko_bsk(ks) + p * bsk_party) == MKTFHE_OK, "load_bsk: %s", mktfhe_last_error(ctx));
        CHECK(mktfhe_load_ksk(ctx, p, mko_ksk(ks) + p * ksk_party) == MKTFHE_OK, "load_ksk: %s", mktfhe_last_error(ctx));
    }
    CHECK(mktfhe_load_bsk(ctx, op.k, mko_bsk(ks)) == MKTFHE_EINVAL, "party out of range must be rejected");
    CHECK(mktfhe_finalize_keys(ctx) == MKTFHE_OK, "finalize: %s", mktfhe_last_error(ctx));

    enum { G = 8 };
    const uint8_t xb_[G] = {0, 0, 1, 1, 0, 1, 0, 1}, yb_[G] = {0, 1, 0, 1, 1, 1, 0, 0}, zb_[G] = {1, 1, 1, 1, 0, 0, 1, 0};
    const size_t an = (size_t)G * op.k * op.n;
    int32_t *xa = malloc(an * 4), *ya = malloc(an * 4), *za = malloc(an * 4), *oa = malloc(an * 4), *ra = malloc(an * 4);
    int32_t xb[G], yb[G], zb[G], ob[G], rb[G], ph[G];
    mko_encrypt(ks, 101, G, xb_, xa, xb);
    mko_encrypt(ks, 102, G, yb_, ya, yb);
    mko_encrypt(ks, 103, G, zb_, za, zb);
    for (int gate = MKTFHE_GATE_NAND; gate <= MKTFHE_GATE_AND3; gate++) {
        const int three = gate == MKTFHE_GATE_AND3;
        rc = mktfhe_gate_batch(ctx, gate, G, xa, xb, ya, yb, three ? za : NULL, three ? zb : NULL, oa, ob);
        CHECK(rc == MKTFHE_OK, "gate %d: %s", gate, mktfhe_last_error(ctx));
        mko_gate_batch(ks, MKO_EXACT_NTT, gate, G, xa, xb, ya, yb, three ? za : NULL, three ? zb : NULL, ra, rb, 8);
        CHECK(memcmp(oa, ra, an * 4) == 0 && memcmp(ob, rb, sizeof ob) == 0, "gate %d differs from the exact oracle", gate);
        mko_phase(ks, G, oa, ob, ph);
        for (int g = 0; g < G; g++) {
            const int x = xb_[g], y = yb_[g], z = zb_[g];
            int want = gate == MKTFHE_GATE_NAND ? !(x && y) : gate == MKTFHE_GATE_OR ? (x || y) : gate == MKTFHE_GATE_AND ? (x && y)
                     : gate == MKTFHE_GATE_XOR ? (x ^ y) : (x && y && z);
            if (three && !x && !y && !z) want = 1;   /* reference quirk: 3AND(false, false, false) wraps to true */
            CHECK((ph[g] > 0) == want, "gate %d sample %d decrypts wrong", gate, g);
        }
    }
    /* bootstrap alone == blind rotate + key switch (mk_bootstrap_3gen = mk_keyswitch_3gen o mk_bootstrap_wo_keyswitch_3gen) */
    const int64_t mu = (int64_t)1 << 61;
    CHECK(mktfhe_bootstrap_batch(ctx, mu, G, xa, xb, oa, ob) == MKTFHE_OK, "bootstrap: %s", mktfhe_last_error(ctx));
    int32_t *ext = malloc((size_t)G * (op.N + 1) * 4);
    CHECK(mktfhe_blind_rotate_batch(ctx, mu, G, xa, xb, ext, NULL) == MKTFHE_OK, "blind_rotate: %s", mktfhe_last_error(ctx));
    CHECK(mktfhe_keyswitch_batch(ctx, G, ext, ra, rb) == MKTFHE_OK, "keyswitch: %s", mktfhe_last_error(ctx));
    CHECK(memcmp(oa, ra, an * 4) == 0 && memcmp(ob, rb, sizeof ob) == 0, "bootstrap != keyswitch(blind_rotate)");
    /* invalid gate id and empty batch */
    CHECK(mktfhe_gate_batch(ctx, 9, G, xa, xb, ya, yb, NULL, NULL, oa, ob) == MKTFHE_EINVAL, "gate id 9 must be rejected");
    CHECK(mktfhe_gate_batch(ctx, MKTFHE_GATE_NAND, 0, xa, xb, ya, yb, NULL, NULL, oa, ob) == MKTFHE_OK, "empty batch is a no-op");
    float br_ms = 0, ks_ms = 0;
    CHECK(mktfhe_last_kernel_ms(ctx, &br_ms, &ks_ms) == MKTFHE_OK && br_ms > 0, "kernel timing");
    printf("cabi_direct: OK (5 gates x %d samples bit-exact vs the oracle; last blind rotate %.3f ms)\n", G, br_ms);

    /* ---- one context spanning several GPUs (mktfhe_create_multi): same keys loaded once, broadcast inside finalize, batches
     * sharded by the library.  Every visible GPU is used; on a one-GPU box the device is listed twice (two replicas sharing
     * it), which runs the same broadcast and sharding code.  Results must equal the single-GPU context's byte for byte. */
    {
        int devs[64], nd = argc > 1 ? atoi(argv[1]) : 0;
        mktfhe_ctx *probe = NULL;
        CHECK(mktfhe_create_multi(&gp, 0, NULL, &probe) == MKTFHE_OK, "create_multi(all): %s", mktfhe_last_error(NULL));
        const int visible = mktfhe_device_count(probe);
        mktfhe_destroy(probe);
        if (nd <= 0) nd = visible > 1 ? visible : 2;
        for (int i = 0; i < nd; i++) devs[i] = i % visible;
        mktfhe_ctx *mc = NULL;
        CHECK(mktfhe_create_multi(&gp, nd, devs, &mc) == MKTFHE_OK && mktfhe_device_count(mc) == nd, "create_multi: %s", mktfhe_last_error(NULL));
        for (int p = 0; p < op.k; p++) {
            CHECK(mktfhe_load_bsk(mc, p, mko_bsk(ks) + p * bsk_party) == MKTFHE_OK, "multi load_bsk: %s", mktfhe_last_error(mc));
            CHECK(mktfhe_load_ksk(mc, p, mko_ksk(ks) + p * ksk_party) == MKTFHE_OK, "multi load_ksk: %s", mktfhe_last_error(mc));
        }
        CHECK(mktfhe_gate_batch(mc, MKTFHE_GATE_NAND, G, xa, xb, ya, yb, NULL, NULL, oa, ob) == MKTFHE_ESTATE, "multi: gates before finalize");
        CHECK(mktfhe_finalize_keys(mc) == MKTFHE_OK, "multi finalize (key broadcast): %s", mktfhe_last_error(mc));
        /* a batch that does not divide evenly, so that slices differ in size; built by repeating the G samples */
        enum { GM = 37 };
        const size_t kn = (size_t)op.k * op.n;
        int32_t *mxa = malloc(GM * kn * 4), *mya = malloc(GM * kn * 4), *mza = malloc(GM * kn * 4), *moa = malloc(GM * kn * 4), *mra = malloc(GM * kn * 4);
        int32_t mxb[GM], myb[GM], mzb[GM], mob[GM], mrb[GM], ids[GM];
        for (int g = 0; g < GM; g++) {
            const int s = (g * 5 + 3) % G;
            memcpy(mxa + g * kn, xa + s * kn, kn * 4); memcpy(mya + g * kn, ya + ((s + 1) % G) * kn, kn * 4); memcpy(mza + g * kn, za + s * kn, kn * 4);
            mxb[g] = xb[s]; myb[g] = yb[(s + 1) % G]; mzb[g] = zb[s]; ids[g] = g % 5;
        }
        size_t covered = 0;
        for (int i = 0; i < nd; i++) {
            size_t lo, hi;
            CHECK(mktfhe_shard_bounds(mc, GM, i, &lo, &hi) == MKTFHE_OK && lo == covered && hi >= lo, "shard_bounds");
            covered = hi;
        }
        CHECK(covered == GM, "shards must cover the batch");
        for (int gate = MKTFHE_GATE_NAND; gate <= MKTFHE_GATE_AND3; gate++) {
            const int three = gate == MKTFHE_GATE_AND3;
            CHECK(mktfhe_gate_batch(mc, gate, GM, mxa, mxb, mya, myb, three ? mza : NULL, three ? mzb : NULL, moa, mob) == MKTFHE_OK, "multi gate %d: %s", gate, mktfhe_last_error(mc));
            CHECK(mktfhe_gate_batch(ctx, gate, GM, mxa, mxb, mya, myb, three ? mza : NULL, three ? mzb : NULL, mra, mrb) == MKTFHE_OK, "single gate %d: %s", gate, mktfhe_last_error(ctx));
            CHECK(memcmp(moa, mra, GM * kn * 4) == 0 && memcmp(mob, mrb, sizeof mob) == 0, "gate %d: %d-device result differs from the 1-device result", gate, nd);
        }
        CHECK(mktfhe_gate_batch_mixed(mc, GM, ids, mxa, mxb, mya, myb, mza, mzb, moa, mob) == MKTFHE_OK, "multi mixed: %s", mktfhe_last_error(mc));
        CHECK(mktfhe_gate_batch_mixed(ctx, GM, ids, mxa, mxb, mya, myb, mza, mzb, mra, mrb) == MKTFHE_OK, "single mixed: %s", mktfhe_last_error(ctx));
        CHECK(memcmp(moa, mra, GM * kn * 4) == 0 && memcmp(mob, mrb, sizeof mob) == 0, "mixed batch: multi != single");
        CHECK(mktfhe_bootstrap_batch(mc, mu, GM, mxa, mxb, moa, mob) == MKTFHE_OK, "multi bootstrap: %s", mktfhe_last_error(mc));
        CHECK(mktfhe_bootstrap_batch(ctx, mu, GM, mxa, mxb, mra, mrb) == MKTFHE_OK, "single bootstrap");
        CHECK(memcmp(moa, mra, GM * kn * 4) == 0 && memcmp(mob, mrb, sizeof mob) == 0, "bootstrap: multi != single");
        /* fewer gates than devices, pinned caller buffers, and the error path of a slice */
        CHECK(mktfhe_pin_host(mxa, GM * kn * 4) == MKTFHE_OK, "pin_host: %s", mktfhe_last_error(NULL));
        CHECK(mktfhe_gate_batch(mc, MKTFHE_GATE_XOR, 1, mxa, mxb, mya, myb, NULL, NULL, moa, mob) == MKTFHE_OK, "multi batch of one");
        CHECK(mktfhe_gate_batch(ctx, MKTFHE_GATE_XOR, 1, mxa, mxb, mya, myb, NULL, NULL, mra, mrb) == MKTFHE_OK, "single batch of one");
        CHECK(memcmp(moa, mra, kn * 4) == 0 && mob[0] == mrb[0], "batch of one: multi != single");
        CHECK(mktfhe_unpin_host(mxa) == MKTFHE_OK, "unpin_host");
        ids[GM - 1] = 77;
        CHECK(mktfhe_gate_batch_mixed(mc, GM, ids, mxa, mxb, mya, myb, mza, mzb, moa, mob) == MKTFHE_EINVAL && strstr(mktfhe_last_error(mc), "device"),
              "a bad gate id in the last slice must surface with the device named: %s", mktfhe_last_error(mc));
        mktfhe_ctx *r1 = NULL; int d1 = -1;
        CHECK(mktfhe_device_ctx(mc, nd - 1, &r1, &d1) == MKTFHE_OK && r1 && d1 == devs[nd - 1], "device_ctx");
        char desc[512];
        CHECK(mktfhe_describe(mc, desc, sizeof desc) == MKTFHE_OK, "describe");
        CHECK(mktfhe_last_kernel_ms(mc, &br_ms, &ks_ms) == MKTFHE_OK && br_ms > 0, "multi kernel timing");
        printf("cabi_direct multi: OK (%d devices, %d visible; %d-gate batches equal the 1-device result byte for byte) %s\n", nd, visible, GM, desc);
        mktfhe_destroy(mc);
        free(mxa); free(mya); free(mza); free(moa); free(mra);
    }
    mktfhe_destroy(ctx);
    mko_keyset_free(ks);
    free(xa); free(ya); free(za); free(oa); free(ra); free(ext);
    return 0;
}
