#!/usr/bin/env python
"""Counterpart of the reference workload 3-gen-mk-tfhe/VolumeMatching.jl (main, :300-360): 2-party key generation, four encrypted buy
and four encrypted sell orders of WIDTH = 8 bits, the dark-pool volume matching circuit, decryption.  The reference fans the
sub-circuits out over 20 Julia worker processes; here every dependency level is one batched launch on the B200
(torus-fhe_b200/workloads.py).  Run on a GPU box: python examples/volume_matching.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torus_fhe_b200 as T  # noqa: E402


def main():
    parties = 2
    params = T.mktfhe_parameters_2party_3gen
    rng = np.random.default_rng()
    print(f"3rd MK-TFHE - {parties} parties\n========\n(3) KEY GENERATION AND PRECOMP ...")
    t0 = time.perf_counter()
    secret_keys = [T.SecretKey_3gen(rng, params) for _ in range(parties)]
    rlwe_keys = [T.RLweKey(rng, T.rlwe_parameters(params), True) for _ in range(parties)]
    crp_a = T.CRP_3gen(rng, T.tgsw_parameters(params), T.rlwe_parameters(params), True)
    pubkeys = [T.PublicKey(rng, rlwe_keys[i], params.gsw_noise_stddev, crp_a, T.tgsw_parameters(params), 1) for i in range(parties)]
    common_pubkey = T.CommonPubKey_3gen(pubkeys, params, parties)
    bk_keys = [T.BootstrapKeyPart_3gen(rng, secret_keys[i].key, params.gsw_noise_stddev, crp_a, common_pubkey, T.tgsw_parameters(params),
                                       T.rlwe_parameters(params), 1) for i in range(parties)]
    bk_keys = [T.TransformedBootstrapKeyPart_3gen(b) for b in bk_keys]
    ks_keys = [T.KeyswitchKey(rng, params.ks_noise_stddev, T.keyswitch_parameters(params), secret_keys[i].key, rlwe_keys[i]) for i in range(parties)]
    T.engine_for(bk_keys, ks_keys)
    print(f"(3) PRECOMP TIME : {time.perf_counter() - t0:.2f} seconds\n")

    WIDTH = 8
    buy, sell = [5, 2, 11, 1], [3, 4, 5, 5]                     # the orders of the reference's main()
    enc = lambda v: T.mk_int_encrypt_3gen(rng, secret_keys, int(v), WIDTH)
    ord_buy, ord_sell = [enc(v) for v in buy], [enc(v) for v in sell]
    zero_arr, acc_buy, acc_sell = enc(0), enc(0), enc(0)
    one, zero = T.mk_encrypt_3gen(rng, secret_keys, True), T.mk_encrypt_3gen(rng, secret_keys, False)
    eng = T.engine_for(bk_keys, ks_keys)
    l0 = eng.ctx.launch_count()
    t0 = time.perf_counter()
    res_buy, res_sell = T.VolumeMatch(bk_keys, ks_keys, ord_buy, ord_sell, acc_buy, acc_sell, zero_arr, one, zero, WIDTH)
    dt = time.perf_counter() - t0
    got_buy = [T.mk_int_decrypt_3gen(secret_keys, r, WIDTH) for r in res_buy]
    got_sell = [T.mk_int_decrypt_3gen(secret_keys, r, WIDTH) for r in res_sell]
    exp_buy, exp_sell = T.volume_match_plain(buy, sell)
    print("buy orders :", buy, "-> matched", got_buy, "(plain model:", exp_buy, ")")
    print("sell orders:", sell, "-> matched", got_sell, "(plain model:", exp_sell, ")")
    print(f"{dt:.3f} seconds, {eng.ctx.launch_count() - l0} launches")


if __name__ == "__main__":
    main()
