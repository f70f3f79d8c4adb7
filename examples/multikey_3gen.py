#!/usr/bin/env python
"""Line-for-line counterpart of the reference demo 3-gen-mk-tfhe/multikey_3gen.jl (:6-94): 2-party key generation, integer
encryption, mk_add_3gen_v2, decryption -- on the B200 engine.  Run on a GPU box: python examples/multikey_3gen.py"""
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
    T.engine_for(bk_keys, ks_keys)          # upload + exact transform on the GPU
    print(f"(3) PRECOMP TIME : {time.perf_counter() - t0:.2f} seconds")
    print(f"(3) BK SIZE : {bk_keys[0].gsw_key.nbytes / 2 ** 20:.1f} MiB, KSK SIZE : {ks_keys[0].key.nbytes / 2 ** 20:.1f} MiB\n")

    WIDTH = 8
    for _ in range(5):
        msg1, msg2 = int(rng.integers(1, 11)), int(rng.integers(1, 11))
        print("msg1:", msg1, "\nmsg2:", msg2)
        ct1 = T.mk_int_encrypt_3gen(rng, secret_keys, msg1, WIDTH)
        ct2 = T.mk_int_encrypt_3gen(rng, secret_keys, msg2, WIDTH)
        print("ct1:", T.mk_int_decrypt_3gen(secret_keys, ct1, WIDTH))
        print("ct2:", T.mk_int_decrypt_3gen(secret_keys, ct2, WIDTH))
        ZERO = T.mk_encrypt_3gen(rng, secret_keys, False)
        print("Cin:", T.mk_decrypt_3gen(secret_keys, ZERO))
        t0 = time.perf_counter()
        ct_res = T.mk_add_3gen_v2(bk_keys, ks_keys, ct1, ct2, ZERO, WIDTH)
        print(f"  {time.perf_counter() - t0:.3f} seconds (40 bootstrapped gates, 17 launches)")
        print("result:", T.mk_int_decrypt_3gen(secret_keys, ct_res, WIDTH))


if __name__ == "__main__":
    main()
