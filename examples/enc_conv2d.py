#!/usr/bin/env python
"""Encrypted convolution layer (BASELINE configs[4]; the reference's enc_conv2d, 3-gen-mk-tfhe/src/3gen_mk_gates.jl:364-397, rebuilt so
that it runs): a small encrypted image, one encrypted 3x3 kernel, every dependency level of the whole layer one launch.
Run on a GPU box: python examples/enc_conv2d.py [image_size] [width_bits]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torus_fhe_b200 as T  # noqa: E402


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    WIDTH = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    parties, params = 2, T.mktfhe_parameters_2party_3gen
    rng = np.random.default_rng(1)
    sk = [T.SecretKey_3gen(rng, params) for _ in range(parties)]
    rk = [T.RLweKey(rng, T.rlwe_parameters(params), True) for _ in range(parties)]
    crp = T.CRP_3gen(rng, T.tgsw_parameters(params), T.rlwe_parameters(params), True)
    pk = [T.PublicKey(rng, rk[i], params.gsw_noise_stddev, crp, T.tgsw_parameters(params), 1) for i in range(parties)]
    cpk = T.CommonPubKey_3gen(pk, params, parties)
    bk = [T.TransformedBootstrapKeyPart_3gen(T.BootstrapKeyPart_3gen(rng, sk[i].key, params.gsw_noise_stddev, crp, cpk, T.tgsw_parameters(params),
                                                                      T.rlwe_parameters(params), 1)) for i in range(parties)]
    ks = [T.KeyswitchKey(rng, params.ks_noise_stddev, T.keyswitch_parameters(params), sk[i].key, rk[i]) for i in range(parties)]
    eng = T.engine_for(bk, ks)

    half = 1 << (WIDTH - 1)
    image, kernel = rng.integers(-half, half, (H, H)), rng.integers(-half, half, (1, 3, 3))
    up = lambda bits: [T.MKLweSampleGPU.from_host(b) for b in bits]          # ciphertexts stay in HBM between levels
    cimg, cker = up(T.mk_int_encrypt_3gen(rng, sk, image, WIDTH)), up(T.mk_int_encrypt_3gen(rng, sk, kernel, WIDTH))
    zero = T.MKLweSampleGPU.from_host(T.mk_encrypt_3gen(rng, sk, False))
    gates = T.conv2d_gate_count((H, H), (1, 3, 3), 1, 0, WIDTH)
    l0, t0 = eng.ctx.launch_count(), time.perf_counter()
    out = T.enc_conv2d(bk, ks, cimg, zero, cker, 1, 0, WIDTH)
    got = T.mk_int_decrypt_3gen(sk, [o.cpu() for o in out], WIDTH)
    dt = time.perf_counter() - t0
    exp = T.conv2d_plain(image, kernel, 1, 0, WIDTH)
    print(f"{H}x{H} image, 3x3 kernel, {WIDTH}-bit: {gates} bootstrapped gates in {eng.ctx.launch_count() - l0} launches, {dt:.2f} s ({gates / dt:.0f} gates/s)")
    print("decrypted == plaintext convolution on", int((got == exp).sum()), "of", exp.size, "outputs (the default parameters fail ~1e-3 of the gates)")
    print(got[0])


if __name__ == "__main__":
    main()
