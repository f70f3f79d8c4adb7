#!/usr/bin/env python
"""Counterpart of the reference's single-key tutorial 3-gen-mk-tfhe/tutorial.jl (:5-83): encrypt two 16-bit words, compute their
minimum homomorphically with the bit comparator of :41-64 (gate_xnor + gate_mux per bit, then one gate_mux per output bit), decrypt.
The reference's copy encrypts only the first bit of each word (:27, :32); this one encrypts all sixteen, as its comments intend, and
-- because every sample may carry a batch -- compares many pairs at once.  Run on a GPU box: python examples/tutorial.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torus_fhe_b200 as T  # noqa: E402

T1 = T.tfhe1


def int_to_bits(x, nbits=16):
    x = np.asarray(x, dtype=np.int64)
    return [((x >> i) & 1) != 0 for i in range(nbits)]


def bits_to_int(bits):
    return sum(np.asarray(b).astype(np.int64) << i for i, b in enumerate(bits))


def encrypted_compare_bit(ck, a, b, lsb_carry):
    """tutorial.jl:41-44: if a == b keep the carry of the lower bits, else the answer is a."""
    tmp = T1.gate_xnor(ck, a, b)
    return T1.gate_mux(ck, tmp, lsb_carry, a)


def encrypted_minimum(ck, a, b):
    """tutorial.jl:47-64."""
    shape = a[0].b.shape
    tmps1 = T1.lwe_noiseless_trivial(T1.encode_message(-1, 8), a[0].params, shape)     # gate_constant(ck, false), batched
    for i in range(len(a)):
        tmps1 = encrypted_compare_bit(ck, a[i], b[i], tmps1)
    # tmps1 = 0 if a is larger, 1 if b is larger: select the smaller word
    return [T1.gate_mux(ck, tmps1, b[i], a[i]) for i in range(len(a))]


def main():
    rng = np.random.default_rng(123)
    t0 = time.perf_counter()
    secret_key, cloud_key = T1.make_key_pair(rng)                  # the reference's default: tfhe_parameters_80
    T1.engine_for(cloud_key)
    print(f"key generation + upload: {time.perf_counter() - t0:.2f} s")
    pairs = 256
    w1 = np.concatenate([[2017], rng.integers(0, 1 << 16, pairs - 1)])
    w2 = np.concatenate([[42], rng.integers(0, 1 << 16, pairs - 1)])
    c1 = [T1.encrypt(rng, secret_key, b) for b in int_to_bits(w1)]
    c2 = [T1.encrypt(rng, secret_key, b) for b in int_to_bits(w2)]
    t0 = time.perf_counter()
    answer = encrypted_minimum(cloud_key, c1, c2)
    dt = time.perf_counter() - t0
    got = bits_to_int([T1.decrypt(secret_key, a) for a in answer])
    print(f"Answer: {int(got[0])}   (minimum of 2017 and 42)")
    ok = int(np.sum(got == np.minimum(w1, w2)))
    print(f"{pairs} comparisons of 16-bit words in {dt:.2f} s ({pairs * 16 * 5 / dt:.0f} bootstraps/s); {ok}/{pairs} equal min(a, b)")


if __name__ == "__main__":
    main()
