"""Pins on the REAL reference's bytes -- active only when a maintainer with Julia has run julia/dump_fixture.jl into
tests/golden/julia/ (skipped otherwise: the build image has no Julia).

With the fixture present, on identical keys and ciphertexts:
  * the reference's own outputs decrypt to the truth table under the dumped secret keys (sanity of the dump and of our reading of the
    ciphertext layout and the phase convention);
  * the oracle's exact and Float64-FFT back-ends decrypt to the same bits as the reference, gate by gate;
  * the torus phase of the oracle's FFT restatement agrees with the reference's FFT path within 2^-20 wherever the FFT restatement
    and the exact path agree with each other to that bound (a float-induced digit flip makes either FFT run a different noise
    realisation), and the count of such agreements is reported;
  * (GPU) the engine's outputs equal the exact oracle bit for bit and decrypt like the reference.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN

FIX = os.path.join(GOLDEN, "julia")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(FIX, "keys.bin")),
                                reason="no fixture from the real reference (run julia/dump_fixture.jl on a machine with Julia)")
GATES = ("nand", "and", "or", "xor")
TRUTH = {"nand": lambda x, y: ~(x & y), "and": lambda x, y: x & y, "or": lambda x, y: x | y, "xor": lambda x, y: x ^ y}


def _load():
    import torus_fhe_b200 as T
    params, bsk, ksk, lwe = T.interchange.read_keys(os.path.join(FIX, "keys.bin"))
    assert lwe is not None, "dump_fixture.jl writes the secret keys"
    ct = {name: T.interchange.read_ciphertexts(os.path.join(FIX, name + ".bin")) for name in ("x", "y", "boot") + GATES}
    plain = np.loadtxt(os.path.join(FIX, "plain.txt"), dtype=np.int64).astype(bool)
    return params, bsk, ksk, lwe, ct, plain


def _phase(lwe, a, b):
    return (b.astype(np.int64) - (a.astype(np.int64) * lwe[None].astype(np.int64)).sum((-1, -2))).astype(np.int32)


def _oracle_keys(oracle, params, bsk, ksk):
    prm = dict(n=params.lwe_size, N=params.rlwe_polynomial_degree, k=params.max_parties, l=params.gsw_decomp_length, bgbit=params.gsw_log2_base,
               t=params.ks_decomp_length, basebit=params.ks_log2_base, sigma_lwe=params.lwe_noise_stddev, sigma_gsw=params.gsw_noise_stddev,
               sigma_ks=params.ks_noise_stddev)
    return oracle.KeySet(prm, raw_bsk=np.stack(bsk), raw_ksk=np.stack(ksk))


def test_reference_outputs_decrypt_to_the_truth_table():
    params, bsk, ksk, lwe, ct, plain = _load()
    x, y = plain[:, 0], plain[:, 1]
    assert np.array_equal(_phase(lwe, *ct["x"]) > 0, x) and np.array_equal(_phase(lwe, *ct["y"]) > 0, y)
    for g in GATES:
        ph = _phase(lwe, *ct[g]).astype(np.float64) / 2 ** 32
        assert np.mean((ph > 0) == TRUTH[g](x, y)) >= 0.9, g          # the scheme's own failure rate is ~1e-3 per gate
        assert np.all(np.abs(np.abs(ph) - 0.125) < 0.12), g


def test_oracle_agrees_with_the_reference_on_its_own_bytes(oracle):
    params, bsk, ksk, lwe, ct, plain = _load()
    ks = _oracle_keys(oracle, params, bsk, ksk)
    gid = {"nand": oracle.GATE_NAND, "and": oracle.GATE_AND, "or": oracle.GATE_OR, "xor": oracle.GATE_XOR}
    close = total = 0
    for g in GATES:
        ref_ph = _phase(lwe, *ct[g])
        ex = ks.gate_batch(oracle.EXACT_NTT, gid[g], ct["x"], ct["y"])
        ff = ks.gate_batch(oracle.FFT, gid[g], ct["x"], ct["y"])
        ex_ph, ff_ph = _phase(lwe, *ex), _phase(lwe, *ff)
        assert np.array_equal(ex_ph > 0, ref_ph > 0) and np.array_equal(ff_ph > 0, ref_ph > 0), g
        d_ref = np.abs((ff_ph - ref_ph).astype(np.int32).astype(np.float64)) / 2 ** 32
        d_own = np.abs((ff_ph - ex_ph).astype(np.int32).astype(np.float64)) / 2 ** 32
        same = d_own <= 2.0 ** -20                                      # runs whose digit streams did not flip
        close += int(np.sum(d_ref[same] <= 2.0 ** -20))
        total += int(np.sum(same))
    print(f"FFT restatement within 2^-20 of the reference's FFT path on {close} of {total} flip-free gates")
    assert total > 0 and close >= 0.5 * total


@pytest.mark.gpu
def test_engine_on_the_reference_bytes(oracle):
    import torus_fhe_b200 as T
    params, bsk, ksk, lwe, ct, plain = _load()
    ks = _oracle_keys(oracle, params, bsk, ksk)
    eng = T.Engine(params, device=0)
    eng.load_keys(bsk, ksk)
    gid = {"nand": T._cabi.GATE_NAND, "and": T._cabi.GATE_AND, "or": T._cabi.GATE_OR, "xor": T._cabi.GATE_XOR}
    oid = {"nand": oracle.GATE_NAND, "and": oracle.GATE_AND, "or": oracle.GATE_OR, "xor": oracle.GATE_XOR}
    for g in GATES:
        oa, ob = eng.ctx.gate_batch(gid[g], ct["x"], ct["y"])
        ra, rb = ks.gate_batch(oracle.EXACT_NTT, oid[g], ct["x"], ct["y"])
        assert np.array_equal(oa, ra) and np.array_equal(ob, rb), g
        assert np.array_equal(_phase(lwe, oa, ob) > 0, _phase(lwe, *ct[g]) > 0), g
    eng.close()
