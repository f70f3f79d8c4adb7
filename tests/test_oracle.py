"""CPU tests of the oracle itself (the checker must be right before it checks anything).

The reference holds no golden vector for the 3gen path ("parity unpinned"), so the
oracle is pinned (i) functionally, as the reference's demos and measurement drivers do
(multikey_3gen.jl:66-92, measurements_us_simplified_3.jl:79-123: decrypt(gate(enc x, enc y)) == gate(x, y)),
(ii) by cross-checking its three multiplication back-ends, (iii) by hand-computed values of the scalar
semantics quoted from the reference source, and (iv) by the committed fixtures in tests/golden/.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_scalar_semantics(oracle):
    L = oracle.lib()
    # encode_message(mu, space) = mu << (32 - log2 space)   numeric-functions.jl:86-95
    assert L.mko_encode_message32(1, 8) == 1 << 29
    assert L.mko_encode_message32(-1, 8) == -(1 << 29)
    assert L.mko_encode_message32(1, 4) == 1 << 30
    assert L.mko_encode_message64(1, 8) == 1 << 61
    # decode_message(x, 2N) = (x + 2^20) >> 21 for N = 1024   numeric-functions.jl:70-73
    assert L.mko_decode_message32(0, 2048) == 0
    assert L.mko_decode_message32((1 << 20) - 1, 2048) == 0
    assert L.mko_decode_message32(1 << 20, 2048) == 1
    assert L.mko_decode_message32(-(1 << 20) - 1, 2048) == -1
    assert L.mko_decode_message32(2 ** 31 - 1, 2048) == -1024          # wrapping add
    assert L.mko_decode_message32(-2 ** 31, 2048) == -1024
    # t64tot32(d) = trunc(Int32, d / 2^32): toward zero, through Float64   numeric-functions.jl:109-111
    assert L.mko_t64tot32(-1) == 0                                      # NOT -1 (an arithmetic shift would give -1)
    assert L.mko_t64tot32(-(1 << 32)) == -1
    assert L.mko_t64tot32((1 << 32) - 1) == 0
    assert L.mko_t64tot32(-(1 << 32) - 1) == -1
    assert L.mko_t64tot32(1 << 61) == 1 << 29
    assert L.mko_t64tot32((5 << 32) + 123) == 5
    assert L.mko_t64tot32(-2 ** 63) == -2 ** 31
    # Float64 rounding of the Int64 happens BEFORE the division: 2^62 + 2^32 - 1 rounds up to 2^62 + 2^32
    assert L.mko_t64tot32((1 << 62) + (1 << 32) - 1) == (1 << 30) + 1
    # gadget offset of the 2-party set: (Bg/2) * (2^57 + 2^50) wrapped = -2^63 + 2^56   tgsw.jl:24-30
    assert L.mko_gadget_offset(2, 7) == -2 ** 63 + 2 ** 56


def test_decompose_matches_formula(oracle, rng):
    # tgsw.jl:125-137, independent numpy restatement
    for (l, bg) in [(2, 7), (3, 6), (4, 4)]:
        poly = rng.integers(-2 ** 63, 2 ** 63 - 1, size=256, dtype=np.int64)
        got = oracle.decompose(poly, l, bg)
        off = np.uint64(sum((1 << (64 - q * bg)) << (bg - 1) for q in range(1, l + 1)) & (2 ** 64 - 1))
        v = (poly.view(np.uint64) + off).view(np.int64)
        for q in range(1, l + 1):
            exp = ((v >> (64 - q * bg)) & ((1 << bg) - 1)) - (1 << (bg - 1))
            assert np.array_equal(got[q - 1], exp)
        assert got.min() >= -(1 << (bg - 1)) and got.max() < (1 << (bg - 1))
        # reconstruction error is below the last gadget level
        rec = sum(got[q - 1].astype(object) * (1 << (64 - q * bg)) for q in range(1, l + 1))
        err = np.array([(int(r) - int(p) + 2 ** 63) % 2 ** 64 - 2 ** 63 for r, p in zip(rec, poly)], dtype=object)
        assert max(abs(int(e)) for e in err) <= 1 << (64 - l * bg)


def test_mul_by_monomial(oracle, rng):
    N = 64
    p = rng.integers(-2 ** 63, 2 ** 63 - 1, size=N, dtype=np.int64)
    assert np.array_equal(oracle.mul_by_monomial(p, 0), p)
    one = oracle.mul_by_monomial(p, 1)
    assert one[0] == np.int64(-p[N - 1]) and np.array_equal(one[1:], p[:-1])
    assert np.array_equal(oracle.mul_by_monomial(oracle.mul_by_monomial(p, 5), -5), p)
    assert np.array_equal(oracle.mul_by_monomial(p, 2 * N + 3), oracle.mul_by_monomial(p, 3))
    with np.errstate(over="ignore"):
        assert np.array_equal(oracle.mul_by_monomial(p, N + 3), -oracle.mul_by_monomial(p, 3))


def test_negacyclic_backends_agree(oracle, rng):
    for N in (64, 1024):
        digits = rng.integers(-64, 64, size=N, dtype=np.int64)
        key = rng.integers(-2 ** 63, 2 ** 63 - 1, size=N, dtype=np.int64)
        exact = oracle.negacyclic_mul(digits, key, oracle.EXACT_SCHOOLBOOK)
        # independent check of the schoolbook itself on a few coefficients with Python integers
        for i in (0, 1, N - 1):
            s = 0
            for j in range(N):
                kk = (i - j) % (2 * N)
                c = int(key[kk % N]) * (-1 if kk >= N else 1)
                s += int(digits[j]) * c
            assert (s - int(exact[i])) % 2 ** 64 == 0
        assert np.array_equal(oracle.negacyclic_mul(digits, key, oracle.EXACT_NTT), exact)
        fft = oracle.negacyclic_mul(digits, key, oracle.FFT)
        with np.errstate(over="ignore"):
            diff = (fft - exact).astype(np.float64)
        assert np.abs(diff).max() < 2.0 ** 26, "Float64 FFT product should be within ~2^-38 of the torus"


def test_toy_truth_tables_all_backends(oracle, toy_keys):
    ks = toy_keys
    gates = {oracle.GATE_NAND: lambda x, y: not (x and y), oracle.GATE_OR: lambda x, y: x or y,
             oracle.GATE_AND: lambda x, y: x and y, oracle.GATE_XOR: lambda x, y: x != y}
    xs = np.array([0, 0, 1, 1] * 4, np.uint8)
    ys = np.array([0, 1, 0, 1] * 4, np.uint8)
    x, y = ks.encrypt(xs, 101), ks.encrypt(ys, 202)
    assert np.array_equal(ks.decrypt(*x), xs.astype(bool))
    ref_out = {}
    for backend in (oracle.EXACT_SCHOOLBOOK, oracle.EXACT_NTT, oracle.FFT):
        for g, fn in gates.items():
            oa, ob = ks.gate_batch(backend, g, x, y, nthreads=4)
            exp = np.array([fn(bool(a), bool(b)) for a, b in zip(xs, ys)])
            assert np.array_equal(ks.decrypt(oa, ob), exp), (backend, g)
            ph = ks.phase(oa, ob).astype(np.float64) / 2 ** 32
            assert np.all(np.abs(np.abs(ph) - 0.125) < 0.04)
            if backend == oracle.EXACT_SCHOOLBOOK:
                ref_out[g] = (oa, ob)
            elif backend == oracle.EXACT_NTT:   # the two exact back-ends must agree bit for bit
                assert np.array_equal(oa, ref_out[g][0]) and np.array_equal(ob, ref_out[g][1])


def test_toy_and3(oracle, toy_keys):
    ks = toy_keys
    bits = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)], np.uint8)
    x, y, z = ks.encrypt(bits[:, 0], 1), ks.encrypt(bits[:, 1], 2), ks.encrypt(bits[:, 2], 3)
    oa, ob = ks.gate_batch(oracle.EXACT_NTT, oracle.GATE_AND3, x, y, z, nthreads=4)
    # Reference behaviour, replicated not fixed: the prologue -1/4 + x + y + z (3gen_mk_gates.jl:55-64) wraps for
    # (0,0,0): -1/4 - 3/8 = -5/8 = +3/8 (mod 1), so mk_gate_3and_3gen(false,false,false) decrypts to TRUE.
    exp = bits.all(axis=1) | ~bits.any(axis=1)
    assert np.array_equal(ks.decrypt(oa, ob), exp)


def test_fullsize_extprod_backends(oracle, keys2, rng):
    """2-party defaults: exact NTT == exact schoolbook bit for bit; Float64 FFT within 2^-38 of the torus (SURVEY H9 i)."""
    acc = rng.integers(-2 ** 63, 2 ** 63 - 1, size=(2, 1024), dtype=np.int64)
    for (party, j) in [(0, 0), (1, 519)]:
        sb = keys2.extprod(oracle.EXACT_SCHOOLBOOK, party, j, acc)
        assert np.array_equal(keys2.extprod(oracle.EXACT_NTT, party, j, acc), sb)
        with np.errstate(over="ignore"):
            d = (keys2.extprod(oracle.FFT, party, j, acc) - sb).astype(np.float64)
        assert np.abs(d).max() < 2.0 ** 26


def test_fullsize_nand_and_golden(oracle, keys2):
    """One full-size NAND truth table through the exact-NTT and the FFT back-end, checked against the committed fixture."""
    with open(os.path.join(GOLDEN, "nand_2party.json")) as f:
        gold = json.load(f)
    assert hashlib.sha256(keys2.bsk.tobytes()).hexdigest() == gold["bsk_sha256"]
    assert hashlib.sha256(keys2.ksk.tobytes()).hexdigest() == gold["ksk_sha256"]
    data = np.load(os.path.join(GOLDEN, "nand_2party.npz"))
    xs, ys = data["x_bits"], data["y_bits"]
    x, y = keys2.encrypt(xs, gold["x_seed"]), keys2.encrypt(ys, gold["y_seed"])
    assert np.array_equal(x[0], data["xa"]) and np.array_equal(x[1], data["xb"])
    oa, ob = keys2.gate_batch(oracle.EXACT_NTT, oracle.GATE_NAND, x, y)
    assert np.array_equal(oa, data["out_a"]) and np.array_equal(ob, data["out_b"])
    exp = ~(xs.astype(bool) & ys.astype(bool))
    assert np.array_equal(keys2.decrypt(oa, ob), exp)
    fa, fb = keys2.gate_batch(oracle.FFT, oracle.GATE_NAND, x, y)
    assert np.array_equal(keys2.decrypt(fa, fb), exp)
    ph = keys2.phase(oa, ob).astype(np.float64) / 2 ** 32
    assert np.all(np.abs(np.abs(ph) - 0.125) < 0.05)


def test_exact_ntt_backend_stays_exact_for_wide_gadget_digits(oracle):
    """The N = 2048 sets have 18..26-bit gadget digits: the Goldilocks back-end then splits the key into three 22-bit limbs
    (ntt_limb_plan) and must still equal the schoolbook back-end bit for bit, bootstraps included."""
    import numpy as np
    for prm in (dict(n=6, N=2048, k=2, l=1, bgbit=26, t=4, basebit=3), dict(n=4, N=2048, k=2, l=2, bgbit=18, t=8, basebit=2)):
        ks = oracle.KeySet(dict(prm, sigma_lwe=2.0 ** -15.34, sigma_gsw=2.0 ** -62, sigma_ks=2.0 ** -15.34), seed=5, nthreads=4)
        r = np.random.default_rng(prm["bgbit"])
        acc = r.integers(-2 ** 63, 2 ** 63 - 1, size=(2, 2048), dtype=np.int64)
        assert np.array_equal(ks.extprod(oracle.EXACT_NTT, 1, 2, acc), ks.extprod(oracle.EXACT_SCHOOLBOOK, 1, 2, acc))
        a = r.integers(-2 ** 31, 2 ** 31, (2, ks.k, ks.n)).astype(np.int32)
        b = r.integers(-2 ** 31, 2 ** 31, 2).astype(np.int32)
        o1, o2 = ks.bootstrap_batch(oracle.EXACT_SCHOOLBOOK, 1 << 61, a, b), ks.bootstrap_batch(oracle.EXACT_NTT, 1 << 61, a, b)
        assert np.array_equal(o1[0], o2[0]) and np.array_equal(o1[1], o2[1])
