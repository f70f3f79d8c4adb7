"""The KMS multi-key scheme's CPU restatement (oracle/kms_oracle.py; `mk_gate_nand_new` / `mk_bootstrap_new`, new_mk_internals.jl) on a toy ring:
the decrypted NAND truth table with two and three parties (the shape of multikey_new.jl:40-60), the TLev blind rotation's phase, and the
hybrid product's algebra.  There is no engine for this scheme yet; this pins the reading of the reference that an
engine would be built against."""
import numpy as np
import pytest

from oracle import kms_oracle as K


def params(k):
    return dict(n=12, N=64, k=k, gsw=(3, 13), lev=(2, 7), uni=(2, 13), t=8, basebit=2, sigma_gsw=2.0 ** -55, sigma_uni=2.0 ** -55, sigma_ks=2.0 ** -22)


@pytest.mark.parametrize("k", [2, 3])
def test_kms_nand_truth_table(k):
    rng = np.random.default_rng(40 + k)
    keys = K.keygen(rng, params(k))
    for a in (False, True):
        for b in (False, True):
            x, y = K.encrypt(rng, keys, a, 2.0 ** -20), K.encrypt(rng, keys, b, 2.0 ** -20)
            out = K.gate_nand(keys, x, y)
            assert out[0].shape == (k, keys["n"])
            assert K.decrypt(keys, out) == (not (a and b)), (a, b, K.phase(keys, out) / 2.0 ** 32)
            assert abs(abs(K.phase(keys, out)) / 2.0 ** 32 - 0.125) < 0.03


def test_tlev_blind_rotation_encrypts_the_monomial():
    """mk_ith_blind_rotate (new_mk_internals.jl:212-225): TLev row q is an RLWE encryption, under the party's own key R, of
    g_q X^(sum_j bara_j s_j) -- the phase body - R mask equals that monomial up to noise far below g_q."""
    rng = np.random.default_rng(7)
    prm = params(2)
    keys = K.keygen(rng, prm)
    N = prm["N"]
    for party in range(2):
        bara = rng.integers(-N, N, prm["n"]).astype(np.int32)
        lev = K.ith_blind_rotate(keys, party, bara)
        e = int((bara.astype(np.int64) * keys["s"][party]).sum())
        g = K.gadget(*prm["lev"])
        for q in range(prm["lev"][0]):
            with np.errstate(over="ignore"):
                phase = lev[q, 1] - K.negacyclic_mul(keys["R"][party], lev[q, 0])
                want = K.mul_by_monomial(np.concatenate([[g[q]], np.zeros(N - 1, np.int64)]), e)
                err = (phase - want).astype(np.float64)
            assert np.abs(err).max() < float(g[q]) * 2.0 ** -20, (party, q)


def test_hybrid_product_multiplies_the_phase_by_R():
    """UniProduct_new (new_mk_internals.jl:85-127): for a multi-key RLWE sample (a_1 .. a_k, b) with phase b - sum_i z_i a_i = m, the
    output has phase R_party m (up to noise) -- which is why mk_lev_rlwe_mul's f - UniProduct_new(e) turns a TLev product under R into a
    sample under the parties' keys."""
    rng = np.random.default_rng(9)
    prm = params(3)
    keys = K.keygen(rng, prm)
    k, N = prm["k"], prm["N"]
    m = rng.integers(-2 ** 62, 2 ** 62, N, dtype=np.int64)
    a = rng.integers(-2 ** 63, 2 ** 63 - 1, (k, N), dtype=np.int64)
    with np.errstate(over="ignore"):
        b = m.copy()
        for i in range(k):
            b = b + K.negacyclic_mul(keys["z"][i], a[i])
        for party in range(k):
            ua, ub = K.uni_product(a, b, keys, party)
            ph = ub.copy()
            for i in range(k):
                ph = ph - K.negacyclic_mul(keys["z"][i], ua[i])
            want = K.negacyclic_mul(keys["R"][party], m)
            err = (ph - want).astype(np.float64)
            # the (floor) truncation of the uni gadget dominates: up to 2^(64 - 26) per coefficient, times the N binary coefficients of R and
            # z, for k + 1 polynomials and the second decomposition: 2^38 N (k + 1) 2 = 2^47 here, against |m| up to 2^62
            assert np.abs(err).max() < 2.0 ** 49, party


def test_decompose_and_products_are_exact():
    rng = np.random.default_rng(3)
    N = 64
    c = rng.integers(-2 ** 63, 2 ** 63 - 1, N, dtype=np.int64)
    for l, bg in ((3, 13), (2, 7)):
        d = K.decompose(c, l, bg)
        assert d.min() >= -(1 << (bg - 1)) and d.max() < (1 << (bg - 1))
        with np.errstate(over="ignore"):
            back = sum(d[q] * K.gadget(l, bg)[q] for q in range(l))
            err = (c - back).astype(np.int64)
        # truncating decomposition: the remainder lies in [-g_l / 2, g_l / 2) up to the floor
        assert np.abs(err.astype(np.float64)).max() <= 2.0 ** (64 - l * bg)
    a = rng.integers(-8, 8, N)
    b = rng.integers(-2 ** 63, 2 ** 63 - 1, N, dtype=np.int64)
    want = [0] * N
    for i in range(N):
        for j in range(N):
            t = int(a[i]) * int(b[j])
            if i + j < N:
                want[i + j] += t
            else:
                want[i + j - N] -= t
    want = np.array([((w + 2 ** 63) % 2 ** 64) - 2 ** 63 for w in want], np.int64)
    assert np.array_equal(K.negacyclic_mul(a, b), want)
    assert np.array_equal(K.mul_by_monomial(b, 5)[5:], b[:-5]) and np.array_equal(K.mul_by_monomial(b, N + 5)[5:], -b[:-5])
