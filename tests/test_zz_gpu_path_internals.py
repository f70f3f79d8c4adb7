"""GPU test of the path's internal entry points in the host mirror (tgsw_extern_mul_3gen, mk_mux_rotate_3gen, mk_blind_rotate_3gen,
rlwe_extract_sample_64, tgsw_encrypt_3gen; 3gen_mk_internals.jl:59-95, tgsw_3gen.jl:41-113): the blind rotation driven one external
product at a time from the host equals the fused kernel bit for bit, at the 2-party default parameters with product-generated keys.
(Runs last: everything it calls on the GPU is covered by earlier files; this one checks the stage-by-stage surface.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    import torus_fhe_b200 as T
    rng = np.random.default_rng(0xB200_0A12)
    params = T.mktfhe_parameters_2party_3gen
    k = params.max_parties
    tg, rl = T.tgsw_parameters(params), T.rlwe_parameters(params)
    sk = [T.SecretKey_3gen(rng, params) for _ in range(k)]
    rk = [T.RLweKey(rng, rl, True) for _ in range(k)]
    crp = T.CRP_3gen(rng, tg, rl, True)
    pk = [T.PublicKey(rng, rk[i], params.gsw_noise_stddev, crp, tg, 1) for i in range(k)]
    cpk = T.CommonPubKey_3gen(pk, params, k)
    bk = [T.TransformedBootstrapKeyPart_3gen(T.BootstrapKeyPart_3gen(rng, sk[i].key, params.gsw_noise_stddev, crp, cpk, tg, rl, 1)) for i in range(k)]
    ks = [T.KeyswitchKey(rng, params.ks_noise_stddev, T.keyswitch_parameters(params), sk[i].key, rk[i]) for i in range(k)]
    T.engine_for(bk, ks)
    return T, rng, params, sk, rk, crp, cpk, bk, ks


def test_stage_by_stage_blind_rotation_equals_the_fused_kernel(world):
    T, rng, params, sk, rk, crp, cpk, bk, ks = world
    N, rl = params.rlwe_polynomial_degree, T.rlwe_parameters(params)
    mu = T.encode_message64(1, 8)
    x = T.mk_encrypt_3gen(rng, sk, True)
    u = T.mk_bootstrap_wo_keyswitch_3gen(bk, mu, x)                       # one launch: the whole loop inside the kernel
    barb, bara = T.decode_message(x.b, 2 * N), T.decode_message(x.a, 2 * N)
    acc = T.rlwe_noiseless_trivial(T.mul_by_monomial(np.full(N, mu, np.int64), -int(barb)), rl)     # 3gen_mk_internals.jl:91-92
    acc = T.mk_blind_rotate_3gen(acc, bk, bara)                           # k n launches of the external-product hook
    v = T.rlwe_extract_sample_64(acc)
    assert np.array_equal(v.a, u.a) and int(v.b) == int(u.b)
    out = T.mk_keyswitch_3gen(ks, v)
    assert bool(T.mk_decrypt_3gen(sk, out)) is True


def test_external_product_scales_the_phase_by_the_key_bit(world):
    """tgsw_extern_mul_3gen(acc, bk[p].gsw_key[j]) multiplies the accumulator's phase (under the joint RLWE key) by LWE key bit j of
    party p, and tgsw_encrypt_3gen builds exactly such samples (tgsw_3gen.jl:41-113)."""
    T, rng, params, sk, rk, crp, cpk, bk, ks = world
    N, tg, rl = params.rlwe_polynomial_degree, T.tgsw_parameters(params), T.rlwe_parameters(params)
    Z = sum(r.key.astype(np.int64) for r in rk)

    def phase(mask, body):
        with np.errstate(over="ignore"):
            return body - T.negacyclic_mul(Z, mask)

    acc = T.RLweSample(rl, rng.integers(-2 ** 63, 2 ** 63 - 1, size=(2, N), dtype=np.int64))
    ph_in = phase(acc.a[0], acc.a[1])
    for p in range(params.max_parties):
        for j in (0, 1, 2, params.lwe_size - 1):
            bit = int(sk[p].key.key[j])
            out = T.tgsw_extern_mul_3gen(acc, bk[p].tgsw_samples[j])
            with np.errstate(over="ignore"):
                err = (phase(out.a[0], out.a[1]) - bit * ph_in).astype(np.float64) / 2.0 ** 64
            assert np.abs(err).max() < 2.0 ** -6, (p, j, bit, np.abs(err).max())
    for m in (0, 1):
        s = T.tgsw_encrypt_3gen(rng, m, params.gsw_noise_stddev, cpk, crp)
        for q in range(tg.decomp_length):
            g = np.int64(tg.gadget_values[q])
            exp = np.zeros(N, np.int64)
            exp[0] = m * g
            with np.errstate(over="ignore"):
                e1 = (phase(s.part_4[q], s.part_1[q]) - exp).astype(np.float64) / 2.0 ** 64
                e2 = (phase(s.part_3[q], s.part_2[q]) + (m * g) * Z).astype(np.float64) / 2.0 ** 64
            assert max(np.abs(e1).max(), np.abs(e2).max()) < 2.0 ** -21, (m, q)
