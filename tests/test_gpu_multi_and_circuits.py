"""GPU tests added in round 2:
  * one C-ABI context spanning several GPUs (mktfhe_create_multi) through the Python binding: keys loaded once and broadcast by the
    library, host-pointer batches sharded by the library, results byte-identical to the one-GPU context.  On a one-GPU box the
    device is listed twice (two replicas sharing it): the same broadcast / sharding / threading code runs.
  * the integer circuits that round 1 only wired on the CPU with mocked gates -- mk_int_mul_3gen, mk_grt_3gen, mk_leq_3gen
    (3gen_mk_gates.jl:258-277, 312-362) -- on real ciphertexts, on a low-noise parameter set so that thousands of gates decrypt
    deterministically."""
import numpy as np
import pytest

from conftest import make_engine

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    n = torch.cuda.device_count()
    return list(range(n)) if n > 1 else [0, 0]


def test_multi_device_context_matches_single_device(oracle, keys2, engine2):
    import torus_fhe_b200 as T
    ks = keys2
    p = ks.prm
    sp = T.SchemeParameters_3gen(p.n, p.sigma_lwe, p.N, 1, False, p.l, p.bgbit, p.sigma_gsw, p.t, p.basebit, p.sigma_ks, p.k)
    devs = _devices()
    eng = T.Engine(sp, devices=devs)
    try:
        assert eng.ctx.device_count() == len(devs) and eng.devices == devs
        with pytest.raises(T.MktfheError):              # not finalized yet
            eng.ctx.gate_batch(T._cabi.GATE_NAND, ks.encrypt(np.zeros(2, np.uint8), 1), ks.encrypt(np.zeros(2, np.uint8), 2))
        eng.load_keys([ks.bsk[i] for i in range(p.k)], [ks.ksk[i] for i in range(p.k)])
        d = eng.ctx.describe()
        assert d["devices"] == devs and d["key_broadcast"] in ("p2p", "nccl")
        G = 2 * len(devs) + 3                            # slices of different sizes
        rng = np.random.default_rng(5)
        bits = rng.integers(0, 2, (3, G)).astype(np.uint8)
        x, y, z = (ks.encrypt(bits[i], 10 + i) for i in range(3))
        covered = 0
        for i in range(len(devs)):
            lo, hi = eng.ctx.shard_bounds(G, i)
            assert lo == covered and hi - lo in (G // len(devs), G // len(devs) + 1)
            covered = hi
        assert covered == G
        for gate, zz in ((T._cabi.GATE_NAND, None), (T._cabi.GATE_XOR, None), (T._cabi.GATE_AND3, z)):
            ma, mb = eng.ctx.gate_batch(gate, x, y, zz)
            sa, sb = engine2.ctx.gate_batch(gate, x, y, zz)
            assert np.array_equal(ma, sa) and np.array_equal(mb, sb), gate
        ids = (np.arange(G) % 4).astype(np.int32)
        ma, mb = eng.ctx.gate_batch_mixed(ids, x, y)
        sa, sb = engine2.ctx.gate_batch_mixed(ids, x, y)
        assert np.array_equal(ma, sa) and np.array_equal(mb, sb)
        ma, mb = eng.ctx.bootstrap_batch(1 << 61, *x)
        sa, sb = engine2.ctx.bootstrap_batch(1 << 61, *x)
        assert np.array_equal(ma, sa) and np.array_equal(mb, sb)
        assert np.array_equal(ks.decrypt(ma, mb), bits[0].astype(bool))
        # parity hooks shard too
        ext_m, acc_m = eng.ctx.blind_rotate_batch(1 << 61, *x, want_acc=True)
        ext_s, acc_s = engine2.ctx.blind_rotate_batch(1 << 61, *x, want_acc=True)
        assert np.array_equal(ext_m, ext_s) and np.array_equal(acc_m, acc_s)
        ka, kb = eng.ctx.keyswitch_batch(ext_m)
        assert np.array_equal(ka, ma) and np.array_equal(kb, mb)
        small, big = rng.integers(-1, 2, (G, 1024)), rng.integers(-2 ** 63, 2 ** 63 - 1, (G, 1024), dtype=np.int64)
        assert np.array_equal(eng.ctx.negacyclic_mul_batch(small, big), engine2.ctx.negacyclic_mul_batch(small, big))
        # the small operand of the exact product is range-checked (|a_i| <= 2^8 at N = 1024)
        with pytest.raises(T.MktfheError) as e:
            engine2.ctx.negacyclic_mul_batch(np.full((1, 1024), 1 << 9), big[:1])
        assert e.value.code == T._cabi.EINVAL
        # device-pointer calls belong to one GPU: the spanning handle addresses its first GPU, the others through their replicas
        import torch
        r = eng.ctx_on(devs[-1])
        xd = T.MKLweSampleGPU.from_host(T.MKLweSample(None, x[0], x[1]), devs[-1])
        yd = T.MKLweSampleGPU.from_host(T.MKLweSample(None, y[0], y[1]), devs[-1])
        oa = torch.empty_like(xd.a); ob = torch.empty_like(xd.b)
        torch.cuda.synchronize(devs[-1])
        r.gate_batch_dev(T._cabi.GATE_NAND, G, xd.a.data_ptr(), xd.b.data_ptr(), yd.a.data_ptr(), yd.b.data_ptr(), 0, 0, oa.data_ptr(), ob.data_ptr())
        torch.cuda.synchronize(devs[-1])
        sa, sb = engine2.ctx.gate_batch(T._cabi.GATE_NAND, x, y)
        assert np.array_equal(oa.cpu().numpy(), sa) and np.array_equal(ob.cpu().numpy(), sb)
        with pytest.raises(ValueError):
            eng.ctx_on(63)
        assert eng.ctx.launch_count() > 0 and eng.ctx.last_kernel_ms()[0] > 0
    finally:
        eng.close()


def test_multi_device_nccl_broadcast(oracle):
    """MKTFHE_B200_BCAST=nccl: grouped ncclBroadcast of the key buffers inside mktfhe_finalize_keys (needs distinct GPUs)."""
    import os
    import torch
    import torus_fhe_b200 as T
    if torch.cuda.device_count() < 2:
        pytest.skip("the NCCL broadcast needs two distinct GPUs")
    prm = dict(oracle.PARAMS_2PARTY, n=16)
    ks = oracle.KeySet(prm, seed=3, nthreads=4)
    sp = T.SchemeParameters_3gen(ks.n, prm["sigma_lwe"], ks.N, 1, False, ks.l, prm["bgbit"], prm["sigma_gsw"], ks.t, prm["basebit"], prm["sigma_ks"], ks.k)
    os.environ["MKTFHE_B200_BCAST"] = "nccl"
    try:
        eng = T.Engine(sp, devices="all")
        eng.load_keys([ks.bsk[i] for i in range(ks.k)], [ks.ksk[i] for i in range(ks.k)])
        assert eng.ctx.describe()["key_broadcast"] == "nccl"
        bits = np.arange(9) % 2
        x, y = ks.encrypt(bits.astype(np.uint8), 1), ks.encrypt((1 - bits).astype(np.uint8), 2)
        oa, ob = eng.ctx.gate_batch(T._cabi.GATE_NAND, x, y)
        ra, rb = ks.gate_batch(oracle.EXACT_NTT, oracle.GATE_NAND, x, y)
        assert np.array_equal(oa, ra) and np.array_equal(ob, rb)
        eng.close()
    finally:
        del os.environ["MKTFHE_B200_BCAST"]


@pytest.fixture(scope="module")
def quiet_world():
    """Three gadget levels and a finer key switch: noise far below the gate margins (as tests/test_gpu_api.py::quiet_world)."""
    import torus_fhe_b200 as T
    rng = np.random.default_rng(0xB200_0C1F)
    params = T.SchemeParameters_3gen(300, 2.0 ** -20, 1024, 1, False, 3, 7, 2.0 ** -45, 5, 3, 2.0 ** -20, 2)
    k = params.max_parties
    sk = [T.SecretKey_3gen(rng, params) for _ in range(k)]
    rk = [T.RLweKey(rng, T.rlwe_parameters(params), True) for _ in range(k)]
    crp = T.CRP_3gen(rng, T.tgsw_parameters(params), T.rlwe_parameters(params), True)
    pk = [T.PublicKey(rng, rk[i], params.gsw_noise_stddev, crp, T.tgsw_parameters(params), 1) for i in range(k)]
    cpk = T.CommonPubKey_3gen(pk, params, k)
    bk = [T.TransformedBootstrapKeyPart_3gen(T.BootstrapKeyPart_3gen(rng, sk[i].key, params.gsw_noise_stddev, crp, cpk,
                                                                      T.tgsw_parameters(params), T.rlwe_parameters(params), 1)) for i in range(k)]
    ks = [T.KeyswitchKey(rng, params.ks_noise_stddev, T.keyswitch_parameters(params), sk[i].key, rk[i]) for i in range(k)]
    yield T, rng, sk, bk, ks
    T.release_engine(bk, ks)


def test_comparators_grt_leq_on_ciphertexts(quiet_world):
    """mk_grt_3gen / mk_leq_3gen (3gen_mk_gates.jl:258-277) together with less / geq, WIDTH = 8, 24 instances per level, including
    equal operands and the extremes of the signed range that does not overflow a - b."""
    T, rng, sk, bk, ks = quiet_world
    W, I = 8, 24
    a, b = rng.integers(-60, 60, I), rng.integers(-60, 60, I)
    a[:4], b[:4] = [5, -5, 63, -64], [5, -5, -64, 63]
    ca, cb = T.mk_int_encrypt_3gen(rng, sk, a, W), T.mk_int_encrypt_3gen(rng, sk, b, W)
    one = T.mk_encrypt_3gen(rng, sk, np.ones(I, bool))
    dec = lambda c: np.asarray(T.mk_decrypt_3gen(sk, c))
    assert np.array_equal(dec(T.mk_grt_3gen(bk, ks, ca, cb, one, W)), a > b)
    assert np.array_equal(dec(T.mk_leq_3gen(bk, ks, ca, cb, one, W)), a <= b)
    assert np.array_equal(dec(T.mk_less_3gen(bk, ks, ca, cb, one, W)), a < b)
    assert np.array_equal(dec(T.mk_geq_3gen(bk, ks, ca, cb, one, W)), a >= b)
    # device-resident operands give the same ciphertext bits
    g = lambda bits: [T.MKLweSampleGPU.from_host(c) for c in bits]
    h = T.mk_leq_3gen(bk, ks, ca, cb, one, W)
    d = T.mk_leq_3gen(bk, ks, g(ca), g(cb), T.MKLweSampleGPU.from_host(one), W).cpu()
    assert np.array_equal(d.a, h.a) and np.array_equal(d.b, h.b)


def test_int_mul_on_ciphertexts(quiet_world):
    """mk_int_mul_3gen (3gen_mk_gates.jl:312-362) on real ciphertexts, WIDTH = 4, against the plaintext model of the reference's own
    wiring (it adds partial-product row `ctr` twice instead of row WIDTH: replicated) -- the same model the CPU
    wiring test uses (tests/test_host_api.py)."""
    T, rng, sk, bk, ks = quiet_world
    Wm, I = 4, 16
    a, b = rng.integers(0, 16, I), rng.integers(0, 16, I)
    ca, cb = T.mk_int_encrypt_3gen(rng, sk, a, Wm), T.mk_int_encrypt_3gen(rng, sk, b, Wm)
    zero = T.mk_encrypt_3gen(rng, sk, np.zeros(I, bool))
    res = T.mk_int_mul_3gen(bk, ks, ca, cb, zero, Wm)
    got = sum(np.asarray(T.mk_decrypt_3gen(sk, res[i])).astype(np.int64) << i for i in range(Wm))
    row = lambda i: ((b >> (i - 1)) & 1) * a                      # BArr[i, :] as an integer, 1-based i
    tmp, ctr, low = row(1) >> 1, 1, row(1) & 1
    for i in range(2, Wm):
        s = tmp + row(i)
        low |= (s & 1) << (i - 1)
        tmp, ctr = s >> 1, i
    s = tmp + row(ctr)
    expect = (low | (s << ctr)) & ((1 << Wm) - 1)
    assert np.array_equal(got, expect)
    # where the reference's wiring coincides with a true product (b < 4: rows 3 and 4 are zero, and row `ctr` = row 3), check a * b
    m = b < 4
    assert np.array_equal(got[m], (a[m] * b[m]) & 15)


def test_keyswitch_key_generated_on_the_device(oracle):
    """mktfhe_generate_ksk (keyswitch.jl:14-41 on the GPU): every row, read back from the device key, is an LWE encryption under s of
    (z_i h) << (32 - j basebit) (:35) with Gaussian noise of the requested deviation, re-centred to zero mean over the key (:28-29);
    the uniform part looks uniform; two parties / two seeds give different rows; and gates through the generated key decrypt."""
    import torch
    import torus_fhe_b200 as T
    rng = np.random.default_rng(0xB200_0D0E)
    params = T.mktfhe_parameters_2party_3gen
    n, N, t, bb = params.lwe_size, params.rlwe_polynomial_degree, params.ks_decomp_length, params.ks_log2_base
    B1, k = (1 << bb) - 1, params.max_parties
    sk = [T.SecretKey_3gen(rng, params) for _ in range(k)]
    rk = [T.RLweKey(rng, T.rlwe_parameters(params), True) for _ in range(k)]
    crp = T.CRP_3gen(rng, T.tgsw_parameters(params), T.rlwe_parameters(params), True)
    pk = [T.PublicKey(rng, rk[i], params.gsw_noise_stddev, crp, T.tgsw_parameters(params), 1) for i in range(k)]
    cpk = T.CommonPubKey_3gen(pk, params, k)
    bk = [T.TransformedBootstrapKeyPart_3gen(T.BootstrapKeyPart_3gen(rng, sk[i].key, params.gsw_noise_stddev, crp, cpk,
                                                                      T.tgsw_parameters(params), T.rlwe_parameters(params), 1)) for i in range(k)]
    ks = [T.KeyswitchKey.on_device(rng, params.ks_noise_stddev, T.keyswitch_parameters(params), sk[i].key, rk[i]) for i in range(k)]
    assert ks[0].key is None
    eng = T.engine_for(bk, ks)
    try:
        _, ksk_t = eng.key_tensors()
        stride = (n + 1 + 3) & ~3
        rows = ksk_t.cpu().numpy().view(np.int32).reshape(k, N, t, B1, stride)
        assert not np.array_equal(rows[0], rows[1])
        for p in range(k):
            s, z = sk[p].key.key.astype(np.int64), rk[p].key.astype(np.int64)
            a, b = rows[p, ..., :n].astype(np.int64), rows[p, ..., n].astype(np.int64)
            assert np.all(rows[p, ..., n + 1:] == 0)                                       # padding
            h = np.arange(1, B1 + 1, dtype=np.int64)[None, None, :]
            sh = (32 - np.arange(1, t + 1) * bb)[None, :, None]
            msg = (z[:, None, None] * h) << sh
            err = ((b - (a * s).sum(-1) - msg + 2 ** 31) % 2 ** 32 - 2 ** 31) / 2.0 ** 32  # phase - message, as a fraction of the torus
            assert abs(err.mean()) < 2.0 ** -32 * 2                                        # re-centred (up to the truncation of dtot32)
            assert 0.97 < err.std() / params.ks_noise_stddev < 1.03
            u = rows[p, ::64, :, :, :n].astype(np.float64) / 2.0 ** 32                     # a sample of the uniform words
            assert abs(u.mean()) < 2e-3 and abs(u.std() - 12 ** -0.5) < 2e-3
            assert len(np.unique(rows[p, :8, :, :, :n])) > 0.999 * rows[p, :8, :, :, :n].size
        bits = rng.integers(0, 2, (2, 64)).astype(bool)
        x, y = T.mk_encrypt_3gen(rng, sk, bits[0]), T.mk_encrypt_3gen(rng, sk, bits[1])
        assert np.array_equal(T.mk_decrypt_3gen(sk, T.mk_gate_nand_3gen(bk, ks, x, y)), ~(bits[0] & bits[1]))
        assert np.array_equal(T.mk_decrypt_3gen(sk, T.mk_gate_xor_3gen(bk, ks, x, y)), bits[0] ^ bits[1])
        # argument checks
        with pytest.raises(T.MktfheError):
            eng.ctx.generate_ksk(5, sk[0].key.key, rk[0].key, 1e-4, 1)
        with pytest.raises(ValueError):
            eng.ctx.generate_ksk(0, sk[0].key.key[:-1], rk[0].key, 1e-4, 1)
    finally:
        T.release_engine(bk, ks)
