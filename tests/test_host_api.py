"""CPU tests of the host-side mirror of the reference API (torus-fhe_b200/tfhe3gen.py, circuits.py, engine.py):
everything that does not launch a kernel."""
import numpy as np
import pytest

import torus_fhe_b200 as T
from torus_fhe_b200 import circuits


def test_parameter_sets_match_reference():
    # 3-gen-mk-tfhe/src/mk_api.jl:32-38, 84-90, 140-146
    p = T.mktfhe_parameters_2party_3gen
    assert (p.lwe_size, p.rlwe_polynomial_degree, p.gsw_decomp_length, p.gsw_log2_base, p.ks_decomp_length, p.ks_log2_base, p.max_parties) == (520, 1024, 2, 7, 3, 3, 2)
    assert not p.rlwe_is32 and p.rlwe_mask_size == 1 and abs(p.lwe_noise_stddev - 2 ** -13.52) < 1e-15
    p = T.mktfhe_parameters_4party_3gen
    assert (p.lwe_size, p.gsw_decomp_length, p.gsw_log2_base, p.ks_decomp_length, p.ks_log2_base, p.max_parties) == (510, 3, 6, 5, 2, 4)
    p = T.mktfhe_parameters_8party_3gen
    assert (p.lwe_size, p.gsw_decomp_length, p.gsw_log2_base, p.ks_decomp_length, p.ks_log2_base, p.max_parties) == (540, 4, 4, 5, 2, 8)
    assert T.tgsw_parameters(T.mktfhe_parameters_2party_3gen).gadget_values == [1 << 57, 1 << 50]   # tgsw.jl:26
    # the N >= 2048 sets by name (mk_api.jl:214-322): N = 2048 is served by csrc/kernels2k.cuh, N = 4096 (512 parties) is rejected at mktfhe_create
    for k, n, N, l, bg in ((16, 590, 2048, 1, 26), (32, 620, 2048, 1, 26), (64, 650, 2048, 1, 25), (128, 670, 2048, 1, 24), (256, 740, 2048, 2, 18),
                           (512, 730, 4096, 1, 27)):
        p = getattr(T, f"mktfhe_parameters_{k}party_3gen")
        assert (p.max_parties, p.lwe_size, p.rlwe_polynomial_degree, p.gsw_decomp_length, p.gsw_log2_base) == (k, n, N, l, bg) and not p.rlwe_is32


def test_decode64_and_noise_calc():
    # numeric-functions.jl:75-78, 117-131
    assert int(T.decode_message64(T.encode_message64(3, 8), 8)) == 3 and int(T.decode_message64(np.int64(-5), 2048)) == 0
    assert int(T.decode_message64(np.int64(2 ** 63 - 1), 4)) == -2                        # wrapping add, arithmetic shift
    e = lambda f: np.int32(int(f * 2 ** 32))
    assert abs(T.noise_calc(e(0.125), e(0.13)) - 0.005) < 1e-9
    assert abs(T.noise_calc(e(0.125), e(-0.49)) - (1.0 - 0.49 - 0.125)) < 1e-9           # d < m - 0.5: wraps upwards
    assert abs(T.noise_calc(e(-0.125), e(-0.12)) - 0.005) < 1e-9
    assert abs(T.noise_calc(e(-0.125), e(0.45)) - (1.0 - 0.45 - 0.125)) < 1e-9           # d > m + 0.5
    assert T.noise_calc(np.int32(0), e(0.25)) == 0.25
    assert T.noise_calc(np.array([e(0.125)] * 2), np.array([e(0.13), e(0.12)])).shape == (2,)


def test_encode_decode_agree_with_oracle(oracle):
    L = oracle.lib()
    for mu, space in [(1, 8), (-1, 8), (1, 4), (-1, 4), (3, 8)]:
        assert int(T.encode_message(mu, space)) == L.mko_encode_message32(mu, space)
        assert int(T.encode_message64(mu, space)) == L.mko_encode_message64(mu, space)
    xs = np.array([0, 1, -1, 2 ** 31 - 1, -2 ** 31, (1 << 20) - 1, 1 << 20, -(1 << 20) - 1, 123456789], np.int32)
    got = T.decode_message(xs, 2048)
    assert [int(v) for v in got] == [L.mko_decode_message32(int(v), 2048) for v in xs]
    assert int(T.decode_message(np.int32(-5), 2048)) == L.mko_decode_message32(-5, 2048)


def test_mklwesample_linear_ops_wrap(rng):
    p = T.LweParams(8)
    x = T.MKLweSample(p, rng.integers(-2 ** 31, 2 ** 31, (2, 8)), np.int32(2 ** 31 - 5), 1.0)
    y = T.MKLweSample(p, rng.integers(-2 ** 31, 2 ** 31, (2, 8)), np.int32(100), 2.0)
    s = x + y
    assert int(s.b) == ((2 ** 31 - 5 + 100 + 2 ** 31) % 2 ** 32) - 2 ** 31 and s.current_variance == 3.0
    d = x - y
    assert np.array_equal(d.a, (x.a.astype(np.int64) - y.a).astype(np.int32))
    n = -x
    assert np.array_equal((n + x).a, np.zeros((2, 8), np.int32)) and int((n + x).b) == 0
    t = np.int32(2) * x                                   # Torus32 * sample, mk_internals.jl:50-51
    assert np.array_equal(t.a, (2 * x.a.astype(np.int64)).astype(np.int32)) and t.current_variance == 4.0
    tr = T.mk_lwe_noiseless_trivial(T.encode_message(1, 8), p, 2)
    assert tr.a.shape == (2, 8) and not tr.a.any() and int(tr.b) == 1 << 29
    st = T.MKLweSample.stack([x, y])
    assert st.b.shape == (2,) and np.array_equal(st[1].a, y.a)


def test_encrypt_decrypt_roundtrip_and_ints(rng):
    params = T.mktfhe_parameters_2party_3gen
    sk = [T.SecretKey_3gen(rng, params) for _ in range(2)]
    assert sk[0].key.key.shape == (520,) and set(np.unique(sk[0].key.key)) <= {0, 1}
    bits = rng.integers(0, 2, 200).astype(bool)
    ct = T.mk_encrypt_3gen(rng, sk, bits)
    assert ct.a.shape == (200, 2, 520) and ct.b.shape == (200,)
    assert np.array_equal(T.mk_decrypt_3gen(sk, ct), bits)
    ph = T.mk_lwe_phase(ct, [s.key for s in sk]).astype(np.float64) / 2 ** 32
    assert np.abs(np.abs(ph) - 0.125).max() < 0.001      # sigma_lwe = 2^-13.52
    one = T.mk_encrypt_3gen(rng, sk, True)
    assert one.b.shape == () and T.mk_decrypt_3gen(sk, one) is True
    # LSB-first two's complement, mk_api.jl:576-589, 612-633
    for v in (0, 1, 5, 127, -1, -128, -77):
        assert T.mk_int_decrypt_3gen(sk, T.mk_int_encrypt_3gen(rng, sk, v, 8), 8) == v
    vals = np.array([3, -4, 100, -100])
    assert np.array_equal(T.mk_int_decrypt_3gen(sk, T.mk_int_encrypt_3gen(rng, sk, vals, 8), 8), vals)


def test_keyswitch_key_generation_is_consistent(rng):
    """KeyswitchKey (keyswitch.jl:14-41) from the host mirror: every row decrypts under the output key to
    (z_i * h) << (32 - j * log2_base) up to the key-switch noise."""
    params = T.mktfhe_parameters_2party_3gen
    sk = T.SecretKey_3gen(rng, params)
    rk = T.RLweKey(rng, T.RLweParams(64, 1, False), True)
    assert set(np.unique(rk.key)) <= {-1, 0, 1}
    ks = T.KeyswitchKey(rng, params.ks_noise_stddev, T.keyswitch_parameters(params), sk.key, rk)
    assert ks.key.shape == (64, 3, 7, 521)
    a, b = ks.key[..., :520].astype(np.int64), ks.key[..., 520].astype(np.int64)
    phase = (b - (a * sk.key.key.astype(np.int64)).sum(-1)).astype(np.int32)
    h = np.arange(1, 8)[None, None, :]
    sh = (32 - np.arange(1, 4) * 3)[None, :, None]
    msg = ((rk.key[:, None, None] * h) << sh).astype(np.int32)
    err = (phase.astype(np.int64) - msg + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.abs(err).max() < 2 ** 32 * params.ks_noise_stddev * 6
    assert abs(err.mean()) < 2 ** 32 * params.ks_noise_stddev * 0.2      # recentred noise, keyswitch.jl:28-29


def test_shard_bounds_cover_batch_exactly():
    for G in (0, 1, 7, 16384, 16385):
        for ws in (1, 2, 3, 8):
            spans = [T.shard_bounds(G, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == G
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _Bit:
    """Plaintext stand-in for MKLweSample (batch of booleans in .b) to check circuit wiring without a GPU."""
    params = None
    current_variance = 0.0

    def __init__(self, b):
        self.b = np.asarray(b, dtype=bool)
        self.a = np.zeros(self.b.shape + (1, 1), np.int32)


def _plain_level(bk, ks, jobs):
    f = {"nand": lambda x, y: ~(x & y), "or": lambda x, y: x | y, "and": lambda x, y: x & y, "xor": lambda x, y: x ^ y}
    return [_Bit(f[k](*np.broadcast_arrays(x.b, y.b))) for k, x, y in jobs]


def _bits(v, W):
    return [_Bit(((np.asarray(v) >> i) & 1).astype(bool)) for i in range(W)]


def _val(bits, W, signed=True):
    v = sum(b.b.astype(np.int64) << i for i, b in enumerate(bits[:W]))
    return np.where(v >= 1 << (W - 1), v - (1 << W), v) if signed else v


def test_circuit_wiring_against_plain_integers(monkeypatch):
    monkeypatch.setattr(circuits, "gate_level", _plain_level)
    monkeypatch.setattr(circuits, "mk_copy_3gen", lambda x: _Bit(x.b.copy()))
    W = 8
    r = np.random.default_rng(5)
    x, y = r.integers(-60, 60, 64), r.integers(-60, 60, 64)
    one, zero = _Bit(np.ones(64, bool)), _Bit(np.zeros(64, bool))
    assert np.array_equal(_val(circuits.mk_add_3gen(None, None, _bits(x, W), _bits(y, W), zero, W), W), x + y)
    assert np.array_equal(_val(circuits.mk_sub_3gen(None, None, _bits(x, W), _bits(y, W), one, W), W), x - y)
    assert np.array_equal(circuits.mk_less_3gen(None, None, _bits(x, W), _bits(y, W), one, W).b, x < y)
    assert np.array_equal(circuits.mk_grt_3gen(None, None, _bits(x, W), _bits(y, W), one, W).b, x > y)
    assert np.array_equal(circuits.mk_leq_3gen(None, None, _bits(x, W), _bits(y, W), one, W).b, x <= y)
    assert np.array_equal(circuits.mk_geq_3gen(None, None, _bits(x, W), _bits(y, W), one, W).b, x >= y)
    assert np.array_equal(_val(circuits.mk_inv_3gen(None, None, _bits(x, W), one, W), W), ~x)
    c = circuits.mk_int_add_with_carry_3gen(None, None, _bits(x & 255, W), _bits(y & 255, W), zero, W)
    assert len(c) == W + 1 and np.array_equal(_val(c, W + 1, signed=False), (x & 255) + (y & 255))
    # mk_int_mul_3gen: literal transcription of the reference loop (1-based, 3gen_mk_gates.jl:312-362) on plain bits
    a, b = r.integers(0, 16, 64), r.integers(0, 16, 64)
    Wm = 4
    got = _val(circuits.mk_int_mul_3gen(None, None, _bits(a, Wm), _bits(b, Wm), zero, Wm), Wm, signed=False)
    row = lambda i: ((b >> (i - 1)) & 1) * a                      # BArr[i, :] as an integer, 1-based i
    tmp, ctr, low = row(1) >> 1, 1, row(1) & 1                    # result[1]; tmpIn = row 1 shifted, top bit ZERO
    for i in range(2, Wm):
        s = tmp + row(i)
        low |= (s & 1) << (i - 1)
        tmp, ctr = s >> 1, i
    s = tmp + row(ctr)                                            # the reference adds row `ctr` again, not row WIDTH
    expect = (low | (s << ctr)) & ((1 << Wm) - 1)
    assert np.array_equal(got, expect)


def test_interchange_roundtrip(tmp_path, toy_keys, oracle):
    """Key / ciphertext files (the format a Julia host writes): byte-exact round trip, and the oracle rebuilt from the file decrypts."""
    ks = toy_keys
    p = ks.prm
    sp = T.SchemeParameters_3gen(p.n, p.sigma_lwe, p.N, 1, False, p.l, p.bgbit, p.sigma_gsw, p.t, p.basebit, p.sigma_ks, p.k)
    kf, cf = str(tmp_path / "keys.bin"), str(tmp_path / "ct.bin")
    T.interchange.write_keys(kf, sp, [ks.bsk[i] for i in range(p.k)], [ks.ksk[i] for i in range(p.k)], ks.lwe_keys)
    sp2, bsk, ksk, lwe = T.interchange.read_keys(kf)
    assert sp2 == sp and np.array_equal(np.stack(bsk), ks.bsk) and np.array_equal(np.stack(ksk), ks.ksk) and np.array_equal(lwe, ks.lwe_keys)
    bits = np.array([1, 0, 1, 1, 0], np.uint8)
    a, b = ks.encrypt(bits, 9)
    T.interchange.write_ciphertexts(cf, a, b)
    a2, b2 = T.interchange.read_ciphertexts(cf)
    assert np.array_equal(a2, a) and np.array_equal(b2, b)
    clone = oracle.KeySet(oracle.params(p.n, p.N, p.k, p.l, p.bgbit, p.t, p.basebit, p.sigma_lwe, p.sigma_gsw, p.sigma_ks),
                          raw_bsk=np.stack(bsk), raw_ksk=np.stack(ksk))
    oa, ob = clone.gate_batch(oracle.EXACT_NTT, oracle.GATE_NAND, (a2, b2), (a2, b2), nthreads=2)
    ph = (ob.astype(np.int64) - (oa.astype(np.int64) * lwe[None]).sum((-1, -2))).astype(np.int32)
    assert np.array_equal(ph > 0, ~bits.astype(bool))          # NAND(x, x) = NOT x
    with pytest.raises(ValueError):
        T.interchange.read_keys(cf)


class _PBit:
    """Plaintext stand-in with MKLweSample's constructor / indexing surface (batch of booleans in .b, dummy .a)."""

    def __init__(self, params, a, b, current_variance=0.0):
        self.params, self.b, self.current_variance = params, np.asarray(b, dtype=bool), 0.0
        self.a = np.zeros(self.b.shape + (1, 1), np.int32)

    @property
    def batch_shape(self):
        return self.b.shape

    def __getitem__(self, idx):
        return _PBit(None, None, self.b[idx])


def _pbits(v, W):
    return [_PBit(None, None, ((np.asarray(v) >> i) & 1).astype(bool)) for i in range(W)]


def _plain_level_p(bk, ks, jobs):
    return [_PBit(None, None, o.b) for o in _plain_level(bk, ks, jobs)]


def test_truncated_multiplier_and_conv_layer_wiring(monkeypatch):
    """mk_int_mul_lo_3gen / add_mod_3gen / enc_conv2d on plain bits == integer arithmetic mod 2^WIDTH, for every shard split."""
    from torus_fhe_b200 import workloads
    calls = []

    def counting_level(bk, ks, jobs):
        calls.append(sum(int(np.prod(np.broadcast_shapes(x.b.shape, y.b.shape), dtype=np.int64)) for _, x, y in jobs))
        return _plain_level_p(bk, ks, jobs)
    monkeypatch.setattr(circuits, "gate_level", counting_level)
    monkeypatch.setattr(workloads, "_cat", lambda xs, axis: _PBit(None, None, np.concatenate([x.b for x in xs], axis)))
    r = np.random.default_rng(9)
    for W in (1, 2, 3, 5, 8):
        a, b = r.integers(-(1 << (W - 1)), 1 << (W - 1), 200) if W > 1 else r.integers(-1, 1, 200), r.integers(-(1 << (W - 1)), max(1 << (W - 1), 1), 200)
        wrap = lambda v: ((v + (1 << (W - 1))) % (1 << W)) - (1 << (W - 1))
        assert np.array_equal(_val(circuits.add_mod_3gen(None, None, _pbits(a, W), _pbits(b, W), W), W), wrap(a + b))
        calls.clear()
        assert np.array_equal(_val(circuits.mk_int_mul_lo_3gen(None, None, _pbits(a, W), _pbits(b, W), W), W), wrap(a * b))
        add = lambda w: 1 if w == 1 else 5 * w - 6
        assert sum(calls) == 200 * (W * (W + 1) // 2 + sum(add(W - i) for i in range(1, W)))
    W, H, C, K = 4, 7, 2, 3
    inp, ker = r.integers(-8, 8, (H, H)), r.integers(-8, 8, (C, K, K))
    zero = _PBit(None, None, False)
    for stride, padding in ((1, 0), (2, 0), (1, 1), (2, 1)):
        exp = workloads.conv2d_plain(inp, ker, stride, padding, W)
        monkeypatch.setattr(workloads, "_pad", lambda x, z, p: _PBit(None, None, np.pad(x.b, p)))
        calls.clear()
        out = workloads.enc_conv2d(None, None, _pbits(inp, W), zero, _pbits(ker, W), stride, padding, W)
        assert out[0].b.shape == exp.shape == workloads.conv2d_output_shape((H, H), (C, K, K), stride, padding)
        assert np.array_equal(_val(out, W), exp)
        assert sum(calls) == workloads.conv2d_gate_count((H, H), (C, K, K), stride, padding, W)
        # sharded over 3 ranks: the slices tile the flattened outputs and equal the unsharded result
        got = np.zeros(exp.size, np.int64)
        for rank in range(3):
            lo, hi, bits = workloads.enc_conv2d(None, None, _pbits(inp, W), zero, _pbits(ker, W), stride, padding, W, shard=(3, rank))
            got[lo:hi] = _val(bits, W)
        assert np.array_equal(got.reshape(exp.shape), exp)
    # plaintext model sanity: one known window
    assert workloads.conv2d_plain(np.arange(9).reshape(3, 3), np.ones((1, 3, 3), int), 1, 0, 8)[0, 0, 0] == 36


def test_blind_rotate_and_keyswitch_entry_points_follow_the_reference_signatures():
    """mk_bootstrap_wo_keyswitch_3gen(bk, mu, x), mk_blind_rotate_and_extract_3gen(v, bk, barb, bara) and mk_keyswitch_3gen(ks, u)
    (3gen_mk_internals.jl:88-109, mk_internals.jl:730-744) with a recording stand-in for the GPU context: argument packing, the
    mod-switch round trip of the rotations, shapes, and the loud errors (no engine, other test vectors, bad shapes)."""
    import types
    import torus_fhe_b200 as T
    prm = T.SchemeParameters_3gen(6, 0.0, 1024, 1, False, 2, 7, 0.0, 3, 3, 0.0, 2)
    k, n, N = 2, 6, 1024
    calls = {}

    class Ctx:
        def blind_rotate_batch(self, mu, a, b):
            calls["br"] = (mu, a.copy(), b.copy())
            G = b.size
            return np.arange(G * (N + 1), dtype=np.int32).reshape(G, N + 1), None

        def keyswitch_batch(self, ext):
            calls["ks"] = ext.copy()
            G = ext.shape[0]
            return np.ones((G, k, n), np.int32), np.full(G, 7, np.int32)

    eng = types.SimpleNamespace(params=prm, ctx=Ctx())
    bk = [types.SimpleNamespace(_engine=None) for _ in range(k)]
    ks = [types.SimpleNamespace() for _ in range(k)]
    mu = T.encode_message64(1, 8)
    x = T.MKLweSample(T.LweParams(n), np.zeros((3, k, n), np.int32), np.zeros(3, np.int32))
    with pytest.raises(RuntimeError, match="engine_for"):
        T.mk_bootstrap_wo_keyswitch_3gen(bk, mu, x)
    with pytest.raises(RuntimeError, match="engine_for"):
        T.mk_keyswitch_3gen(ks, T.LweSample(T.LweParams(N), np.zeros((3, N), np.int32), np.zeros(3, np.int32)))
    T.attach_engine(bk, ks, eng)
    u = T.mk_bootstrap_wo_keyswitch_3gen(bk, mu, x)
    assert isinstance(u, T.LweSample) and u.params.size == N and u.a.shape == (3, N) and u.b.shape == (3,)
    assert calls["br"][0] == int(mu) and u.b[1] == 2 * (N + 1) - 1 and u.a[1, 0] == N + 1
    # reference form: rotations already mod-switched; the kernel's own mod-switch must give them back
    r = np.random.default_rng(3)
    barb, bara = r.integers(-N, N, 3), r.integers(-N, N, (3, k, n))
    barb[0], bara[0, 0, 0], bara[0, 0, 1] = -N, N - 1, 0
    u = T.mk_blind_rotate_and_extract_3gen(np.full(N, mu, np.int64), bk, barb, bara)
    _, a_in, b_in = calls["br"]
    assert a_in.dtype == np.int32 and a_in.shape == (3, k, n)
    assert np.array_equal(T.decode_message(a_in, 2 * N), bara) and np.array_equal(T.decode_message(b_in, 2 * N), barb)
    assert u.a.shape == (3, N)
    tv = np.full(N, mu, np.int64); tv[5] += 1
    with pytest.raises(ValueError, match="constant test vector"):
        T.mk_blind_rotate_and_extract_3gen(tv, bk, barb, bara)
    with pytest.raises(ValueError, match="rotations"):
        T.mk_blind_rotate_and_extract_3gen(np.full(N, mu, np.int64), bk, barb, bara + N)
    with pytest.raises(ValueError, match="rotations"):
        T.mk_blind_rotate_and_extract_3gen(np.full(N, mu, np.int64), bk, barb, bara[:, :1])
    out = T.mk_keyswitch_3gen(ks, u)
    assert isinstance(out, T.MKLweSample) and out.a.shape == (3, k, n) and out.params.size == n and out.current_variance == 0.0
    assert np.array_equal(calls["ks"][:, :N], u.a) and np.array_equal(calls["ks"][:, N], u.b) and np.all(out.b == 7)
    with pytest.raises(ValueError, match="dimension N"):
        T.mk_keyswitch_3gen(ks, T.LweSample(T.LweParams(n), np.zeros((3, n), np.int32), np.zeros(3, np.int32)))


def test_exported_name_surface_is_the_reference_3gen_surface():
    """every 3gen name 3-gen-mk-tfhe/src/TFHE.jl exports (lines 11-15, 32-96, 113-119, 135-165, 177-196) resolves in
    the host mirror, plus the un-exported mk_bootstrap_wo_keyswitch_3gen the parity tests use; gates keep the reference's positional
    arguments (bk, ks, x, y[, z]) and circuits (bk, ks, a, b, carry-or-one, WIDTH)."""
    import inspect
    import torus_fhe_b200 as T
    names = """SecretKey_3gen RLweKey KeyswitchKey MKLweSample CRP_3gen PublicKey CommonPubKey_3gen BootstrapKeyPart_3gen
        TransformedBootstrapKeyPart_3gen lwe_parameters rlwe_parameters tgsw_parameters keyswitch_parameters encode_message encode_message64
        decode_message decode_message64 noise_calc mk_lwe_noiseless_trivial mk_keyswitch_3gen mk_lwe_phase mk_blind_rotate_and_extract_3gen
        mk_bootstrap_3gen GenCRP_3gen mk_encrypt_3gen mk_int_encrypt_3gen mk_decrypt_3gen mk_int_decrypt_3gen mk_gate_nand_3gen mk_gate_or_3gen
        mk_gate_xor_3gen mk_gate_and_3gen mk_gate_3and_3gen mk_gate_not_3gen mk_gate_mux_3gen mk_copy_3gen mk_add_3gen mk_add_3gen_v2 mk_inv_3gen
        mk_sub_3gen mk_less_3gen mk_grt_3gen mk_leq_3gen mk_geq_3gen mk_int_add_with_carry_3gen mk_int_mul_3gen
        mk_bootstrap_wo_keyswitch_3gen mk_gate_nand_3gen_wb mk_gate_or_3gen_wb mk_gate_and_3gen_wb mk_gate_xor_3gen_wb mk_gate_xor_3gen_gpu
        mk_int_add_3gen_gpu xor_3gen_gpu mk_lwe_noiseless_trivial_gpu enc_conv2d
        TGswSample_3gen TransformedTGswSample_3gen tgsw_encrypt_3gen tgsw_extern_mul_3gen mk_mux_rotate_3gen mk_ith_blind_rotate_3gen
        mk_blind_rotate_3gen RLweSample rlwe_noiseless_trivial rlwe_extract_sample_64 mul_by_monomial t64tot32""".split()
    # = every function and type of 3gen_mk_gates.jl, 3gen_mk_internals.jl, tgsw_3gen.jl and gpu_circuits.jl, plus what they call in rlwe.jl
    names += [f"mktfhe_parameters_{k}party_3gen" for k in (2, 3, 4, 5, 8, 16, 32, 64, 128, 256, 512)]
    assert [n for n in names if not hasattr(T, n)] == []
    arity = {"mk_gate_nand_3gen": 4, "mk_gate_or_3gen": 4, "mk_gate_and_3gen": 4, "mk_gate_xor_3gen": 4, "mk_gate_3and_3gen": 5, "mk_gate_mux_3gen": 5,
             "mk_gate_not_3gen": 1, "mk_bootstrap_3gen": 4, "mk_bootstrap_wo_keyswitch_3gen": 3, "mk_blind_rotate_and_extract_3gen": 4,
             "mk_keyswitch_3gen": 2, "tgsw_extern_mul_3gen": 2, "mk_mux_rotate_3gen": 3, "mk_ith_blind_rotate_3gen": 3, "mk_blind_rotate_3gen": 3,
             "tgsw_encrypt_3gen": 5, "mk_add_3gen": 6, "mk_add_3gen_v2": 6, "mk_sub_3gen": 6, "mk_inv_3gen": 5, "mk_less_3gen": 6, "mk_grt_3gen": 6,
             "mk_leq_3gen": 6, "mk_geq_3gen": 6, "mk_int_add_with_carry_3gen": 6, "mk_int_mul_3gen": 6}
    for name, n in arity.items():
        params = [p for p in inspect.signature(getattr(T, name)).parameters.values() if p.default is inspect.Parameter.empty]
        assert len(params) == n, (name, [p.name for p in params])


def test_unbootstrapped_gate_variants_are_the_gate_prologues(oracle, rng):
    """mk_gate_{nand,or,and,xor}_3gen_wb (3gen_mk_gates.jl:16-21, 32-37, 48-53, 76-81) return the gate's linear prologue: the same
    sample the oracle (and the kernel's fused prologue) feeds to the bootstrap."""
    import ctypes as C
    import torus_fhe_b200 as T
    L = oracle.lib()
    L.mko_gate_prologue.restype = None
    L.mko_gate_prologue.argtypes = [C.POINTER(oracle.Params), C.c_int] + [C.c_void_p, C.c_int32] * 3 + [C.c_void_p, C.c_void_p]
    k, n = 3, 17
    prm = oracle.params(n, 64, k, 2, 7, 3, 3, 0.0, 0.0, 0.0)
    lp = T.LweParams(n)
    mk = lambda: T.MKLweSample(lp, rng.integers(-2 ** 31, 2 ** 31, size=(k, n), dtype=np.int64).astype(np.int32),
                               np.int32(rng.integers(-2 ** 31, 2 ** 31)), 0.25)
    x, y = mk(), mk()
    bk = [None] * k        # only its length is read, as in the reference
    for fn, gate in ((T.mk_gate_nand_3gen_wb, oracle.GATE_NAND), (T.mk_gate_or_3gen_wb, oracle.GATE_OR),
                     (T.mk_gate_and_3gen_wb, oracle.GATE_AND), (T.mk_gate_xor_3gen_wb, oracle.GATE_XOR)):
        out = fn(bk, None, x, y)
        ta, tb = np.empty((k, n), np.int32), C.c_int32(0)
        L.mko_gate_prologue(C.byref(prm), gate, x.a.ctypes.data, int(x.b), y.a.ctypes.data, int(y.b), None, 0, ta.ctypes.data, C.byref(tb))
        assert np.array_equal(out.a, ta) and int(out.b) == tb.value, fn.__name__
        assert out.current_variance == (0.5 if gate != oracle.GATE_XOR else 2.0)      # mk_internals.jl:40-51
    # batched samples too
    xb = T.MKLweSample.stack([x, y, x])
    yb = T.MKLweSample.stack([y, y, x])
    out = T.mk_gate_nand_3gen_wb(bk, None, xb, yb)
    assert out.a.shape == (3, k, n) and np.array_equal(out[0].a, T.mk_gate_nand_3gen_wb(bk, None, x, y).a)


def test_path_internals_follow_the_reference_step_by_step(oracle, toy_keys):
    """tgsw_extern_mul_3gen / mk_mux_rotate_3gen / mk_ith_blind_rotate_3gen / mk_blind_rotate_3gen / rlwe_extract_sample_64
    (3gen_mk_internals.jl:59-95, tgsw_3gen.jl:102-113, rlwe.jl:70-74) in the host mirror: the host-side steps (monomial product,
    sample arithmetic, loop order, zero skip, extraction) against the oracle, with the GPU external product replaced by the
    oracle's -- on the B200 the same functions are checked against the fused kernel (tests/test_zz_gpu_path_internals.py)."""
    import torus_fhe_b200 as T
    ks = toy_keys
    N, n, k = ks.N, ks.n, ks.k

    class FakeCtx:
        calls = 0

        def extprod_batch(self, elem, acc):
            out = np.empty_like(acc)
            for g, e in enumerate(np.asarray(elem).reshape(-1)):
                out[g] = ks.extprod(oracle.EXACT_SCHOOLBOOK, int(e) // n, int(e) % n, acc[g])
                FakeCtx.calls += 1
            return out

    class FakeEngine:
        params = T.SchemeParameters_3gen(n, 0.0, N, 1, False, ks.l, ks.prm.bgbit, 0.0, ks.t, ks.prm.basebit, 0.0, k)
        ctx = FakeCtx()

    tg, rl = T.TGswParams(ks.l, ks.prm.bgbit, False), T.RLweParams(N, 1, False)
    bk = [T.TransformedBootstrapKeyPart_3gen(T.BootstrapKeyPart_3gen.from_array(ks.bsk[p], tg, rl)) for p in range(k)]
    sample = bk[1].tgsw_samples[3]
    with pytest.raises(RuntimeError, match="engine_for"):
        T.tgsw_extern_mul_3gen(T.RLweSample(rl, np.zeros((2, N), np.int64)), sample)
    kk = [T.KeyswitchKey.from_array(ks.ksk[p], T.KeyswitchParameters(ks.t, ks.prm.basebit)) for p in range(k)]
    T.attach_engine(bk, kk, FakeEngine())
    assert np.array_equal(sample.part_3, ks.bsk[1, 3, 2]) and sample._locate()[1] == 1 * n + 3

    r = np.random.default_rng(5)
    poly = r.integers(-2 ** 63, 2 ** 63 - 1, size=N, dtype=np.int64)
    for s in (0, 1, -1, 7, N, N + 5, -N - 5, 2 * N, 5 * N + 3):
        assert np.array_equal(T.mul_by_monomial(poly, s), oracle.mul_by_monomial(poly, s)), s
    acc = T.RLweSample(rl, r.integers(-2 ** 63, 2 ** 63 - 1, size=(2, N), dtype=np.int64))
    assert np.array_equal(T.tgsw_extern_mul_3gen(acc, sample).a, ks.extprod(oracle.EXACT_SCHOOLBOOK, 1, 3, acc.a))
    assert np.array_equal(T.mk_mux_rotate_3gen(acc, sample, np.int32(-9)).a, ks.mux_rotate(oracle.EXACT_SCHOOLBOOK, 1, 3, -9, acc.a))
    batch = T.RLweSample(rl, np.stack([acc.a, -acc.a]))        # a leading batch dimension goes through one call
    out = T.mk_mux_rotate_3gen(batch, sample, 17)
    assert np.array_equal(out.a[0], ks.mux_rotate(oracle.EXACT_SCHOOLBOOK, 1, 3, 17, acc.a))

    # the whole blind rotation, stage by stage, equals the oracle's bootstrap without key switch (accumulator and extracted sample)
    mu = T.encode_message64(1, 8)
    a, b = ks.encrypt(np.array([1], np.uint8), 5)
    x = T.MKLweSample(T.LweParams(n), a[0], b[0])
    barb, bara = T.decode_message(x.b, 2 * N), T.decode_message(x.a, 2 * N)
    bara = bara.copy()
    bara[0, 2] = 0                                             # a zero rotation is skipped (3gen_mk_internals.jl:69)
    FakeCtx.calls = 0
    accum = T.rlwe_noiseless_trivial(T.mul_by_monomial(np.full(N, mu, np.int64), -int(barb)), rl)
    accum = T.mk_blind_rotate_3gen(accum, bk, bara)
    assert FakeCtx.calls == int(np.count_nonzero(bara))
    # the oracle mod-switches itself: hand it torus elements that decode to the same rotations
    sh = 32 - int(np.log2(2 * N))
    ext_a, ext_b, acc_ref, _ = ks.bootstrap_wo_keyswitch(oracle.EXACT_SCHOOLBOOK, int(mu), (bara.astype(np.int64) << sh).astype(np.int32),
                                                         np.int32(int(barb) << sh), want_acc=True)
    assert np.array_equal(accum.a, acc_ref)
    u = T.rlwe_extract_sample_64(accum)
    assert u.params.size == N and np.array_equal(u.a, ext_a) and int(u.b) == int(ext_b)
    L = oracle.lib()
    for v in (-1, -(1 << 32), (1 << 32) - 1, -(1 << 32) - 1, 1 << 61, -2 ** 63, (1 << 62) + (1 << 32) - 1):
        assert int(T.t64tot32(np.int64(v))) == L.mko_t64tot32(v)


def test_key_generation_formulas_with_the_exact_product_swapped_for_the_oracle(oracle, monkeypatch):
    """tgsw_encrypt_3gen / BootstrapKeyPart_3gen / PublicKey / CommonPubKey_3gen of the host mirror (tgsw_3gen.jl:41-95,
    mk_internals.jl:266-345) produce TGSW encryptions of the LWE key bits under the joint RLWE key -- checked with the secret keys
    in hand, the GPU's exact product replaced here by the oracle's schoolbook (on the B200: tests/test_gpu_api.py)."""
    import torus_fhe_b200 as T
    from torus_fhe_b200 import tfhe3gen
    from test_oracle_semantics import rlwe_phase, torus

    def cpu_mul(small, big):
        small, big = np.asarray(small, np.int64), np.asarray(big, np.int64)
        shape = np.broadcast_shapes(small.shape, big.shape)
        s = np.broadcast_to(small, shape).reshape(-1, shape[-1])
        b = np.broadcast_to(big, shape).reshape(-1, shape[-1])
        return np.stack([oracle.negacyclic_mul(s[i], b[i], oracle.EXACT_SCHOOLBOOK) for i in range(s.shape[0])]).reshape(shape)

    monkeypatch.setattr(tfhe3gen, "negacyclic_mul", cpu_mul)
    rng = np.random.default_rng(9)
    params = T.SchemeParameters_3gen(6, 2.0 ** -18, 64, 1, False, 2, 7, 2.0 ** -40, 3, 3, 2.0 ** -18, 2)
    tg, rl = T.tgsw_parameters(params), T.rlwe_parameters(params)
    sk = [T.SecretKey_3gen(rng, params) for _ in range(2)]
    rk = [T.RLweKey(rng, rl, True) for _ in range(2)]
    crp = T.CRP_3gen(rng, tg, rl, True)
    pk = [T.PublicKey(rng, rk[i], params.gsw_noise_stddev, crp, tg, 1) for i in range(2)]
    cpk = T.CommonPubKey_3gen(pk, params, 2)
    Z = (rk[0].key + rk[1].key).astype(np.int64)
    assert set(np.unique(rk[0].key)) <= {-1, 0, 1}

    def check(parts, m):
        for q in range(tg.decomp_length):
            g = np.int64(tg.gadget_values[q])
            exp = np.zeros(64, np.int64)
            exp[0] = m * g
            assert np.abs(torus(rlwe_phase(np.stack([parts[3][q], parts[0][q]]), Z) - exp)).max() < 2.0 ** -30      # body-digit row
            assert np.abs(torus(rlwe_phase(np.stack([parts[2][q], parts[1][q]]), Z) + (m * g) * Z)).max() < 2.0 ** -30   # mask-digit row

    for m in (0, 1):
        s = T.tgsw_encrypt_3gen(rng, m, params.gsw_noise_stddev, cpk, crp)
        assert isinstance(s, T.TGswSample_3gen) and s.part_1.shape == (2, 64)
        check([s.part_1, s.part_2, s.part_3, s.part_4], m)
    bkp = T.BootstrapKeyPart_3gen(rng, sk[0].key, params.gsw_noise_stddev, crp, cpk, tg, rl, 1)
    assert bkp.gsw_key.shape == (6, 4, 2, 64)
    for j in range(6):
        check(bkp.gsw_key[j], int(sk[0].key.key[j]))
