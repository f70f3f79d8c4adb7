"""Single-key TFHE behind the 3gen engine (first slice): torus-fhe_b200/tfhe1.py.

CPU tests: the numpy Torus32 restatement of the reference's single-key path (oracle/tfhe1_oracle.py) reproduces the reference's own
test shape -- the decrypted truth table of its twelve gates (3-gen-mk-tfhe/test/runtests.jl:10-42) -- and the embedding the product
uses (Torus32 values carried as v << 32 through the Torus64 3gen path, standard TGSW rows as the four 3gen parts) is exact: the 3gen
exact oracle run on the embedded keys returns the single-key oracle's bits.
GPU tests: the twelve gates through the product's host mirror, bit for bit against the single-key oracle on the same key bytes,
and the reference's truth-table test on tfhe_parameters_128."""
import numpy as np
import pytest

from oracle import tfhe1_oracle as O1

PLAIN = {"NAND": lambda x, y: not (x and y), "OR": lambda x, y: x or y, "AND": lambda x, y: x and y, "XOR": lambda x, y: x != y,
         "XNOR": lambda x, y: x == y, "NOR": lambda x, y: not (x or y), "ANDNY": lambda x, y: (not x) and y, "ANDYN": lambda x, y: x and not y,
         "ORNY": lambda x, y: (not x) or y, "ORYN": lambda x, y: x or not y}


def test_single_key_oracle_truth_tables_toy_ring():
    """runtests.jl:10-42 on a toy ring (N = 128, n = 12): every gate, every input combination, decrypts to the plain gate."""
    rng = np.random.default_rng(123)
    n, N, l, bg, t, bb = 12, 128, 3, 7, 8, 2
    s, z, bk, ksk = O1.keygen(rng, n, N, l, bg, t, bb, 2.0 ** -25, 2.0 ** -15)
    prm = (l, bg, t, bb)
    enc = lambda b: O1.encrypt(rng, s, b, 2.0 ** -15)
    for name, plain in PLAIN.items():
        for xb in (False, True):
            for yb in (False, True):
                assert O1.decrypt(s, O1.gate(name, bk, ksk, prm, enc(xb), enc(yb))) == plain(xb, yb), (name, xb, yb)
    for xb in (False, True):
        x = enc(xb)
        assert O1.decrypt(s, O1.gate("NOT", bk, ksk, prm, x)) == (not xb)
        for yb in (False, True):
            for zb in (False, True):
                assert O1.decrypt(s, O1.gate("MUX", bk, ksk, prm, x, enc(yb), enc(zb))) == (yb if xb else zb)


def test_oracle_primitives_against_definitions():
    rng = np.random.default_rng(5)
    N = 64
    a, b = rng.integers(-64, 64, N), rng.integers(-2 ** 31, 2 ** 31, N).astype(np.int32)
    ref = np.zeros(N, object)
    for i in range(N):
        for j in range(N):
            k, sgn = (i + j) % N, -1 if i + j >= N else 1
            ref[k] += sgn * int(a[i]) * int(b[j])
    assert np.array_equal(O1.negacyclic_mul(a, b), np.array([((int(v) + 2 ** 31) % 2 ** 32) - 2 ** 31 for v in ref], np.int32))
    # decomposition recomposes to the input rounded DOWN to a multiple of 2^(32 - l*bgbit) = 2^11 (tgsw.jl:105-110: it floors)
    p = rng.integers(-2 ** 31, 2 ** 31, N).astype(np.int32)
    d = O1.decompose(p, 3, 7)
    assert d.min() >= -64 and d.max() <= 63
    rec = sum(d[q].astype(np.int64) << (32 - 7 * (q + 1)) for q in range(3))
    err = (p.astype(np.int64) - rec + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.all((err >= 0) & (err < (1 << 11)))
    # monomial: X^N = -1, X^(2N) = 1
    assert np.array_equal(O1.mul_by_monomial(p, N), O1.wrap32(-p.astype(np.int64))) and np.array_equal(O1.mul_by_monomial(p, 2 * N), p)
    assert np.array_equal(O1.mul_by_monomial(O1.mul_by_monomial(p, 5), -5), p)


def test_host_mirror_names_parameters_and_linear_prologues():
    """api.jl:76-113 parameter values; the twelve gates of gates.jl under their names; each gate's linear prologue in the product
    equals the oracle's table (gates.jl:16-142); LweSample arithmetic wraps as Int32 (lwe.jl:62-76)."""
    import inspect
    import torus_fhe_b200.tfhe1 as T1
    p = T1.tfhe_parameters_128()
    assert (p.lwe_size, p.rlwe_polynomial_degree, p.rlwe_mask_size, p.bs_decomp_length, p.bs_log2_base, p.ks_decomp_length, p.ks_log2_base) == (630, 1024, 1, 3, 7, 8, 2)
    assert p.lwe_noise_stddev == 1 / 2 ** 15 and p.bs_noise_stddev == 1 / 2 ** 25 and p.rlwe_is32
    p = T1.tfhe_parameters_80()
    assert (p.lwe_size, p.bs_decomp_length, p.bs_log2_base, p.ks_decomp_length, p.ks_log2_base) == (500, 2, 10, 8, 2)
    assert set(T1.GATES) == {"NAND", "OR", "AND", "XOR", "XNOR", "NOT", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN", "MUX"}      # runtests.jl:10-23
    for name in ("make_key_pair", "encrypt", "decrypt", "SecretKey", "CloudKey", "BootstrapKey", "bootstrap", "bootstrap_wo_keyswitch", "keyswitch",
                 "gate_constant", "lwe_noiseless_trivial", "lwe_encrypt", "lwe_phase"):
        assert hasattr(T1, name), name
    # the (mu0, cx, cy) each gate hands to mktfhe_affine_bootstrap_batch, read from its source, against the oracle's table
    for name, (m, space, cx, cy) in O1.GATE_LINEAR.items():
        src = inspect.getsource(T1.GATES[name][0])
        assert f"encode_message({m}, {space}), {cx}, {cy}, 0" in src, (name, src)
    a = T1.LweSample(T1.LweParams(3), np.array([2 ** 31 - 1, -2 ** 31, 5], np.int32), np.int32(2 ** 31 - 1))
    s2 = a + a
    assert list(s2.a) == [-2, 0, 10] and int(s2.b) == -2 and list((-a).a[:1]) == [-(2 ** 31 - 1)] and list((a * 2).a) == [-2, 0, 10] and list((2 * a).a) == [-2, 0, 10]
    assert int((a - a).b) == 0 and T1.lwe_noiseless_trivial(7, T1.LweParams(4), (2,)).a.shape == (2, 4)
    # engine_parts: part order and the << 32 embedding
    bk = np.arange(2 * 3 * 2 * 2 * 4, dtype=np.int32).reshape(2, 3, 2, 2, 4) - 40
    parts = T1.BootstrapKey(None, 0.0, None, None, None, samples=bk).engine_parts()
    assert parts.shape == (2, 4, 3, 4) and parts.dtype == np.int64
    assert np.array_equal(parts[:, 0], bk[:, :, 1, 1].astype(np.int64) << 32) and np.array_equal(parts[:, 1], bk[:, :, 0, 1].astype(np.int64) << 32)
    assert np.array_equal(parts[:, 2], bk[:, :, 0, 0].astype(np.int64) << 32) and np.array_equal(parts[:, 3], bk[:, :, 1, 0].astype(np.int64) << 32)


def _engine_parts(bk):
    """The product's mapping of a standard TGSW key onto the engine's four parts, through its own class."""
    import torus_fhe_b200.tfhe1 as T1
    return T1.BootstrapKey(None, 0.0, None, None, None, samples=bk).engine_parts()


def test_embedding_into_the_3gen_path_is_exact(oracle):
    """Full ring (N = 1024), reduced LWE dimension: the exact 3gen oracle (Torus64, four parts) on the embedded key bytes gives the
    single-key Torus32 oracle's bootstrap output bit for bit -- the claim tfhe1.py's module docstring makes, without a GPU."""
    rng = np.random.default_rng(7)
    n, N, l, bg, t, bb = 6, 1024, 3, 7, 8, 2
    s, z, bk, ksk = O1.keygen(rng, n, N, l, bg, t, bb, 2.0 ** -25, 2.0 ** -15)
    prm3 = dict(n=n, N=N, k=1, l=l, bgbit=bg, t=t, basebit=bb, sigma_lwe=2.0 ** -15, sigma_gsw=2.0 ** -25, sigma_ks=2.0 ** -15)
    ks3 = oracle.KeySet(prm3, raw_bsk=_engine_parts(bk)[None], raw_ksk=ksk[None])
    mu = O1.encode_message(1, 8)
    for bit in (False, True):
        x = O1.encrypt(rng, s, bit, 2.0 ** -15)
        ra, rb = O1.bootstrap(bk, ksk, mu, x[0], x[1], l, bg, t, bb)
        ga, gb = ks3.bootstrap_batch(oracle.EXACT_SCHOOLBOOK, mu << 32, x[0].reshape(1, 1, n), np.array([x[1]], np.int32), nthreads=1)
        assert np.array_equal(ga.reshape(-1), ra) and int(gb[0]) == int(rb)
        assert O1.decrypt(s, (ra, rb)) == bit


@pytest.fixture(scope="module", params=[(3, 7), (2, 10), (3, 9)], ids=["l3_bg7_default_kernels", "l2_bg10_torus32_mode", "l3_bg9_torus32_mode"])
def small_single_key(request):
    """Key bytes from the single-key oracle's own generator (N = 1024, n = 16), loaded into the product through its classes: the gadget
    shape of tfhe_parameters_128 (default kernels, keys << 32), of tfhe_parameters_80 (10-bit digits: the library's Torus32 mode) and
    of the CCS 2-party set (l = 3, Bg = 2^9: Torus32 mode with three levels)."""
    import torus_fhe_b200.tfhe1 as T1
    rng = np.random.default_rng(11)
    l, bg = request.param
    n, N, t, bb = 16, 1024, 8, 2
    s, z, bk, ksk = O1.keygen(rng, n, N, l, bg, t, bb, 2.0 ** -25, 2.0 ** -15)
    params = T1.SchemeParameters(n, 2.0 ** -15, N, 1, True, l, bg, 2.0 ** -25, t, bb, 2.0 ** -15, 1)
    sk = T1.SecretKey(None, params, key=s)
    ck = T1.CloudKey.__new__(T1.CloudKey)
    ck.params, ck._engine = params, None
    ck.rlwe_key = T1.RLweKey(None, T1.rlwe_parameters(params), key=z)
    ck.bootstrap_key = T1.BootstrapKey(None, 0.0, None, None, T1.tgsw_parameters(params), samples=bk)
    ck.keyswitch_key = T1.KeyswitchKey.from_array(ksk, T1.keyswitch_parameters(params))
    yield T1, rng, sk, ck, (s, bk, ksk, (l, bg, t, bb))
    if ck._engine is not None:
        ck._engine.close()


@pytest.mark.gpu
def test_single_key_gates_bit_exact_against_the_oracle(small_single_key):
    T1, rng, sk, ck, (s, bk, ksk, prm) = small_single_key
    bits = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)], bool)
    x, y, z = (T1.encrypt(rng, sk, bits[:, i]) for i in range(3))
    for name, (fn, nargs, plain) in T1.GATES.items():
        args = (x, y, z)[:nargs]
        out = fn(ck, *args)
        assert out.a.shape == x.a.shape and out.b.shape == x.b.shape
        for g in range(len(bits)):
            ra, rb = O1.gate(name, bk, ksk, prm, *[(v.a[g], v.b[g]) for v in args])
            assert np.array_equal(out.a[g], ra) and int(out.b[g]) == int(rb), (name, g)
        assert np.array_equal(T1.decrypt(sk, out), plain(*[bits[:, i] for i in range(nargs)])), name
    # one external product alone (tgsw_extern_mul, tgsw.jl:143-147) through the parity hook: Torus32 accumulators ride in the top halves
    eng = T1.engine_for(ck)
    acc = rng.integers(-2 ** 31, 2 ** 31, (6, 2, 1024)).astype(np.int32)
    acc[0], acc[1] = 0, -1
    elem = np.array([0, 3, 7, 15, 1, 2], np.int32)
    got = eng.ctx.extprod_batch(elem, acc.astype(np.int64) << 32)
    assert np.all((got & 0xFFFFFFFF) == 0)
    l, bgbit = prm[0], prm[1]
    for g in range(6):
        assert np.array_equal((got[g] >> 32).astype(np.int32), O1.tgsw_extern_mul(acc[g], bk[elem[g]], l, bgbit)), g
    # bootstrap = keyswitch o bootstrap_wo_keyswitch (bootstrap.jl:97-100); scalar call = batch of one
    mu = T1.encode_message(1, 8)
    b1 = T1.bootstrap(ck, None, mu, x)
    b2 = T1.keyswitch(ck, T1.bootstrap_wo_keyswitch(ck, mu, x))
    assert np.array_equal(b1.a, b2.a) and np.array_equal(b1.b, b2.b) and np.array_equal(T1.decrypt(sk, b1), bits[:, 0])
    one = T1.gate_nand(ck, x[3], y[3])
    assert one.b.shape == () and T1.decrypt(sk, one) == (not (bits[3, 0] and bits[3, 1]))
    assert T1.decrypt(sk, T1.gate_constant(ck, True)) is True


@pytest.mark.gpu
def test_reference_truth_table_test_on_tfhe_parameters_128():
    """test/runtests.jl:28-59 ("gate" and "single party, custom parameters"): make_key_pair, then every gate on every input
    combination, on tfhe_parameters_128 (keys generated by the product: exact key products on the GPU)."""
    import torus_fhe_b200.tfhe1 as T1
    rng = np.random.default_rng(123)
    sk, ck = T1.make_key_pair(rng, T1.tfhe_parameters_128())
    try:
        bits = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)], bool)
        x, y, z = (T1.encrypt(rng, sk, bits[:, i]) for i in range(3))
        for name, (fn, nargs, plain) in T1.GATES.items():
            out = fn(ck, *(x, y, z)[:nargs])
            assert np.array_equal(T1.decrypt(sk, out), plain(*[bits[:, i] for i in range(nargs)])), name
    finally:
        ck._engine.close()


@pytest.mark.gpu
def test_reference_gate_testcase_on_its_default_parameters():
    """test/runtests.jl:28-42 ("gate"): `make_key_pair(rng)` with the DEFAULT parameters -- tfhe_parameters_80, Bg = 2^10 -- then every
    gate on every input combination.  Runs in the library's Torus32 mode (10-bit gadget digits)."""
    import torus_fhe_b200 as T
    import torus_fhe_b200.tfhe1 as T1
    rng = np.random.default_rng(123)
    sk, ck = T1.make_key_pair(rng)
    assert sk.params == T1.tfhe_parameters_80()
    try:
        eng = T1.engine_for(ck)
        assert eng.ctx.flags == T._cabi.FLAG_TORUS32
        bits = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)] * 4, bool)
        x, y, z = (T1.encrypt(rng, sk, bits[:, i]) for i in range(3))
        for name, (fn, nargs, plain) in T1.GATES.items():
            out = fn(ck, *(x, y, z)[:nargs])
            assert np.array_equal(T1.decrypt(sk, out), plain(*[bits[:, i] for i in range(nargs)])), name
        # flag validation: Torus32 mode is an N = 1024, l = 2..4 mode; unknown flag bits are refused
        for bad in ((500, 1024, 1, 5, 6, 8, 2, 1), (500, 1024, 1, 2, 17, 8, 2, 1), (500, 2048, 1, 1, 10, 8, 2, 1), (500, 1024, 1, 2, 10, 8, 2, 2)):
            with pytest.raises(T.MktfheError):
                T._cabi.Context(*bad[:7], flags=bad[7])
    finally:
        ck._engine.close()
