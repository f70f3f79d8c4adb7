"""Stage-by-stage algebraic pins of the oracle, written in numpy only (nothing below calls the oracle's own multiplication, rotation or
phase code to build an expectation).

The reference holds no golden vector for the 3gen path, so besides the truth tables of tests/test_oracle.py the
oracle is held to what the scheme's algebra demands of every intermediate value, with the secret keys in hand:

  * a bootstrapping-key element is a TGSW encryption of the LWE key bit under the joint RLWE key Z = sum_p z_p: its body rows have
    phase s * g_q, its mask rows phase -s * g_q * Z (tgsw_3gen.jl:80-83 read together with :102-113);
  * an external product multiplies the accumulator's phase by that bit (tgsw_3gen.jl:102-113);
  * after the blind rotation the accumulator's phase is X^(-(barb - sum bara * s)) * testvector (3gen_mk_internals.jl:59-95);
  * the extracted sample has, under Z as an LWE key of dimension N, the phase of coefficient 0 (rlwe.jl:70-74);
  * the key switch keeps that phase under the parties' LWE keys (mk_internals.jl:730-744, keyswitch.jl:45-80).

Each of these fixes a sign, an index direction or a layout convention that a truth table alone could leave to cancel out.
"""
import numpy as np


def negacyclic(a, b):
    """a * b mod (X^N + 1, 2^64), schoolbook on wrapped uint64 -- independent of oracle/mk_oracle.c."""
    a, b = np.asarray(a).astype(np.uint64), np.asarray(b).astype(np.uint64)
    N = a.size
    out = np.zeros(N, np.uint64)
    with np.errstate(over="ignore"):
        for i in range(N):
            if a[i] == 0:
                continue
            prod = a[i] * b
            out[i:] += prod[:N - i]
            out[:i] -= prod[N - i:]
    return out.astype(np.int64)


def rlwe_phase(acc, Z):
    """body - mask * Z for acc = [mask, body] (3gen accumulators: accum.a[1] = mask, accum.a[2] = body)."""
    with np.errstate(over="ignore"):
        return (acc[1].astype(np.uint64) - negacyclic(Z, acc[0]).astype(np.uint64)).astype(np.int64)


def monomial(poly, s):
    """poly * X^s mod X^N + 1 for any integer s."""
    N = poly.size
    s %= 2 * N
    out = np.empty_like(poly)
    for i in range(N):
        j = i + s
        sign = 1
        while j >= N:
            j -= N
            sign = -sign
        with np.errstate(over="ignore"):
            out[j] = poly[i] if sign > 0 else (np.uint64(0) - poly[i].astype(np.uint64)).astype(np.int64)
    return out


def torus(v):
    return np.asarray(v, np.int64).astype(np.float64) / 2.0 ** 64


def joint_key(ks):
    return ks.rlwe_keys.sum(axis=0).astype(np.int64)       # Z = sum of the parties' ternary RLWE keys


def test_bootstrapping_key_elements_are_tgsw_encryptions_of_the_lwe_key_bits(toy_keys):
    ks = toy_keys
    Z = joint_key(ks)
    bgbit = ks.prm.bgbit
    worst = 0.0
    for p in range(ks.k):
        for j in (0, 1, ks.n // 2, ks.n - 1):
            s = int(ks.lwe_keys[p, j])
            assert s in (0, 1)
            e = ks.bsk[p, j]                                 # [4][l][N]: part_1..part_4
            for q in range(ks.l):
                g = np.int64(1) << np.int64(64 - (q + 1) * bgbit)
                # row of a body digit: (mask, body) = (part_4, part_1) has phase s * g
                ph = rlwe_phase(np.stack([e[3, q], e[0, q]]), Z)
                exp = np.zeros(ks.N, np.int64)
                exp[0] = s * g
                worst = max(worst, np.abs(torus(ph - exp)).max())
                # row of a mask digit: (mask, body) = (part_3, part_2) has phase -s * g * Z
                ph = rlwe_phase(np.stack([e[2, q], e[1, q]]), Z)
                exp = -(s * g) * Z
                worst = max(worst, np.abs(torus(ph - exp)).max())
    assert worst < 2.0 ** -30, worst        # sigma_gsw = 2^-40 times at most (1 + |r . e|) ~ a few hundred


def test_external_product_multiplies_the_phase_by_the_key_bit(oracle, toy_keys, rng):
    ks = toy_keys
    Z = joint_key(ks)
    acc = rng.integers(-2 ** 63, 2 ** 63 - 1, size=(2, ks.N), dtype=np.int64)
    ph_in = rlwe_phase(acc, Z)
    seen = set()
    for p in range(ks.k):
        for j in range(ks.n):
            s = int(ks.lwe_keys[p, j])
            if (p, s) in seen:
                continue
            seen.add((p, s))
            out = ks.extprod(oracle.EXACT_SCHOOLBOOK, p, j, acc)
            err = torus(rlwe_phase(out, Z) - s * ph_in)
            # truncating decomposition (l * bgbit = 14 bits kept, floor -> bias) times the key norm, plus the key noise times the digits
            assert np.abs(err).max() < 2.0 ** -10, (p, j, s, np.abs(err).max())
    assert len(seen) == 2 * ks.k


def test_blind_rotation_turns_the_test_vector_by_the_phase_of_the_input(oracle, toy_keys):
    ks = toy_keys
    Z = joint_key(ks)
    N = ks.N
    mu = np.int64(1) << np.int64(61)
    bits = np.array([0, 1, 1, 0, 1], np.uint8)
    a, b = ks.encrypt(bits, 77)
    log2_2N = int(np.log2(2 * N))
    for g in range(bits.size):
        # mod switch, numeric-functions.jl:70-73: (x + 2^(31 - log2 2N)) >> (32 - log2 2N), wrapping add, arithmetic shift
        def bar(x):
            with np.errstate(over="ignore"):
                return (np.asarray(x, np.int32) + np.int32(1 << (31 - log2_2N))) >> np.int32(32 - log2_2N)
        bara, barb = bar(a[g]), int(bar(b[g]))
        rot = -barb + int((bara.astype(np.int64) * ks.lwe_keys).sum())
        ext_a, ext_b, acc, _ = ks.bootstrap_wo_keyswitch(oracle.EXACT_SCHOOLBOOK, int(mu), a[g], b[g], want_acc=True)
        expect = monomial(np.full(N, mu, np.int64), rot)
        err = torus(rlwe_phase(acc, Z) - expect)
        assert np.abs(err).max() < 0.03, (g, np.abs(err).max())
        # the sign of coefficient 0 is the bit: rot = -round(2N * phase), phase = +-1/8
        assert (expect[0] > 0) == bool(bits[g])
        # sample extraction (rlwe.jl:70-74, polynomials.jl:69-72): under Z as an LWE key the phase is coefficient 0, on 32 bits
        with np.errstate(over="ignore"):
            lwe_phase = np.int32(ext_b) - np.int32((ext_a.astype(np.int64) * Z).sum() & 0xFFFFFFFF)
        ph0 = rlwe_phase(acc, Z)[0]
        assert abs(float(np.int32(lwe_phase)) / 2 ** 32 - float(ph0) / 2 ** 64) < (N + 2) * 2.0 ** -32
        # t64tot32 truncates toward zero through Float64 (numeric-functions.jl:109-111)
        assert int(ext_b) == int(np.trunc(float(acc[1][0]) / 2.0 ** 32))
        assert int(ext_a[0]) == int(np.trunc(float(acc[0][0]) / 2.0 ** 32))
        assert int(ext_a[1]) == int(np.trunc(float(-acc[0][N - 1]) / 2.0 ** 32))
        # key switch (mk_internals.jl:730-744): same phase under the parties' LWE keys, up to the key-switch noise and rounding
        oa, ob = ks.keyswitch(ext_a, ext_b)
        with np.errstate(over="ignore"):
            out_phase = np.int32(ob) - np.int32((oa.astype(np.int64) * ks.lwe_keys).sum() & 0xFFFFFFFF)
        assert abs(float(np.int32(out_phase)) / 2 ** 32 - float(np.int32(lwe_phase)) / 2 ** 32) < 0.02
        assert np.array_equal(ks.phase(oa[None], np.array([ob])), np.array([out_phase], np.int32))


def test_rotation_step_is_a_cmux_on_the_key_bit(oracle, toy_keys, rng):
    """mk_mux_rotate_3gen (3gen_mk_internals.jl:59-62): acc + ExtProd(X^bara * acc - acc, bsk[p][j]) has the phase of
    X^(bara * s) * acc -- the exponent is +bara for a key bit of one and the accumulator is untouched (in phase) for zero."""
    ks = toy_keys
    Z = joint_key(ks)
    acc = np.zeros((2, ks.N), np.int64)
    acc[1] = rng.integers(-2 ** 61, 2 ** 61, size=ks.N, dtype=np.int64)      # trivial sample: phase == body
    for p in range(ks.k):
        for want in (0, 1):
            j = int(np.flatnonzero(ks.lwe_keys[p] == want)[0])
            for bara in (5, -9, ks.N + 3):
                out = ks.mux_rotate(oracle.EXACT_SCHOOLBOOK, p, j, bara, acc)
                expect = monomial(acc[1], bara * want)
                err = torus(rlwe_phase(out, Z) - expect)
                assert np.abs(err).max() < 2.0 ** -12, (p, j, want, bara, np.abs(err).max())
