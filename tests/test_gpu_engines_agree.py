"""The two external-product engines of the library -- the exact FP64 FFT channel (csrc/fft64.cuh, the default of the N = 1024 Torus64 sets) and the
three-prime RNS NTT kernels (csrc/kernels.cuh, MKTFHE_B200_FFT=0) -- on identical keys and inputs: every output must agree bit for bit, at the
full 2-party parameters and on a 4-party (l = 3) and an 8-party-shaped (l = 4) set.  Both are separately held against the CPU oracle in
test_gpu_parity.py; this test keeps the RNS kernels exercised now that they are no longer the default, and covers batch shapes (full waves, a
tail, single gates) where the two engines launch different kernels."""
import os

import numpy as np
import pytest

from conftest import make_engine

pytestmark = pytest.mark.gpu


def engine_with(ks, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        eng = make_engine(ks)       # the library reads its environment in mktfhe_create
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return eng


def check_engines(ks, G, rng, extprods=8):
    import torus_fhe_b200 as T
    p = ks.prm
    fft, ntt = engine_with(ks, MKTFHE_B200_FFT=1), engine_with(ks, MKTFHE_B200_FFT=0)
    try:
        assert fft.ctx.describe()["external_product"] == "fft64" and ntt.ctx.describe()["external_product"] == "ntt_rns"
        # single external products on arbitrary accumulators (extremes included)
        acc = rng.integers(-2 ** 63, 2 ** 63 - 1, size=(extprods, 2, p.N), dtype=np.int64)
        acc[0], acc[1], acc[2] = 0, np.int64(2 ** 63 - 1), np.int64(-2 ** 63)
        elem = rng.integers(0, p.k * p.n, size=extprods).astype(np.int32)
        assert np.array_equal(fft.ctx.extprod_batch(elem, acc), ntt.ctx.extprod_batch(elem, acc))
        # whole gates on uniformly random ciphertext words (every rotation amount occurs), all five bootstrapped gates
        xa, ya, za = (rng.integers(-2 ** 31, 2 ** 31 - 1, size=(G, p.k, p.n), dtype=np.int64).astype(np.int32) for _ in range(3))
        xb, yb, zb = (rng.integers(-2 ** 31, 2 ** 31 - 1, size=G, dtype=np.int64).astype(np.int32) for _ in range(3))
        for gate in (T._cabi.GATE_NAND, T._cabi.GATE_XOR, T._cabi.GATE_AND3):
            z = ((za, zb),) if gate == T._cabi.GATE_AND3 else ()
            a1, b1 = fft.ctx.gate_batch(gate, (xa, xb), (ya, yb), *z)
            a2, b2 = ntt.ctx.gate_batch(gate, (xa, xb), (ya, yb), *z)
            assert np.array_equal(a1, a2) and np.array_equal(b1, b2), gate
        # accumulators and extracted samples of the blind rotation alone, a single gate (the one-gate-per-CTA launches)
        e1, c1 = fft.ctx.blind_rotate_batch(1 << 61, xa[:1], xb[:1], want_acc=True)
        e2, c2 = ntt.ctx.blind_rotate_batch(1 << 61, xa[:1], xb[:1], want_acc=True)
        assert np.array_equal(e1, e2) and np.array_equal(c1, c2)
    finally:
        fft.close()
        ntt.close()


def test_engines_agree_2party_full_parameters(keys2, rng):
    """296 gates per wave on a B200: 2 full waves + a tail of 40 gates + (inside check_engines) single-gate launches."""
    check_engines(keys2, 2 * 296 + 40, rng)


@pytest.mark.parametrize("l,bgbit,k", [(3, 6, 4), (4, 4, 3), (1, 8, 1)])
def test_engines_agree_other_gadget_shapes(oracle, rng, l, bgbit, k):
    prm = dict(n=40, N=1024, k=k, l=l, bgbit=bgbit, t=5, basebit=2, sigma_lwe=2.0 ** -20, sigma_gsw=2.0 ** -40, sigma_ks=2.0 ** -20)
    ks = oracle.KeySet(prm, seed=100 + l, nthreads=os.cpu_count() or 8)
    check_engines(ks, 300, rng, extprods=6)


@pytest.mark.parametrize("l,bgbit", [(2, 10), (3, 9), (4, 8)])
def test_engines_agree_torus32_mode(rng, l, bgbit):
    """Torus32 mode (MKTFHE_FLAG_TORUS32: unshifted 32-bit keys, 16-bit digit fields, products added as R << 32) -- the gadget shapes of
    tfhe_parameters_80 and of the CCS sets: external products and whole bootstraps of the FFT channel (two 16-bit key limbs) against the RNS
    kernels on uniformly random key words (nothing of a real key's structure is needed for the two engines to have to agree)."""
    import torus_fhe_b200 as T
    n, N, k, t, bb = 24, 1024, 1, 8, 2
    sp = T.SchemeParameters_3gen(n, 2.0 ** -15, N, 1, False, l, bgbit, 2.0 ** -25, t, bb, 2.0 ** -15, k)
    bsk = rng.integers(-2 ** 31, 2 ** 31, size=(n, 4, l, N), dtype=np.int64)          # 32-bit values in int64 words
    bsk[0, 0, 0, :8] = [2 ** 31 - 1, -2 ** 31, 0, -1, 1, 2 ** 15, -2 ** 15, 2 ** 15 - 1]
    ksk = rng.integers(-2 ** 31, 2 ** 31, size=(N, t, (1 << bb) - 1, n + 1), dtype=np.int64).astype(np.int32)
    engines = []
    try:
        for f in (1, 0):
            os.environ["MKTFHE_B200_FFT"] = str(f)
            try:
                eng = T.Engine(sp, device=0, flags=T._cabi.FLAG_TORUS32)
            finally:
                os.environ.pop("MKTFHE_B200_FFT", None)
            eng.load_keys([bsk], [ksk])
            engines.append(eng)
        fft, ntt = engines
        assert fft.ctx.describe()["external_product"] == "fft64" and ntt.ctx.describe()["external_product"] == "ntt_rns"
        acc = rng.integers(-2 ** 31, 2 ** 31, size=(16, 2, N), dtype=np.int64) << 32
        acc[0], acc[1], acc[2] = 0, np.int64(-1) << 32, np.int64(-2 ** 31) << 32
        elem = rng.integers(0, n, size=16).astype(np.int32)
        assert np.array_equal(fft.ctx.extprod_batch(elem, acc), ntt.ctx.extprod_batch(elem, acc))
        for G in (5, 300):          # one gate per CTA / two gates per CTA
            xa = rng.integers(-2 ** 31, 2 ** 31, size=(G, k, n), dtype=np.int64).astype(np.int32)
            xb = rng.integers(-2 ** 31, 2 ** 31, size=G, dtype=np.int64).astype(np.int32)
            a1, b1 = fft.ctx.bootstrap_batch((1 << 29) << 32, xa, xb)
            a2, b2 = ntt.ctx.bootstrap_batch((1 << 29) << 32, xa, xb)
            assert np.array_equal(a1, a2) and np.array_equal(b1, b2), G
    finally:
        for e in engines:
            e.close()
