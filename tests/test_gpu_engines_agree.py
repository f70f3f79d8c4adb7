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
