"""Prototype (numpy) of the FP64 FFT channel of the external product, with exactly the butterfly networks, index conventions and
tables the CUDA kernel (torus-fhe_b200/csrc/fft64.cuh) uses.  Checks exactness against the wrap-around integer product.

Ring: real negacyclic polynomials mod X^N + 1, N = 1024, folded to C[X]/(X^M - i), M = 512:  a~[j] = a[j] + i a[j + M].
forward : Cooley-Tukey butterflies, natural -> bit-reversed, twiddle per group (twist merged): stage d = 0..8, span 256 >> d,
          group g = pos >> (9 - d), s(d, g) = exp(i theta / 2), theta = pi / 2^(d+1) + 2 pi bitrev_d(g) / 2^d
inverse : classic decimation-in-time, bit-reversed -> natural, twiddle per position: span sp = 1..256, w = exp(-2 pi i (pos mod sp) / (2 sp)),
          then untwist by zeta^-j, zeta = exp(i pi / N); the 1 / M is folded into the key spectrum
key     : Torus64 coefficient split in three balanced limbs of 22 / 21 / 21 bits; every limb product is an integer below 2^39, recovered by
          rounding; R = r0 + (r1 << 22) + (r2 << 43) mod 2^64
"""
import numpy as np

N, M = 1024, 512
LIMB_BITS = (22, 21, 21)
LIMB_SHIFT = (0, 22, 43)


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def fwd_twiddle(d, g):
    theta = np.pi / 2 ** (d + 1) + 2 * np.pi * bitrev(g, d) / 2 ** d
    return np.exp(0.5j * theta)


def forward(x):
    """x: complex [M] folded coefficients (natural order) -> spectrum in bit-reversed order."""
    x = x.astype(np.complex128).copy()
    pos = np.arange(M)
    for d in range(9):
        sp = 256 >> d
        lo = pos[(pos & sp) == 0]
        s = np.array([fwd_twiddle(d, p >> (9 - d)) for p in lo])
        t = s * x[lo + sp]
        x[lo + sp] = x[lo] - t
        x[lo] = x[lo] + t
    return x


def inverse(X):
    """spectrum (bit-reversed order) -> M * folded coefficients (natural order), untwisted."""
    x = X.astype(np.complex128).copy()
    pos = np.arange(M)
    sp = 1
    while sp < M:
        lo = pos[(pos & sp) == 0]
        w = np.exp(-2j * np.pi * (lo % sp) / (2 * sp))
        t = w * x[lo + sp]
        x[lo + sp] = x[lo] - t
        x[lo] = x[lo] + t
        sp *= 2
    return x * np.exp(-1j * np.pi * pos / N)


def fold(a):
    return a[:M].astype(np.float64) + 1j * a[M:].astype(np.float64)


def split_limbs(k):
    """int64 [..] -> three balanced limbs (float64-exact ints): k = l0 + l1 2^22 + l2 2^43 mod 2^64."""
    k = k.astype(np.int64)
    out = []
    for bits in LIMB_BITS[:-1]:
        half = 1 << (bits - 1)
        l = ((k + half) & ((1 << bits) - 1)) - half
        out.append(l)
        k = (k - l) >> bits
    out.append(k)          # top limb: what is left (signed, 21 bits)
    return out


def negacyclic_exact(a, b):
    """wrap-around int64 negacyclic product."""
    full = np.convolve(a.astype(np.int64), b.astype(np.int64))
    res = full[:N].copy()
    res[: N - 1] -= full[N:]
    return res


def extprod_fft(digs, keys):
    """sum_s digs[s] * keys[s] mod (X^N + 1, 2^64) through the FFT channel; returns (result int64 [N], max rounding distance)."""
    spec = [forward(fold(d)) for d in digs]
    worst = 0.0
    R = np.zeros(N, np.int64)
    limbs = [split_limbs(k) for k in keys]
    for li in range(3):
        acc = np.zeros(M, np.complex128)
        for s in range(len(digs)):
            kspec = forward(fold(limbs[s][li])) / M
            acc += spec[s] * kspec
        c = inverse(acc)
        r = np.concatenate([c.real, c.imag])
        ri = np.rint(r)
        worst = max(worst, np.abs(r - ri).max())
        R += ri.astype(np.int64) << LIMB_SHIFT[li]
    return R, worst


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    # network sanity: forward output = evaluations at zeta omega^k in bit-reversed order
    a = rng.integers(-64, 64, N)
    A = forward(fold(a))
    j = np.arange(M)
    for k in (0, 1, 5, 300):
        xk = np.exp(1j * np.pi / N) * np.exp(2j * np.pi * k / M)
        assert abs(A[bitrev(k, 9)] - (fold(a) * xk ** j).sum()) < 1e-6
    assert np.abs(inverse(A) / M - fold(a)).max() < 1e-9
    worst_all = 0.0
    for trial in range(20):
        L2 = 4
        if trial < 3:      # extreme operands: every digit at -64, every key word at +-2^63 / limb maxima
            digs = [np.full(N, -64, np.int64) for _ in range(L2)]
            keys = [np.full(N, v, np.int64) for v in (np.int64(-2 ** 63), np.int64(2 ** 63 - 1), np.int64(0x1FFFFF1FFFFF), np.int64(-1))]
            if trial == 1:
                sg = rng.integers(0, 2, (L2, N)) * 2 - 1
                keys = [k * s for k, s in zip(keys, sg)]
            if trial == 2:
                digs = [np.where(rng.integers(0, 2, N) == 1, 63, -64) for _ in range(L2)]
        else:
            digs = [rng.integers(-64, 64, N) for _ in range(L2)]
            keys = [rng.integers(-2 ** 63, 2 ** 63 - 1, N, dtype=np.int64) for _ in range(L2)]
        want = np.zeros(N, np.int64)
        for d, k in zip(digs, keys):
            want += negacyclic_exact(d, k)
        got, worst = extprod_fft(digs, keys)
        worst_all = max(worst_all, worst)
        assert np.array_equal(got, want), f"trial {trial}: mismatch"
    print(f"FFT channel exact on 20 trials; worst distance to an integer {worst_all:.3e} (2^{np.log2(worst_all):.1f})")
