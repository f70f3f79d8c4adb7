"""TEST INFRASTRUCTURE -- CPU restatement (numpy, Torus32) of the reference's single-key TFHE bootstrapped-gate path.
Only tests/ may import this file; the product never does.

Follows, line by line, 3-gen-mk-tfhe/src/:
  numeric-functions.jl:70-73, 86-89   decode_message / encode_message
  tgsw.jl:1-34, 112-138, 143-147      TGswParams (gadget values, offset), decompose, tgsw_extern_mul
  rlwe.jl:64-68, 113-131              rlwe_extract_sample, rlwe_noiseless_trivial, mul_by_monomial on samples
  polynomials.jl:69-72                reverse_polynomial
  bootstrap.jl:20-100                 mux_rotate, blind_rotate, blind_rotate_and_extract, bootstrap_wo_keyswitch, bootstrap
  keyswitch.jl:45-80                  keyswitch
  gates.jl:16-177                     the twelve gates
The reference multiplies through a Float64 FFT (polynomials.jl:208-242); for Torus32 operands and gadget digits of at most 10 bits
that transform is exact after rounding (53-bit significand against 10 + 32 + 10 bits), so the exact integer product used here IS
the reference's arithmetic -- unlike the Torus64 3gen path, no tolerance is involved.
Parity pin: the reference holds no golden vector for this path either (test/runtests.jl:10-42 is a decrypted truth table with
MersenneTwister(123), reproducible only in Julia): PARITY UNPINNED, checked functionally through the same truth table.
Pure numpy loops: meant for reduced LWE dimensions (a bootstrap costs n * 4l products of degree N).
"""
import numpy as np


def wrap32(x):
    return np.asarray(x).astype(np.int64).astype(np.uint32).view(np.int32) if np.ndim(x) else np.int32(np.uint32(int(x) & 0xFFFFFFFF))


def encode_message(mu, space):
    lg = int(space).bit_length() - 1
    return int(wrap32(int(mu) << (32 - lg)))                       # numeric-functions.jl:86-89


def decode_message(phase, space):
    lg = int(space).bit_length() - 1                                # numeric-functions.jl:70-73
    p = np.asarray(phase, dtype=np.int64)
    return (wrap32(p + (1 << (32 - lg - 1))).astype(np.int64) >> (32 - lg)).astype(np.int64)


def negacyclic_mul(small, big):
    """Exact product mod (X^N + 1, 2^32): `small` has |coeff| < 2^11, `big` is Torus32 (the sum stays below 2^53)."""
    N = small.shape[-1]
    full = np.convolve(small.astype(np.int64), big.astype(np.int64))
    out = full[:N].copy()
    out[:N - 1] -= full[N:]
    return wrap32(out)


def mul_by_monomial(poly, s):
    """X^s * poly mod X^N + 1, s any integer (DarkIntegers' mul_by_monomial; rlwe.jl:130-131)."""
    N = poly.shape[-1]
    s %= 2 * N
    neg = s >= N
    s %= N
    out = np.concatenate([-poly[..., N - s:], poly[..., :N - s]], axis=-1) if s else poly.copy()
    return wrap32(-out.astype(np.int64) if neg else out)


def decompose(poly, l, bgbit):
    """tgsw.jl:112-138 with bit = 32: l digit polynomials in [-Bg/2, Bg/2)."""
    offset = sum(1 << (32 - q * bgbit) for q in range(1, l + 1)) << (bgbit - 1)           # tgsw.jl:26-30
    t = wrap32(poly.astype(np.int64) + offset).astype(np.int64)
    return np.stack([((t >> (32 - q * bgbit)) & ((1 << bgbit) - 1)) - (1 << (bgbit - 1)) for q in range(1, l + 1)])


def tgsw_extern_mul(accum, sample, l, bgbit):
    """tgsw.jl:143-147: accum int32 [2 (mask, body)][N]; sample int32 [l][2 (row j)][2 (mask, body)][N] = TGswSample.samples[q, j].a.
    result = sum_{q, j} decompose(accum.a[j])[q] * samples[q, j]."""
    out = np.zeros((2, accum.shape[-1]), np.int64)
    for j in range(2):
        d = decompose(accum[j], l, bgbit)
        for q in range(l):
            for c in range(2):
                out[c] += negacyclic_mul(d[q], sample[q, j, c])
    return wrap32(out)


def mux_rotate(accum, bki, barai, l, bgbit):
    """bootstrap.jl:20-24."""
    temp = wrap32(mul_by_monomial(accum, int(barai)).astype(np.int64) - accum)
    return wrap32(accum.astype(np.int64) + tgsw_extern_mul(temp, bki, l, bgbit))


def blind_rotate_and_extract(mu, bk, barb, bara, l, bgbit):
    """bootstrap.jl:37-67 + rlwe_extract_sample (rlwe.jl:64-68) + reverse_polynomial (polynomials.jl:69-72)."""
    N = bk.shape[-1]
    accum = np.stack([np.zeros(N, np.int32), mul_by_monomial(np.full(N, mu, np.int32), -int(barb))])
    for i in range(bk.shape[0]):
        if bara[i] != 0:
            accum = mux_rotate(accum, bk[i], bara[i], l, bgbit)
    a = np.concatenate([accum[0][:1], wrap32(-accum[0][:0:-1].astype(np.int64))])
    return a, np.int32(accum[1][0])


def bootstrap_wo_keyswitch(bk, mu, xa, xb, l, bgbit):
    """bootstrap.jl:73-86."""
    N = bk.shape[-1]
    return blind_rotate_and_extract(mu, bk, decode_message(xb, 2 * N), decode_message(xa, 2 * N), l, bgbit)


def keyswitch(ksk, a, b, t, basebit):
    """keyswitch.jl:45-80: ksk int32 [N][t][base-1][n+1]."""
    n = ksk.shape[-1] - 1
    res = np.zeros(n + 1, np.int64)
    res[n] = b
    aibar = wrap32(a.astype(np.int64) + (1 << (32 - (1 + basebit * t)))).astype(np.int64)
    for j in range(1, t + 1):
        dig = (aibar >> (32 - j * basebit)) & ((1 << basebit) - 1)
        nz = np.nonzero(dig)[0]
        res -= ksk[nz, j - 1, dig[nz] - 1].astype(np.int64).sum(0)
    res = wrap32(res)
    return res[:n], np.int32(res[n])


def bootstrap(bk, ksk, mu, xa, xb, l, bgbit, t, basebit):
    """bootstrap.jl:97-100."""
    a, b = bootstrap_wo_keyswitch(bk, mu, xa, xb, l, bgbit)
    return keyswitch(ksk, a, b, t, basebit)


# gates.jl:16-142: (mu0 numerator, message space, cx, cy)
GATE_LINEAR = {"NAND": (1, 8, -1, -1), "OR": (1, 8, 1, 1), "AND": (-1, 8, 1, 1), "XOR": (1, 4, 2, 2), "XNOR": (-1, 4, -2, -2), "NOR": (-1, 8, -1, -1),
               "ANDNY": (-1, 8, -1, 1), "ANDYN": (-1, 8, 1, -1), "ORNY": (1, 8, -1, 1), "ORYN": (1, 8, 1, -1)}


def gate(name, bk, ksk, prm, x, y=None, z=None):
    """One gate on single samples x = (a int32 [n], b); prm = (l, bgbit, t, basebit)."""
    l, bgbit, t, basebit = prm
    mu = encode_message(1, 8)
    lin = lambda mu0, *terms: (wrap32(sum(c * s[0].astype(np.int64) for c, s in terms)), wrap32(mu0 + sum(c * int(s[1]) for c, s in terms)))
    if name == "NOT":                                             # gates.jl:80-83
        return wrap32(-x[0].astype(np.int64)), wrap32(-int(x[1]))
    if name == "MUX":                                             # gates.jl:166-177
        t1 = lin(encode_message(-1, 8), (1, x), (1, y))
        t2 = lin(encode_message(-1, 8), (-1, x), (1, z))
        u1 = bootstrap_wo_keyswitch(bk, mu, t1[0], t1[1], l, bgbit)
        u2 = bootstrap_wo_keyswitch(bk, mu, t2[0], t2[1], l, bgbit)
        t3 = lin(mu, (1, u1), (1, u2))
        return keyswitch(ksk, t3[0], t3[1], t, basebit)
    m, space, cx, cy = GATE_LINEAR[name]
    temp = lin(encode_message(m, space), (cx, x), (cy, y))
    return bootstrap(bk, ksk, mu, temp[0], temp[1], l, bgbit, t, basebit)


# ---- keys and samples for the oracle's own functional pin (numpy; independent of the product's key generation) -------------------
def keygen(rng, n, N, l, bgbit, t, basebit, sigma_bs, sigma_ks):
    """api.jl:212-228: LWE key, binary RLWE key, BootstrapKey (tgsw_encrypt per key bit, tgsw.jl:85-108), KeyswitchKey (keyswitch.jl:14-41)."""
    s = rng.integers(0, 2, n).astype(np.int32)
    z = rng.integers(0, 2, N).astype(np.int32)
    bk = np.empty((n, l, 2, 2, N), np.int32)
    for i in range(n):
        for q in range(l):
            for j in range(2):
                mask = rng.integers(-2 ** 31, 2 ** 31, N).astype(np.int32)
                e = np.trunc(rng.standard_normal(N) * sigma_bs * 2.0 ** 32).astype(np.int64)
                body = wrap32(e + negacyclic_mul(z, mask))
                row = np.stack([mask, body])
                row[j, 0] = wrap32(int(row[j, 0]) + int(s[i]) * (1 << (32 - (q + 1) * bgbit)))
                bk[i, q, j] = row
    B1 = (1 << basebit) - 1
    noise = rng.standard_normal((N, t, B1)) * sigma_ks
    noise -= noise.mean()
    a = rng.integers(-2 ** 31, 2 ** 31, (N, t, B1, n)).astype(np.int32)
    h = np.arange(1, B1 + 1, dtype=np.int64)[None, None, :]
    sh = (32 - np.arange(1, t + 1) * basebit)[None, :, None]
    msg = (z.astype(np.int64)[:, None, None] * h) << sh
    b = wrap32(msg + np.trunc(noise * 2.0 ** 32).astype(np.int64) + (a.astype(np.int64) * s).sum(-1))
    return s, z, bk, np.concatenate([a, b[..., None]], axis=-1)


def encrypt(rng, s, bit, sigma):
    a = rng.integers(-2 ** 31, 2 ** 31, s.size).astype(np.int32)
    e = int(np.trunc(rng.standard_normal() * sigma * 2.0 ** 32))
    return a, wrap32(encode_message(1 if bit else -1, 8) + e + int((a.astype(np.int64) * s).sum()))


def decrypt(s, x):
    return int(wrap32(int(x[1]) - int((x[0].astype(np.int64) * s).sum()))) > 0
