"""Single-key TFHE (the reference's base scheme: 3-gen-mk-tfhe/src/api.jl, bootstrap.jl, gates.jl, tgsw.jl, keyswitch.jl)
behind the SAME engine and kernels as the 3gen multi-key path -- SURVEY.md section 8(f) rank 4, first slice.

How the scheme maps onto the engine (no new kernel):
  * one party (k = 1); an `LweSample` is the k = 1 case of the engine's ciphertext batch layout int32 a[G][1][n], b[G];
  * the scheme is Torus32 (`rlwe_is32 = true`); the engine's accumulator is Torus64.  Every Torus32 value v is carried as v << 32:
    the gadget digits of the top l*Bgbit <= 32 bits are the same numbers (tgsw.jl:112-138 with bit = 32 vs 64: the offset's low 32
    bits are zero), products with keys stored as K << 32 are exact mod 2^64 and their top halves are the Torus32 products mod 2^32,
    and extraction's t64tot32 (numeric-functions.jl:109-111) of an exact multiple of 2^32 is the Torus32 value itself, i.e.
    rlwe_extract_sample (rlwe.jl:64-68);
  * a standard TGSW sample (tgsw.jl:36-46; samples[q, j], j = 1 the row with the gadget on the mask, j = 2 on the body) fills the
    engine's four parts: part_1 (body <- body digits) = samples[q, 2].a[2], part_2 (body <- mask digits) = samples[q, 1].a[2],
    part_3 (mask <- mask digits) = samples[q, 1].a[1], part_4 (mask <- body digits) = samples[q, 2].a[1]
    (tgsw_extern_mul, tgsw.jl:143-147, against tgsw_extern_mul_3gen, tgsw_3gen.jl:102-113);
  * the key switch is keyswitch.jl:45-80 for one party, which is what the multi-key one runs per party.
Every bootstrapped gate is one `mktfhe_affine_bootstrap_batch` call (linear prologue fused into the kernel); gate_mux follows
gates.jl:166-177: two bootstraps without key switch, the OR in the extracted domain, one key switch.

Supported: parameter sets with N = 1024 and l*Bgbit <= 32.  Bgbit <= 8 (tfhe_parameters_128: l = 3, Bg = 2^7) runs on the default
kernels as described above.  tfhe_parameters_80 (l = 2, Bg = 2^10) has 10-bit digits and, embedded as v << 32, would leave the exact
range of the three-prime CRT (2l N Bg/2 2^63 = 2^84): it runs in the library's Torus32 mode (MKTFHE_FLAG_TORUS32: 16-bit digit
fields, keys loaded unshifted, every external product added as R << 32 -- kernels.cuh, blind_rotate_t32_kernel).  In the fork
neither constructor runs as written (api.jl:76-113 pass 11 values to the 12-field struct after `rlwe_is32` was added, :50-67);
here `rlwe_is32 = true` is filled in.  Names, arguments and behaviour follow the reference; samples may carry leading batch
dimensions (a scalar call is a batch of one).
"""
from dataclasses import dataclass

import numpy as np

from . import _cabi
from .engine import Engine
from .tfhe3gen import (KeyswitchKey, KeyswitchParameters, LweKey, LweParams, RLweKey, RLweParams, SchemeParameters_3gen, TGswParams, dtot32,
                       encode_message, negacyclic_mul, rand_uniform_torus32)


@dataclass(frozen=True)
class SchemeParameters:
    """api.jl:50-67 (field order of the struct, with rlwe_is32 where the struct has it)."""
    lwe_size: int
    lwe_noise_stddev: float
    rlwe_polynomial_degree: int
    rlwe_mask_size: int
    rlwe_is32: bool
    bs_decomp_length: int
    bs_log2_base: int
    bs_noise_stddev: float
    ks_decomp_length: int
    ks_log2_base: int
    ks_noise_stddev: float
    max_parties: int


def tfhe_parameters_80(rlwe_mask_size=1):
    """api.jl:76-91 (CGGI16, ~80 bits; the reference's default): Bg = 2^10, served by the library's Torus32 mode."""
    return SchemeParameters(500, 1 / 2 ** 15 * np.sqrt(2 / np.pi), 1024, rlwe_mask_size, True, 2, 10, 9e-9 * np.sqrt(2 / np.pi), 8, 2,
                            1 / 2 ** 15 * np.sqrt(2 / np.pi), 1)


def tfhe_parameters_128(rlwe_mask_size=1):
    """api.jl:100-113 (CGGI19, ~128 bits)."""
    return SchemeParameters(630, 1 / 2 ** 15, 1024, rlwe_mask_size, True, 3, 7, 1 / 2 ** 25, 8, 2, 1 / 2 ** 15, 1)


def lwe_parameters(p):
    return LweParams(p.lwe_size)


def rlwe_parameters(p):
    return RLweParams(p.rlwe_polynomial_degree, p.rlwe_mask_size, p.rlwe_is32)


def tgsw_parameters(p):
    return TGswParams(p.bs_decomp_length, p.bs_log2_base, p.rlwe_is32)


def keyswitch_parameters(p):
    return KeyswitchParameters(p.ks_decomp_length, p.ks_log2_base)


class LweSample:
    """lwe.jl:23-33 with the linear operations of :62-76.  `a` is int32 [.., n], `b` int32 [..]."""
    __slots__ = ("params", "a", "b", "current_variance")
    __array_ufunc__ = None

    def __init__(self, params, a, b, current_variance=0.0):
        self.params = params
        self.a = np.asarray(a, dtype=np.int32)
        self.b = np.asarray(b, dtype=np.int32)
        self.current_variance = current_variance

    def __getitem__(self, idx):
        return LweSample(self.params, self.a[idx], self.b[idx], self.current_variance)

    def __add__(self, y):
        with np.errstate(over="ignore"):
            return LweSample(self.params, self.a + y.a, self.b + y.b, self.current_variance + y.current_variance)

    def __sub__(self, y):
        with np.errstate(over="ignore"):
            return LweSample(self.params, self.a - y.a, self.b - y.b, self.current_variance + y.current_variance)

    def __neg__(self):
        with np.errstate(over="ignore"):
            return LweSample(self.params, -self.a, -self.b, self.current_variance)

    def __mul__(self, y):
        y = np.int32(y)
        with np.errstate(over="ignore"):
            return LweSample(self.params, self.a * y, self.b * y, self.current_variance * float(y) ** 2)

    __rmul__ = __mul__


def lwe_noiseless_trivial(mu, params, batch_shape=()):
    """lwe.jl:58-59."""
    return LweSample(params, np.zeros(tuple(batch_shape) + (params.size,), np.int32), np.full(batch_shape, np.int32(mu), np.int32), 0.0)


def lwe_encrypt(rng, message, alpha, key, noise=None):
    """lwe.jl:36-53 on an array of messages (Torus32)."""
    msg = np.asarray(message, dtype=np.int32)
    n = key.params.size
    a = rand_uniform_torus32(rng, msg.size * n).reshape(msg.shape + (n,))
    e = dtot32(rng.standard_normal(msg.shape) * alpha if noise is None else noise).astype(np.int64)
    b = (msg.astype(np.int64) + e + (a.astype(np.int64) * key.key.astype(np.int64)).sum(-1)).astype(np.int32)
    return LweSample(key.params, a, b, alpha ** 2)


def lwe_phase(x, key):
    """lwe.jl:56."""
    return (x.b.astype(np.int64) - (x.a.astype(np.int64) * key.key.astype(np.int64)).sum(-1)).astype(np.int32)


class SecretKey:   # api.jl:176-184
    def __init__(self, rng, params, key=None):
        self.params = params
        self.key = LweKey(rng, lwe_parameters(params))
        if key is not None:
            self.key.key = np.asarray(key, dtype=np.int32)


class BootstrapKey:
    """bootstrap.jl:1-17: n TGSW encryptions (tgsw.jl:85-108) of the LWE key bits under the RLWE key.  `samples` is the integer key,
    int32 [n][l][2 (row j)][2 (mask, body)][N]; the reference's forward_transform of it is what the engine does at load time."""

    def __init__(self, rng, alpha, lwe_key, rlwe_key, tgsw_params, samples=None):
        self.tgsw_params, self.rlwe_params = tgsw_params, rlwe_key.params if rlwe_key is not None else None
        if samples is not None:
            self.samples = np.asarray(samples, dtype=np.int32)
            return
        n, l, N = lwe_key.params.size, tgsw_params.decomp_length, rlwe_key.params.polynomial_degree
        # rlwe_encrypt_zero (rlwe.jl:77-103): mask uniform, body = e + key * mask (exact product; Torus32 wrap)
        mask = rand_uniform_torus32(rng, n * l * 2 * N).reshape(n, l, 2, N)
        e = dtot32(rng.standard_normal((n, l, 2, N)) * alpha)
        prod = negacyclic_mul(rlwe_key.key.astype(np.int64)[None, None, None, :], mask.astype(np.int64))      # exact mod 2^64: low 32 bits are the
        body = (e.astype(np.int64) + prod).astype(np.int32)                                                # product mod 2^32
        s = np.stack([mask, body], axis=3)                                                                   # [n][l][j][mask/body][N]
        # tgsw_add_gadget_times_message (tgsw.jl:57-81): row j gets message * gadget[q] on polynomial j (constant coefficient)
        g = np.array([1 << (32 - q * tgsw_params.log2_base) for q in range(1, l + 1)], dtype=np.int64)
        mg = (lwe_key.key.astype(np.int64)[:, None] * g[None, :]).astype(np.int32)
        with np.errstate(over="ignore"):
            s[:, :, 0, 0, 0] += mg
            s[:, :, 1, 1, 0] += mg
        self.samples = s

    def engine_parts(self, shift=32):
        """int64 [n][4][l][N] in the engine's part order; every Torus32 value carried as v << 32 for the default kernels (module
        docstring), or unshifted (shift = 0) for the library's Torus32 mode (MKTFHE_FLAG_TORUS32: gadget bases above 2^8)."""
        s = self.samples.astype(np.int64) << shift
        return np.ascontiguousarray(np.stack([s[:, :, 1, 1], s[:, :, 0, 1], s[:, :, 0, 0], s[:, :, 1, 0]], axis=1))


class CloudKey:   # api.jl:212-228
    def __init__(self, rng, secret_key, rlwe_key=None):
        params = secret_key.params
        self.params = params
        self.rlwe_key = rlwe_key if rlwe_key is not None else RLweKey(rng, rlwe_parameters(params), negative_random=False)
        self.bootstrap_key = BootstrapKey(rng, params.bs_noise_stddev, secret_key.key, self.rlwe_key, tgsw_parameters(params))
        self.keyswitch_key = KeyswitchKey(rng, params.ks_noise_stddev, keyswitch_parameters(params), secret_key.key, self.rlwe_key)
        self._engine = None


def make_key_pair(rng, params=None):
    """api.jl:237-245 (default parameters: tfhe_parameters_80, as in the reference)."""
    if params is None:
        params = tfhe_parameters_80()
    sk = SecretKey(rng, params)
    return sk, CloudKey(rng, sk)


def encrypt(rng, key, message):
    """api.jl:253-256; `message` may be an array of bools."""
    m = np.asarray(message, dtype=bool)
    return lwe_encrypt(rng, np.where(m, int(encode_message(1, 8)), int(encode_message(-1, 8))), key.params.lwe_noise_stddev, key.key)


def decrypt(key, sample):
    """api.jl:264-266."""
    r = lwe_phase(sample, key.key) > 0
    return bool(r) if r.ndim == 0 else r


def engine_for(ck, device=None, devices=None):
    """The GPU engine holding `ck`'s keys (created on first use): the 3gen engine with one party."""
    if ck._engine is None:
        p = ck.params
        if not p.rlwe_is32 or p.rlwe_mask_size != 1:
            raise NotImplementedError("single-key sets are Torus32 with rlwe_mask_size = 1 (api.jl:76-113)")
        sp = SchemeParameters_3gen(p.lwe_size, p.lwe_noise_stddev, p.rlwe_polynomial_degree, 1, False, p.bs_decomp_length, p.bs_log2_base,
                                   p.bs_noise_stddev, p.ks_decomp_length, p.ks_log2_base, p.ks_noise_stddev, 1)
        # gadget digits of up to 8 bits ride the default kernels (keys << 32, table-driven first transform stage); wider ones
        # (tfhe_parameters_80: Bg = 2^10) need the library's Torus32 mode: unshifted keys, 16-bit digit fields
        t32 = p.bs_log2_base > 8
        eng = Engine(sp, device=device, devices=devices, flags=_cabi.FLAG_TORUS32 if t32 else 0)
        eng.load_keys([ck.bootstrap_key.engine_parts(0 if t32 else 32)], [ck.keyswitch_key.key])
        ck._engine = eng
    return ck._engine


_MU = int(encode_message(1, 8)) << 32          # test-vector message encode_message(1, 8), carried as v << 32


def _affine(ck, mu0, cx, cy, cz, x, y=None, z=None):
    eng = engine_for(ck)
    n = eng.params.lwe_size
    f = lambda s: (s.a.reshape(-1, 1, n), s.b.reshape(-1))
    oa, ob = eng.ctx.affine_bootstrap_batch(int(mu0), cx, cy, cz, _MU, f(x), f(y) if y is not None else None, f(z) if z is not None else None)
    return LweSample(x.params, oa.reshape(x.b.shape + (n,)), ob.reshape(x.b.shape), 0.0)


def bootstrap(bk_or_ck, ks, mu, x):
    """bootstrap.jl:97-100.  The engine holds both keys, so the first argument is the CloudKey (or its bootstrap key after a gate
    created the engine); `mu` is the Torus32 output message."""
    ck = bk_or_ck
    eng = engine_for(ck)
    n = eng.params.lwe_size
    oa, ob = eng.ctx.bootstrap_batch(int(np.int32(mu)) << 32, x.a.reshape(-1, 1, n), x.b.reshape(-1))
    return LweSample(x.params, oa.reshape(x.b.shape + (n,)), ob.reshape(x.b.shape), 0.0)


def bootstrap_wo_keyswitch(ck, mu, x):
    """bootstrap.jl:73-86: the extracted sample of dimension N."""
    eng = engine_for(ck)
    n, N = eng.params.lwe_size, eng.params.rlwe_polynomial_degree
    ext, _ = eng.ctx.blind_rotate_batch(int(np.int32(mu)) << 32, x.a.reshape(-1, 1, n), x.b.reshape(-1))
    return LweSample(LweParams(N), ext[:, :N].reshape(x.b.shape + (N,)), ext[:, N].reshape(x.b.shape), 0.0)


def keyswitch(ck, sample):
    """keyswitch.jl:45-80."""
    eng = engine_for(ck)
    n, N = eng.params.lwe_size, eng.params.rlwe_polynomial_degree
    ext = np.concatenate([sample.a.reshape(-1, N), sample.b.reshape(-1, 1)], axis=1)
    oa, ob = eng.ctx.keyswitch_batch(ext)
    return LweSample(LweParams(n), oa.reshape(sample.b.shape + (n,)), ob.reshape(sample.b.shape), 0.0)


# gates.jl:16-142 -- linear prologue constants: (message space 8 unless noted)
def gate_nand(ck, x, y):
    return _affine(ck, encode_message(1, 8), -1, -1, 0, x, y)


def gate_or(ck, x, y):
    return _affine(ck, encode_message(1, 8), 1, 1, 0, x, y)


def gate_and(ck, x, y):
    return _affine(ck, encode_message(-1, 8), 1, 1, 0, x, y)


def gate_xor(ck, x, y):
    return _affine(ck, encode_message(1, 4), 2, 2, 0, x, y)


def gate_xnor(ck, x, y):
    return _affine(ck, encode_message(-1, 4), -2, -2, 0, x, y)


def gate_not(ck, x):
    return -x                               # gates.jl:80-83: not bootstrapped


def gate_constant(ck, value):
    return lwe_noiseless_trivial(encode_message(1 if value else -1, 8), lwe_parameters(ck.params))


def gate_nor(ck, x, y):
    return _affine(ck, encode_message(-1, 8), -1, -1, 0, x, y)


def gate_andny(ck, x, y):
    return _affine(ck, encode_message(-1, 8), -1, 1, 0, x, y)


def gate_andyn(ck, x, y):
    return _affine(ck, encode_message(-1, 8), 1, -1, 0, x, y)


def gate_orny(ck, x, y):
    return _affine(ck, encode_message(1, 8), -1, 1, 0, x, y)


def gate_oryn(ck, x, y):
    return _affine(ck, encode_message(1, 8), 1, -1, 0, x, y)


def gate_mux(ck, x, y, z):
    """gates.jl:166-177."""
    mu = encode_message(1, 8)
    t1 = lwe_noiseless_trivial(encode_message(-1, 8), x.params, x.b.shape) + x + y
    t2 = lwe_noiseless_trivial(encode_message(-1, 8), x.params, x.b.shape) - x + z
    u = bootstrap_wo_keyswitch(ck, mu, LweSample(x.params, np.stack([t1.a, t2.a]), np.stack([t1.b, t2.b])))
    t3 = lwe_noiseless_trivial(mu, u.params, x.b.shape) + u[0] + u[1]
    return keyswitch(ck, t3)


GATES = {"NAND": (gate_nand, 2, lambda x, y: ~(x & y)), "OR": (gate_or, 2, lambda x, y: x | y), "AND": (gate_and, 2, lambda x, y: x & y),
         "XOR": (gate_xor, 2, lambda x, y: x ^ y), "XNOR": (gate_xnor, 2, lambda x, y: ~(x ^ y)), "NOT": (gate_not, 1, lambda x: ~x),
         "NOR": (gate_nor, 2, lambda x, y: ~(x | y)), "ANDNY": (gate_andny, 2, lambda x, y: ~x & y), "ANDYN": (gate_andyn, 2, lambda x, y: x & ~y),
         "ORNY": (gate_orny, 2, lambda x, y: ~x | y), "ORYN": (gate_oryn, 2, lambda x, y: x | ~y),
         "MUX": (gate_mux, 3, lambda x, y, z: np.where(x, y, z))}      # the reference's gate_tests table, test/runtests.jl:10-23
