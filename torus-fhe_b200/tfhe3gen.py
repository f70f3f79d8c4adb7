"""Host-side mirror of the reference's 3gen multi-key gate API (module TFHE,
3-gen-mk-tfhe/src/TFHE.jl:56-196) over the C ABI of libmktfhe_b200.so.

Same exported names, positional arguments and meaning as the Julia functions; the
bootstrapped gates forward to the sm_100a kernels (no CPU path).  The Julia shim
`julia/TFHE_B200.jl` is the `ccall` twin of this file; this Python twin exists
because the image has no Julia, so it is what the tests and benchmarks drive.

Batching: every `MKLweSample` may carry leading batch dimensions (`a` of shape
batch + (k, n), `b` of shape batch); a plain sample is batch = ().  The
reference's `a::Array{Int32,2}` of size (n, k), column-major, is the same memory
as our C-order (k, n).

Citations are relative to 3-gen-mk-tfhe/src/.
"""
from dataclasses import dataclass

import numpy as np

from . import _cabi
from .engine import Engine

Torus32 = np.int32
Torus64 = np.int64


# ---------------------------------------------------------------------------------
# numeric-functions.jl
# ---------------------------------------------------------------------------------
def _log2(x):
    lg = int(x).bit_length() - 1
    if (1 << lg) != x:
        raise ValueError(f"{x} is not a power of two")
    return lg


def encode_message(mu, space):
    """numeric-functions.jl:86-89: Torus32(mu) << (32 - log2(space)), wrapping."""
    return np.int32(np.uint32((int(mu) << (32 - _log2(space))) & 0xFFFFFFFF).astype(np.int32))


def encode_message64(mu, space):
    """numeric-functions.jl:92-95."""
    return np.int64(np.uint64((int(mu) << (64 - _log2(space))) & 0xFFFFFFFFFFFFFFFF).astype(np.int64))


def decode_message(phase, space):
    """numeric-functions.jl:70-73: (phase + 2^(32-log2(space)-1)) >> (32 - log2(space)), wrapping add, arithmetic shift."""
    lg = _log2(space)
    p = np.asarray(phase, dtype=np.int32)
    s = (p.view(np.uint32) + np.uint32(1 << (32 - lg - 1))).view(np.int32) if p.ndim else \
        np.uint32((int(p) + (1 << (32 - lg - 1))) & 0xFFFFFFFF).astype(np.int32)
    return s >> (32 - lg)


def decode_message64(phase, space):
    """numeric-functions.jl:75-78: the same on Torus64."""
    lg = _log2(space)
    p = np.asarray(phase, dtype=np.int64)
    return (p.view(np.uint64) + np.uint64(1 << (64 - lg - 1))).view(np.int64) >> (64 - lg)


def noise_calc(m_torus, d_torus):
    """numeric-functions.jl:117-131: distance on the torus (as a fraction of it) between a decrypted phase `d_torus` and the message
    `m_torus` it should carry, with the reference's wrap rules (note its asymmetric sign for negative messages, kept)."""
    m = np.asarray(m_torus, dtype=np.int32).astype(np.float64) / 2.0 ** 32
    d = np.asarray(d_torus, dtype=np.int32).astype(np.float64) / 2.0 ** 32
    pos = np.where((d < 0) & (d < m - 0.5), 1.0 + d - m, d - m)
    neg = np.where(d > m + 0.5, 1.0 - d + m, d - m)
    out = np.where(m > 0, pos, np.where(m < 0, neg, d))
    return float(out) if out.ndim == 0 else out


def dtot32(d):
    """numeric-functions.jl:101-103: trunc(Int32, d * 2^32)."""
    return np.trunc(np.asarray(d, dtype=np.float64) * 4294967296.0).astype(np.int64).astype(np.int32)


def dtot64(d):
    """numeric-functions.jl:105-107: trunc(Int64, d * 2^64) (|d| < 0.5)."""
    return np.trunc(np.asarray(d, dtype=np.float64) * 18446744073709551616.0).astype(np.int64)


_TERNARY_W = 0.113546097609674   # numeric-functions.jl:26-28


def rand_negative_binary64(rng, n):
    u = rng.random(n)
    return np.where(u < _TERNARY_W, -1, np.where(u < 1.0 - _TERNARY_W, 0, 1)).astype(np.int64)


def rand_uniform_torus32(rng, n):
    return rng.integers(-2 ** 31, 2 ** 31, size=n, dtype=np.int64).astype(np.int32)


def rand_uniform_torus64(rng, n):
    return rng.integers(-2 ** 63, 2 ** 63 - 1, size=n, dtype=np.int64, endpoint=True)


def rand_gaussian_torus32(rng, mean, sigma, n=None):
    return (np.int32(mean) + dtot32(rng.standard_normal(n) * sigma)).astype(np.int32)


def rand_gaussian_torus64(rng, mean, sigma, n=None):
    return np.int64(mean) + dtot64(rng.standard_normal(n) * sigma)


# ---------------------------------------------------------------------------------
# parameters (api.jl:50-67, 156-166; mk_api.jl:32-146)
# ---------------------------------------------------------------------------------
@dataclass(frozen=True)
class SchemeParameters_3gen:
    lwe_size: int
    lwe_noise_stddev: float
    rlwe_polynomial_degree: int
    rlwe_mask_size: int
    rlwe_is32: bool
    gsw_decomp_length: int
    gsw_log2_base: int
    gsw_noise_stddev: float
    ks_decomp_length: int
    ks_log2_base: int
    ks_noise_stddev: float
    max_parties: int


@dataclass(frozen=True)
class LweParams:
    size: int


@dataclass(frozen=True)
class RLweParams:
    polynomial_degree: int
    mask_size: int
    is32: bool


@dataclass(frozen=True)
class TGswParams:
    decomp_length: int
    log2_base: int
    is32: bool

    @property
    def gadget_values(self):   # tgsw.jl:26
        bit = 32 if self.is32 else 64
        return [1 << (bit - q * self.log2_base) for q in range(1, self.decomp_length + 1)]


@dataclass(frozen=True)
class KeyswitchParameters:
    decomp_length: int
    log2_base: int


def lwe_parameters(p):
    return LweParams(p.lwe_size)


def rlwe_parameters(p):
    return RLweParams(p.rlwe_polynomial_degree, p.rlwe_mask_size, p.rlwe_is32)


def tgsw_parameters(p):
    return TGswParams(p.gsw_decomp_length, p.gsw_log2_base, p.rlwe_is32)


def keyswitch_parameters(p):
    return KeyswitchParameters(p.ks_decomp_length, p.ks_log2_base)


mktfhe_parameters_2party_3gen = SchemeParameters_3gen(520, 2 ** -13.52, 1024, 1, False, 2, 7, 2 ** -30.70, 3, 3, 2 ** -13.52, 2)
mktfhe_parameters_3party_3gen = SchemeParameters_3gen(510, 2 ** -13.26, 1024, 1, False, 2, 7, 2 ** -30.70, 5, 2, 2 ** -13.26, 3)
mktfhe_parameters_4party_3gen = SchemeParameters_3gen(510, 2 ** -13.26, 1024, 1, False, 3, 6, 2 ** -30.70, 5, 2, 2 ** -13.26, 4)
mktfhe_parameters_5party_3gen = SchemeParameters_3gen(520, 2 ** -13.52, 1024, 1, False, 3, 6, 2 ** -30.70, 5, 2, 2 ** -13.52, 5)
mktfhe_parameters_8party_3gen = SchemeParameters_3gen(540, 2 ** -14.04, 1024, 1, False, 4, 4, 2 ** -30.70, 5, 2, 2 ** -14.04, 8)
# 16 parties and up use N = 2048 and an 18..26-bit gadget base with l = 1 or 2 (mk_api.jl:214-310): served by the four-prime N = 2048
# kernels (csrc/kernels2k.cuh).  The 512-party set (N = 4096) is defined for API completeness and rejected by mktfhe_create
# (MKTFHE_EINVAL).
mktfhe_parameters_16party_3gen = SchemeParameters_3gen(590, 2 ** -15.34, 2048, 1, False, 1, 26, 2 ** -62.00, 4, 3, 2 ** -15.34, 16)
mktfhe_parameters_32party_3gen = SchemeParameters_3gen(620, 2 ** -16.12, 2048, 1, False, 1, 26, 2 ** -62.00, 4, 3, 2 ** -16.12, 32)      # :246-252
mktfhe_parameters_64party_3gen = SchemeParameters_3gen(650, 2 ** -16.90, 2048, 1, False, 1, 25, 2 ** -62.00, 4, 3, 2 ** -16.90, 64)      # :268-274
mktfhe_parameters_128party_3gen = SchemeParameters_3gen(670, 2 ** -17.42, 2048, 1, False, 1, 24, 2 ** -62.00, 5, 3, 2 ** -17.42, 128)   # :292-298
mktfhe_parameters_256party_3gen = SchemeParameters_3gen(740, 2 ** -19.24, 2048, 1, False, 2, 18, 2 ** -62.00, 8, 2, 2 ** -19.24, 256)   # :304-310
mktfhe_parameters_512party_3gen = SchemeParameters_3gen(730, 2 ** -18.98, 4096, 1, False, 1, 27, 2 ** -62.00, 5, 3, 2 ** -18.98, 512)   # :316-322


# ---------------------------------------------------------------------------------
# MKLweSample (mk_internals.jl:23-51, 94-96)
# ---------------------------------------------------------------------------------
def _w32(x):
    return np.asarray(x).astype(np.int64).astype(np.int32) if not isinstance(x, np.ndarray) or x.dtype != np.int32 else x


class MKLweSample:
    __slots__ = ("params", "a", "b", "current_variance")
    __array_ufunc__ = None      # numpy scalars (Torus32(2) * sample) defer to __rmul__

    def __init__(self, params, a, b, current_variance=0.0):
        self.params = params
        self.a = np.asarray(a, dtype=np.int32)
        self.b = np.asarray(b, dtype=np.int32)
        self.current_variance = current_variance

    @property
    def batch_shape(self):
        return self.b.shape

    def __len__(self):
        return self.b.shape[0]

    def __getitem__(self, idx):
        return MKLweSample(self.params, self.a[idx], self.b[idx], self.current_variance)

    def __sub__(self, y):
        with np.errstate(over="ignore"):
            return MKLweSample(self.params, self.a - y.a, self.b - y.b, self.current_variance + y.current_variance)

    def __add__(self, y):
        with np.errstate(over="ignore"):
            return MKLweSample(self.params, self.a + y.a, self.b + y.b, self.current_variance + y.current_variance)

    def __neg__(self):
        with np.errstate(over="ignore"):
            return MKLweSample(self.params, -self.a, -self.b, self.current_variance)

    def __rmul__(self, x):   # Torus32 * sample, mk_internals.jl:50-51
        x = np.int32(x)
        with np.errstate(over="ignore"):
            return MKLweSample(self.params, x * self.a, x * self.b, float(x) * float(x) * self.current_variance)

    @staticmethod
    def stack(samples):
        return MKLweSample(samples[0].params, np.stack([s.a for s in samples]), np.stack([s.b for s in samples]),
                           max(s.current_variance for s in samples))


class MKLweSampleGPU:
    """mk_internals.jl:56-82: the reference's CuArray twin of MKLweSample (it only ever carries the linear ops there; the
    bootstrap call in gpu_circuits.jl:27-36 is commented out).  Here `a` / `b` are torch int32 tensors resident on the GPU of the
    engine; every gate below accepts it and runs through the device-pointer C-ABI entries, so ciphertexts stay in HBM between
    gates and dependency levels."""
    __slots__ = ("params", "a", "b", "current_variance")
    __array_ufunc__ = None

    def __init__(self, params, a, b, current_variance=0.0):
        self.params, self.a, self.b, self.current_variance = params, a, b, current_variance

    @staticmethod
    def from_host(x, device=None):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        return MKLweSampleGPU(x.params, torch.from_numpy(np.ascontiguousarray(x.a)).to(dev), torch.from_numpy(np.ascontiguousarray(x.b)).to(dev),
                              x.current_variance)

    def cpu(self):
        return MKLweSample(self.params, self.a.cpu().numpy(), self.b.cpu().numpy(), self.current_variance)

    @property
    def batch_shape(self):
        return tuple(self.b.shape)

    def __getitem__(self, idx):
        return MKLweSampleGPU(self.params, self.a[idx], self.b[idx], self.current_variance)

    def __sub__(self, y):
        return MKLweSampleGPU(self.params, self.a - y.a, self.b - y.b, self.current_variance + y.current_variance)

    def __add__(self, y):
        return MKLweSampleGPU(self.params, self.a + y.a, self.b + y.b, self.current_variance + y.current_variance)

    def __neg__(self):
        return MKLweSampleGPU(self.params, -self.a, -self.b, self.current_variance)

    def __rmul__(self, x):
        x = int(x)
        return MKLweSampleGPU(self.params, self.a * x, self.b * x, float(x) * float(x) * self.current_variance)


def mk_lwe_noiseless_trivial_gpu(mu, params, parties, batch_shape=(), device=None):
    """gpu_circuits.jl:23-25."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    return MKLweSampleGPU(params, torch.zeros(tuple(batch_shape) + (parties, params.size), dtype=torch.int32, device=dev),
                          torch.full(tuple(batch_shape), int(mu), dtype=torch.int32, device=dev), 0.0)


def mk_lwe_noiseless_trivial(mu, params, parties, batch_shape=()):
    return MKLweSample(params, np.zeros(tuple(batch_shape) + (parties, params.size), np.int32),
                       np.full(batch_shape, np.int32(mu), np.int32), 0.0)


# ---------------------------------------------------------------------------------
# keys
# ---------------------------------------------------------------------------------
class LweKey:   # lwe.jl:7-14
    def __init__(self, rng, params):
        self.params = params
        self.key = rng.integers(0, 2, size=params.size, dtype=np.int64).astype(np.int32) if rng is not None else None


class SecretKey_3gen:   # api.jl:196-204
    def __init__(self, rng, params, key=None):
        self.params = params
        self.key = LweKey(rng, lwe_parameters(params))
        if key is not None:
            self.key.key = np.asarray(key, dtype=np.int32)


class RLweKey:   # rlwe.jl:13-31 (negative_random = true: sparse ternary)
    def __init__(self, rng, params, negative_random=True, key=None):
        self.params = params
        if key is not None:
            self.key = np.asarray(key, dtype=np.int64)
        elif negative_random:
            self.key = rand_negative_binary64(rng, params.polynomial_degree)
        else:
            self.key = rng.integers(0, 2, size=params.polynomial_degree, dtype=np.int64)


_mul_ctx_cache = {}


def _mul_context(N, device=0):
    """A key-less context used only for exact negacyclic products during key generation."""
    key = (N, device)
    if key not in _mul_ctx_cache:
        _mul_ctx_cache[key] = _cabi.Context(1, N, 1, 1, 1, 1, 1, device=device)
    return _mul_ctx_cache[key]


def negacyclic_mul(small, big):
    """Exact product mod (X^N + 1, 2^64) on the GPU.  `small` must satisfy |coeff| <= 2^8 at N = 1024 (2^25 at N = 2048): beyond that
    N * |small| * 2^63 leaves the exact range of the CRT lift, and the library rejects the call (MKTFHE_EINVAL)."""
    small, big = np.asarray(small, np.int64), np.asarray(big, np.int64)
    N = small.shape[-1]
    bshape = np.broadcast_shapes(small.shape, big.shape)
    s = np.ascontiguousarray(np.broadcast_to(small, bshape)).reshape(-1, N)
    b = np.ascontiguousarray(np.broadcast_to(big, bshape)).reshape(-1, N)
    return _mul_context(N).negacyclic_mul_batch(s, b).reshape(bshape)


class CRP_3gen:   # mk_internals.jl:177-196
    def __init__(self, rng, tgsw_params, rlwe_params, a_same=False):
        l, N = tgsw_params.decomp_length, rlwe_params.polynomial_degree
        self.tgsw_params, self.rlwe_params = tgsw_params, rlwe_params
        if a_same:
            self.a = np.repeat(rand_uniform_torus64(rng, N)[None], l, axis=0)
        else:
            self.a = np.stack([rand_uniform_torus64(rng, N) for _ in range(l)])


def GenCRP_3gen(rng, tgsw_params, rlwe_params, a_same=False):
    return CRP_3gen(rng, tgsw_params, rlwe_params, a_same)


class PublicKey:   # mk_internals.jl:266-298: b[q] = z * a[q] + e
    def __init__(self, rng, rlwe_key, alpha, crp, tgsw_params, wo_FFT=0):
        l, N = tgsw_params.decomp_length, rlwe_key.params.polynomial_degree
        self.tgsw_params, self.rlwe_params = tgsw_params, rlwe_key.params
        prod = negacyclic_mul(rlwe_key.key[None, :], crp.a)
        with np.errstate(over="ignore"):
            self.b = prod + np.stack([rand_gaussian_torus64(rng, 0, alpha, N) for _ in range(l)])


class CommonPubKey_3gen:   # mk_internals.jl:325-345
    def __init__(self, pubkeys, params, parties):
        self.params, self.parties = params, parties
        b = pubkeys[0].b.copy()
        with np.errstate(over="ignore"):
            for i in range(1, parties):
                b = b + pubkeys[i].b
        self.b = b


class BootstrapKeyPart_3gen:
    """3gen_mk_internals.jl:10-43: n TGSW encryptions (tgsw_encrypt_3gen, tgsw_3gen.jl:41-95) of the
    LWE key bits under the common public key.  `gsw_key` is int64 [n][4][l][N] (part_1..part_4)."""

    def __init__(self, rng, lwe_key, alpha, crp_a, common_pubkey, tgsw_params, rlwe_params, wo_FFT=0, gsw_key=None):
        self.tgsw_params, self.rlwe_params = tgsw_params, rlwe_params
        if gsw_key is not None:
            self.gsw_key = np.asarray(gsw_key, dtype=np.int64)
            self.key_size = self.gsw_key.shape[0]
            return
        n, l, N = lwe_key.params.size, tgsw_params.decomp_length, rlwe_params.polynomial_degree
        self.key_size = n
        r1 = rand_negative_binary64(rng, n * l * N).reshape(n, l, N)
        r2 = rand_negative_binary64(rng, n * l * N).reshape(n, l, N)
        err = rand_gaussian_torus64(rng, 0, alpha, n * 4 * l * N).reshape(n, 4, l, N)
        B, A = common_pubkey.b[None], crp_a.a[None]
        mg = lwe_key.key.astype(np.int64)[:, None] * np.array(tgsw_params.gadget_values, dtype=np.uint64).view(np.int64)[None, :]
        with np.errstate(over="ignore"):
            gsw = err
            gsw[:, 0] += negacyclic_mul(r1, B)
            gsw[:, 1] += negacyclic_mul(r2, B)
            gsw[:, 2] += negacyclic_mul(r2, A)
            gsw[:, 3] += negacyclic_mul(r1, A)
            gsw[:, 0, :, 0] += mg   # Polynomial .+ scalar: constant coefficient
            gsw[:, 2, :, 0] += mg
        self.gsw_key = gsw

    @classmethod
    def from_array(cls, gsw_key, tgsw_params, rlwe_params):
        return cls(None, None, 0.0, None, None, tgsw_params, rlwe_params, gsw_key=gsw_key)


class TransformedBootstrapKeyPart_3gen:
    """3gen_mk_internals.jl:45-55.  The reference keeps Complex{Float64} FFTs here; this build keeps
    the integer polynomials and transforms them on the GPU (exact NTT) when the engine loads them."""

    def __init__(self, bk):
        self.tgsw_params, self.rlwe_params = bk.tgsw_params, bk.rlwe_params
        self.gsw_key = bk.gsw_key
        self.key_size = bk.key_size
        self._engine = None
        self._owner = None

    @property
    def tgsw_samples(self):
        """The reference's view of this key, Array{TransformedTGswSample_3gen,1} (3gen_mk_internals.jl:47): element j is the TGSW
        encryption of LWE key bit j.  (`gsw_key` itself stays the flat int64 [n][4][l][N] array the engine loads.)"""
        return [TransformedTGswSample_3gen(self, j) for j in range(self.key_size)]


class KeyswitchKey:
    """keyswitch.jl:7-42.  `key` is int32 [N][t][base-1][n+1] (row = LweSample a[0..n-1], b), i.e. the
    reference's Array{LweSample,3} of dims (base-1, t, N) in memory order."""

    def __init__(self, rng, alpha, params, out_key, in_key, key=None):
        self.params = params
        if key is not None:
            self.key = np.asarray(key, dtype=np.int32)
            return
        s = out_key.key.astype(np.int64)
        z = in_key.key
        n, N, t, bb = s.size, z.size, params.decomp_length, params.log2_base
        B1 = (1 << bb) - 1
        noise = rng.standard_normal((N, t, B1)) * alpha
        noise -= noise.mean()                                              # :28-29
        a = rand_uniform_torus32(rng, N * t * B1 * n).reshape(N, t, B1, n)
        h = np.arange(1, B1 + 1, dtype=np.int64)[None, None, :]
        sh = (32 - np.arange(1, t + 1) * bb)[None, :, None]
        msg = (z[:, None, None] * h) << sh                                 # :35
        dot = (a.astype(np.int64) * s).sum(-1)
        b = (msg + dtot32(noise).astype(np.int64) + dot).astype(np.int32)  # lwe.jl:47-53
        self.key = np.concatenate([a, b[..., None]], axis=-1)

    @classmethod
    def from_array(cls, key, params):
        return cls(None, 0.0, params, None, None, key=key)

    @classmethod
    def on_device(cls, rng, alpha, params, out_key, in_key):
        """The same key, generated ON the GPU when an engine loads it (mktfhe_generate_ksk: Philox4x32-10 keyed by a seed drawn
        from `rng`): only the two secret vectors cross PCIe instead of 45 MB (2 parties) .. 140 MB (16 parties) per party of rows
        built with numpy.  `key` stays None: the rows exist on the device only."""
        self = cls.__new__(cls)
        self.params, self.key = params, None
        self._device_gen = (np.asarray(out_key.key, np.int32).copy(), np.asarray(in_key.key, np.int64).copy(), float(alpha),
                            int(rng.integers(0, 2 ** 63)))
        return self

    def engine_part(self):
        """What Engine.load_keys takes for this party: the host rows, or the recipe for generating them on the device."""
        gen = getattr(self, "_device_gen", None)
        return ("generate",) + gen if gen is not None else self.key


# ---------------------------------------------------------------------------------
# encrypt / decrypt (mk_api.jl:519-536, 576-633; mk_internals.jl:85-91)
# ---------------------------------------------------------------------------------
def mk_encrypt_3gen(rng, secret_keys, message):
    """`message` may be a bool or an array of bools (batch)."""
    params = secret_keys[0].params
    msg = np.asarray(message, dtype=bool)
    n, k = params.lwe_size, len(secret_keys)
    a = rand_uniform_torus32(rng, msg.size * k * n).reshape(msg.shape + (k, n))
    s = np.stack([sk.key.key for sk in secret_keys]).astype(np.int64)
    mu = np.where(msg, int(encode_message(1, 8)), int(encode_message(-1, 8)))
    e = dtot32(rng.standard_normal(msg.shape) * params.lwe_noise_stddev).astype(np.int64)
    b = (mu + e + (a.astype(np.int64) * s).sum((-1, -2))).astype(np.int32)
    return MKLweSample(lwe_parameters(params), a, b, params.lwe_noise_stddev ** 2)


def mk_lwe_phase(sample, lwe_keys):
    s = np.stack([k.key for k in lwe_keys]).astype(np.int64)
    return (sample.b.astype(np.int64) - (sample.a.astype(np.int64) * s).sum((-1, -2))).astype(np.int32)


def mk_decrypt_3gen(secret_keys, sample):
    r = mk_lwe_phase(sample, [sk.key for sk in secret_keys]) > 0
    return bool(r) if r.ndim == 0 else r


def mk_int_encrypt_3gen(rng, secret_keys, message, WIDTH):
    """LSB-first bit vector; `message` may be an int or an integer array (batch of instances)."""
    m = np.asarray(message, dtype=np.int64)
    return [mk_encrypt_3gen(rng, secret_keys, ((m >> i) & 1) == 1) for i in range(WIDTH)]


def mk_int_decrypt_3gen(secret_keys, sample, WIDTH):
    msb = np.asarray(mk_decrypt_3gen(secret_keys, sample[WIDTH - 1]))
    result = np.zeros(msb.shape, np.int64)
    for i in range(WIDTH - 1):
        bit = np.asarray(mk_decrypt_3gen(secret_keys, sample[i]))
        result += (bit ^ msb).astype(np.int64) << i
    result = np.where(msb, -(result + 1), result)
    return int(result) if result.ndim == 0 else result


# ---------------------------------------------------------------------------------
# engine lookup: the reference passes (bk, ks) to every gate; the GPU context that
# holds their transformed copies is created on first use and cached on bk[0].
# ---------------------------------------------------------------------------------
def _scheme_params_of(bk, ks):
    tg, ksp = bk[0].tgsw_params, ks[0].params
    n = bk[0].key_size
    return SchemeParameters_3gen(n, 0.0, bk[0].rlwe_params.polynomial_degree, bk[0].rlwe_params.mask_size, bk[0].rlwe_params.is32,
                                 tg.decomp_length, tg.log2_base, 0.0, ksp.decomp_length, ksp.log2_base, 0.0, len(bk))


def default_devices():
    """GPUs a new engine spans when the caller does not say: MKTFHE_B200_DEVICES = "all" | "0,1,2,3" | unset (one GPU: LOCAL_RANK or 0)."""
    import os
    v = os.environ.get("MKTFHE_B200_DEVICES", "")
    if not v:
        return None
    return "all" if v == "all" else [int(x) for x in v.split(",")]


def _same_keys(tag, bk, ks):
    """The cache is keyed on the key OBJECTS (held by the tag, compared with `is`): ids alone can alias after garbage collection."""
    tb, tk = tag
    return len(tb) == len(bk) and len(tk) == len(ks) and all(a is b for a, b in zip(tb, bk)) and all(a is b for a, b in zip(tk, ks))


def engine_for(bk, ks, device=None, devices=None):
    """The engine holding (bk, ks), created on first use.  `devices` (a list of GPU ordinals or "all") makes it one context
    spanning those GPUs (mktfhe_create_multi): keys broadcast inside the library, batched gate calls sharded over the GPUs."""
    cached = getattr(bk[0], "_engine", None)
    if cached is not None and _same_keys(cached[0], bk, ks):
        return cached[1]
    if bk[0].rlwe_params.is32:
        raise NotImplementedError("rlwe_is32 = true parameter sets are not part of the 3gen path (every 3gen set is Torus64)")
    eng = Engine(_scheme_params_of(bk, ks), device=device, devices=devices if devices is not None else default_devices())
    eng.load_keys([b.gsw_key for b in bk], [k.engine_part() for k in ks])
    return attach_engine(bk, ks, eng)


def release_engine(bk, ks):
    """Frees the device key replicas held for (bk, ks) (the engine is otherwise kept alive by the key objects)."""
    cached = getattr(bk[0], "_engine", None)
    if cached is not None:
        cached[1].close()
        bk[0]._engine = ks[0]._engine = None


def _tag_parts(bk, eng):
    """Each key part remembers the engine that holds it and its party index: the per-element entry points of the reference
    (tgsw_extern_mul_3gen, mk_mux_rotate_3gen, mk_ith_blind_rotate_3gen) receive one part or one of its TGSW samples, not the array."""
    for i, b in enumerate(bk):
        try:
            b._owner = (eng, i)
        except AttributeError:      # RemoteKeys and other stand-ins without the slot
            pass


class RemoteKeys:
    """Stand-in for (bk, ks) on a rank whose engine received the keys by broadcast (Engine.broadcast_keys) instead of from host
    arrays: the gate API finds the engine through it."""

    def __init__(self, eng):
        self._engine = None
        self.tgsw_params, self.rlwe_params = tgsw_parameters(eng.params), rlwe_parameters(eng.params)

    @staticmethod
    def pair(eng):
        r = RemoteKeys(eng)
        bk, ks = [r] * eng.params.max_parties, [r] * eng.params.max_parties
        r._engine = ((tuple(bk), tuple(ks)), eng)
        return bk, ks


def attach_engine(bk, ks, eng):
    """Make `eng` (already holding these keys) the engine the gate API uses for (bk, ks)."""
    bk[0]._engine = ks[0]._engine = ((tuple(bk), tuple(ks)), eng)
    _tag_parts(bk, eng)
    return eng


def _flat(x, k, n):
    return x.a.reshape(-1, k, n), x.b.reshape(-1)


def _gate_gpu(eng, gate, x, y, z):
    """Device-resident operands: mktfhe_gate_batch_dev on torch's current stream; nothing crosses PCIe."""
    import torch
    k, n = eng.params.max_parties, eng.params.lwe_size
    ops = [(t.a.reshape(-1, k, n).contiguous(), t.b.reshape(-1).contiguous()) for t in (x, y) + ((z,) if z is not None else ())]
    G = ops[0][1].numel()
    dev = ops[0][0].device
    if any(t.device != dev for op in ops for t in op):
        raise ValueError("gate operands live on different devices")
    ctx = eng.ctx_on(dev.index)                 # raises when the engine holds no key replica on the operands' GPU
    oa = torch.empty((G, k, n), dtype=torch.int32, device=ops[0][0].device)
    ob = torch.empty(G, dtype=torch.int32, device=ops[0][0].device)
    za, zb = (ops[2][0].data_ptr(), ops[2][1].data_ptr()) if z is not None else (0, 0)
    stream = torch.cuda.current_stream(ops[0][0].device).cuda_stream
    if not stream:                      # legacy default stream: the library then works on the context's own (non-blocking) stream, which
        torch.cuda.synchronize(ops[0][0].device)     # does not wait for the default stream -- the operands must be complete first
    ctx.gate_batch_dev(gate, G, ops[0][0].data_ptr(), ops[0][1].data_ptr(), ops[1][0].data_ptr(), ops[1][1].data_ptr(), za, zb,
                       oa.data_ptr(), ob.data_ptr(), stream=stream)
    if not stream:                      # legacy default stream: the context's own stream did the work
        torch.cuda.synchronize(ops[0][0].device)
    shape = tuple(x.b.shape)
    return MKLweSampleGPU(x.params, oa.reshape(shape + (k, n)), ob.reshape(shape), 0.0)


def _gate(bk, ks, gate, x, y, z=None):
    eng = engine_for(bk, ks)
    if isinstance(x, MKLweSampleGPU):
        return _gate_gpu(eng, gate, x, y, z)
    k, n = eng.params.max_parties, eng.params.lwe_size
    shape = x.b.shape
    oa, ob = eng.ctx.gate_batch(gate, _flat(x, k, n), _flat(y, k, n), _flat(z, k, n) if z is not None else None)
    return MKLweSample(x.params, oa.reshape(shape + (k, n)), ob.reshape(shape), 0.0)   # variance 0.0: mk_internals.jl:738-743


# ---------------------------------------------------------------------------------
# the hot path (3gen_mk_internals.jl:99-116, mk_internals.jl:730-744)
# ---------------------------------------------------------------------------------
def mk_bootstrap_3gen(bk, ks, mu, x):
    eng = engine_for(bk, ks)
    k, n = eng.params.max_parties, eng.params.lwe_size
    oa, ob = eng.ctx.bootstrap_batch(int(mu), *_flat(x, k, n))
    return MKLweSample(x.params, oa.reshape(x.b.shape + (k, n)), ob.reshape(x.b.shape), 0.0)


class LweSample:
    """lwe.jl:23-33: a single-key LWE sample; here the sample extracted from the accumulator, of dimension N under the sum of the
    parties' extracted RLWE keys (rlwe.jl:70-74).  `a` is int32 [.., size], `b` int32 [..] (leading dimensions = batch)."""
    __slots__ = ("params", "a", "b", "current_variance")

    def __init__(self, params, a, b, current_variance=0.0):
        self.params = params
        self.a = np.asarray(a, dtype=np.int32)
        self.b = np.asarray(b, dtype=np.int32)
        self.current_variance = current_variance


def _cached_engine(keys, what):
    """Engine of a call that receives only one of (bk, ks), as mk_bootstrap_wo_keyswitch_3gen and mk_keyswitch_3gen do in the
    reference: the GPU context always holds both keys, so it must have been created from the pair before."""
    cached = getattr(keys[0], "_engine", None)
    if cached is None:
        raise RuntimeError(f"no GPU engine holds these {what} yet: call engine_for(bk, ks) (or any gate) with the key pair first")
    return cached[1]


def mk_blind_rotate_and_extract_3gen(v, bk, barb, bara):
    """3gen_mk_internals.jl:88-95: acc = X^{-barb} v, blind rotation by `bara` over every party's key, sample extraction.
    `v` must be the constant test vector [mu] * N (the only one the reference's callers build, :105-108); `barb` int32 [..] and
    `bara` int32 [.., k, n] are rotations in [-N, N) as decode_message(x, 2N) returns them.  Returns the extracted LweSample."""
    eng = _cached_engine(bk, "bootstrapping keys")
    k, n, N = eng.params.max_parties, eng.params.lwe_size, eng.params.rlwe_polynomial_degree
    v = np.asarray(v, dtype=np.int64).reshape(-1)
    if v.size != N or np.any(v != v[0]):
        raise ValueError("the engine rotates the constant test vector repeat([mu], N) only (3gen_mk_internals.jl:105-108)")
    barb, bara = np.asarray(barb, np.int64), np.asarray(bara, np.int64)
    if bara.shape != barb.shape + (k, n) or any(np.any((r < -N) | (r >= N)) for r in (barb, bara)):
        raise ValueError(f"need rotations in [-N, N) of shapes [..] and [.., {k}, {n}]")
    # the kernel mod-switches its input itself: hand it the torus elements bar * 2^32 / 2N, which decode_message maps back to bar
    sh = 32 - _log2(2 * N)
    ext, _ = eng.ctx.blind_rotate_batch(int(v[0]), (bara << sh).astype(np.int32).reshape(-1, k, n), (barb << sh).astype(np.int32).reshape(-1))
    return LweSample(LweParams(N), ext[:, :N].reshape(barb.shape + (N,)), ext[:, N].reshape(barb.shape), 0.0)


def mk_bootstrap_wo_keyswitch_3gen(bk, mu, x):
    """3gen_mk_internals.jl:99-109: mod-switch, blind rotation of the test vector [mu] * N, extraction; no key switch."""
    eng = _cached_engine(bk, "bootstrapping keys")
    k, n, N = eng.params.max_parties, eng.params.lwe_size, eng.params.rlwe_polynomial_degree
    ext, _ = eng.ctx.blind_rotate_batch(int(mu), *_flat(x, k, n))
    return LweSample(LweParams(N), ext[:, :N].reshape(x.b.shape + (N,)), ext[:, N].reshape(x.b.shape), 0.0)


def mk_keyswitch_3gen(ks, sample):
    """mk_internals.jl:730-744: the extracted LweSample (dimension N) to an MKLweSample under the parties' LWE keys."""
    eng = _cached_engine(ks, "key-switching keys")
    k, n, N = eng.params.max_parties, eng.params.lwe_size, eng.params.rlwe_polynomial_degree
    if sample.a.shape != sample.b.shape + (N,):
        raise ValueError(f"mk_keyswitch_3gen expects an LweSample of dimension N = {N}")
    ext = np.concatenate([sample.a.reshape(-1, N), sample.b.reshape(-1, 1)], axis=1)
    oa, ob = eng.ctx.keyswitch_batch(ext)
    return MKLweSample(LweParams(n), oa.reshape(sample.b.shape + (k, n)), ob.reshape(sample.b.shape), 0.0)


# ---------------------------------------------------------------------------------
# the path's internal entry points, same names and arguments as the reference; each forwards to
# the parity hook of the C ABI that runs that stage on the GPU
# ---------------------------------------------------------------------------------
class RLweSample:
    """rlwe.jl:43-50 with mask_size = 1: `a` is int64 [.., 2, N]; a[.., 0, :] is the mask (the reference's accum.a[1]) and
    a[.., 1, :] the body (accum.a[2]).  Leading dimensions are a batch."""
    __slots__ = ("params", "a", "current_variance")

    def __init__(self, params, a, current_variance=0.0):
        self.params = params
        self.a = np.asarray(a, dtype=np.int64)
        self.current_variance = current_variance

    def __add__(self, y):      # rlwe.jl:121-122
        with np.errstate(over="ignore"):
            return RLweSample(self.params, self.a + y.a, self.current_variance + y.current_variance)

    def __sub__(self, y):      # rlwe.jl:125-126
        with np.errstate(over="ignore"):
            return RLweSample(self.params, self.a - y.a, self.current_variance + y.current_variance)


def mul_by_monomial(x, s):
    """X^s * x mod X^N + 1 for any integer s (DarkIntegers.mul_by_monomial on a polynomial, rlwe.jl:130-131 on a sample):
    coefficient i moves to i + s, changing sign every time it wraps past N."""
    if isinstance(x, RLweSample):
        return RLweSample(x.params, mul_by_monomial(x.a, s), x.current_variance)
    x = np.asarray(x, dtype=np.int64)
    N = x.shape[-1]
    s = int(s) % (2 * N)
    with np.errstate(over="ignore"):
        if s >= N:
            x, s = -x, s - N
        return np.concatenate([-x[..., N - s:], x[..., :N - s]], axis=-1) if s else x.copy()


def rlwe_noiseless_trivial(mu, params):
    """rlwe.jl:113-119: mask 0, body = the polynomial mu."""
    mu = np.asarray(mu, dtype=np.int64)
    return RLweSample(params, np.stack([np.zeros_like(mu), mu], axis=-2), 0.0)


def t64tot32(d):
    """numeric-functions.jl:109-111: trunc(Int32, Float64(d) / 2^32) -- toward zero, through Float64."""
    q = np.trunc(np.asarray(d, dtype=np.int64).astype(np.float64) / 2.0 ** 32)
    if np.any(q >= 2.0 ** 31):
        raise OverflowError("InexactError: trunc(Int32, 2.147483648e9)")      # what the reference throws (probability ~2^-54)
    return q.astype(np.int32)


def rlwe_extract_sample_64(x):
    """rlwe.jl:70-74 with polynomials.jl:69-72: a'_0 = mask_0, a'_j = -mask_{N-j}, b' = body_0, each through t64tot32."""
    mask, body = x.a[..., 0, :], x.a[..., 1, :]
    N = mask.shape[-1]
    with np.errstate(over="ignore"):
        rev = np.concatenate([mask[..., :1], -mask[..., :0:-1]], axis=-1)
    return LweSample(LweParams(N), t64tot32(rev), t64tot32(body[..., 0]), 0.0)


class TGswSample_3gen:
    """tgsw_3gen.jl:3-20: part_1..part_4, each l integer polynomials (int64 [l][N])."""

    def __init__(self, tgsw_params, rlwe_params, parts):
        self.tgsw_params, self.rlwe_params = tgsw_params, rlwe_params
        self.parts = np.asarray(parts, dtype=np.int64)       # [4][l][N]

    part_1 = property(lambda self: self.parts[0])
    part_2 = property(lambda self: self.parts[1])
    part_3 = property(lambda self: self.parts[2])
    part_4 = property(lambda self: self.parts[3])


def tgsw_encrypt_3gen(rng, message, alpha, common_pubkey, crp_a, negative_random=True, wo_FFT=0):
    """tgsw_3gen.jl:41-95, same positional arguments: part_1 = r1 B + m g + e, part_4 = r1 a + e (the row pair of a body digit),
    part_2 = r2 B + e, part_3 = r2 a + m g + e (of a mask digit); r1, r2 sparse ternary (uniform bits when negative_random is
    false), `+ m g` on the constant coefficient.  As in the reference the noise is params.gsw_noise_stddev of the common public key's
    parameter set (`alpha` is accepted and not read, :64-67), and both values of wo_FFT mean the same here: the products are exact,
    on the GPU (mktfhe_negacyclic_mul_batch)."""
    params = common_pubkey.params
    tgsw_params, rlwe_params = tgsw_parameters(params), rlwe_parameters(params)
    l, N = tgsw_params.decomp_length, rlwe_params.polynomial_degree
    draw = rand_negative_binary64 if negative_random else (lambda r, n: r.integers(0, 2, size=n, dtype=np.int64))
    r1 = draw(rng, l * N).reshape(l, N)
    r2 = draw(rng, l * N).reshape(l, N)
    parts = rand_gaussian_torus64(rng, 0, params.gsw_noise_stddev, 4 * l * N).reshape(4, l, N)
    mg = np.int64(message) * np.array(tgsw_params.gadget_values, dtype=np.uint64).view(np.int64)
    with np.errstate(over="ignore"):
        parts[0] += negacyclic_mul(r1, common_pubkey.b)
        parts[1] += negacyclic_mul(r2, common_pubkey.b)
        parts[2] += negacyclic_mul(r2, crp_a.a)
        parts[3] += negacyclic_mul(r1, crp_a.a)
        parts[0, :, 0] += mg
        parts[2, :, 0] += mg
    return TGswSample_3gen(tgsw_params, rlwe_params, parts)


class TransformedTGswSample_3gen:
    """tgsw_3gen.jl:23-39: element j of a party's bootstrapping key.  The reference holds its FFTs; here it is a handle on the
    element already transformed (exact NTT) and resident on the GPU of the engine that loaded the key part."""

    def __init__(self, part, j):
        if not 0 <= j < part.key_size:
            raise IndexError(f"key element {j} outside 0..{part.key_size - 1}")
        self.part, self.j = part, int(j)
        self.tgsw_params, self.rlwe_params = part.tgsw_params, part.rlwe_params

    part_1 = property(lambda self: self.part.gsw_key[self.j, 0])
    part_2 = property(lambda self: self.part.gsw_key[self.j, 1])
    part_3 = property(lambda self: self.part.gsw_key[self.j, 2])
    part_4 = property(lambda self: self.part.gsw_key[self.j, 3])

    def _locate(self):
        owner = getattr(self.part, "_owner", None)
        if owner is None:
            raise RuntimeError("no GPU engine holds this bootstrapping key part yet: call engine_for(bk, ks) (or any gate) with the key pair first")
        eng, party = owner
        return eng, party * eng.params.lwe_size + self.j


def tgsw_extern_mul_3gen(accum, sample):
    """tgsw_3gen.jl:102-113 on the GPU (mktfhe_extprod_batch): body' = sum_q dec(body)_q part_1[q] + dec(mask)_q part_2[q],
    mask' = sum_q dec(body)_q part_4[q] + dec(mask)_q part_3[q], exact mod 2^64.  `accum` may carry a batch."""
    eng, elem = sample._locate()
    N = eng.params.rlwe_polynomial_degree
    if accum.a.shape[-2:] != (2, N):
        raise ValueError(f"accumulator must be [.., 2, {N}]")
    acc = accum.a.reshape(-1, 2, N)
    out = eng.ctx.extprod_batch(np.full(acc.shape[0], elem, np.int32), acc)
    return RLweSample(accum.params, out.reshape(accum.a.shape), 0.0)


def mk_mux_rotate_3gen(accum, bki, barai):
    """3gen_mk_internals.jl:59-62: accum + ExtProd(X^barai accum - accum, bki)."""
    return accum + tgsw_extern_mul_3gen(mul_by_monomial(accum, int(barai)) - accum, bki)


def mk_ith_blind_rotate_3gen(acc, gsw_key, bara):
    """3gen_mk_internals.jl:66-75: one party's n steps; zero rotations are skipped.  `gsw_key` is the party's
    `tgsw_samples` (or the key part itself); `bara` int32 [n]."""
    if isinstance(gsw_key, TransformedBootstrapKeyPart_3gen):
        gsw_key = gsw_key.tgsw_samples
    for i, barai in enumerate(np.asarray(bara).reshape(-1)):
        if barai != 0:
            acc = mk_mux_rotate_3gen(acc, gsw_key[i], barai)
    return acc


def mk_blind_rotate_3gen(accum, bk, bara):
    """3gen_mk_internals.jl:78-84: parties outer, coefficients inner.  `bara` is int32 [k][n] (the reference's (n, parties)
    column-major array).  One external-product launch per step: the stage-by-stage form of the path, for parity work -- gates and
    bootstraps run the whole loop inside one kernel (mk_bootstrap_3gen)."""
    bara = np.asarray(bara)
    for i in range(len(bk)):
        accum = mk_ith_blind_rotate_3gen(accum, bk[i], bara[i])
    return accum


def xor_3gen_gpu(bk, ks, x, y):
    """gpu_circuits.jl:27-36 (the reference's function builds the linear part and stops); finished here."""
    return mk_gate_xor_3gen(bk, ks, x, y)


# gates, 3gen_mk_gates.jl:8-150
def mk_gate_nand_3gen(bk, ks, x, y):
    return _gate(bk, ks, _cabi.GATE_NAND, x, y)


def mk_gate_or_3gen(bk, ks, x, y):
    return _gate(bk, ks, _cabi.GATE_OR, x, y)


def mk_gate_and_3gen(bk, ks, x, y):
    return _gate(bk, ks, _cabi.GATE_AND, x, y)


def mk_gate_3and_3gen(bk, ks, x, y, z):
    return _gate(bk, ks, _cabi.GATE_AND3, x, y, z)


def mk_gate_xor_3gen(bk, ks, x, y):
    return _gate(bk, ks, _cabi.GATE_XOR, x, y)


def mk_gate_xor_3gen_gpu(bk, ks, x, y):
    """3gen_mk_gates.jl:437-445 / gpu_circuits.jl:27-36, finished: the reference stops after the linear part."""
    return mk_gate_xor_3gen(bk, ks, x, y)


def mk_gate_not_3gen(x):
    return -x


def _gate_wb(bk, mu0, cx, x, y):
    """The `_wb` ("without bootstrap") variants, 3gen_mk_gates.jl:16-21, 32-37, 48-53, 76-81: the linear prologue of the gate alone,
    returned un-bootstrapped.  Host (or torch) arithmetic on the sample; bk only supplies the party count, ks is unused, as there."""
    if isinstance(x, MKLweSampleGPU):
        t = mk_lwe_noiseless_trivial_gpu(mu0, x.params, len(bk), tuple(x.b.shape), x.a.device.index)
    else:
        t = mk_lwe_noiseless_trivial(mu0, x.params, len(bk), x.b.shape)
    if cx == -1:
        return t - x - y
    if cx == 1:
        return t + x + y
    return t + cx * x + cx * y


def mk_gate_nand_3gen_wb(bk, ks, x, y):
    return _gate_wb(bk, encode_message(1, 8), -1, x, y)


def mk_gate_or_3gen_wb(bk, ks, x, y):
    return _gate_wb(bk, encode_message(1, 8), 1, x, y)


def mk_gate_and_3gen_wb(bk, ks, x, y):
    return _gate_wb(bk, encode_message(-1, 8), 1, x, y)


def mk_gate_xor_3gen_wb(bk, ks, x, y):
    return _gate_wb(bk, encode_message(1, 4), 2, x, y)


def mk_gate_mux_3gen(bk, ks, x, y, z):
    """3gen_mk_gates.jl:133-150: AND(x, y) and AND(-x, z) bootstrapped (one launch: the two ANDs are
    independent), then 1/8 + t1 + t2 NOT bootstrapped, exactly as the reference."""
    k = len(bk)
    if isinstance(x, MKLweSampleGPU):
        import torch
        both = _gate(bk, ks, _cabi.GATE_AND, MKLweSampleGPU(x.params, torch.stack([x.a, -x.a]), torch.stack([x.b, -x.b])),
                     MKLweSampleGPU(y.params, torch.stack([y.a, z.a]), torch.stack([y.b, z.b])))
        return mk_lwe_noiseless_trivial_gpu(encode_message(1, 8), x.params, k, tuple(x.b.shape), x.a.device.index) + both[0] + both[1]
    both = _gate(bk, ks, _cabi.GATE_AND, MKLweSample.stack([x, -x]), MKLweSample.stack([y, z]))
    t1, t2 = both[0], both[1]
    return mk_lwe_noiseless_trivial(encode_message(1, 8), t1.params, k, t1.b.shape) + t1 + t2


def mk_copy_3gen(x):
    if isinstance(x, MKLweSampleGPU):
        return MKLweSampleGPU(x.params, x.a.clone(), x.b.clone(), x.current_variance)
    return MKLweSample(x.params, x.a.copy(), x.b.copy(), x.current_variance)
