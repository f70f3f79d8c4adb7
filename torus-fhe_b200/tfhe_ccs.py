"""The CCS multi-key scheme (Chen-Chillotti-Song; the reference's `mk_gate_nand` / `mk_bootstrap`, src/mk_api.jl:333-459,
src/mk_internals.jl:105-171, 215-262, 351-535, 703-719, 793-858, src/mk_gates.jl) on the SAME external-product kernels as the 3gen
path -- SURVEY.md section 8(f) rank 4, second slice.

CCS keeps a (k+1)-polynomial accumulator (b, a_1..a_k) and its hybrid product (UniProduct, mk_internals.jl:471-535) is, per polynomial,
two rounds of "decompose, multiply the digits with two rows of key material".  Each such round IS an external product of the engine with
the polynomial as body, a zero mask, and a key element whose part_1 / part_4 hold the two rows (part_2 = part_3 = 0):
    round 1, polynomial i:  body' = <g^-1(acc_i), d>        = u_i      mask' = <g^-1(acc_i), b_i>  (i >= 1)  or  <g^-1(b), -a>  (i = 0)   = v_i
    round 2, polynomial i:  body' = <g^-1(v_i), f0>         = w0_i     mask' = <g^-1(v_i), f1>                                                = w1_i
    a_i += u_i,   a_party += sum_i w1_i,   b += u_0 + sum_i w0_i.
So one blind-rotate step of a batch of G gates is two launches of G (k+1) external products (Torus32 mode: 9-bit gadget digits) with
small rotation / accumulation kernels between them -- all inside the library (`mktfhe_ccs_blind_rotate_batch`); accumulators stay in HBM
for the whole blind rotation.  Extraction (mk_rlwe_extract_sample) gives one mask per party, which
`mktfhe_mk_keyswitch_batch` switches with party p's key (mk_keyswitch, :703-719).  This is a composition, not a fused kernel: it is
bit-exact and three orders of magnitude faster than the CPU path, but it leaves the factor a fused (k+1)-polynomial kernel would
bring (half of each external product multiplies zeros) on the table -- DESIGN.md section 6/7.

Names and arguments follow the reference; samples may carry a leading batch dimension.  Served: `mktfhe_parameters_2party` (l = 3,
Bg = 2^9) and `mktfhe_parameters_4party` (l = 4, Bg = 2^8); the 8- and 16-party CCS sets use l = 5 and 12, beyond the N = 1024 kernels'
l <= 4.
"""
import numpy as np

from . import _cabi
from .engine import Engine
from .tfhe1 import SchemeParameters, keyswitch_parameters, lwe_parameters, rlwe_parameters, tgsw_parameters
from .tfhe3gen import (KeyswitchKey, LweKey, LweParams, MKLweSample, RLweKey, SchemeParameters_3gen, dtot32, encode_message, mk_lwe_noiseless_trivial,
                       mk_lwe_phase, negacyclic_mul, rand_uniform_torus32)

# mk_api.jl:4-10, 56-62
mktfhe_parameters_2party = SchemeParameters(560, 3.05e-5, 1024, 1, True, 3, 9, 3.72e-9, 8, 2, 3.05e-5, 2)
mktfhe_parameters_4party = SchemeParameters(560, 3.05e-5, 1024, 1, True, 4, 8, 3.72e-9, 8, 2, 3.05e-5, 4)


class SecretKey:   # api.jl:176-184
    def __init__(self, rng, params, key=None):
        self.params = params
        self.key = LweKey(rng, lwe_parameters(params))
        if key is not None:
            self.key.key = np.asarray(key, dtype=np.int32)


class SharedKey:   # mk_internals.jl:159-171, mk_api.jl:333-339: l uniform Torus32 polynomials shared by the parties
    def __init__(self, rng, params, a=None):
        self.tgsw_params, self.rlwe_params = tgsw_parameters(params), rlwe_parameters(params)
        l, N = self.tgsw_params.decomp_length, self.rlwe_params.polynomial_degree
        self.a = np.asarray(a, np.int32) if a is not None else rand_uniform_torus32(rng, l * N).reshape(l, N)


def _gauss32(rng, sigma, shape):
    return dtot32(rng.standard_normal(shape) * sigma).astype(np.int64)


class CloudKeyPart:
    """mk_api.jl:362-381: what one party contributes -- its public key b = e + z a (mk_internals.jl:215-262), the uni-encryptions of its
    LWE key bits (mk_tgsw_encrypt, :351-437; of the six rows only d1, f0, f1 are ever read by the hybrid product, so only they are
    kept: d, f0, f1 int32 [n][l][N]) and its key-switching key."""

    def __init__(self, rng, secret_key, shared_key, arrays=None):
        params = secret_key.params
        self.params = params
        if arrays is not None:                                        # key bytes made elsewhere (the oracle's generator, in the tests)
            self.pk, self.d, self.f0, self.f1, self.ks = arrays
            return
        tp = tgsw_parameters(params)
        n, l, N, alpha = params.lwe_size, tp.decomp_length, params.rlwe_polynomial_degree, params.bs_noise_stddev
        rlwe_key = RLweKey(rng, rlwe_parameters(params), negative_random=False)
        z, a = rlwe_key.key.astype(np.int64), shared_key.a.astype(np.int64)
        g = np.array([1 << (32 - (q + 1) * tp.log2_base) for q in range(l)], np.int64)
        self.pk = (negacyclic_mul(z[None], a) + _gauss32(rng, alpha, (l, N))).astype(np.int32)
        r = rng.integers(0, 2, (n, 1, N), dtype=np.int64)              # the shared randomness of each uni-encryption
        d = negacyclic_mul(r, a[None]) + _gauss32(rng, alpha, (n, l, N))
        d[:, :, 0] += secret_key.key.key.astype(np.int64)[:, None] * g[None, :]
        f1 = rand_uniform_torus32(rng, n * l * N).reshape(n, l, N)
        f0 = negacyclic_mul(z[None, None], f1.astype(np.int64)) + _gauss32(rng, alpha, (n, l, N)) + r * g[None, :, None]
        self.d, self.f0, self.f1 = d.astype(np.int32), f0.astype(np.int32), f1
        self.ks = KeyswitchKey(rng, params.ks_noise_stddev, keyswitch_parameters(params), secret_key.key, rlwe_key)


def build_elements(ck_parts, shared_key):
    """The key elements of the hybrid product in the engine's layout: int64 [k (k + 2) pseudo-parties][n][4 parts][l][N], values unshifted
    (Torus32 mode); the engine addresses element (pseudo-party q, j) as q n + j.  Pseudo-party party (k + 1) + i, i = 0..k, serves round 1 of
    polynomial i in the steps of `party`'s key bits: part_1 = d[party][j], part_4 = -a (i = 0: the b polynomial) or pk_i (i >= 1);
    pseudo-party k (k + 1) + party serves round 2: part_1 = f0[party][j], part_4 = f1[party][j].  part_2 = part_3 = 0 (the mask input of
    these products is zero)."""
    k = len(ck_parts)
    n, l, N = ck_parts[0].d.shape
    elems = np.zeros((k * (k + 2), n, 4, l, N), np.int64)
    second = [-shared_key.a.astype(np.int64)] + [part.pk.astype(np.int64) for part in ck_parts]
    for pi, part in enumerate(ck_parts):
        for i in range(k + 1):
            elems[pi * (k + 1) + i, :, 0] = part.d
            elems[pi * (k + 1) + i, :, 3] = second[i][None]
        elems[k * (k + 1) + pi, :, 0], elems[k * (k + 1) + pi, :, 3] = part.f0, part.f1
    return elems


class MKCloudKey:
    """mk_api.jl:390-405.  Building it loads the GPU: one Torus32-mode context holding the k (k + 2) n key elements of the hybrid
    product (as k (k + 2) pseudo-parties of n elements), one context holding the parties' key-switching keys."""

    def __init__(self, ck_parts, shared_key, device=0):
        self.parties, self.params = len(ck_parts), ck_parts[0].params
        p = self.params
        tp = tgsw_parameters(p)
        k, n, l, N = self.parties, p.lwe_size, tp.decomp_length, p.rlwe_polynomial_degree
        if k > p.max_parties:
            raise ValueError("more key parts than max_parties")
        if not p.rlwe_is32:
            raise NotImplementedError("CCS sets are Torus32")
        self.k, self.n, self.l, self.N = k, n, l, N
        elems = build_elements(ck_parts, shared_key)
        sp = SchemeParameters_3gen(n, 0.0, N, 1, False, l, tp.log2_base, 0.0, p.ks_decomp_length, p.ks_log2_base, 0.0, elems.shape[0])
        self.products = Engine(sp, device=device, flags=_cabi.FLAG_TORUS32)
        for q in range(elems.shape[0]):
            self.products.ctx.load_bsk(q, elems[q])
        self.products.ctx.mark_keys_received()                        # this context multiplies only: it holds no key-switching key
        self.products.ctx.finalize_keys()
        sk = SchemeParameters_3gen(n, 0.0, N, 1, False, 2, 7, 0.0, p.ks_decomp_length, p.ks_log2_base, 0.0, k)
        self.switch = Engine(sk, device=device)
        for pi, part in enumerate(ck_parts):
            self.switch.ctx.load_ksk(pi, part.ks.key)
        self.switch.ctx.mark_keys_received()                          # ... and this one switches only
        self.switch.ctx.finalize_keys()
        self.device = device

    def close(self):
        self.products.close()
        self.switch.close()


def mk_encrypt(rng, secret_keys, message):
    """mk_api.jl:447-459; `message` may be an array of bools."""
    params = secret_keys[0].params
    msg = np.asarray(message, dtype=bool)
    n, k = params.lwe_size, len(secret_keys)
    a = rand_uniform_torus32(rng, msg.size * k * n).reshape(msg.shape + (k, n))
    s = np.stack([sk.key.key for sk in secret_keys]).astype(np.int64)
    mu = np.where(msg, int(encode_message(1, 8)), int(encode_message(-1, 8)))
    e = dtot32(rng.standard_normal(msg.shape) * params.lwe_noise_stddev).astype(np.int64)
    b = (mu + e + (a.astype(np.int64) * s).sum((-1, -2))).astype(np.int32)
    return MKLweSample(lwe_parameters(params), a, b, params.lwe_noise_stddev ** 2)


def mk_decrypt(secret_keys, sample):
    """mk_api.jl:597-600."""
    r = mk_lwe_phase(sample, [sk.key for sk in secret_keys]) > 0
    return bool(r) if r.ndim == 0 else r


def mk_bootstrap_wo_keyswitch(ck, mu, x):
    """mk_internals.jl:839-850 on a batch: the extracted sample, one mask per party -- (ext_a int32 [G][k][N], ext_b int32 [G]).
    One library call (mktfhe_ccs_blind_rotate_batch): mod-switch, the k n steps of two rounds of G (k + 1) external products each with
    the rotation / accumulation kernels between them, extraction; accumulators never leave HBM."""
    return ck.products.ctx.ccs_blind_rotate_batch(ck.k, mu, x.a.reshape(-1, ck.k, ck.n), x.b.reshape(-1))


def mk_keyswitch(ck, ext_a, ext_b):
    """mk_internals.jl:703-719."""
    return ck.switch.ctx.mk_keyswitch_batch(ext_a, ext_b)


def mk_bootstrap(ck, mu, x):
    """mk_internals.jl:853-856 on a batch."""
    oa, ob = mk_keyswitch(ck, *mk_bootstrap_wo_keyswitch(ck, mu, x))
    return MKLweSample(x.params, oa.reshape(x.b.shape + (ck.k, ck.n)), ob.reshape(x.b.shape), 0.0)


def mk_gate_nand(ck, x, y):
    """mk_gates.jl:7-13."""
    temp = mk_lwe_noiseless_trivial(encode_message(1, 8), x.params, ck.parties, x.b.shape) - x - y
    return mk_bootstrap(ck, encode_message(1, 8), temp)
