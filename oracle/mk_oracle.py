"""ctypes binding of the CPU oracle (oracle/mk_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(torus-fhe_b200/) must never import this module.

PARITY UNPINNED: see oracle/mk_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_mk3gen.so")

EXACT_SCHOOLBOOK, EXACT_NTT, FFT = 0, 1, 2
GATE_NAND, GATE_OR, GATE_AND, GATE_XOR, GATE_AND3 = 0, 1, 2, 3, 4


class Params(C.Structure):
    """SchemeParameters_3gen (3-gen-mk-tfhe/src/api.jl:50-67)."""
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("k", C.c_int32), ("l", C.c_int32),
                ("bgbit", C.c_int32), ("t", C.c_int32), ("basebit", C.c_int32), ("_pad", C.c_int32),
                ("sigma_lwe", C.c_double), ("sigma_gsw", C.c_double), ("sigma_ks", C.c_double)]


def params(n, N, k, l, bgbit, t, basebit, sigma_lwe, sigma_gsw, sigma_ks):
    return Params(n, N, k, l, bgbit, t, basebit, 0, sigma_lwe, sigma_gsw, sigma_ks)


# 3-gen-mk-tfhe/src/mk_api.jl:32-38, 44-50, 84-90, 98-104, 140-146
PARAMS_2PARTY = dict(n=520, N=1024, k=2, l=2, bgbit=7, t=3, basebit=3,
                     sigma_lwe=2.0 ** -13.52, sigma_gsw=2.0 ** -30.70, sigma_ks=2.0 ** -13.52)
PARAMS_3PARTY = dict(n=510, N=1024, k=3, l=2, bgbit=7, t=5, basebit=2,
                     sigma_lwe=2.0 ** -13.26, sigma_gsw=2.0 ** -30.70, sigma_ks=2.0 ** -13.26)
PARAMS_5PARTY = dict(n=520, N=1024, k=5, l=3, bgbit=6, t=5, basebit=2,
                     sigma_lwe=2.0 ** -13.52, sigma_gsw=2.0 ** -30.70, sigma_ks=2.0 ** -13.52)
PARAMS_4PARTY = dict(n=510, N=1024, k=4, l=3, bgbit=6, t=5, basebit=2,
                     sigma_lwe=2.0 ** -13.26, sigma_gsw=2.0 ** -30.70, sigma_ks=2.0 ** -13.26)
PARAMS_8PARTY = dict(n=540, N=1024, k=8, l=4, bgbit=4, t=5, basebit=2,
                     sigma_lwe=2.0 ** -14.04, sigma_gsw=2.0 ** -30.70, sigma_ks=2.0 ** -14.04)


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("mk_oracle.c", "mk_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    PP = C.POINTER(Params)
    sigs = {
        "mko_encode_message32": (i32, [i64, C.c_int]),
        "mko_encode_message64": (i64, [i64, C.c_int]),
        "mko_decode_message32": (i32, [i32, C.c_int]),
        "mko_dtot32": (i32, [dbl]),
        "mko_dtot64": (i64, [dbl]),
        "mko_t64tot32": (i32, [i64]),
        "mko_gadget_offset": (i64, [C.c_int, C.c_int]),
        "mko_decompose": (None, [vp, C.c_int, C.c_int, C.c_int, vp]),
        "mko_mul_by_monomial": (None, [vp, C.c_int, i64, vp]),
        "mko_negacyclic_mul_schoolbook": (None, [vp, vp, C.c_int, vp]),
        "mko_negacyclic_mul_ntt": (None, [vp, vp, C.c_int, vp]),
        "mko_negacyclic_mul_fft": (None, [vp, vp, C.c_int, vp]),
        "mko_keygen": (vp, [PP, u64, C.c_int]),
        "mko_keyset_from_raw": (vp, [PP, vp, vp]),
        "mko_keyset_free": (None, [vp]),
        "mko_bsk": (vp, [vp]), "mko_ksk": (vp, [vp]), "mko_lwe_keys": (vp, [vp]), "mko_rlwe_keys": (vp, [vp]),
        "mko_bsk_len": (C.c_size_t, [vp]), "mko_ksk_len": (C.c_size_t, [vp]),
        "mko_prepare_fft_key": (None, [vp]), "mko_prepare_ntt_key": (None, [vp]),
        "mko_encrypt": (None, [vp, u64, C.c_int, vp, vp, vp]),
        "mko_phase": (None, [vp, C.c_int, vp, vp, vp]),
        "mko_extprod": (None, [vp, C.c_int, C.c_int, C.c_int, vp, vp]),
        "mko_mux_rotate": (None, [vp, C.c_int, C.c_int, C.c_int, i32, vp]),
        "mko_bootstrap_wo_keyswitch": (None, [vp, C.c_int, i64, vp, i32, vp, vp, vp, vp]),
        "mko_keyswitch": (None, [vp, vp, i32, vp, vp]),
        "mko_bootstrap_batch": (None, [vp, C.c_int, i64, C.c_int, vp, vp, vp, vp, C.c_int]),
        "mko_gate_batch": (None, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class KeySet:
    """Key material of multikey_3gen.jl:15-30, oracle-generated (or raw synthetic keys)."""

    def __init__(self, prm, seed=None, nthreads=8, raw_bsk=None, raw_ksk=None):
        self.prm = prm if isinstance(prm, Params) else params(**prm)
        L = lib()
        if raw_bsk is not None:
            self._bsk_keep, self._ksk_keep = _c(raw_bsk, np.int64), _c(raw_ksk, np.int32)
            self.h = L.mko_keyset_from_raw(C.byref(self.prm), _p(self._bsk_keep), _p(self._ksk_keep))
            self.has_secrets = False
        else:
            self.h = L.mko_keygen(C.byref(self.prm), seed, nthreads)
            self.has_secrets = True
        p = self.prm
        self.n, self.N, self.k, self.l, self.t = p.n, p.N, p.k, p.l, p.t
        self.B1 = (1 << p.basebit) - 1

    def __del__(self):
        if getattr(self, "h", None):
            lib().mko_keyset_free(self.h)
            self.h = None

    def _view(self, ptr, count, dt):
        buf = (C.c_char * (count * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt, count=count)

    @property
    def bsk(self):  # int64 [k][n][4][l][N]
        L = lib()
        return self._view(L.mko_bsk(self.h), L.mko_bsk_len(self.h), np.int64).reshape(self.k, self.n, 4, self.l, self.N)

    @property
    def ksk(self):  # int32 [k][N][t][B-1][n+1]
        L = lib()
        return self._view(L.mko_ksk(self.h), L.mko_ksk_len(self.h), np.int32).reshape(self.k, self.N, self.t, self.B1, self.n + 1)

    @property
    def lwe_keys(self):
        return self._view(lib().mko_lwe_keys(self.h), self.k * self.n, np.int32).reshape(self.k, self.n)

    @property
    def rlwe_keys(self):
        return self._view(lib().mko_rlwe_keys(self.h), self.k * self.N, np.int64).reshape(self.k, self.N)

    # -- encrypt / decrypt (mk_api.jl:519-536, 607-610)
    def encrypt(self, bits, seed):
        bits = _c(bits, np.uint8).ravel()
        G = bits.size
        a = np.empty((G, self.k, self.n), np.int32)
        b = np.empty(G, np.int32)
        lib().mko_encrypt(self.h, seed, G, _p(bits), _p(a), _p(b))
        return a, b

    def phase(self, a, b):
        a, b = _c(a, np.int32), _c(b, np.int32)
        ph = np.empty(b.size, np.int32)
        lib().mko_phase(self.h, b.size, _p(a), _p(b), _p(ph))
        return ph

    def decrypt(self, a, b):
        return self.phase(a, b) > 0

    # -- hot path
    def extprod(self, backend, party, j, acc):
        acc = _c(acc, np.int64)
        out = np.empty_like(acc)
        lib().mko_extprod(self.h, backend, party, j, _p(acc), _p(out))
        return out

    def mux_rotate(self, backend, party, j, bara, acc):
        acc = _c(acc, np.int64).copy()
        lib().mko_mux_rotate(self.h, backend, party, j, int(bara), _p(acc))
        return acc

    def bootstrap_wo_keyswitch(self, backend, mu, a, b, want_acc=False, want_digits=False):
        a = _c(a, np.int32)
        ext_a = np.empty(self.N, np.int32)
        ext_b = C.c_int32(0)
        acc = np.empty((2, self.N), np.int64) if want_acc else None
        dl = np.empty(self.k * self.n, np.uint64) if want_digits else None
        lib().mko_bootstrap_wo_keyswitch(self.h, backend, mu, _p(a), int(b), _p(ext_a), C.byref(ext_b), _p(acc), _p(dl))
        return ext_a, np.int32(ext_b.value), acc, dl

    def keyswitch(self, ext_a, ext_b):
        ext_a = _c(ext_a, np.int32)
        oa = np.empty((self.k, self.n), np.int32)
        ob = C.c_int32(0)
        lib().mko_keyswitch(self.h, _p(ext_a), int(ext_b), _p(oa), C.byref(ob))
        return oa, np.int32(ob.value)

    def bootstrap_batch(self, backend, mu, a, b, nthreads=8):
        a, b = _c(a, np.int32), _c(b, np.int32)
        G = b.size
        oa, ob = np.empty((G, self.k, self.n), np.int32), np.empty(G, np.int32)
        lib().mko_bootstrap_batch(self.h, backend, mu, G, _p(a), _p(b), _p(oa), _p(ob), nthreads)
        return oa, ob

    def gate_batch(self, backend, gate, x, y, z=None, nthreads=8):
        xa, xb = _c(x[0], np.int32), _c(x[1], np.int32)
        ya, yb = _c(y[0], np.int32), _c(y[1], np.int32)
        za = zb = None
        if z is not None:
            za, zb = _c(z[0], np.int32), _c(z[1], np.int32)
        G = xb.size
        oa, ob = np.empty((G, self.k, self.n), np.int32), np.empty(G, np.int32)
        lib().mko_gate_batch(self.h, backend, gate, G, _p(xa), _p(xb), _p(ya), _p(yb), _p(za), _p(zb), _p(oa), _p(ob), nthreads)
        return oa, ob


def decompose(poly, l, bgbit):
    poly = _c(poly, np.int64)
    out = np.empty((l, poly.size), np.int64)
    lib().mko_decompose(_p(poly), poly.size, l, bgbit, _p(out))
    return out


def mul_by_monomial(poly, shift):
    poly = _c(poly, np.int64)
    out = np.empty_like(poly)
    lib().mko_mul_by_monomial(_p(poly), poly.size, int(shift), _p(out))
    return out


def negacyclic_mul(a, b, backend=EXACT_SCHOOLBOOK):
    a, b = _c(a, np.int64), _c(b, np.int64)
    out = np.empty_like(a)
    fn = {EXACT_SCHOOLBOOK: lib().mko_negacyclic_mul_schoolbook, EXACT_NTT: lib().mko_negacyclic_mul_ntt,
          FFT: lib().mko_negacyclic_mul_fft}[backend]
    fn(_p(a), _p(b), a.size, _p(out))
    return out
