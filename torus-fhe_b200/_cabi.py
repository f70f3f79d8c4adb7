"""ctypes binding of libmktfhe_b200.so (C ABI declared in include/mktfhe_b200.h).

This is the same boundary the Julia shim (julia/TFHE_B200.jl) binds with `ccall`.
There is no fallback: if the shared library is missing or no sm_100a GPU is
present, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MKTFHE_B200_LIB") or os.path.join(_HERE, "libmktfhe_b200.so")   # env override: kernel experiments

OK, EINVAL, ECUDA, ESTATE, ENOMEM = 0, -1, -2, -3, -4
GATE_NAND, GATE_OR, GATE_AND, GATE_XOR, GATE_AND3 = 0, 1, 2, 3, 4
FLAG_TORUS32 = 1          # mktfhe_params.reserved: Torus32 mode (unshifted 32-bit keys, gadget digits of up to 16 bits)

# every symbol include/mktfhe_b200.h declares
EXPORTS = (
    "mktfhe_create", "mktfhe_create_multi", "mktfhe_destroy", "mktfhe_last_error",
    "mktfhe_device_count", "mktfhe_device_ctx", "mktfhe_shard_bounds", "mktfhe_pin_host", "mktfhe_unpin_host",
    "mktfhe_load_bsk", "mktfhe_load_ksk", "mktfhe_generate_ksk", "mktfhe_finalize_keys", "mktfhe_key_buffers", "mktfhe_mark_keys_received",
    "mktfhe_bootstrap_batch", "mktfhe_gate_batch", "mktfhe_bootstrap_batch_dev", "mktfhe_gate_batch_dev",
    "mktfhe_gate_batch_mixed", "mktfhe_gate_batch_mixed_dev", "mktfhe_affine_bootstrap_batch", "mktfhe_affine_bootstrap_batch_dev",
    "mktfhe_extprod_batch", "mktfhe_extprod_batch_dev", "mktfhe_mk_keyswitch_batch", "mktfhe_ccs_blind_rotate_batch", "mktfhe_blind_rotate_batch", "mktfhe_keyswitch_batch", "mktfhe_negacyclic_mul_batch",
    "mktfhe_launch_count", "mktfhe_last_kernel_ms", "mktfhe_algorithmic_bytes", "mktfhe_build_id", "mktfhe_describe",
)


class MktfheError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmktfhe_b200 error {code}: {msg}")
        self.code = code


class CParams(C.Structure):
    """mktfhe_params: integer part of SchemeParameters_3gen (3-gen-mk-tfhe/src/api.jl:50-67)."""
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("k", C.c_int32), ("l", C.c_int32),
                ("bgbit", C.c_int32), ("t", C.c_int32), ("basebit", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def build(verbose=False):
    """Compile libmktfhe_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    import subprocess
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-s"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode:
        raise RuntimeError("nvcc build of libmktfhe_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    sigs = {
        "mktfhe_create": (C.c_int, [C.POINTER(CParams), C.c_int, C.POINTER(vp)]),
        "mktfhe_create_multi": (C.c_int, [C.POINTER(CParams), C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]),
        "mktfhe_destroy": (None, [vp]),
        "mktfhe_device_count": (C.c_int, [vp]),
        "mktfhe_device_ctx": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int)]),
        "mktfhe_shard_bounds": (C.c_int, [vp, sz, C.c_int, C.POINTER(sz), C.POINTER(sz)]),
        "mktfhe_pin_host": (C.c_int, [vp, sz]),
        "mktfhe_unpin_host": (C.c_int, [vp]),
        "mktfhe_build_id": (C.c_char_p, []),
        "mktfhe_describe": (C.c_int, [vp, C.c_char_p, sz]),
        "mktfhe_last_error": (C.c_char_p, [vp]),
        "mktfhe_load_bsk": (C.c_int, [vp, C.c_int, vp]),
        "mktfhe_load_ksk": (C.c_int, [vp, C.c_int, vp]),
        "mktfhe_generate_ksk": (C.c_int, [vp, C.c_int, vp, vp, C.c_double, C.c_uint64]),
        "mktfhe_finalize_keys": (C.c_int, [vp]),
        "mktfhe_key_buffers": (C.c_int, [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(sz)]),
        "mktfhe_mark_keys_received": (C.c_int, [vp]),
        "mktfhe_bootstrap_batch": (C.c_int, [vp, i64, sz, vp, vp, vp, vp]),
        "mktfhe_gate_batch": (C.c_int, [vp, C.c_int, sz, vp, vp, vp, vp, vp, vp, vp, vp]),
        "mktfhe_bootstrap_batch_dev": (C.c_int, [vp, i64, sz, vp, vp, vp, vp, vp]),
        "mktfhe_gate_batch_dev": (C.c_int, [vp, C.c_int, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "mktfhe_affine_bootstrap_batch": (C.c_int, [vp, i32, i32, i32, i32, i64, sz, vp, vp, vp, vp, vp, vp, vp, vp]),
        "mktfhe_affine_bootstrap_batch_dev": (C.c_int, [vp, i32, i32, i32, i32, i64, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "mktfhe_gate_batch_mixed": (C.c_int, [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "mktfhe_gate_batch_mixed_dev": (C.c_int, [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "mktfhe_extprod_batch": (C.c_int, [vp, sz, vp, vp, vp]),
        "mktfhe_extprod_batch_dev": (C.c_int, [vp, sz, vp, vp, vp, vp]),
        "mktfhe_mk_keyswitch_batch": (C.c_int, [vp, sz, vp, vp, vp, vp]),
        "mktfhe_ccs_blind_rotate_batch": (C.c_int, [vp, C.c_int, i32, sz, vp, vp, vp, vp]),
        "mktfhe_blind_rotate_batch": (C.c_int, [vp, i64, sz, vp, vp, vp, vp]),
        "mktfhe_keyswitch_batch": (C.c_int, [vp, sz, vp, vp, vp]),
        "mktfhe_negacyclic_mul_batch": (C.c_int, [vp, sz, vp, vp, vp]),
        "mktfhe_launch_count": (C.c_uint64, [vp]),
        "mktfhe_last_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "mktfhe_algorithmic_bytes": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    }
    assert set(sigs) == set(EXPORTS)
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Context:
    """One mktfhe_ctx: its keys and, per GPU it spans, one stream.  `device` = one GPU (mktfhe_create); `devices` = a list of
    GPUs, or "all", behind one handle (mktfhe_create_multi): keys are broadcast inside finalize_keys and the host-pointer
    batch calls shard their batch over the GPUs."""

    def __init__(self, n, N, k, l, bgbit, t, basebit, device=0, devices=None, _borrowed=None, flags=0):
        L = lib()
        self.prm = CParams(n, N, k, l, bgbit, t, basebit, flags)
        self.flags = flags
        self.n, self.N, self.k, self.l, self.bgbit, self.t, self.basebit = n, N, k, l, bgbit, t, basebit
        self._owned = _borrowed is None
        if _borrowed is not None:                       # a replica of a multi-device context (owned by its parent)
            self.h, self.device, self.devices = C.c_void_p(_borrowed), device, [device]
            return
        h = C.c_void_p()
        if devices is None:
            rc = L.mktfhe_create(C.byref(self.prm), device, C.byref(h))
        elif isinstance(devices, str):
            if devices != "all":
                raise ValueError('devices must be a list of GPU ordinals or "all"')
            rc = L.mktfhe_create_multi(C.byref(self.prm), 0, None, C.byref(h))
        else:
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = L.mktfhe_create_multi(C.byref(self.prm), len(devices), devs, C.byref(h))
        if rc:
            raise MktfheError(rc, L.mktfhe_last_error(None).decode())
        self.h = h
        self.devices = [self.replica_device(i) for i in range(self.device_count())]
        self.device = self.devices[0]

    def close(self):
        if getattr(self, "h", None) and getattr(self, "_owned", False):
            lib().mktfhe_destroy(self.h)
        self.h = None

    # -- multi-device
    def device_count(self):
        return int(lib().mktfhe_device_count(self.h))

    def replica_device(self, i):
        d = C.c_int()
        self._chk(lib().mktfhe_device_ctx(self.h, i, None, C.byref(d)))
        return d.value

    def replica(self, i):
        """Replica i as a single-device Context (borrowed: valid while this context lives) for the *_dev calls."""
        r, d = C.c_void_p(), C.c_int()
        self._chk(lib().mktfhe_device_ctx(self.h, i, C.byref(r), C.byref(d)))
        return Context(self.n, self.N, self.k, self.l, self.bgbit, self.t, self.basebit, device=d.value, _borrowed=r.value, flags=self.flags)

    def shard_bounds(self, G, i):
        lo, hi = C.c_size_t(), C.c_size_t()
        self._chk(lib().mktfhe_shard_bounds(self.h, G, i, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def describe(self):
        import json
        buf = C.create_string_buffer(1024)
        self._chk(lib().mktfhe_describe(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())

    __del__ = close

    def _chk(self, rc):
        if rc:
            raise MktfheError(rc, lib().mktfhe_last_error(self.h).decode())

    # -- keys
    def load_bsk(self, party, polys):
        """polys: int64 [n][4][l][N] = bk.gsw_key[j].part_{1..4}[q].coeffs (3gen_mk_internals.jl:10-43)."""
        polys = _c(polys, np.int64)
        if polys.size != self.n * 4 * self.l * self.N:
            raise ValueError(f"bsk of party {party}: expected {(self.n, 4, self.l, self.N)}, got {polys.shape}")
        self._chk(lib().mktfhe_load_bsk(self.h, party, _p(polys)))

    def load_ksk(self, party, rows):
        """rows: int32 [N][t][base-1][n+1] (keyswitch.jl:7-42)."""
        rows = _c(rows, np.int32)
        B1 = (1 << self.basebit) - 1
        if rows.size != self.N * self.t * B1 * (self.n + 1):
            raise ValueError(f"ksk of party {party}: expected {(self.N, self.t, B1, self.n + 1)}, got {rows.shape}")
        self._chk(lib().mktfhe_load_ksk(self.h, party, _p(rows)))

    def generate_ksk(self, party, lwe_key, rlwe_key, sigma, seed):
        """keyswitch.jl:14-41 on the GPU, straight into the device key (mktfhe_generate_ksk)."""
        s, z = _c(lwe_key, np.int32).reshape(-1), _c(rlwe_key, np.int64).reshape(-1)
        if s.size != self.n or z.size != self.N:
            raise ValueError(f"generate_ksk: expected an LWE key of {self.n} and an RLWE key of {self.N} coefficients")
        self._chk(lib().mktfhe_generate_ksk(self.h, party, _p(s), _p(z), float(sigma), int(seed) & 0xFFFFFFFFFFFFFFFF))

    def finalize_keys(self):
        self._chk(lib().mktfhe_finalize_keys(self.h))

    def key_buffers(self):
        bp, bb, kp, kb = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        self._chk(lib().mktfhe_key_buffers(self.h, C.byref(bp), C.byref(bb), C.byref(kp), C.byref(kb)))
        return (bp.value, bb.value), (kp.value, kb.value)

    def mark_keys_received(self):
        self._chk(lib().mktfhe_mark_keys_received(self.h))

    # -- hot path, host buffers
    def _ab(self, a, b):
        b = _c(b, np.int32).reshape(-1)
        a = _c(a, np.int32)
        if a.size != b.size * self.k * self.n:
            raise ValueError(f"ciphertext batch: a has {a.size} words, expected {b.size}*{self.k}*{self.n}")
        return a, b

    def bootstrap_batch(self, mu, a, b):
        a, b = self._ab(a, b)
        G = b.size
        oa, ob = np.empty((G, self.k, self.n), np.int32), np.empty(G, np.int32)
        self._chk(lib().mktfhe_bootstrap_batch(self.h, int(mu), G, _p(a), _p(b), _p(oa), _p(ob)))
        return oa, ob

    def gate_batch(self, gate, x, y, z=None, out=None):
        xa, xb = self._ab(*x)
        ya, yb = self._ab(*y)
        G = xb.size
        if yb.size != G:
            raise ValueError("gate operands have different batch sizes")
        za = zb = None
        if z is not None:
            za, zb = self._ab(*z)
        if out is None:
            oa, ob = np.empty((G, self.k, self.n), np.int32), np.empty(G, np.int32)
        else:
            oa, ob = out
        self._chk(lib().mktfhe_gate_batch(self.h, gate, G, _p(xa), _p(xb), _p(ya), _p(yb), _p(za), _p(zb), _p(oa), _p(ob)))
        return oa, ob

    def affine_bootstrap_batch(self, mu0, cx, cy, cz, mu, x, y=None, z=None):
        """bootstrap(mu0 + cx x + cy y + cz z) with test-vector message mu: the general form of every bootstrapped gate."""
        xa, xb = self._ab(*x)
        G = xb.size
        ya = yb = za = zb = None
        if y is not None:
            ya, yb = self._ab(*y)
        if z is not None:
            za, zb = self._ab(*z)
        if (cy and y is None) or (cz and z is None):
            raise ValueError("a non-zero coefficient needs its operand")
        oa, ob = np.empty((G, self.k, self.n), np.int32), np.empty(G, np.int32)
        w = lambda v: int(np.int32(np.uint32(int(v) & 0xFFFFFFFF)))
        self._chk(lib().mktfhe_affine_bootstrap_batch(self.h, w(mu0), w(cx), w(cy), w(cz), int(mu), G, _p(xa), _p(xb), _p(ya), _p(yb), _p(za), _p(zb),
                                                      _p(oa), _p(ob)))
        return oa, ob

    def gate_batch_mixed(self, gate_ids, x, y, z=None):
        """One launch for gates of different kinds: gate_ids[g] selects the prologue of gate g."""
        gate_ids = _c(gate_ids, np.int32).reshape(-1)
        xa, xb = self._ab(*x)
        ya, yb = self._ab(*y)
        G = xb.size
        if yb.size != G or gate_ids.size != G:
            raise ValueError("gate operands / ids have different batch sizes")
        za = zb = None
        if z is not None:
            za, zb = self._ab(*z)
        oa, ob = np.empty((G, self.k, self.n), np.int32), np.empty(G, np.int32)
        self._chk(lib().mktfhe_gate_batch_mixed(self.h, G, _p(gate_ids), _p(xa), _p(xb), _p(ya), _p(yb), _p(za), _p(zb), _p(oa), _p(ob)))
        return oa, ob

    # -- hot path, device pointers (ints), asynchronous on `stream` (0 = context stream)
    def bootstrap_batch_dev(self, mu, G, a_in, b_in, a_out, b_out, stream=0):
        self._chk(lib().mktfhe_bootstrap_batch_dev(self.h, int(mu), G, a_in, b_in, a_out, b_out, stream or None))

    def gate_batch_dev(self, gate, G, xa, xb, ya, yb, za, zb, oa, ob, stream=0):
        self._chk(lib().mktfhe_gate_batch_dev(self.h, gate, G, xa, xb, ya, yb, za or None, zb or None, oa, ob, stream or None))

    def affine_bootstrap_batch_dev(self, mu0, cx, cy, cz, mu, G, xa, xb, ya, yb, za, zb, oa, ob, stream=0):
        w = lambda v: int(np.int32(np.uint32(int(v) & 0xFFFFFFFF)))
        self._chk(lib().mktfhe_affine_bootstrap_batch_dev(self.h, w(mu0), w(cx), w(cy), w(cz), int(mu), G, xa, xb, ya or None, yb or None,
                                                          za or None, zb or None, oa, ob, stream or None))

    def gate_batch_mixed_dev(self, G, gate_ids, xa, xb, ya, yb, za, zb, oa, ob, stream=0):
        self._chk(lib().mktfhe_gate_batch_mixed_dev(self.h, G, gate_ids, xa, xb, ya, yb, za or None, zb or None, oa, ob, stream or None))

    # -- parity hooks
    def extprod_batch(self, elem, acc):
        elem = _c(elem, np.int32).reshape(-1)
        acc = _c(acc, np.int64)
        G = elem.size
        if acc.size != G * 2 * self.N:
            raise ValueError("acc must be int64 [G][2][N]")
        out = np.empty((G, 2, self.N), np.int64)
        self._chk(lib().mktfhe_extprod_batch(self.h, G, _p(elem), _p(acc), _p(out)))
        return out

    def extprod_batch_dev(self, G, elem, acc_in, acc_out, stream=0):
        """Device pointers (ints): acc_out[g] = ExtProd(acc_in[g], element elem[g]); asynchronous on `stream`."""
        self._chk(lib().mktfhe_extprod_batch_dev(self.h, G, elem, acc_in, acc_out, stream or None))

    def ccs_blind_rotate_batch(self, parties, mu, a, b):
        """CCS mk_bootstrap_wo_keyswitch on a batch (this context holds the hybrid product's key elements): a int32 [G][parties][n], b [G]
        -> (ext_a int32 [G][parties][N], ext_b [G])."""
        a, b = _c(a, np.int32), _c(b, np.int32).reshape(-1)
        G = b.size
        if a.size != G * parties * self.n:
            raise ValueError("a must be int32 [G][parties][n]")
        ea, eb = np.empty((G, parties, self.N), np.int32), np.empty(G, np.int32)
        self._chk(lib().mktfhe_ccs_blind_rotate_batch(self.h, parties, int(np.int32(mu)), G, _p(a), _p(b), _p(ea), _p(eb)))
        return ea, eb

    def mk_keyswitch_batch(self, ext_a, ext_b):
        """CCS key switch: ext_a int32 [G][k][N] (one mask per party), ext_b int32 [G]."""
        ext_a, ext_b = _c(ext_a, np.int32), _c(ext_b, np.int32).reshape(-1)
        G = ext_b.size
        if ext_a.size != G * self.k * self.N:
            raise ValueError("ext_a must be int32 [G][k][N]")
        oa, ob = np.empty((G, self.k, self.n), np.int32), np.empty(G, np.int32)
        self._chk(lib().mktfhe_mk_keyswitch_batch(self.h, G, _p(ext_a), _p(ext_b), _p(oa), _p(ob)))
        return oa, ob

    def blind_rotate_batch(self, mu, a, b, want_acc=False):
        a, b = self._ab(a, b)
        G = b.size
        ext = np.empty((G, self.N + 1), np.int32)
        acc = np.empty((G, 2, self.N), np.int64) if want_acc else None
        self._chk(lib().mktfhe_blind_rotate_batch(self.h, int(mu), G, _p(a), _p(b), _p(ext), _p(acc)))
        return ext, acc

    def keyswitch_batch(self, ext):
        ext = _c(ext, np.int32)
        G = ext.size // (self.N + 1)
        if ext.size != G * (self.N + 1):
            raise ValueError("ext must be int32 [G][N+1]")
        oa, ob = np.empty((G, self.k, self.n), np.int32), np.empty(G, np.int32)
        self._chk(lib().mktfhe_keyswitch_batch(self.h, G, _p(ext), _p(oa), _p(ob)))
        return oa, ob

    def negacyclic_mul_batch(self, a, b):
        a, b = _c(a, np.int64), _c(b, np.int64)
        G = a.size // self.N
        if a.size != G * self.N or b.size != a.size:
            raise ValueError("operands must be int64 [G][N]")
        out = np.empty((G, self.N), np.int64)
        self._chk(lib().mktfhe_negacyclic_mul_batch(self.h, G, _p(a), _p(b), _p(out)))
        return out

    # -- introspection
    def launch_count(self):
        return int(lib().mktfhe_launch_count(self.h))

    def last_kernel_ms(self):
        br, ks = C.c_float(), C.c_float()
        self._chk(lib().mktfhe_last_kernel_ms(self.h, C.byref(br), C.byref(ks)))
        return br.value, ks.value

    @staticmethod
    def build_id():
        return lib().mktfhe_build_id().decode()

    def algorithmic_bytes(self):
        b, k = C.c_double(), C.c_double()
        self._chk(lib().mktfhe_algorithmic_bytes(self.h, C.byref(b), C.byref(k)))
        return b.value, k.value


def pin_host(arr):
    """Page-lock a numpy array in place (mktfhe_pin_host); returns the array.  Unpin with unpin_host before freeing it."""
    rc = lib().mktfhe_pin_host(_p(arr), arr.nbytes)
    if rc:
        raise MktfheError(rc, lib().mktfhe_last_error(None).decode())
    return arr


def unpin_host(arr):
    rc = lib().mktfhe_unpin_host(_p(arr))
    if rc:
        raise MktfheError(rc, lib().mktfhe_last_error(None).decode())
