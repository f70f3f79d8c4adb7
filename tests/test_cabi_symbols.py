"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls: no GPU here)."""
import ctypes
import glob
import os
import re

from conftest import ROOT


def declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        syms |= set(re.findall(r"\b(mktfhe_[a-z0-9_]+)\s*\(", text))
    return syms


def test_header_and_binding_agree():
    import torus_fhe_b200 as T
    assert declared_symbols() == set(T._cabi.EXPORTS)


def test_library_exports_every_declared_symbol():
    import torus_fhe_b200 as T
    if not os.path.exists(T._cabi.LIB_PATH):
        T._cabi.build()          # nvcc cross-compiles sm_100a without a GPU (about a minute)
    L = ctypes.CDLL(T._cabi.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), s
    T._cabi.lib()   # the full argtypes table binds


def test_create_fails_loudly_without_gpu_or_bad_params():
    import torch
    import torus_fhe_b200 as T
    L = T._cabi.lib()
    bad = T._cabi.CParams(520, 512, 2, 2, 7, 3, 3, 0)
    h = ctypes.c_void_p()
    assert L.mktfhe_create(ctypes.byref(bad), 0, ctypes.byref(h)) == T._cabi.EINVAL
    assert b"N=512" in L.mktfhe_last_error(None)
    if not torch.cuda.is_available():
        ok = T._cabi.CParams(520, 1024, 2, 2, 7, 3, 3, 0)
        assert L.mktfhe_create(ctypes.byref(ok), 0, ctypes.byref(h)) == T._cabi.ECUDA   # no CPU fallback
        assert not h.value


def test_oracle_is_not_reachable_from_the_product():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "torus-fhe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "mk_oracle" not in text and "liboracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def _create_rc(L, T, sp_or_tuple):
    import ctypes as C
    v = sp_or_tuple
    if not isinstance(v, tuple):
        v = (v.lwe_size, v.rlwe_polynomial_degree, v.max_parties, v.gsw_decomp_length, v.gsw_log2_base, v.ks_decomp_length, v.ks_log2_base)
    prm = T._cabi.CParams(*v, 0)
    h = C.c_void_p()
    rc = L.mktfhe_create(C.byref(prm), 0, C.byref(h))
    if rc == 0:                 # a GPU box: the set was accepted for real
        L.mktfhe_destroy(h)
    return rc, L.mktfhe_last_error(None).decode()


def test_every_reference_parameter_set_is_classified_as_the_header_says():
    """mk_api.jl:32-322: the 2..256-party sets pass parameter validation (without a GPU the call then stops at the device probe,
    MKTFHE_ECUDA -- never a CPU fallback); the 512-party set (N = 4096) is MKTFHE_EINVAL.  Validation runs before any CUDA call."""
    import torch
    import torus_fhe_b200 as T
    L = T._cabi.lib()
    accepted = T._cabi.OK if torch.cuda.is_available() else T._cabi.ECUDA
    for parties in (2, 3, 4, 5, 8, 16, 32, 64, 128, 256):
        # the big sets would allocate their full keys on a GPU box: validate those through a reduced LWE dimension there
        sp = getattr(T, f"mktfhe_parameters_{parties}party_3gen")
        v = (sp.lwe_size if parties <= 8 or not torch.cuda.is_available() else 8, sp.rlwe_polynomial_degree, min(sp.max_parties, 16)
             if torch.cuda.is_available() else sp.max_parties, sp.gsw_decomp_length, sp.gsw_log2_base, sp.ks_decomp_length, sp.ks_log2_base)
        rc, msg = _create_rc(L, T, v)
        assert rc == accepted, (parties, rc, msg)
    rc, msg = _create_rc(L, T, T.mktfhe_parameters_512party_3gen)
    assert rc == T._cabi.EINVAL and "N=4096" in msg


def test_degenerate_parameters_are_rejected_not_undefined():
    """Fields that feed shifts, table sizes and loop bounds: every out-of-range value is MKTFHE_EINVAL with a message."""
    import torus_fhe_b200 as T
    L = T._cabi.lib()
    good = (520, 1024, 2, 2, 7, 3, 3)
    names = ("n", "N", "k", "l", "bgbit", "t", "basebit")
    bad_values = {"n": (0, -1, 1 << 20), "N": (0, 512, 4096, -1024), "k": (0, -2, 1 << 20), "l": (0, 5, -1), "bgbit": (0, 9, 31, 32, 33, -7),
                  "t": (0, -3, 11), "basebit": (0, 17, 31, 32, -1)}
    for i, name in enumerate(names):
        for bad in bad_values[name]:
            v = list(good)
            v[i] = bad
            rc, msg = _create_rc(L, T, tuple(v))
            assert rc == T._cabi.EINVAL and msg, (name, bad, rc, msg)
    # N = 2048 shapes: l = 3, and a gadget digit that no longer fits a 28-bit residue
    for v in ((590, 2048, 16, 3, 18, 4, 3), (590, 2048, 16, 1, 28, 4, 3)):
        rc, msg = _create_rc(L, T, v)
        assert rc == T._cabi.EINVAL and msg, (v, rc, msg)
    assert L.mktfhe_create(None, 0, None) == T._cabi.EINVAL
