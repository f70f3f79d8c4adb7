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
