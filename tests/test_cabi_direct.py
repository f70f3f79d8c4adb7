"""The C ABI driven from a plain C program (tests/cabi/cabi_direct.c), i.e. the binding a cgo / ccall / JNI host makes: no Python,
no torch on the path.  Without a GPU the library must refuse loudly (no CPU fallback); on a B200 all five gates are bit-exact
against the oracle and the error codes behave."""
import os
import subprocess

import pytest

from conftest import ROOT


def _build(tmp_path, oracle):
    import torus_fhe_b200 as T
    T._cabi.build(verbose=False)
    exe = str(tmp_path / "cabi_direct")
    lib_dir, ora_dir = os.path.join(ROOT, "torus-fhe_b200"), os.path.join(ROOT, "oracle")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", ora_dir, os.path.join(ROOT, "tests", "cabi", "cabi_direct.c"),
                           "-o", exe, os.path.join(lib_dir, "libmktfhe_b200.so"), os.path.join(ora_dir, "liboracle_mk3gen.so"),
                           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{ora_dir}", "-lm"])
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_plain_c_host_links_and_fails_loudly_without_gpu(tmp_path, oracle):
    exe = _build(tmp_path, oracle)
    if _has_gpu():
        pytest.skip("GPU present: covered by the gpu-marked test")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 1 and "no CUDA device" in out.stderr and "no CPU fallback" in out.stderr, out.stderr


@pytest.mark.gpu
def test_plain_c_host_all_gates_bit_exact(tmp_path, oracle):
    exe = _build(tmp_path, oracle)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "cabi_direct: OK" in out.stdout and "cabi_direct multi: OK" in out.stdout, out.stdout + out.stderr
