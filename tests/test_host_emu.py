"""Compiles tests/host_emu/ntt_emu.cpp, which runs the very per-thread NTT passes the CUDA kernels use
(torus-fhe_b200/csrc/ntt1024.cuh) on the CPU for 32 emulated lanes, against the O(N^2) definition and an
exact schoolbook product."""
import os
import subprocess

from conftest import ROOT


import pytest


@pytest.mark.parametrize("crt_float", [0, 1])
def test_ntt_host_emulation(tmp_path, crt_float):
    exe = str(tmp_path / "ntt_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", f"-DMK_CRT_FLOAT={crt_float}", "-I", os.path.join(ROOT, "torus-fhe_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_emu", "ntt_emu.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ntt_emu: OK" in out.stdout, out.stdout + out.stderr
