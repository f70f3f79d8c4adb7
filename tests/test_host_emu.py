"""Compiles tests/host_emu/ntt_emu.cpp, which runs the very per-thread NTT passes the CUDA kernels use
(torus-fhe_b200/csrc/ntt1024.cuh) on the CPU for 32 emulated lanes, against the O(N^2) definition and an
exact schoolbook product."""
import os
import subprocess
import sys

from conftest import ROOT


import pytest


@pytest.mark.parametrize("crt_float", [0, 1])
def test_ntt_host_emulation(tmp_path, crt_float):
    exe = str(tmp_path / "ntt_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", f"-DMK_CRT_FLOAT={crt_float}", "-I", os.path.join(ROOT, "torus-fhe_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_emu", "ntt_emu.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ntt_emu: OK" in out.stdout, out.stdout + out.stderr


def test_ntt2048_host_emulation(tmp_path):
    """Arithmetic core of the N = 2048 parameter sets (ntt2048.cuh): transform vs definition, four-prime Garner lift,
    exact products of 26-bit digits with 64-bit keys vs schoolbook."""
    exe = str(tmp_path / "ntt2048_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "torus-fhe_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_emu", "ntt2048_emu.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ntt2048_emu: OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_ntt2048_on_gpu(tmp_path):
    """The same core as a warp-level CUDA kernel: exact negacyclic products of degree 2048 on the B200 vs schoolbook."""
    exe = str(tmp_path / "ntt2048_gpu")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", os.path.join(ROOT, "torus-fhe_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_emu", "ntt2048_gpu.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ntt2048_gpu: OK" in out.stdout, out.stdout + out.stderr


def test_f64_channel_prototype(tmp_path):
    """Round-2 candidate: one RNS channel carried in doubles so its butterflies leave the binding IMAD pipe.
    Host emulation of tools/f64_channel/ntt_f64.cuh: same residues as the u32 channel position by position, exact three-prime
    product with the unchanged Garner lift, all intermediates exact integers far below 2^53.  Not linked into the library."""
    exe = str(tmp_path / "f64_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "torus-fhe_b200", "csrc"),
                           "-I", os.path.join(ROOT, "tools", "f64_channel"), os.path.join(ROOT, "tools", "f64_channel", "emu.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "f64_emu: OK" in out.stdout, out.stdout + out.stderr


def test_fft64_host_emulation(tmp_path):
    """Arithmetic core of the FP64 FFT channel (fft64_core.cuh, tables_fft.h): the kernels' per-thread passes for 32 emulated lanes against the
    definition of the transform and against the wrap-around integer external product (extreme operands included); the rounding margin
    of the limb results is printed and bounded."""
    exe = str(tmp_path / "fft64_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "torus-fhe_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_emu", "fft64_emu.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "fft64_emu: OK" in out.stdout, out.stdout + out.stderr


def test_fft64_numpy_prototype():
    """tools/fft_channel/proto.py: the same networks in numpy -- exact on random and extreme operands."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fft_channel", "proto.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "exact on 20 trials" in out.stdout, out.stdout + out.stderr
