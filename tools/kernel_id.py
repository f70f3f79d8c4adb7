#!/usr/bin/env python
"""Identity of the hot kernels of a built libmktfhe_b200.so: sha256 over the SASS text (cuobjdump) of every blind_rotate kernel.
Source edits that leave the machine code unchanged keep the id; bench.py compares it with the id recorded beside each ncu capture
under profiles/ and refuses to quote a capture taken from other machine code.
   python tools/kernel_id.py [path/to/libmktfhe_b200.so]"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hot_kernels(N, l, engine="ntt_rns"):
    """Substrings (of mangled names) selecting the kernels that serve a parameter set: the N = 1024 kernels of gadget length l -- the FP64 FFT
    channel (engine "fft64": both launch shapes are blind_rotate_fft_kernel<l, .>) or the three-prime NTT throughput and latency kernels --
    or the N = 2048 kernels."""
    if N == 2048:
        return ["N4mk2k"]
    if engine == "fft64":
        return [f"N3mkf23blind_rotate_fft_kernelILi{l}E"]
    return [f"N2mk19blind_rotate_kernelILi{l}E", f"N2mk23blind_rotate_lat_kernelILi{l}E"]


def kernel_id(lib=None, match=("blind_rotate",)):
    lib = lib or os.path.join(ROOT, "torus-fhe_b200", "libmktfhe_b200.so")
    try:
        out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, timeout=300).stdout
    except Exception:
        return None
    h, keep, n = hashlib.sha256(), False, 0
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            keep = any(x in m.group(1) for x in ([match] if isinstance(match, str) else match))
            if keep:
                h.update(m.group(1).encode()); n += 1
            continue
        if keep:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
            if m:
                h.update(m.group(1).encode())
    return h.hexdigest()[:16] if n else None


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else None
    print("all blind_rotate kernels:", kernel_id(lib), " N=1024 l=2 fft64:", kernel_id(lib, hot_kernels(1024, 2, "fft64")),
          " N=1024 l=2 ntt_rns:", kernel_id(lib, hot_kernels(1024, 2)), " N=2048:", kernel_id(lib, hot_kernels(2048, 1)))
