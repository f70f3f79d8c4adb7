"""The rule by which mktfhe_create picks the exact FP64 FFT channel (csrc/mktfhe_b200.cu), restated in Python and applied to
every parameter set of the reference the host mirror defines: all N = 1024 sets (Torus64 3gen sets, the single-key sets, the CCS gadget
shapes in Torus32 mode) must be served by it with limb products inside the range a double-precision transform reproduces exactly, and the
N = 2048 sets must not be (their 18..26-bit digits leave no room).  Also pins the limb arithmetic's ranges."""
import math
import re

import os

from conftest import ROOT


def fft_eligible(N, l, bgbit, torus32):
    if N != 1024:
        return False
    return math.log2(2 * l * N) + (bgbit - 1) + (15 if torus32 else 21) <= 40.0


def test_rule_matches_the_library_source():
    src = open(os.path.join(ROOT, "torus-fhe_b200", "csrc", "mktfhe_b200.cu")).read()
    m = re.search(r"c->fft = !big && std::log2\(\(double\)\(2 \* params->l\) \* params->N\) \+ \(params->bgbit - 1\) \+ \(c->t32 \? ([\d.]+) : ([\d.]+)\) <= ([\d.]+);", src)
    assert m and (float(m.group(1)), float(m.group(2)), float(m.group(3))) == (15.0, 21.0, 40.0)


def test_reference_parameter_sets():
    import torus_fhe_b200 as T
    sets3 = {k: getattr(T, f"mktfhe_parameters_{k}party_3gen") for k in (2, 3, 4, 5, 8, 16, 32, 64, 128, 256)}
    for k, p in sets3.items():
        N, l, bg = p.rlwe_polynomial_degree, p.gsw_decomp_length, p.gsw_log2_base
        want = N == 1024
        assert fft_eligible(N, l, bg, False) == want, (k, N, l, bg)
        if want:
            # limb products: 2 l N digit-key terms of at most (Bg / 2) 2^21 each -- far inside the 2^53 of a double, 2^-13 measured rounding distance
            assert 2 * l * N * (1 << (bg - 1)) * (1 << 21) <= 1 << 40
    T1 = T.tfhe1
    for name, t32 in (("tfhe_parameters_128", False), ("tfhe_parameters_80", True)):
        p = getattr(T1, name)()
        assert (p.bs_log2_base > 8) == t32                      # the host mirror's own switch to Torus32 mode (tfhe1.engine_for)
        assert fft_eligible(p.rlwe_polynomial_degree, p.bs_decomp_length, p.bs_log2_base, t32), name
    TC = T.tfhe_ccs
    for p in (TC.mktfhe_parameters_2party, TC.mktfhe_parameters_4party):
        assert fft_eligible(p.rlwe_polynomial_degree, p.bs_decomp_length, p.bs_log2_base, True)
        assert 2 * p.bs_decomp_length * 1024 * (1 << (p.bs_log2_base - 1)) * (1 << 15) <= 1 << 40


def test_limb_ranges():
    """key_limb (fft64_core.cuh): 22 / 21 / 21-bit balanced limbs of a 64-bit word, 16 / 16 (+1) of a 32-bit word."""
    import random
    rnd = random.Random(5)
    words = [0, 1, -1, 2 ** 63 - 1, -2 ** 63, 2 ** 21, 2 ** 21 - 1, -2 ** 21, 2 ** 43 - 1] + [rnd.randrange(-2 ** 63, 2 ** 63) for _ in range(2000)]
    for k in words:
        l0 = ((k + (1 << 21)) & ((1 << 22) - 1)) - (1 << 21)
        k1 = (((k - l0) + 2 ** 63) % 2 ** 64 - 2 ** 63) >> 22
        l1 = ((k1 + (1 << 20)) & ((1 << 21) - 1)) - (1 << 20)
        l2 = (k1 - l1) >> 21
        assert -2 ** 21 <= l0 < 2 ** 21 and -2 ** 20 <= l1 < 2 ** 20 and -2 ** 20 <= l2 <= 2 ** 20
        assert (l0 + (l1 << 22) + (l2 << 43) - k) % 2 ** 64 == 0
    for v in [0, 1, -1, 2 ** 31 - 1, -2 ** 31, 2 ** 15, 2 ** 15 - 1, -2 ** 15] + [rnd.randrange(-2 ** 31, 2 ** 31) for _ in range(2000)]:
        l0 = ((v + (1 << 15)) & 0xFFFF) - (1 << 15)
        l1 = (((v - l0) + 2 ** 31) % 2 ** 32 - 2 ** 31) >> 16
        assert -2 ** 15 <= l0 < 2 ** 15 and -2 ** 15 <= l1 <= 2 ** 15
        assert (l0 + (l1 << 16) - v) % 2 ** 32 == 0
