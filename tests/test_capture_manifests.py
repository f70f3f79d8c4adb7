"""The ncu capture manifests under profiles/ that bench.py quotes in its `roofline` object must belong to the machine code of the library that is
built from this tree: tools/kernel_id.py hashes the SASS of the kernels a manifest covers.  A kernel edit that changes the machine code without a
new capture makes bench.py drop the quoted figures (they become null); this test makes that visible on the CPU, for the headline set."""
import json
import os
import shutil
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_headline_capture_manifest_matches_the_built_library():
    from kernel_id import kernel_id
    import torus_fhe_b200 as T
    T._cabi.build(verbose=False)
    man = json.load(open(os.path.join(ROOT, "profiles", "ncu_capture_2party_fft64.json")))
    assert man["engine"] == "fft64" and man["parties"] == 2
    assert kernel_id(T._cabi.LIB_PATH, man["kernel_id_covers"]) == man["kernel_id"], \
        "blind_rotate_fft_kernel<2, .> changed since profiles/%s was captured: re-run tools/gpu_round.sh and tools/ncu_manifest.py" % os.path.basename(man["summary_file"])
    summ = os.path.join(ROOT, man["summary_file"])
    import hashlib
    assert hashlib.sha256(open(summ, "rb").read()).hexdigest() == man["summary_sha256"]
