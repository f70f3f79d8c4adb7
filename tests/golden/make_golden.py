"""Generates tests/golden/nand_2party.{json,npz} from the CPU oracle at fixed seeds.

The reference has no golden vectors for the 3gen path and cannot run here (Julia), so these fixtures
pin the ORACLE (and through it the GPU path) against silent change; they are not reference outputs.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mk_oracle as O  # noqa: E402

KEYGEN_SEED = 0xB20000A1
X_SEED, Y_SEED = 0xB20000D1, 0xB20000D2

O.build()
ks = O.KeySet(O.PARAMS_2PARTY, seed=KEYGEN_SEED, nthreads=os.cpu_count() or 8)
xs = np.array([0, 0, 1, 1, 1, 0, 1, 0], np.uint8)
ys = np.array([0, 1, 0, 1, 1, 1, 0, 0], np.uint8)
x, y = ks.encrypt(xs, X_SEED), ks.encrypt(ys, Y_SEED)
ta = (-x[0].astype(np.int64) - y[0]).astype(np.int32)
tb = ((1 << 29) - x[1].astype(np.int64) - y[1]).astype(np.int32)
ext = np.empty((xs.size, 1025), np.int32)
acc = np.empty((xs.size, 2, 1024), np.int64)
for g in range(xs.size):
    ea, eb, ac, _ = ks.bootstrap_wo_keyswitch(O.EXACT_NTT, 1 << 61, ta[g], tb[g], want_acc=True)
    ext[g, :1024], ext[g, 1024], acc[g] = ea, eb, ac
oa, ob = ks.gate_batch(O.EXACT_NTT, O.GATE_NAND, x, y)
here = os.path.dirname(os.path.abspath(__file__))
np.savez_compressed(os.path.join(here, "nand_2party.npz"), x_bits=xs, y_bits=ys, xa=x[0], xb=x[1], ya=y[0], yb=y[1],
                    ext=ext, acc_sha256=np.array([hashlib.sha256(a.tobytes()).hexdigest() for a in acc]), out_a=oa, out_b=ob)
with open(os.path.join(here, "nand_2party.json"), "w") as f:
    json.dump({"params": "mktfhe_parameters_2party_3gen (mk_api.jl:32-38)", "keygen_seed": KEYGEN_SEED, "x_seed": X_SEED, "y_seed": Y_SEED,
               "bsk_sha256": hashlib.sha256(ks.bsk.tobytes()).hexdigest(), "ksk_sha256": hashlib.sha256(ks.ksk.tobytes()).hexdigest(),
               "backend": "EXACT_NTT (bit-identical to EXACT_SCHOOLBOOK, tests/test_oracle.py)"}, f, indent=1)
print("wrote golden fixtures; decrypt:", ks.decrypt(oa, ob), "expected", ~(xs.astype(bool) & ys.astype(bool)))
