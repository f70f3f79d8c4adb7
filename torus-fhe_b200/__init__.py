"""torus-fhe_b200: B200-native engine behind the 3gen multi-key TFHE gate API of
Animesh005/Torus-FHE (3-gen-mk-tfhe/).  Import as `torus_fhe_b200` (the root shim
`torus_fhe_b200.py` maps the importable name onto this directory).

  _cabi.py     ctypes binding of the C ABI (include/mktfhe_b200.h) -> libmktfhe_b200.so
  engine.py    GPU context + key loading/broadcast + batch sharding
  tfhe3gen.py  the reference's exported names (gates, bootstrap, keys, encrypt/decrypt)
  circuits.py  level-batched integer circuits (mk_add_3gen, mk_sub_3gen, ...)
  interchange.py binary key / ciphertext files a Julia host can write (fixtures from the real reference)
  workloads.py the reference's VolumeMatching workload and the encrypted convolution layer on batched circuit instances
  tfhe1.py     single-key TFHE (api.jl, gates.jl, bootstrap.jl) through the same engine with one party (first slice)
  tfhe_ccs.py  the CCS multi-key scheme (mk_gate_nand / mk_bootstrap) composed from batched external products of the same kernels
  csrc/        hand-written sm_100a kernels and the C ABI implementation
"""
from . import _cabi
from ._cabi import MktfheError
from .engine import Engine, shard_bounds, shard_batch
from .tfhe3gen import *  # noqa: F401,F403
from .tfhe3gen import engine_for, negacyclic_mul, attach_engine, RemoteKeys
from .circuits import (gate_level, mk_int_add_3gen_gpu, mk_add_3gen, mk_add_3gen_v2, mk_inv_3gen, mk_sub_3gen, mk_less_3gen, mk_grt_3gen, mk_leq_3gen,
                       mk_geq_3gen, mk_int_add_with_carry_3gen, mk_int_mul_3gen, mk_int_add_3gen, mk_int_mul_lo_3gen, add_mod_3gen)
from .workloads import VolumeMatch, volume_match_plain, enc_conv2d, conv2d_plain, conv2d_gate_count, conv2d_output_shape
from . import interchange
from . import tfhe1
from . import tfhe_ccs
