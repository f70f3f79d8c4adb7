"""World-size-2 CPU test (gloo) of the multi-GPU host logic: rank 0 owns the keys and broadcasts them once, the gate batch is
sharded contiguously, every rank processes only its slice, and the gathered result equals the single-process result.  The
"engine" here is the CPU oracle standing in for a GPU (the real path differs only in the buffers being device memory and the
backend being NCCL): what is under test is sharding, key hand-off and result assembly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torus_fhe_b200 as T
    from torus_fhe_b200.engine import broadcast_host_keys
    from oracle import mk_oracle as O
    prm = dict(n=16, N=64, k=2, l=2, bgbit=7, t=3, basebit=3, sigma_lwe=2.0 ** -18, sigma_gsw=2.0 ** -40, sigma_ks=2.0 ** -18)
    G = 11                                     # ragged: 6 + 5
    bits = np.random.default_rng(1).integers(0, 2, (2, G)).astype(np.uint8)
    if rank == 0:
        ks = O.KeySet(prm, seed=3, nthreads=2)
        bsk, ksk = [ks.bsk[i] for i in range(2)], [ks.ksk[i] for i in range(2)]
        x, y = ks.encrypt(bits[0], 10), ks.encrypt(bits[1], 11)
        full = ks.gate_batch(O.EXACT_NTT, O.GATE_NAND, x, y, nthreads=2)
        assert np.array_equal(ks.decrypt(*full), ~(bits[0].astype(bool) & bits[1].astype(bool)))
        payload = [torch.from_numpy(np.ascontiguousarray(t)) for t in (x[0], x[1], y[0], y[1])]
    else:
        bsk = ksk = None
        payload = [torch.empty((G, 2, 16), dtype=torch.int32), torch.empty(G, dtype=torch.int32),
                   torch.empty((G, 2, 16), dtype=torch.int32), torch.empty(G, dtype=torch.int32)]
    # one-time key broadcast (the only collective of the path)
    bsk, ksk = broadcast_host_keys(bsk, ksk, (2, (16, 4, 2, 64), (64, 3, 7, 17)))
    for t in payload:
        dist.broadcast(t, src=0)                # test scaffolding: every rank needs the inputs it will slice
    xa, xb, ya, yb = [t.numpy() for t in payload]
    lo, hi = T.shard_bounds(G, world, rank)
    sl = T.shard_batch([xa, xb, ya, yb], world, rank)
    assert sl[0].shape[0] == hi - lo
    local = O.KeySet(prm, raw_bsk=np.stack(bsk), raw_ksk=np.stack(ksk))      # replica built from the received key bytes
    oa, ob = local.gate_batch(O.EXACT_NTT, O.GATE_NAND, (sl[0], sl[1]), (sl[2], sl[3]), nthreads=1)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, oa=oa, ob=ob)
    dist.barrier()
    if rank == 0:
        ga, gb = np.empty_like(full[0]), np.empty_like(full[1])
        for r in range(world):
            d = np.load(os.path.join(out_dir, f"rank{r}.npz"))
            ga[d["lo"]:d["hi"]], gb[d["lo"]:d["hi"]] = d["oa"], d["ob"]
        assert np.array_equal(ga, full[0]) and np.array_equal(gb, full[1])
    dist.destroy_process_group()


def test_two_rank_sharded_gate_batch(tmp_path):
    from oracle import mk_oracle as O
    O.build()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def _conv_worker(rank, world, port, out_dir):
    """Encrypted convolution layer sharded over two ranks (BASELINE configs[4] host logic): rank 0 makes keys and ciphertexts and
    broadcasts them, every rank evaluates only its slice of the outputs (no exchange), rank 0 assembles and decrypts.  The gate
    evaluator is the CPU oracle on a toy ring standing in for the GPU engine."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torus_fhe_b200 as T
    from torus_fhe_b200 import circuits
    from oracle import mk_oracle as O
    prm = dict(n=16, N=64, k=2, l=3, bgbit=7, t=5, basebit=3, sigma_lwe=2.0 ** -22, sigma_gsw=2.0 ** -45, sigma_ks=2.0 ** -22)
    W, H = 3, 4
    payload = [None]
    if rank == 0:
        ks = O.KeySet(prm, seed=5, nthreads=2)
        r = np.random.default_rng(2)
        inp, ker = r.integers(-4, 4, (H, H)), r.integers(-4, 4, (1, 3, 3))

        def enc_int(v, seed):
            v = np.asarray(v)
            out = []
            for q in range(W):
                a, b = ks.encrypt(((v.reshape(-1) >> q) & 1).astype(np.uint8), seed + q)
                out.append(T.MKLweSample(None, a.reshape(v.shape + a.shape[-2:]), b.reshape(v.shape)))
            return out
        za, zb = ks.encrypt(np.zeros(1, np.uint8), 99)
        payload = [(np.stack([ks.bsk[i] for i in range(2)]), np.stack([ks.ksk[i] for i in range(2)]), inp, ker, enc_int(inp, 100), enc_int(ker, 200),
                    T.MKLweSample(None, za[0], zb[0]))]
    dist.broadcast_object_list(payload, src=0)
    bsk, ksk, inp, ker, cin, cker, zero = payload[0]
    local = ks if rank == 0 else O.KeySet(prm, raw_bsk=bsk, raw_ksk=ksk)
    gid = {"nand": O.GATE_NAND, "or": O.GATE_OR, "and": O.GATE_AND, "xor": O.GATE_XOR}

    def oracle_level(bk, ks_, jobs):
        outs = []
        for kind, x, y in jobs:
            shp = np.broadcast_shapes(x.b.shape, y.b.shape)
            xa, ya = np.broadcast_to(x.a, shp + x.a.shape[-2:]).reshape(-1, 2, 16), np.broadcast_to(y.a, shp + y.a.shape[-2:]).reshape(-1, 2, 16)
            xb, yb = np.broadcast_to(x.b, shp).reshape(-1), np.broadcast_to(y.b, shp).reshape(-1)
            oa, ob = local.gate_batch(O.EXACT_NTT, gid[kind], (np.ascontiguousarray(xa), np.ascontiguousarray(xb)),
                                      (np.ascontiguousarray(ya), np.ascontiguousarray(yb)), nthreads=2)
            outs.append(T.MKLweSample(None, oa.reshape(shp + (2, 16)), ob.reshape(shp)))
        return outs
    circuits.gate_level = oracle_level
    lo, hi, bits = T.enc_conv2d(None, None, cin, zero, cker, 1, 0, W, shard=(world, rank))
    assert (lo, hi) == T.shard_bounds(4, world, rank)
    np.savez(os.path.join(out_dir, f"conv{rank}.npz"), lo=lo, hi=hi, a=np.stack([b.a for b in bits]), b=np.stack([b.b for b in bits]))
    dist.barrier()
    if rank == 0:
        got = np.zeros(4, np.int64)
        for r_ in range(world):
            d = np.load(os.path.join(out_dir, f"conv{r_}.npz"))
            dec = [ks.decrypt(np.ascontiguousarray(d["a"][q]), np.ascontiguousarray(d["b"][q])) for q in range(W)]
            v = sum(dec[q].astype(np.int64) << q for q in range(W))
            got[d["lo"]:d["hi"]] = np.where(v >= 1 << (W - 1), v - (1 << W), v)
        assert np.array_equal(got.reshape(1, 2, 2), T.conv2d_plain(inp, ker, 1, 0, W)), (got, T.conv2d_plain(inp, ker, 1, 0, W))
    dist.destroy_process_group()


def test_two_rank_sharded_conv_layer(tmp_path):
    from oracle import mk_oracle as O
    O.build()
    port = 29950 + os.getpid() % 300
    mp.spawn(_conv_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
