"""Engine: a GPU context with the parties' keys loaded, plus the multi-GPU plumbing.

Multi-GPU model: every GPU holds a full replica of the bootstrapping
and key-switching keys, gate batches are sharded across GPUs, and the only exchange is
the one-time key broadcast.  Two forms:
  * inside the library (`Engine(params, devices=[0, 1, ...])` or `devices="all"` ->
    `mktfhe_create_multi`): one process, one handle; the C ABI broadcasts the keys GPU
    to GPU in `mktfhe_finalize_keys` and shards every host-pointer batch call.  This is
    what a Julia caller of `mk_gate_nand_3gen(bk, ks, xs, ys)` gets;
  * one process per GPU (torchrun): rank 0 loads, `broadcast_keys` moves the key
    buffers with NCCL (`gloo` on CPU in the tests, which exercises the same host logic
    on plain byte buffers), the host shards the batch with `shard_bounds`.
"""
import os

import numpy as np

from . import _cabi


class Engine:
    """Owns a `_cabi.Context`.  Keys come either from the host (`load_keys`) or, on
    non-root ranks, from rank 0 through `broadcast_keys`."""

    def __init__(self, params, device=None, devices=None, flags=0):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.params = params
        self.ctx = _cabi.Context(params.lwe_size, params.rlwe_polynomial_degree, params.max_parties,
                                 params.gsw_decomp_length, params.gsw_log2_base,
                                 params.ks_decomp_length, params.ks_log2_base, device=device, devices=devices, flags=flags)
        self.device = self.ctx.device          # first (or only) GPU
        self.devices = list(self.ctx.devices)
        self.ready = False

    # bsk: per party int64 [n][4][l][N]; ksk: per party int32 [N][t][B-1][n+1], or ("generate", lwe_key, rlwe_key, sigma, seed)
    def load_keys(self, bsk_parts, ksk_parts):
        k = self.params.max_parties
        if len(bsk_parts) != k or len(ksk_parts) != k:
            raise ValueError(f"expected keys of {k} parties, got {len(bsk_parts)} / {len(ksk_parts)}")
        for p in range(k):
            self.ctx.load_bsk(p, bsk_parts[p])
            if isinstance(ksk_parts[p], tuple) and ksk_parts[p][0] == "generate":     # KeyswitchKey.on_device: rows made on the GPU
                self.ctx.generate_ksk(p, *ksk_parts[p][1:])
            else:
                self.ctx.load_ksk(p, ksk_parts[p])
        self.ctx.finalize_keys()
        self.ready = True
        return self

    def key_tensors(self):
        """The device-resident key buffers as torch uint8 tensors (zero-copy views)."""
        import torch
        (bp, bb), (kp, kb) = self.ctx.key_buffers()
        return (_device_view(bp, bb, self.device, torch), _device_view(kp, kb, self.device, torch))

    def broadcast_keys(self, src=0):
        """One-time broadcast of the transformed bsk and the ksk from rank `src` (NCCL)."""
        import torch
        import torch.distributed as dist
        bsk, ksk = self.key_tensors()
        torch.cuda.synchronize(self.device)
        dist.broadcast(bsk, src=src)
        dist.broadcast(ksk, src=src)
        torch.cuda.synchronize(self.device)
        if dist.get_rank() != src:
            self.ctx.mark_keys_received()
            self.ctx.finalize_keys()
        self.ready = True
        return self

    def ctx_on(self, device):
        """The single-device context holding the keys on GPU `device` (for the *_dev calls, whose pointers live on one GPU):
        the context itself for a one-GPU engine, the matching replica of a multi-device one."""
        if len(self.devices) == 1:
            if device != self.device:
                raise ValueError(f"operands live on cuda:{device}, the engine holds its keys on cuda:{self.device}")
            return self.ctx
        if not hasattr(self, "_replicas"):
            self._replicas = {}
        if device not in self._replicas:
            if device not in self.devices:
                raise ValueError(f"operands live on cuda:{device}, the engine spans GPUs {self.devices}")
            self._replicas[device] = self.ctx.replica(self.devices.index(device))
        return self._replicas[device]

    def close(self):
        self.ctx.close()


class _CudaView:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _device_view(ptr, nbytes, device, torch):
    with torch.cuda.device(device):
        return torch.as_tensor(_CudaView(ptr, nbytes), device=f"cuda:{device}")


# ---------------------------------------------------------------------------------
# sharding of a gate batch over ranks (pure host logic; tested with gloo on CPU)
# ---------------------------------------------------------------------------------
def shard_bounds(G, world_size, rank):
    """Contiguous slice [lo, hi) of a batch of G gates owned by `rank`; sizes differ by at most 1."""
    base, rem = divmod(G, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(arrays, world_size, rank):
    G = arrays[0].shape[0]
    lo, hi = shard_bounds(G, world_size, rank)
    return [a[lo:hi] for a in arrays]


def broadcast_host_keys(bsk_parts, ksk_parts, shapes, src=0):
    """Host-side key hand-off for process groups without a GPU (gloo): rank `src` passes
    its numpy keys, the other ranks pass None and receive copies.  `shapes` =
    (k, bsk_shape, ksk_shape) so receivers can allocate."""
    import torch
    import torch.distributed as dist
    k, bshape, kshape = shapes
    rank = dist.get_rank()
    out_b, out_k = [], []
    for p in range(k):
        tb = torch.from_numpy(np.ascontiguousarray(bsk_parts[p], dtype=np.int64)) if rank == src else torch.empty(bshape, dtype=torch.int64)
        tk = torch.from_numpy(np.ascontiguousarray(ksk_parts[p], dtype=np.int32)) if rank == src else torch.empty(kshape, dtype=torch.int32)
        dist.broadcast(tb, src=src)
        dist.broadcast(tk, src=src)
        out_b.append(tb.numpy())
        out_k.append(tk.numpy())
    return out_b, out_k
