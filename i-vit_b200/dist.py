"""Multi-GPU plumbing for the frozen INT8 forward.

The hot path shards naturally: images are independent and no operator mixes batch elements, so
the batch is split into contiguous per-rank slices and the forward has NO collective.  The only
communication is a ONE-TIME broadcast of the frozen parameter pack (int8 weights, int32 biases,
dyadic tables; ~86 MB for DeiT-B) from rank 0: a single contiguous byte blob, one
``torch.distributed.broadcast`` (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU
tests).  One process per GPU, launched by torchrun.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .pack import Pack

_ALIGN = 256


def pack_to_blob(pack: Pack):
    """Flatten a pack into (manifest, uint8 ndarray); every array starts 256-byte aligned so that
    device views of the blob satisfy the kernels' 16-byte alignment requirements."""
    entries, off = [], 0
    for k in sorted(pack.arrays):
        a = np.ascontiguousarray(pack.arrays[k])
        entries.append((k, a.dtype.str, tuple(a.shape), off, a.nbytes))
        off += (a.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    blob = np.zeros(max(off, _ALIGN), np.uint8)
    for (k, _, _, o, nb) in entries:
        blob[o:o + nb] = np.frombuffer(np.ascontiguousarray(pack.arrays[k]).tobytes(), np.uint8)
    return {"meta": pack.meta, "entries": entries, "total": int(blob.size)}, blob


def blob_to_pack(manifest, blob_host: np.ndarray, blob_device: torch.Tensor = None) -> Pack:
    arrays = {}
    for (k, dt, shape, o, nb) in manifest["entries"]:
        arrays[k] = np.frombuffer(blob_host[o:o + nb].tobytes(), dtype=np.dtype(dt)).reshape(shape).copy()
    p = Pack(manifest["meta"], arrays)
    if blob_device is not None:
        dev = {}
        for (k, dt, shape, o, nb) in manifest["entries"]:
            tdt = {"|i1": torch.int8, "<i2": torch.int16, "<i4": torch.int32, "<f4": torch.float32, "|u1": torch.uint8}[dt]
            dev[k] = blob_device[o:o + nb].view(tdt).reshape(shape)
        p.device_tensors = dev
    return p


def broadcast_pack(pack, src: int = 0, device=None) -> Pack:
    """One-time broadcast of the frozen parameters from ``src`` to every rank.  ``pack`` is needed
    on ``src`` only.  On CUDA devices the returned pack carries ``device_tensors`` (views of the
    received blob), so the engine does not upload anything again."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return pack
    rank = dist.get_rank()
    device = torch.device(device) if device is not None else torch.device("cpu")
    obj = [None]
    blob = None
    if rank == src:
        manifest, blob = pack_to_blob(pack)
        obj = [manifest]
    dist.broadcast_object_list(obj, src=src)
    manifest = obj[0]
    buf = torch.empty(manifest["total"], dtype=torch.uint8, device=device)
    if rank == src:
        buf.copy_(torch.from_numpy(blob))
    dist.broadcast(buf, src=src)                     # the only collective of the whole job
    host = buf.cpu().numpy() if rank != src else blob
    return blob_to_pack(manifest, host, buf if device.type == "cuda" else None)


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a global batch owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_logits(local: torch.Tensor, n_items: int) -> torch.Tensor:
    """Optional all-gather of per-rank logits (outside any timed region): returns [n_items, classes]
    on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_batch(n_items, r, world) for r in range(world)]
    mx = max(h - l for l, h in sizes)
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:h - l] for o, (l, h) in zip(outs, sizes)], dim=0)
