"""Deterministic synthetic parameters, keyed by parameter NAME (not construction order).

BASELINE.json fixes the workload as random-init weights of the named architecture
(no network for checkpoints).  The reference recipe perturbs
biases ~N(0, 0.1) and LayerNorm gamma ~U(0.5, 1.5) to avoid the degenerate zero-bias /
unit-gamma init.  Seeding per name makes the same tensors appear in the reference's
model object (tests/golden/make_golden.py, this container) and in this package's
graphs on the GPU box, independent of module construction order."""
from __future__ import annotations

import hashlib
import zlib

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


@torch.no_grad()
def synth_parameters(model: torch.nn.Module, seed: int = 0) -> str:
    """Overwrite every parameter of ``model`` in place; returns a sha256 over all of them."""
    h = hashlib.sha256()
    for name, p in sorted(model.named_parameters(), key=lambda kv: kv[0]):
        g = _gen(name, seed)
        shape = tuple(p.shape)
        if name.endswith("bias"):
            v = torch.randn(shape, generator=g) * 0.1
        elif "norm" in name and name.endswith("weight"):
            v = torch.rand(shape, generator=g) + 0.5
        else:  # linear / conv weights, cls_token, pos_embed, relative_position_bias_table
            v = (torch.randn(shape, generator=g) * 0.02).clamp_(-0.04, 0.04)
        p.copy_(v.to(p.dtype))
        h.update(name.encode())
        h.update(v.to(torch.float32).contiguous().numpy().tobytes())
    return h.hexdigest()


def synth_images(batch: int, seed: int = 0, img_size: int = 224) -> torch.Tensor:
    """Synthetic 3 x img x img fp32 batch (BASELINE: random 3x224x224), CPU tensor."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1_000_003 * (seed + 1))
    return torch.randn((batch, 3, img_size, img_size), generator=g)
