"""SparseLatents: the [B, k] (values, indices) form of a top-k latent matrix.

The reference returns dense [B, H] float32 latents (latent * mask, sae/binary.py:96-99;
zeros_like + scatter_, sae/baseline.py:38-39). The B200 path never materialises that matrix
unless asked to: `to_dense()` runs the densify kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class SparseLatents:
    values: torch.Tensor   # [B, k] float32, ordered by (value desc, index asc)
    indices: torch.Tensor  # [B, k] int32 (-1 = empty slot)
    shape: tuple           # (B, H)

    def to_dense(self) -> torch.Tensor:
        return _lib.densify(self.values, self.indices, self.shape[1])

    @property
    def device(self):
        return self.values.device

    def sum(self, dim: int = -1) -> torch.Tensor:
        """latent.sum(-1) as the reference's logging uses it (training/trainer.py:189)."""
        if dim not in (-1, 1):
            raise ValueError("SparseLatents.sum only reduces over the latent axis")
        return (self.values * (self.indices >= 0)).sum(-1)

    def l0(self) -> torch.Tensor:
        return ((self.indices >= 0) & (self.values != 0)).sum(-1)
