"""Base class with the reference's encode/decode/forward contract (sae/base.py:5-29) plus the
weight-preparation cache shared by the B200 modules.

Prepared (packed / bf16 / transposed) copies of parameters are device-resident and keyed on the
parameters' identity, storage pointer and in-place version counter, so `load_state_dict`,
`.to(device)`, optimizer steps and `param.data = ...` all invalidate them.

Limitation: an in-place edit made THROUGH `.data` (`w.data.mul_(m)`, `w.data.copy_(v)`; the reference's own
`STEWeights.init_mask` / `update_mask` do `self.weight.data *= mask`) changes none of the three -- `.data` is a
detached alias with its own version counter -- so the prepared copies would go stale silently. After such an edit
call `module.invalidate()` (every module of this package has it; it also reaches the sub-modules' caches).
Tensors created under `torch.inference_mode()` have no version counter: for them the key is identity + storage
pointer only, and `invalidate()` is the way to announce any in-place change.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _version(t: torch.Tensor) -> int:
    try:
        return t._version
    except RuntimeError:          # inference tensors do not track a version counter
        return -1


def param_key(*tensors: torch.Tensor | None) -> tuple:
    return tuple(None if t is None else (id(t), t.data_ptr(), _version(t), tuple(t.shape), str(t.device))
                 for t in tensors)


def invalidate_prepared(module: nn.Module) -> None:
    """Drop every prepared (packed / bf16 / transposed / sampled) copy held by `module` and its sub-modules; the next
    forward rebuilds them from the current parameter values."""
    for m in module.modules():
        prep = getattr(m, "_prep", None)
        if isinstance(prep, PreparedCache):
            prep.clear()
        for name in ("_regime", "_pending", "_pol_cache"):
            if hasattr(m, name):
                setattr(m, name, None)
        ops = getattr(m, "ops", None)
        if ops is not None and isinstance(getattr(ops, "_prep", None), PreparedCache):
            ops._prep.clear()


class PreparedCache:
    """name -> (key, value); value is rebuilt by `make()` when the key changes."""

    def __init__(self):
        self._slots: dict = {}

    def get(self, name: str, key: tuple, make):
        slot = self._slots.get(name)
        if slot is None or slot[0] != key:
            slot = (key, make())
            self._slots[name] = slot
        return slot[1]

    def clear(self):
        self._slots.clear()


def require_cuda_input(x: torch.Tensor, module: nn.Module) -> torch.Tensor:
    """The product path has no CPU implementation: fail loudly instead of falling back."""
    p = next(module.parameters())
    if not p.is_cuda or not x.is_cuda:
        raise RuntimeError(
            f"{type(module).__name__}.forward runs only on a CUDA (sm_100a) device through "
            "libqsae_b200.so; move the module and its input to the GPU (there is no CPU fallback)")
    if x.dim() != 2:
        raise RuntimeError(f"expected a [batch, input_dim] matrix, got shape {tuple(x.shape)}")
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


class SparseAutoencoder(nn.Module):
    """Same surface as the reference base class: `encode`, `decode`, `forward -> (latent, recon)`;
    unset encoder/decoder raise NotImplementedError (sae/base.py:16-24)."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.encoder = None
        self.decoder = None

    def invalidate(self) -> None:
        """Forget the prepared copies of the weights (see the module docstring: needed after `.data` edits)."""
        invalidate_prepared(self)

    refresh = invalidate

    def encode(self, x):
        if self.encoder is None:
            raise NotImplementedError("Encoder has not been implemented.")
        return self.encoder(x)

    def decode(self, h):
        if self.decoder is None:
            raise NotImplementedError("Decoder has not been implemented.")
        return self.decoder(h)

    def forward(self, x):
        latent = self.encode(x)
        return latent, self.decode(latent)
