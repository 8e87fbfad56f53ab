"""Sparse-native analysis consumers: the functions of the reference's scripts/analysis/dynamic_analysis.py
with the same names, arguments and return structure, computed from the sparse active sets the B200 forward
already produces.

The reference (dynamic_analysis.py) re-densifies everything: `_activation_mask` runs a second forward and
returns a [B, H] boolean matrix on the CPU (:30-73), `compute_activation_stats` / `analyze_dataset` form the
co-activation counts as a dense [H, B] x [B, H] product per batch (:344-345, :391) and ship a 4 GB matrix to the
host chunk by chunk. Here one forward yields (values, indices) [B, k] (b_sae, baseline) or per-row active lists
(q_sae, rq_sae); activation counts, the co-activation matrix and the squared-error sums are accumulated by
libqsae_b200 kernels into HBM-resident buffers (qsae_activation_counts, qsae_coactivation,
qsae_sq_error_accumulate) and leave the device once, at the end.

`sae` is a quantizedsae_b200.inference.SAEWrapper (or any object with `.model`, `.to`, `.eval`); `loader` is any
iterable of [B, D] tensors or 1-tuples of them. CUDA only -- there is no CPU path.
"""
from __future__ import annotations

from typing import Any, Iterable, Optional

import torch

from . import _lib
from .sae.baseline import BaselineSparseAutoencoder
from .sae.binary import BinarySAE
from .sae.quantized_matryoshka import QuantizedMatryoshkaSAE
from .sae.residual_quantized import ResidualQuantizedSAE
from .sparse import SparseLatents


def _batch_tensor(batch: Any) -> torch.Tensor:
    if isinstance(batch, (list, tuple)):
        batch = batch[0]
    return batch


def _hidden_dim(sae) -> int:
    """dynamic_analysis.py:17-27"""
    model = sae.model
    if hasattr(model, "hidden_dim"):
        return int(model.hidden_dim)
    if isinstance(model, BaselineSparseAutoencoder):
        return int(model.decoder.weight.shape[1])
    if isinstance(model, ResidualQuantizedSAE):
        return int(sum(s.hidden_dim for s in model.saes))
    raise ValueError(f"Unable to determine hidden_dim for model type {type(model)}")


def forward_with_active(sae, x: torch.Tensor) -> dict:
    """One forward of the wrapped model -> dict(reconstruction, reconstruction_levels | None,
    active_idx [B, cap] int32 (-1 = empty), active_vals [B, cap] | None, level_counts | None).

    Activity follows `_activation_mask` (:30-73): b_sae / baseline: top-k latent > 0; q_sae: sigmoid(z) > 0.5;
    rq_sae: the stages' activities concatenated along the latent axis, residual updated as in forward."""
    model = sae.model
    x = x.contiguous().float()
    with torch.no_grad():
        if isinstance(model, (BinarySAE, BaselineSparseAutoencoder)):
            keep = model.return_dense
            model.return_dense = False
            try:
                out = model(x)
            finally:
                model.return_dense = keep
            latents: SparseLatents = out[0]
            return {"reconstruction": out[1], "reconstruction_levels": None, "active_idx": latents.indices,
                    "active_vals": latents.values, "level_counts": None}
        if isinstance(model, QuantizedMatryoshkaSAE):
            r = model.forward_active(x)
            return {"reconstruction": r["reconstruction_levels"][-1], "reconstruction_levels": r["reconstruction_levels"],
                    "active_idx": r["active_idx"], "active_vals": None, "level_counts": r["level_counts"]}
        if isinstance(model, ResidualQuantizedSAE):
            residual, start = x, 0
            levels, lists, counts = [], [], []
            for sub, size in zip(model.saes, model.sae_hidden_dims):
                r = sub.forward_active(residual)
                recon = r["reconstruction_levels"][-1]
                a = r["active_idx"]
                lists.append(torch.where(a >= 0, a + start, a))
                counts.append(r["level_counts"][-1])
                levels.append(recon)
                residual = _lib.residual_update(residual, recon.contiguous())
                start += size
            return {"reconstruction": levels[-1], "reconstruction_levels": levels,
                    "active_idx": torch.cat(lists, dim=1).contiguous(), "active_vals": None,
                    "level_counts": torch.stack(counts)}
    raise TypeError(f"Unsupported SAE model type: {type(model)}")


def _activation_mask(sae, x: torch.Tensor) -> torch.Tensor:
    """Boolean [batch, hidden_dim] mask on the CPU, as the reference returns it (:30-73). Dense by
    definition -- the accumulating functions below never build it."""
    r = forward_with_active(sae, x.to(_device_of(sae)))
    idx, vals = r["active_idx"], r["active_vals"]
    H = _hidden_dim(sae)
    active = idx >= 0 if vals is None else (idx >= 0) & (vals > 0)
    mask = torch.zeros((idx.shape[0], H + 1), dtype=torch.bool, device=idx.device)
    mask.scatter_(1, torch.where(active, idx, torch.full_like(idx, H)).long(), True)
    return mask[:, :H].cpu()


def _device_of(sae) -> torch.device:
    return next(sae.model.parameters()).device


def _prepare(sae, device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("quantizedsae_b200.analysis runs on CUDA only (no CPU fallback)")
    sae.to(dev).eval()
    return dev


def compute_reconstruction_error(sae, loader: Iterable, device: str = "cuda") -> float:
    """mean (recon - x)^2 over all tokens and dimensions (:76-100)."""
    dev = _prepare(sae, device)
    acc = torch.zeros((), dtype=torch.float64, device=dev)
    n = 0
    for batch in loader:
        x = _batch_tensor(batch).to(dev).contiguous().float()
        recon = forward_with_active(sae, x)["reconstruction"]
        _lib.sq_error_accumulate(recon.contiguous(), x, acc)
        n += x.numel()
    return float(acc.item()) / n


def compute_reconstruction_error_by_level(sae, loader: Iterable, device: str = "cuda") -> torch.Tensor:
    """Per-level MSE (:103-165): q_sae levels vs x; rq_sae stage t vs the residual it was given."""
    dev = _prepare(sae, device)
    model = sae.model
    if not isinstance(model, (QuantizedMatryoshkaSAE, ResidualQuantizedSAE)):
        return torch.tensor([compute_reconstruction_error(sae, loader, device=device)], dtype=torch.float64)
    sums: Optional[torch.Tensor] = None
    n = 0
    for batch in loader:
        x = _batch_tensor(batch).to(dev).contiguous().float()
        levels = forward_with_active(sae, x)["reconstruction_levels"]
        if sums is None:
            sums = torch.zeros(len(levels), dtype=torch.float64, device=dev)
        target = x
        for i, recon in enumerate(levels):
            recon = recon.contiguous()
            _lib.sq_error_accumulate(recon, target, sums[i])
            if isinstance(model, ResidualQuantizedSAE):
                target = _lib.residual_update(target, recon)
        n += x.numel()
    return (sums / n).cpu()


def compute_l0_by_level(sae, loader: Iterable, device: str = "cuda") -> torch.Tensor:
    """Average number of active latents per token and level (:168-252)."""
    dev = _prepare(sae, device)
    total: Optional[torch.Tensor] = None
    n_tokens = 0
    for batch in loader:
        x = _batch_tensor(batch).to(dev)
        r = forward_with_active(sae, x)
        if r["level_counts"] is not None:
            c = r["level_counts"].to(torch.float64)
        else:
            idx, vals = r["active_idx"], r["active_vals"]
            c = ((idx >= 0) & (vals > 0)).sum().to(torch.float64).reshape(1)
        total = c.clone() if total is None else total + c
        n_tokens += x.shape[0]
    return (total / max(n_tokens, 1.0)).cpu()


class _StatsAccumulator:
    """activation counts [H] int64, co-activation [H, H] int32 (HBM resident), (feature, token) pairs."""

    def __init__(self, H: int, dev: torch.device, token_ids: Optional[torch.Tensor], tokens_per_context: int):
        self.H, self.dev = H, dev
        self.counts = torch.zeros(H, dtype=torch.int64, device=dev)
        self.cooc = torch.zeros((H, H), dtype=torch.int32, device=dev)
        self.token_ids = None if token_ids is None else token_ids.to(dev)
        self.tpc = tokens_per_context
        self.pairs = []
        self.global_index = 0

    def add(self, idx: torch.Tensor, vals: Optional[torch.Tensor]) -> None:
        idx = idx.contiguous()
        vals = None if vals is None else vals.contiguous()
        _lib.activation_counts(idx, vals, self.counts)
        _lib.coactivation(idx, vals, self.cooc)
        B = idx.shape[0]
        if self.token_ids is not None:
            active = (idx >= 0) if vals is None else (idx >= 0) & (vals > 0)
            rows, slots = active.nonzero(as_tuple=True)
            feat = idx[rows, slots].long()
            self.pairs.append(feat * (1 << 40) + (rows + self.global_index))      # sort key: (feature, global token)
        self.global_index += B

    def tokens_per_feature(self) -> list:
        out = [[] for _ in range(self.H)]
        if self.token_ids is None or not self.pairs:
            return out
        keys = torch.sort(torch.cat(self.pairs)).values
        feat = keys >> 40
        gidx = keys & ((1 << 40) - 1)
        toks = self.token_ids[torch.div(gidx, self.tpc, rounding_mode="floor"), gidx % self.tpc]
        per = torch.bincount(feat, minlength=self.H).cpu().tolist()
        toks = toks.cpu().tolist()
        pos = 0
        for f, c in enumerate(per):
            if c:
                out[f] = [int(t) for t in toks[pos:pos + c]]
                pos += c
        return out


def compute_activation_stats(sae, loader: Iterable, *, token_ids: torch.Tensor, tokens_per_context: int,
                             device: str = "cuda") -> dict:
    """activation_counts [H], coactivation [H, H] = A^T A, tokens_per_feature (:255-314); CPU tensors / lists
    like the reference's."""
    dev = _prepare(sae, device)
    acc = _StatsAccumulator(_hidden_dim(sae), dev, token_ids, tokens_per_context)
    for batch in loader:
        r = forward_with_active(sae, _batch_tensor(batch).to(dev))
        acc.add(r["active_idx"], r["active_vals"])
    return {"activation_counts": acc.counts.cpu(), "coactivation": acc.cooc.cpu(),
            "tokens_per_feature": acc.tokens_per_feature()}


def analyze_dataset(sae, loader: Iterable, *, token_ids: torch.Tensor, tokens_per_context: int, device: str) -> dict:
    """One pass: final-reconstruction MSE + activation statistics (:317-440), one forward per batch."""
    dev = _prepare(sae, device)
    acc = _StatsAccumulator(_hidden_dim(sae), dev, token_ids, tokens_per_context)
    sq = torch.zeros((), dtype=torch.float64, device=dev)
    n = 0
    for batch in loader:
        x = _batch_tensor(batch).to(dev).contiguous().float()
        r = forward_with_active(sae, x)
        _lib.sq_error_accumulate(r["reconstruction"].contiguous(), x, sq)
        n += x.numel()
        acc.add(r["active_idx"], r["active_vals"])
    return {"mse_final": float(sq.item()) / max(n, 1), "mse_per_level": None, "l0_per_level": None,
            "activation_counts": acc.counts.cpu(), "coactivation": acc.cooc.cpu(),
            "tokens_per_feature": acc.tokens_per_feature()}
