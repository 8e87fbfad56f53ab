"""rq_sae: ResidualQuantizedSAE (sae/residual_quantized.py) on libqsae_b200.so.

forward(x) -> (all_latent_groups: list[n_bits] of 0-d tensors, all_reconstruction_levels: list[n_bits] of [B, D])

A cascade of n_bits one-bit QuantizedMatryoshkaSAE stages of widths [1, 1, 2, 4, ...] scaled to
hidden_dim (:22-35); stage t encodes and reconstructs the residual r_t, and r_{t+1} = 2 (r_t - recon_t)
(:59-67). Only stage 0 carries the decoder bias in its forward (:46). The stages are serial by
construction; each one is the fused tcgen05 threshold encoder + sparse level decoder of q_sae
(n_levels = 1), and the residual step is one elementwise kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from .base import SparseAutoencoder, require_cuda_input
from .quantized_matryoshka import QuantizedMatryoshkaSAE, nested_sizes


class ResidualQuantizedSAE(SparseAutoencoder):
    def __init__(self, input_dim, hidden_dim, top_k, abs_range=4, n_bits=8):
        super().__init__(input_dim, hidden_dim)
        self.n_bits = n_bits
        self.abs_range = abs_range
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.top_k = top_k
        self.sae_hidden_dims = nested_sizes(hidden_dim, n_bits)
        self.saes = nn.ModuleList(
            QuantizedMatryoshkaSAE(input_dim=input_dim, hidden_dim=self.sae_hidden_dims[i], top_k=top_k,
                                   abs_range=abs_range, n_bits=1, allow_bias=(i == 0))
            for i in range(n_bits))

    @property
    def exact(self):
        return all(s.exact for s in self.saes)

    @exact.setter
    def exact(self, value):
        for s in self.saes:
            s.exact = bool(value)

    def _weights_key(self):
        return tuple(s._weights_key() for s in self.saes)

    def forward(self, x):
        x = require_cuda_input(x, self)
        key = self._weights_key()
        regime = self._regime[1] if (getattr(self, "_regime", None) is not None and self._regime[0] == key) else None
        capturing = torch.cuda.is_current_stream_capturing()
        pending = getattr(self, "_pending", None)
        if pending is not None and not capturing and pending[1].query():
            self._pending = None
            if int(pending[0][0]) != 0:
                import warnings

                warnings.warn("rq_sae: the previous forward overflowed a stage's sparse survivor lists (its outputs were NaN); "
                              "switching this weight version to the stage-by-stage path with the dense fallback")
                self._regime = (key, "dense")
                regime = "dense"
        if all(s.dense_mode == "auto" for s in self.saes) and regime != "dense":
            # every stage on the sparse path; no host synchronisation in the steady state: the first forward of a
            # weight version reads the stages' overflow flags (one sync) and records the regime, later forwards copy
            # the flags to pinned memory behind the forward and look at them at the start of the next one (an overflowing
            # stage poisons its outputs and the residual it hands on with NaN, so nothing wrong is ever returned quietly)
            residual = x
            groups, levels, flags = [], [], []
            B = x.shape[0]
            for sae in self.saes:
                result, counts, overflow, residual = sae._forward_sparse(residual, want_residual=True)
                groups.append(counts[-1].to(torch.float32) / float(max(B, 1)))
                levels.append(result[-1])
                flags.append(overflow)
            any_flag = torch.cat(flags).sum().to(torch.int32).reshape(1)
            self.last_overflow = any_flag
            ok = True
            if regime is None and not capturing:
                ok = int(any_flag.item()) == 0
                self._regime = (key, "sparse" if ok else "dense")
            elif not capturing:
                host = torch.empty((1,), dtype=torch.int32).pin_memory() if pending is None else pending[0]
                host.copy_(any_flag, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                self._pending = (host, ev)
            if ok:
                for sae in self.saes:
                    sae.last_path = "sparse"
                return groups, levels
        residual = x
        all_latent_groups, all_reconstruction_levels = [], []
        for sae in self.saes:
            latent_group, reconstructions = sae(residual)
            all_latent_groups.append(latent_group[-1])
            all_reconstruction_levels.append(reconstructions[-1])
            residual = _lib.residual_update(residual, reconstructions[-1].contiguous())
        return all_latent_groups, all_reconstruction_levels
