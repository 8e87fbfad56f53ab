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

    def forward(self, x):
        x = require_cuda_input(x, self)
        if all(s.dense_mode == "auto" for s in self.saes):
            # every stage on the sparse path, the overflow flags read once at the end (one host sync per forward
            # instead of one per stage); any overflow -> the forward is redone stage by stage with the dense fallback
            residual = x
            groups, levels, flags = [], [], []
            B = x.shape[0]
            for sae in self.saes:
                result, counts, overflow, residual = sae._forward_sparse(residual, want_residual=True)
                groups.append(counts[-1].to(torch.float32) / float(max(B, 1)))
                levels.append(result[-1])
                flags.append(overflow)
            if int(torch.cat(flags).sum().item()) == 0:
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
