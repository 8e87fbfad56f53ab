"""Shared pieces of the dictionary-sharded tests: seeded inputs and a numpy `ops` backend (built on the
oracle) that lets the collective choreography of quantizedsae_b200.sharded run on CPU under gloo."""
from __future__ import annotations

import numpy as np
import torch

from oracle import qsae_oracle as O
from tests.golden import cases


def sharded_case(D=64, H=2048, B=37, n_bits=4, seed=123):
    cfg = dict(D=D, H=H, n_bits=n_bits, gamma=4.0, B=B, polar=True, bf16=True, seed=seed)
    return cfg, cases.bsae_inputs(cfg)


def full_state_dict(inp):
    return {"encoder.0.weight": torch.from_numpy(inp["We"]), "encoder.0.bias": torch.from_numpy(inp["be"]),
            "decoder.weight": torch.from_numpy(inp["logits"]), "decoder.bias": torch.from_numpy(inp["bd"])}


class NumpyShardOps:
    """CPU restatement of CudaShardOps' four steps (test infrastructure only)."""

    def __init__(self, module):
        self.m = module

    def local_candidates(self, x, k_local):
        lin = self.m.encoder[0]
        z = O.encode_pre(x.numpy(), lin.weight.detach().numpy(), lin.bias.detach().numpy())
        vals, idx = O.topk_rows(z, k_local)
        out = np.empty(vals.shape + (2,), dtype=np.int32)
        out[..., 0] = vals.view(np.int32)
        out[..., 1] = idx
        return torch.from_numpy(out)

    def merge(self, cand_all, shard_latents, k, truncated=False):
        c = cand_all.numpy()
        G, B, kin, _ = c.shape
        vals = np.ascontiguousarray(c[..., 0]).view(np.float32)                       # [G, B, kin]
        gidx = c[..., 1].astype(np.int64) + (np.arange(G, dtype=np.int64) * shard_latents)[:, None, None]
        v = vals.transpose(1, 0, 2).reshape(B, G * kin)
        i = gidx.transpose(1, 0, 2).reshape(B, G * kin)
        order = np.lexsort((i, -v.astype(np.float64)), axis=1)[:, :k]
        out = (torch.from_numpy(np.take_along_axis(v, order, 1).astype(np.float32)),
               torch.from_numpy(np.take_along_axis(i, order, 1).astype(np.int32)))
        if not truncated:
            return out
        # incomplete: some shard had all of its kin entries selected (positions g * kin .. (g + 1) * kin - 1 of the row)
        chosen = np.zeros((B, G * kin), dtype=bool)
        np.put_along_axis(chosen, order, True, axis=1)
        used_up = chosen.reshape(B, G, kin).all(axis=2).any()
        return out + (torch.tensor([int(used_up)], dtype=torch.int32),)

    def decode_partial(self, vals, idx, plan, with_bias):
        rows = O.dequant_hard(self.m.decoder_weight.detach().numpy(), self.m.n_bits).astype(np.float32)
        a, b = plan.latent_range()
        i = idx.numpy().astype(np.int64)
        mine = (i >= a) & (i < b)
        v = np.where(mine, vals.numpy(), np.float32(0))
        li = np.where(mine, i - a, 0)
        bias = self.m.decoder_bias.detach().numpy() if with_bias else None
        return torch.from_numpy(O.decode_rows(v, li, rows, self.m.quantization_step, bias))

    def polarize_numerator(self):
        logits = self.m.decoder_weight.detach().numpy()
        return O.polarize_loss(logits, self.m.n_bits) * logits.size
