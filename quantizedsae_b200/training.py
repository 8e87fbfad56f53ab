"""Training-side pieces adjacent to the forward, on libqsae_b200.so (csrc/train.cu).

The reference trains with eager autograd over dense [B, H] latents and dense [H, D n_bits] soft-bit tensors
(training/trainer.py:88-173). The pieces of that backward that touch the quantised decoders are sparse in the same
way the forward is, and are provided here on the sparse forward quantities:

b_sae   `bsae_forward(model, x)` (what `BinarySAE.forward` runs when `model.autograd = True`): the reference's
        forward tuple, with `reconstruction` and `polarize_loss` attached to an autograd node, so the trainer's
        own lines -- loss = 0.5 * mse(recon, batch) + polarize_lambda * polarize_loss; loss.backward()
        (training/trainer.py:143-151) -- fill .grad of encoder.0.weight / encoder.0.bias / decoder.weight /
        decoder.bias. Backward of sae/binary.py:24-47 and :91-99: sparse outer products (16-byte vector reductions),
        row dots against the soft dictionary, one streaming pass over the bit logits that applies the sigmoid chain
        rule and adds the polarize gradient. The returned latent carries no gradient (the trainer only logs it).
q_sae   `QuantizedMatryoshkaDecoder.ste_backward(grad_levels)`: what loss.backward() leaves in decoder.weight.grad,
        weight_mirror.grad and bias.grad (STE through the sign, sae/quantized_matryoshka.py:94-124, joint_gradient =
        False), from the active lists of `QuantizedMatryoshkaSAE.forward_active`; `apply_secant_grad()` (:145-190).
        `qsae_trainer_decoder_grads` strings them together under the trainer's loss (training/trainer.py:88-113).
        The encoder side of q_sae's backward is dense (the STE passes a gradient to every latent, :99) and is not
        built: it is two plain dense GEMMs with no quantisation-specific structure.
t_sae   `STEWeights.init_mask / update_mask / mask_grad` (sae/ternary.py:27-90): exact RigL drop / grow by radix
        select on the device.
No CPU fallback: CUDA tensors only.
"""
from __future__ import annotations

import torch

from . import _lib
from .sparse import SparseLatents


def _acc_grad(p: torch.nn.Parameter) -> torch.Tensor:
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


class _BinarySAEFn(torch.autograd.Function):
    """(x, W_enc, b_enc, logits, b_dec) -> (values, indices, reconstruction, polarize_loss)."""

    @staticmethod
    def forward(ctx, x, w_enc, b_enc, logits, b_dec, model):
        dec = model.decoder
        lat = model.encode_topk(x)
        rows = dec._soft_rows()
        H, D = dec.in_features, dec.out_features
        recon = _lib.decode_rows_f32(lat.values, lat.indices, rows, H, D, dec.quantization_step, b_dec.detach())
        pol = dec.polarize_loss().detach().clone()
        ctx.save_for_backward(x, lat.values, lat.indices, logits, w_enc, rows)
        ctx.dims = (H, D, dec.n_bits, float(dec.quantization_step))
        ctx.mark_non_differentiable(lat.values, lat.indices)
        return lat.values, lat.indices, recon, pol

    @staticmethod
    def backward(ctx, _g_vals, _g_idx, g_recon, g_pol):
        x, vals, idx, logits, w_enc, rows = ctx.saved_tensors
        H, D, n_bits, q = ctx.dims
        dev = x.device
        B = x.shape[0]
        g = (torch.zeros((B, D), dtype=torch.float32, device=dev) if g_recon is None
             else g_recon.contiguous().float())
        gp = 0.0 if g_pol is None else g_pol.detach().reshape(()).float().contiguous()
        G = torch.zeros((H, D), dtype=torch.float32, device=dev)
        _lib.rows_scatter_add(vals, idx, g, G, scale=q)                       # d / d int_w
        grad_logits = torch.empty_like(logits, memory_format=torch.contiguous_format)
        _lib.bsae_logit_grad(logits.detach().contiguous(), G, D, n_bits, gp, grad_logits, accumulate=False)
        grad_bd = _lib.column_sum(g)
        grad_vals = _lib.rows_gather_dot(g, rows, idx, scale=q)               # d / d latent at the kept positions
        grad_we = torch.zeros((H, D), dtype=torch.float32, device=dev)
        grad_be = torch.zeros((H,), dtype=torch.float32, device=dev)
        _lib.rows_scatter_add(grad_vals, idx, x.detach().contiguous(), grad_we, scale=1.0, dst_col=grad_be)
        grad_x = None
        if ctx.needs_input_grad[0]:
            grad_x = _lib.decode_rows_f32(grad_vals, idx, w_enc.detach().contiguous(), H, D, 1.0, None)
        return grad_x, grad_we, grad_be, grad_logits, grad_bd, None


def bsae_forward(model, x):
    """BinarySAE.forward with an autograd node behind `reconstruction` and `polarize_loss` (see module docstring).
    Always decodes with the soft (sigmoid) dictionary: that is the function the reference differentiates
    (sae/binary.py:26-38), whatever `decode_mode` says about inference."""
    from .sae.base import require_cuda_input

    x = require_cuda_input(x, model)
    lin, dec = model.encoder[0], model.decoder
    vals, idx, recon, pol = _BinarySAEFn.apply(x, lin.weight, lin.bias, dec.weight, dec.bias, model)
    latents = SparseLatents(vals, idx, (x.shape[0], model.hidden_dim))
    return (latents.to_dense() if model.return_dense else latents), recon, pol


# ---- q_sae --------------------------------------------------------------------------------------------

def qsae_ste_backward(decoder, active_idx: torch.Tensor, grad_levels, batch_size: int | None = None) -> torch.Tensor:
    """Accumulate into decoder.weight.grad / weight_mirror.grad / bias.grad what autograd leaves there for upstream
    gradients grad_levels[i] = d loss / d result[i] (sae/quantized_matryoshka.py:94-124). Returns z2 [H] int32 (the
    per-latent activity counts, :137) and stashes it for apply_secant_grad."""
    H, D = decoder.in_features, decoder.out_features
    dev = decoder.weight.device
    gl = [g.detach().contiguous().float() for g in grad_levels]
    if len(gl) != decoder.n_bits:
        raise ValueError(f"expected {decoder.n_bits} level gradients, got {len(gl)}")
    M = torch.zeros((H, D), dtype=torch.float32, device=dev)
    z2 = torch.zeros((H,), dtype=torch.int32, device=dev)
    starts = decoder._level_starts_host()
    _lib.matryoshka_backward_scatter(active_idx, gl, starts, M, z2)
    _, alpha = decoder._packed()
    _lib.matryoshka_backward_finish(decoder.weight.detach(), decoder.weight_mirror.detach(), M, None, alpha, starts, 0.0, 0,
                                    _acc_grad(decoder.weight), _acc_grad(decoder.weight_mirror))
    if decoder.allow_bias:
        _lib.column_sum(gl[0], 1.0, _acc_grad(decoder.bias))
    decoder._ctx = {"z2": z2, "batch_size": int(batch_size if batch_size is not None else active_idx.shape[0])}
    return z2


def qsae_apply_secant_grad(decoder) -> None:
    """decoder.apply_secant_grad() (sae/quantized_matryoshka.py:145-190): grad -= c m z2 alpha^2 Bsign s'(w)."""
    ctx = decoder._ctx
    if not isinstance(ctx, dict) or ctx.get("z2") is None:
        raise RuntimeError("apply_secant_grad: no training context; call forward_active + ste_backward first "
                           "(the reference stashes it in forward, sae/quantized_matryoshka.py:131-141)")
    _, alpha = decoder._packed()
    c = 1.0 / ctx["batch_size"] / decoder.out_features
    _lib.matryoshka_backward_finish(decoder.weight.detach(), decoder.weight_mirror.detach(), None, ctx["z2"], alpha,
                                    decoder._level_starts_host(), c, decoder.n_bits if decoder.joint_gradient else 0,
                                    _acc_grad(decoder.weight), _acc_grad(decoder.weight_mirror))


def qsae_trainer_decoder_grads(model, x, active_cap: int = 256):
    """One q_sae step of the reference trainer as far as the decoder is concerned (training/trainer.py:88-113):
    forward, recon_loss = sum_i 0.5 * mse(result_i, x), decoder gradients of it, apply_secant_grad.
    -> dict(recon_losses [n_bits] device tensor, latent_groups, reconstruction_levels)."""
    out = model.forward_active(x, active_cap)
    B, D = x.shape
    levels = out["reconstruction_levels"]
    grads = [(r - x) / float(B * D) for r in levels]
    qsae_ste_backward(model.decoder, out["active_idx"], grads, B)
    qsae_apply_secant_grad(model.decoder)
    losses = torch.stack([0.5 * ((r - x) ** 2).mean() for r in levels])
    return {"recon_losses": losses, "latent_groups": out["latent_groups"], "reconstruction_levels": levels}


# ---- t_sae: RigL mask maintenance ---------------------------------------------------------------------

def rigl_init_mask(ste, sparsity: float) -> None:
    """STEWeights.init_mask (sae/ternary.py:27-39)."""
    w = ste.weight.data
    if not w.is_cuda:
        raise RuntimeError("STEWeights.init_mask runs only on CUDA (no CPU fallback)")
    n_inactive = int(w.numel() * sparsity)
    mask = torch.ones_like(w, memory_format=torch.contiguous_format)
    wc = w if w.is_contiguous() else w.contiguous()
    _lib.rigl_init_mask(wc, mask, n_inactive)
    if wc is not w:
        w.copy_(wc)
    ste.mask = mask
    ste.invalidate()


def rigl_update_mask(ste, f_decay: float, sparsity_rate: float = 0.7) -> None:
    """STEWeights.update_mask (sae/ternary.py:54-87)."""
    w = ste.weight.data
    if not w.is_cuda:
        raise RuntimeError("STEWeights.update_mask runs only on CUDA (no CPU fallback)")
    n_drop = int(f_decay * (1 - sparsity_rate) * w.numel())
    n_grow = n_drop
    a_mean = d_mean = None
    if n_grow > 0 and ste.input_activations is not None:
        if ste.output_grad is None:
            raise RuntimeError("update_mask: output_grad is not set (the reference captures it in a backward hook, "
                               "sae/ternary.py:24-25); assign decoder.output_grad = d loss / d reconstruction")
        a = ste.input_activations.detach().contiguous().float()
        d = ste.output_grad.detach().contiguous().float()
        a_mean = _lib.column_sum(a, 1.0 / a.shape[0])
        d_mean = _lib.column_sum(d, 1.0 / d.shape[0])
    mask = ste.mask.contiguous().float()
    wc = w if w.is_contiguous() else w.contiguous()
    _lib.rigl_update_mask(wc, mask, a_mean, d_mean, n_drop, n_grow if a_mean is not None else 0)
    if wc is not w:
        w.copy_(wc)
    ste.mask = mask
    ste.invalidate()


def rigl_mask_grad(ste) -> None:
    """STEWeights.mask_grad (sae/ternary.py:89-90)."""
    g = ste.weight.grad
    if g is None:
        raise AttributeError("mask_grad: weight.grad is None")
    gc = g.data if g.is_contiguous() else g.data.contiguous()
    _lib.mul_inplace(gc, ste.mask.contiguous().float())
    if gc is not g.data:
        g.data.copy_(gc)
