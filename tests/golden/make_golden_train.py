"""Generate tests/golden/train_*.npz: gradients / mask updates produced by the UNMODIFIED reference modules
(via oracle/ref_shim) under the reference trainer's own losses (training/trainer.py:88-162). Pins the training-side
restatement oracle/train_oracle.py. Authoring container only (needs /root/reference):

    python tests/golden/make_golden_train.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402
from tests.golden import cases  # noqa: E402

OUT = Path(__file__).resolve().parent
T = torch.from_numpy


def _np(t):
    return t.detach().cpu().numpy()


def make_bsae(ref, name, cfg, lam):
    """trainer.py:143-151: loss = 0.5 * mse(recon, batch) + polarize_lambda * polarize_loss; loss.backward()."""
    inp = cases.bsae_inputs(cfg)
    m = ref.BinarySAE(cfg["D"], cfg["H"], cfg["gamma"], cfg["n_bits"])
    m.load_state_dict({"encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]),
                       "decoder.weight": T(inp["logits"]), "decoder.bias": T(inp["bd"])}, strict=True)
    m.train()
    x = T(inp["x"])
    latent, recon, pol = m(x)
    recon_loss = 0.5 * F.mse_loss(recon, x)
    loss = recon_loss + lam * pol
    loss.backward()
    vals, idx = cases.sparse_from_dense(_np(latent))
    rng = np.random.default_rng(cfg["seed"] + 1000)
    touched = np.unique(idx)
    others = np.setdiff1d(rng.choice(cfg["H"], size=48, replace=False), touched)
    rows = np.concatenate([touched, others]).astype(np.int64)
    gl = _np(m.decoder.weight.grad)
    gw = _np(m.encoder[0].weight.grad)
    np.savez_compressed(
        OUT / f"train_{name}.npz", input_sha=cases.checksum(inp), polarize_lambda=np.float64(lam),
        rows=rows, n_touched=np.int64(touched.size),
        grad_logits_rows=gl[rows], grad_logits_abs_sum=np.float64(np.abs(gl.astype(np.float64)).sum()),
        grad_We_rows=gw[rows], grad_We_abs_sum=np.float64(np.abs(gw.astype(np.float64)).sum()),
        grad_be=_np(m.encoder[0].bias.grad), grad_bd=_np(m.decoder.bias.grad),
        recon_loss=np.float64(recon_loss.item()), polarize=np.float64(pol.item()), latent_idx=idx,
    )


def make_qsae(ref, name, cfg, sparsity_lambda):
    """trainer.py:88-113: loss = sum_i 0.5 * mse(recon_i, batch) + sum(latent_group) * sparsity_lambda;
    loss.backward(); decoder.apply_secant_grad()."""
    inp = cases.qsae_inputs(cfg)
    m = ref.QuantizedMatryoshkaSAE(cfg["D"], cfg["H"], 32, cfg["abs_range"], cfg["n_bits"], cfg["allow_bias"])
    m.load_state_dict({"encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]), "decoder.weight": T(inp["W"]),
                       "decoder.weight_mirror": T(inp["Wm"]), "decoder.bias": T(inp["bd"])}, strict=True)
    m.train()
    x = T(inp["x"])
    latent_group, recon_groups = m(x)
    loss = sum(0.5 * F.mse_loss(r, x) for r in recon_groups) + sum(latent_group) * sparsity_lambda
    loss.backward()
    ste_W, ste_Wm = _np(m.decoder.weight.grad).copy(), _np(m.decoder.weight_mirror.grad).copy()
    gb = _np(m.decoder.bias.grad).copy() if m.decoder.bias.grad is not None else np.zeros(cfg["D"], np.float32)
    z2 = np.concatenate([_np(c["z2"]) for c in m.decoder._ctx if c is not None])
    m.decoder.apply_secant_grad()
    # big dictionaries: every 16th row + float64 sums of |grad| over the whole matrix
    rows = np.arange(0, cfg["H"], 16 if cfg["H"] * cfg["D"] > 200_000 else 1)
    sec_W, sec_Wm = _np(m.decoder.weight.grad), _np(m.decoder.weight_mirror.grad)
    np.savez_compressed(
        OUT / f"train_{name}.npz", input_sha=cases.checksum(inp), rows=rows,
        ste_W=ste_W[rows], ste_Wm=ste_Wm[rows], grad_bias=gb, z2=z2,
        secant_W=sec_W[rows], secant_Wm=sec_Wm[rows],
        abs_sums=np.array([np.abs(a.astype(np.float64)).sum() for a in (ste_W, ste_Wm, sec_W, sec_Wm)]),
    )


def make_rigl(ref, name, cfg):
    """STEWeights.init_mask / update_mask / mask_grad on the reference module (sae/ternary.py:27-90)."""
    inp = cases.rigl_inputs(cfg)
    ste = ref.STEWeights(cfg["H"], cfg["D"])
    with torch.no_grad():
        ste.weight.copy_(T(inp["w"]))
    # init_mask's topk tie choice is unspecified: only tie-free inputs are pinned for it
    init_mask = None
    if not cfg.get("quantise"):
        ste.init_mask(cfg["sparsity"])
        init_mask = _np(ste.mask).copy()
        w_after_init = _np(ste.weight).copy()
    else:
        rng = np.random.default_rng(cfg["seed"] + 7)
        m0 = (rng.random((cfg["D"], cfg["H"])) >= cfg["sparsity"]).astype(np.float32)
        ste.mask.data = T(m0)
        ste.weight.data *= ste.mask.data
        init_mask = m0
        w_after_init = _np(ste.weight).copy()
    ste.input_activations = T(inp["act"])
    ste.output_grad = T(inp["grad"])
    ste.update_mask(cfg["f_decay"], cfg["sparsity"])
    np.savez_compressed(
        OUT / f"train_{name}.npz", input_sha=cases.checksum(inp),
        init_mask=np.packbits(init_mask.astype(bool)), w_after_init=w_after_init,
        new_mask=np.packbits(_np(ste.mask).astype(bool)), w_after_update=_np(ste.weight),
        a_mean=_np(T(inp["act"]).mean(dim=0)), d_mean=_np(T(inp["grad"]).mean(dim=0)),
    )




def main():
    ref = ref_shim.load()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    for name, lam in cases.TRAIN_BSAE.items():
        make_bsae(ref, name, cases.BSAE_CASES[name], lam)
        print("wrote", name)
    for name, lam in cases.TRAIN_QSAE.items():
        make_qsae(ref, name, cases.QSAE_CASES[name], lam)
        print("wrote", name)
    for name, cfg in cases.RIGL_CASES.items():
        make_rigl(ref, name, cfg)
        print("wrote", name)


if __name__ == "__main__":
    main()
