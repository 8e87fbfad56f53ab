"""CPU: the training-side restatement (oracle/train_oracle.py) against the fixtures produced by torch
autograd on the UNMODIFIED reference modules under the reference trainer's losses (tests/golden/make_golden_train.py)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import qsae_oracle as fo
from oracle import train_oracle as to
from tests.golden import cases

GOLD = Path(__file__).resolve().parent / "golden"


def _load(name, inp):
    g = np.load(GOLD / f"train_{name}.npz")
    assert str(g["input_sha"]) == cases.checksum(inp), "fixture was generated from different inputs"
    return g


def _close(got, want, rtol=2e-4, atol_frac=2e-6):
    want = np.asarray(want, dtype=np.float64)
    atol = atol_frac * max(1e-30, float(np.abs(want).max()))
    np.testing.assert_allclose(np.asarray(got, dtype=np.float64), want, rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", sorted(cases.TRAIN_BSAE))
def test_bsae_training_gradients_match_reference_autograd(name):
    cfg = cases.BSAE_CASES[name]
    inp = cases.bsae_inputs(cfg)
    g = _load(name, inp)
    k = fo.bsae_k(cfg["H"])
    out = to.bsae_training_grads(inp["x"], inp["We"], inp["be"], inp["logits"], inp["bd"], n_bits=cfg["n_bits"],
                                 gamma=cfg["gamma"], k=k, polarize_lambda=float(g["polarize_lambda"]))
    assert np.array_equal(np.sort(out["idx"], 1), np.sort(g["latent_idx"], 1))
    rows = g["rows"]
    _close(out["decoder.weight"][rows], g["grad_logits_rows"])
    _close(out["encoder.0.weight"][rows], g["grad_We_rows"])
    _close(out["encoder.0.bias"], g["grad_be"])
    _close(out["decoder.bias"], g["grad_bd"])
    assert abs(np.abs(out["decoder.weight"]).sum() / float(g["grad_logits_abs_sum"]) - 1) < 1e-5
    assert abs(np.abs(out["encoder.0.weight"]).sum() / float(g["grad_We_abs_sum"]) - 1) < 1e-5
    assert abs(out["recon_loss"] / float(g["recon_loss"]) - 1) < 1e-5
    # rows no sample touched: only the polarize term
    untouched = rows[int(g["n_touched"]):]
    assert np.all(out["encoder.0.weight"][untouched] == 0)


@pytest.mark.parametrize("name", sorted(cases.TRAIN_QSAE))
def test_qsae_decoder_gradients_and_secant_match_reference(name):
    cfg = cases.QSAE_CASES[name]
    inp = cases.qsae_inputs(cfg)
    g = _load(name, inp)
    out = to.qsae_training_decoder_grads(inp["x"], inp["We"], inp["be"], inp["W"], inp["Wm"], inp["bd"],
                                         n_bits=cfg["n_bits"], abs_range=cfg["abs_range"], allow_bias=cfg["allow_bias"])
    rows = g["rows"]
    assert np.array_equal(out["z2"], g["z2"])
    _close(out["ste_W"][rows], g["ste_W"])
    _close(out["ste_Wm"][rows], g["ste_Wm"])
    _close(out["secant_W"][rows], g["secant_W"])
    _close(out["secant_Wm"][rows], g["secant_Wm"])
    _close(out["bias"], g["grad_bias"])
    sums = [np.abs(out[k]).sum() for k in ("ste_W", "ste_Wm", "secant_W", "secant_Wm")]
    np.testing.assert_allclose(sums, g["abs_sums"], rtol=1e-5)


@pytest.mark.parametrize("name", sorted(cases.RIGL_CASES))
def test_rigl_mask_updates_match_reference(name):
    cfg = cases.RIGL_CASES[name]
    inp = cases.rigl_inputs(cfg)
    g = _load(name, inp)
    n = cfg["D"] * cfg["H"]
    init_mask = np.unpackbits(g["init_mask"])[:n].reshape(cfg["D"], cfg["H"]).astype(np.float32)
    if not cfg.get("quantise"):
        m0, w0 = to.rigl_init_mask(inp["w"], cfg["sparsity"])
        assert np.array_equal(m0, init_mask)
        assert np.array_equal(w0, g["w_after_init"])
    m1, w1 = to.rigl_update_mask(g["w_after_init"], init_mask, inp["act"], inp["grad"], cfg["f_decay"], cfg["sparsity"])
    want = np.unpackbits(g["new_mask"])[:n].reshape(cfg["D"], cfg["H"]).astype(np.float32)
    assert np.array_equal(m1, want)
    assert np.array_equal(w1, g["w_after_update"])
