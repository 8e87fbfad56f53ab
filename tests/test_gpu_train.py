"""GPU parity of the training-side kernels (csrc/train.cu through the C ABI) against
  * the fixtures produced by torch autograd on the UNMODIFIED reference under the reference trainer's losses
    (tests/golden/train_*.npz, tests/golden/make_golden_train.py), and
  * the numpy restatement oracle/train_oracle.py on seeded inputs.
Tolerances: gradients are fp32 sums whose order is not fixed (vector atomics): rtol 2e-4, atol 2e-6 * max|grad|;
masks, activity counts and selections are bit exact (tie-free inputs) or exact up to the documented tie rule."""
import numpy as np
import pytest
import torch

import quantizedsae_b200 as Q
from oracle import qsae_oracle as O
from oracle import train_oracle as TO
from quantizedsae_b200 import _lib as L
from quantizedsae_b200 import training
from tests.golden import cases

pytestmark = pytest.mark.gpu


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def close(got, want, rtol=2e-4, atol_frac=2e-6):
    want = np.asarray(want, dtype=np.float64)
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    atol = atol_frac * max(1e-30, float(np.abs(want).max()))
    np.testing.assert_allclose(got.astype(np.float64), want, rtol=rtol, atol=atol)


# ------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,k,H,D", [(1, 1, 16, 4), (37, 5, 300, 30), (64, 33, 2048, 64), (130, 65, 4096, 512), (9, 7, 128, 640)])
def test_rows_scatter_add_and_gather_dot_vs_numpy(cuda_device, B, k, H, D):
    rng = np.random.default_rng(B + k)
    coef = rng.standard_normal((B, k)).astype(np.float32)
    idx = rng.integers(0, H, size=(B, k)).astype(np.int32)       # repeats across (and inside) rows: sums must add up
    idx[0, -1] = -1
    src = rng.standard_normal((B, D)).astype(np.float32)
    dst0 = rng.standard_normal((H, D)).astype(np.float32)
    want = dst0.astype(np.float64)
    col = np.zeros(H)
    for b in range(B):
        for j in range(k):
            if idx[b, j] >= 0:
                want[idx[b, j]] += 0.5 * coef[b, j] * src[b].astype(np.float64)
                col[idx[b, j]] += 0.5 * coef[b, j]
    dst = T(dst0, cuda_device)
    dcol = torch.zeros(H, device=cuda_device)
    L.rows_scatter_add(T(coef, cuda_device), T(idx, cuda_device), T(src, cuda_device), dst, scale=0.5, dst_col=dcol)
    close(dst, want, rtol=1e-4, atol_frac=1e-6)
    close(dcol, col, rtol=1e-4, atol_frac=1e-6)
    ones = torch.zeros((H, D), device=cuda_device)
    L.rows_scatter_add(None, T(idx, cuda_device), T(src, cuda_device), ones)          # coef = 1
    w1 = np.zeros((H, D))
    for b in range(B):
        for j in range(k):
            if idx[b, j] >= 0:
                w1[idx[b, j]] += src[b]
    close(ones, w1, rtol=1e-4, atol_frac=1e-6)
    rows = rng.standard_normal((H, D)).astype(np.float32)
    got = L.rows_gather_dot(T(src, cuda_device), T(rows, cuda_device), T(idx, cuda_device), scale=2.0)
    wd = 2.0 * np.einsum("bd,bkd->bk", src.astype(np.float64), rows[np.maximum(idx, 0)].astype(np.float64))
    wd[idx < 0] = 0
    close(got, wd, rtol=1e-4, atol_frac=1e-6)


@pytest.mark.parametrize("R,C", [(1, 1), (33, 70), (4096, 512), (300, 2049)])
def test_column_sum_vs_numpy(cuda_device, R, C):
    a = np.random.default_rng(R).standard_normal((R, C)).astype(np.float32)
    got = L.column_sum(T(a, cuda_device), 1.0 / R)
    close(got, a.astype(np.float64).mean(0), rtol=1e-4, atol_frac=1e-5)


@pytest.mark.parametrize("H,D,n_bits", [(64, 8, 4), (33, 5, 3), (128, 32, 8), (50, 6, 1)])
def test_logit_grad_kernel_vs_oracle_formula(cuda_device, H, D, n_bits):
    rng = np.random.default_rng(H)
    logits = (rng.standard_normal((H, D * n_bits)) * 2).astype(np.float32)
    G = rng.standard_normal((H, D)).astype(np.float32)
    lam = 0.37
    p = 1 / (1 + np.exp(-logits.astype(np.float64))).reshape(H, D, n_bits)
    c = O.bit_coefficients(n_bits).astype(np.float64)
    pw = 2.0 ** np.arange(n_bits)
    want = ((G[:, :, None] * c + lam * pw * (1 - 2 * p) / logits.size) * p * (1 - p)).reshape(H, -1)
    out = torch.full((H, D * n_bits), 7.0, device=cuda_device)
    L.bsae_logit_grad(T(logits, cuda_device), T(G, cuda_device), D, n_bits, lam, out, accumulate=False)
    close(out, want, rtol=1e-4, atol_frac=1e-6)
    L.bsae_logit_grad(T(logits, cuda_device), None, D, n_bits, torch.tensor(lam, device=cuda_device), out, accumulate=True)
    want2 = want + (lam * pw * (1 - 2 * p) / logits.size * p * (1 - p)).reshape(H, -1)
    close(out, want2, rtol=1e-4, atol_frac=1e-6)


# ------------------------------------------------------------------------------------------
# b_sae: the reference trainer's step through autograd (training/trainer.py:143-151)
# ------------------------------------------------------------------------------------------
def _bsae_model(cfg, inp, dev):
    m = Q.BinarySAE(cfg["D"], cfg["H"], cfg["gamma"], cfg["n_bits"]).to(dev)
    m.load_state_dict({"encoder.0.weight": T(inp["We"], dev), "encoder.0.bias": T(inp["be"], dev),
                       "decoder.weight": T(inp["logits"], dev), "decoder.bias": T(inp["bd"], dev)}, strict=True)
    m.autograd = True
    return m


@pytest.mark.parametrize("name", sorted(cases.TRAIN_BSAE))
def test_bsae_trainer_step_gradients_match_reference_autograd(cuda_device, golden_dir, name):
    cfg = cases.BSAE_CASES[name]
    inp = cases.bsae_inputs(cfg)
    g = np.load(golden_dir / f"train_{name}.npz")
    assert str(g["input_sha"]) == cases.checksum(inp)
    lam = float(g["polarize_lambda"])
    m = _bsae_model(cfg, inp, cuda_device)
    x = T(inp["x"], cuda_device)
    latent, recon, pol = m(x)
    recon_loss = 0.5 * torch.nn.functional.mse_loss(recon, x)
    loss = recon_loss + lam * pol
    loss.backward()
    assert latent.shape == (cfg["B"], cfg["H"]) and not latent.requires_grad
    assert float(recon_loss) == pytest.approx(float(g["recon_loss"]), rel=1e-5)
    assert float(pol) == pytest.approx(float(g["polarize"]), rel=1e-5)
    rows = g["rows"]
    close(m.decoder.weight.grad[rows], g["grad_logits_rows"])
    close(m.encoder[0].weight.grad[rows], g["grad_We_rows"])
    close(m.encoder[0].bias.grad, g["grad_be"])
    close(m.decoder.bias.grad, g["grad_bd"])
    assert float(m.decoder.weight.grad.double().abs().sum()) == pytest.approx(float(g["grad_logits_abs_sum"]), rel=1e-5)
    assert float(m.encoder[0].weight.grad.double().abs().sum()) == pytest.approx(float(g["grad_We_abs_sum"]), rel=1e-5)
    # and against the full restatement
    want = TO.bsae_training_grads(inp["x"], inp["We"], inp["be"], inp["logits"], inp["bd"], n_bits=cfg["n_bits"],
                                  gamma=cfg["gamma"], k=O.bsae_k(cfg["H"]), polarize_lambda=lam)
    for key, p in (("decoder.weight", m.decoder.weight), ("encoder.0.weight", m.encoder[0].weight),
                   ("encoder.0.bias", m.encoder[0].bias), ("decoder.bias", m.decoder.bias)):
        close(p.grad, want[key])
    # a second backward accumulates like autograd does
    latent, recon, pol = m(x)
    (0.5 * torch.nn.functional.mse_loss(recon, x) + lam * pol).backward()
    close(m.decoder.bias.grad, 2 * want["decoder.bias"])


def test_bsae_autograd_input_gradient_and_no_grad_path(cuda_device):
    cfg = cases.BSAE_CASES["bsae_soft_d64_h2048"]
    inp = cases.bsae_inputs(cfg)
    m = _bsae_model(cfg, inp, cuda_device)
    x = T(inp["x"], cuda_device).requires_grad_(True)
    _, recon, _ = m(x)
    recon.sum().backward()
    vals, idx, _, _ = O.bsae_forward(inp["x"], inp["We"], inp["be"], inp["logits"], inp["bd"], n_bits=cfg["n_bits"],
                                     gamma=cfg["gamma"], k=O.bsae_k(cfg["H"]))
    q = cfg["gamma"] / 2 ** (cfg["n_bits"] - 1)
    soft = O.dequant_soft(inp["logits"], cfg["n_bits"]).astype(np.float64)
    gz = q * soft[idx.astype(np.int64)].sum(-1)                              # d sum(recon) / d z at the kept latents
    want = np.einsum("bk,bkd->bd", gz, inp["We"][idx.astype(np.int64)].astype(np.float64))
    close(x.grad, want, rtol=1e-4, atol_frac=1e-5)
    with torch.no_grad():                                                    # inference path untouched by the switch
        lat, recon2, _ = m(x)
    assert not recon2.requires_grad
    np.testing.assert_allclose(recon2.cpu().numpy(), recon.detach().cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_bsae_backward_full_size_properties(cuda_device):
    """BASELINE config-1 shape (512 -> 32768, n_bits 4, k 65), B = 1024: size-independent properties + sampled rows."""
    D, H, nb, B = 512, 32768, 4, 1024
    rng = np.random.default_rng(5)
    m = Q.BinarySAE(D, H, 4.0, nb).to(cuda_device)
    with torch.no_grad():
        m.decoder.weight.copy_(T((rng.standard_normal((H, D * nb)) * 1.5).astype(np.float32), cuda_device))
        m.decoder.bias.copy_(T(rng.standard_normal(D).astype(np.float32), cuda_device))
    m.autograd = True
    m.return_dense = False
    x = T(rng.standard_normal((B, D)).astype(np.float32), cuda_device)
    lat, recon, pol = m(x)
    lam = 0.25
    (0.5 * torch.nn.functional.mse_loss(recon, x) + lam * pol).backward()
    g = ((recon - x) / (B * D)).detach()
    k = lat.indices.shape[1]
    assert k == 65
    # bias gradients are column sums; encoder bias gradient sums to the total latent gradient
    close(m.decoder.bias.grad, g.double().sum(0).cpu().numpy(), rtol=1e-4, atol_frac=1e-5)
    q = 4.0 / 8
    soft = L.dequant_soft(m.decoder.weight.detach(), D, nb)
    gz = q * torch.einsum("bd,bkd->bk", g.double(), soft[lat.indices.long()].double())
    assert float(m.encoder[0].bias.grad.double().sum()) == pytest.approx(float(gz.sum()), rel=1e-4, abs=1e-9)
    # untouched rows: zero encoder gradient, polarize-only logit gradient
    touched = torch.zeros(H, dtype=torch.bool, device=cuda_device)
    touched[lat.indices.long().reshape(-1)] = True
    un = torch.nonzero(~touched).reshape(-1)[:64]
    assert un.numel() > 0
    assert float(m.encoder[0].weight.grad[un].abs().max()) == 0.0
    wl = m.decoder.weight.detach()[un].double().cpu().numpy()
    p = (1 / (1 + np.exp(-wl))).reshape(len(un), D, nb)
    pw = 2.0 ** np.arange(nb)
    close(m.decoder.weight.grad[un], (lam * pw * (1 - 2 * p) / (H * D * nb) * p * (1 - p)).reshape(len(un), -1), rtol=2e-4, atol_frac=1e-5)
    # sampled touched rows against the closed form
    rows = torch.nonzero(touched).reshape(-1)[::97][:48]
    idx_c, vals_c, g_c = lat.indices.cpu().numpy(), lat.values.detach().double().cpu().numpy(), g.double().cpu().numpy()
    xc = x.double().cpu().numpy()
    gzc = gz.cpu().numpy()
    for h in rows.cpu().numpy():
        bb, jj = np.nonzero(idx_c == h)
        Gh = q * (vals_c[bb, jj][:, None] * g_c[bb]).sum(0)
        wl = m.decoder.weight.detach()[h].double().cpu().numpy().reshape(D, nb)
        p = 1 / (1 + np.exp(-wl))
        c = O.bit_coefficients(nb).astype(np.float64)
        want = ((Gh[:, None] * c + lam * pw * (1 - 2 * p) / (H * D * nb)) * p * (1 - p)).reshape(-1)
        close(m.decoder.weight.grad[h], want, rtol=3e-4, atol_frac=1e-5)
        close(m.encoder[0].weight.grad[h], (gzc[bb, jj][:, None] * xc[bb]).sum(0), rtol=3e-4, atol_frac=1e-5)


# ------------------------------------------------------------------------------------------
# q_sae decoder: STE backward + secant correction (training/trainer.py:88-113)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(cases.TRAIN_QSAE))
def test_qsae_decoder_gradients_and_secant_match_reference(cuda_device, golden_dir, name):
    cfg = cases.QSAE_CASES[name]
    inp = cases.qsae_inputs(cfg)
    g = np.load(golden_dir / f"train_{name}.npz")
    assert str(g["input_sha"]) == cases.checksum(inp)
    dev = cuda_device
    m = Q.QuantizedMatryoshkaSAE(cfg["D"], cfg["H"], 32, cfg["abs_range"], cfg["n_bits"], cfg["allow_bias"]).to(dev)
    m.load_state_dict({"encoder.0.weight": T(inp["We"], dev), "encoder.0.bias": T(inp["be"], dev),
                       "decoder.weight": T(inp["W"], dev), "decoder.weight_mirror": T(inp["Wm"], dev),
                       "decoder.bias": T(inp["bd"], dev)}, strict=True)
    x = T(inp["x"], dev)
    try:
        out = m.forward_active(x, active_cap=64)
    except RuntimeError:
        pytest.skip("dense activity has no sparse active-list form (sparse path only)")
    B, D = x.shape
    grads = [(r - x) / float(B * D) for r in out["reconstruction_levels"]]
    z2 = m.decoder.ste_backward(out["active_idx"], grads)
    rows = g["rows"]
    assert np.array_equal(z2.cpu().numpy(), g["z2"].astype(np.int64))
    close(m.decoder.weight.grad[rows], g["ste_W"])
    close(m.decoder.weight_mirror.grad[rows], g["ste_Wm"])
    if cfg["allow_bias"]:
        close(m.decoder.bias.grad, g["grad_bias"])
    else:
        assert m.decoder.bias.grad is None
    m.decoder.apply_secant_grad()
    close(m.decoder.weight.grad[rows], g["secant_W"])
    close(m.decoder.weight_mirror.grad[rows], g["secant_Wm"])
    sums = [float(t.double().abs().sum()) for t in (m.decoder.weight.grad, m.decoder.weight_mirror.grad)]
    np.testing.assert_allclose(sums, g["abs_sums"][2:], rtol=1e-5)
    # the one-call trainer mirror gives the same thing on fresh gradients
    m.zero_grad(set_to_none=True)
    res = training.qsae_trainer_decoder_grads(m, x, active_cap=64)
    close(m.decoder.weight.grad[rows], g["secant_W"])
    want = TO.qsae_training_decoder_grads(inp["x"], inp["We"], inp["be"], inp["W"], inp["Wm"], inp["bd"], n_bits=cfg["n_bits"],
                                          abs_range=cfg["abs_range"], allow_bias=cfg["allow_bias"])
    close(m.decoder.weight_mirror.grad, want["secant_Wm"])
    assert res["recon_losses"].shape == (cfg["n_bits"],)


def test_qsae_secant_joint_gradient_factor(cuda_device):
    cfg = cases.QSAE_CASES["qsae_d64_h2048"]
    inp = cases.qsae_inputs(cfg)
    dev = cuda_device
    m = Q.QuantizedMatryoshkaSAE(cfg["D"], cfg["H"], 32, cfg["abs_range"], cfg["n_bits"], True).to(dev)
    m.load_state_dict({"encoder.0.weight": T(inp["We"], dev), "encoder.0.bias": T(inp["be"], dev),
                       "decoder.weight": T(inp["W"], dev), "decoder.weight_mirror": T(inp["Wm"], dev),
                       "decoder.bias": T(inp["bd"], dev)}, strict=True)
    x = T(inp["x"], dev)
    out = m.forward_active(x, active_cap=64)
    B, D = x.shape
    z2 = m.decoder.ste_backward(out["active_idx"], [torch.zeros_like(x) for _ in range(cfg["n_bits"])])
    m.zero_grad(set_to_none=True)
    m.decoder.joint_gradient = True
    m.decoder.apply_secant_grad()
    dW, dWm = TO.qsae_secant_term(z2.cpu().numpy(), inp["W"], inp["Wm"], B, n_bits=cfg["n_bits"], abs_range=cfg["abs_range"],
                                  joint_gradient=True)
    close(m.decoder.weight.grad, dW)
    close(m.decoder.weight_mirror.grad, dWm)


# ------------------------------------------------------------------------------------------
# t_sae: RigL mask maintenance (sae/ternary.py:27-90)
# ------------------------------------------------------------------------------------------
def _unpack(bits, D, H):
    return np.unpackbits(bits)[: D * H].reshape(D, H).astype(np.float32)


@pytest.mark.parametrize("name", sorted(cases.RIGL_CASES))
def test_rigl_init_and_update_mask_match_reference(cuda_device, golden_dir, name):
    cfg = cases.RIGL_CASES[name]
    inp = cases.rigl_inputs(cfg)
    g = np.load(golden_dir / f"train_{name}.npz")
    assert str(g["input_sha"]) == cases.checksum(inp)
    D, H = cfg["D"], cfg["H"]
    dev = cuda_device
    ste = Q.STEWeights(H, D).to(dev)
    init_mask = _unpack(g["init_mask"], D, H)
    if not cfg.get("quantise"):
        with torch.no_grad():
            ste.weight.copy_(T(inp["w"], dev))
        ste.init_mask(cfg["sparsity"])
        assert np.array_equal(ste.mask.cpu().numpy(), init_mask)
        assert np.array_equal(ste.weight.detach().cpu().numpy(), g["w_after_init"])
    else:
        with torch.no_grad():
            ste.weight.copy_(T(g["w_after_init"], dev))
        ste.mask = T(init_mask, dev)
    # exact selection with the reference's own batch means
    w = T(g["w_after_init"], dev).clone()
    mask = T(init_mask, dev).clone()
    n_drop = int(cfg["f_decay"] * (1 - cfg["sparsity"]) * w.numel())
    L.rigl_update_mask(w, mask, T(g["a_mean"], dev), T(g["d_mean"], dev), n_drop, n_drop)
    want_mask = _unpack(g["new_mask"], D, H)
    assert np.array_equal(mask.cpu().numpy(), want_mask)
    assert np.array_equal(w.cpu().numpy(), g["w_after_update"])
    # module API: the means come from the column-sum kernel (different summation order than torch.mean), so entries
    # whose score is within rounding of the cut may differ; everything else must agree
    ste.input_activations = T(inp["act"], dev)
    ste.output_grad = T(inp["grad"], dev)
    ste.update_mask(cfg["f_decay"], cfg["sparsity"])
    got = ste.mask.cpu().numpy()
    assert int(got.sum()) == int(want_mask.sum())
    diff = np.argwhere(got != want_mask)
    assert len(diff) <= max(2, int(2e-4 * got.size))
    scores = np.outer(np.abs(g["d_mean"]), np.abs(g["a_mean"]))
    grown = (want_mask == 1) & (init_mask == 0)
    if grown.any() and len(diff):
        cut = scores[grown].min()
        assert np.all(np.abs(scores[diff[:, 0], diff[:, 1]] - cut) <= 1e-5 * cut)
    assert float((ste.weight.detach() * (1 - ste.mask)).abs().max()) == 0.0
    # mask_grad
    ste.weight.grad = torch.ones_like(ste.weight)
    ste.mask_grad()
    assert np.array_equal(ste.weight.grad.cpu().numpy(), got)


def test_rigl_tie_rule_and_edge_counts(cuda_device):
    """Ties at the grow cut: lowest flat index wins (documented rule); drop removes whole tie groups (ternary.py:68)."""
    dev = cuda_device
    D, H = 4, 64
    w = np.full((D, H), 0.25, np.float32)
    w[0, :8] = 0.125                                   # 8 smallest active weights, all tied
    mask = np.ones((D, H), np.float32)
    mask[3, :] = 0                                     # one inactive row: 64 grow candidates
    w[3, :] = 0
    a = np.ones(H, np.float32)                         # all scores equal within a row of d
    d = np.array([1, 1, 1, 2], np.float32)
    wt, mt = T(w, dev), T(mask, dev)
    L.rigl_update_mask(wt, mt, T(a, dev), T(d, dev), 3, 3)      # 3rd smallest active |w| = 0.125 -> all 8 ties dropped
    got = mt.cpu().numpy()
    assert got[0, :8].sum() == 0 and got[0, 8:].all() and got[1].all() and got[2].all()
    # inactive now: row 3 (score 2) and the 8 dropped (score 1): the 3 grown are the lowest flat indices of row 3
    assert got[3, :3].all() and got[3, 3:].sum() == 0
    om, ow = TO.rigl_update_mask(w, mask, np.ones((1, H), np.float32), d[None, :], 3 / (0.3 * w.size), 0.7)
    assert np.array_equal(got, om)
    assert np.array_equal(wt.cpu().numpy(), ow)
    # n_drop = 0: nothing changes except weight *= mask
    wt2, mt2 = T(w + 1, dev), T(mask, dev)
    L.rigl_update_mask(wt2, mt2, None, None, 0, 0)
    assert np.array_equal(mt2.cpu().numpy(), mask)
    assert np.array_equal(wt2.cpu().numpy(), (w + 1) * mask)


def test_rigl_full_size_against_oracle(cuda_device):
    """BASELINE config-3 shape: decoder.weight [512, 32768] (16.7 M entries), RigL target density 0.3."""
    rng = np.random.default_rng(9)
    D, H, B = 512, 32768, 64
    w = (0.4824 * rng.standard_normal((D, H))).astype(np.float32)
    act = (np.maximum(rng.standard_normal((B, H)), 0) + 1e-3).astype(np.float32)
    grad = rng.standard_normal((B, D)).astype(np.float32)
    m0, w0 = TO.rigl_init_mask(w, 0.7)
    dev = cuda_device
    wt = T(w, dev)
    mt = torch.ones_like(wt)
    L.rigl_init_mask(wt, mt, int(w.size * 0.7))
    assert np.array_equal(mt.cpu().numpy(), m0)
    assert np.array_equal(wt.cpu().numpy(), w0)
    a_mean = act.mean(0, dtype=np.float32)
    d_mean = grad.mean(0, dtype=np.float32)
    f_decay = 0.2
    n_drop = int(f_decay * (1 - 0.7) * w.size)
    L.rigl_update_mask(wt, mt, T(a_mean, dev), T(d_mean, dev), n_drop, n_drop)
    # oracle with the same means (it recomputes them from one-row "batches")
    m1, w1 = TO.rigl_update_mask(w0, m0, a_mean[None, :], d_mean[None, :], f_decay, 0.7)
    assert np.array_equal(mt.cpu().numpy(), m1)
    assert np.array_equal(wt.cpu().numpy(), w1)
    assert int(m1.sum()) == int(m0.sum())              # drop n, grow n (no ties in continuous data)


@pytest.mark.parametrize("D,H", [(3, 37), (5, 64), (16, 250)])
def test_rigl_odd_shapes_against_oracle(cuda_device, D, H):
    """Shapes that take the scalar (H % 4 != 0) and the 16-byte paths of the selection kernels."""
    rng = np.random.default_rng(D * H)
    w = rng.standard_normal((D, H)).astype(np.float32)
    m0, w0 = TO.rigl_init_mask(w, 0.6)
    wt, mt = T(w, cuda_device), torch.ones((D, H), device=cuda_device)
    L.rigl_init_mask(wt, mt, int(w.size * 0.6))
    assert np.array_equal(mt.cpu().numpy(), m0) and np.array_equal(wt.cpu().numpy(), w0)
    a = (rng.random(H) + 0.01).astype(np.float32)
    d = rng.standard_normal(D).astype(np.float32)
    f_decay = 0.25
    n = int(f_decay * (1 - 0.7) * w.size)
    L.rigl_update_mask(wt, mt, T(a, cuda_device), T(d, cuda_device), n, n)
    m1, w1 = TO.rigl_update_mask(w0, m0, a[None, :], d[None, :], f_decay, 0.7)
    assert np.array_equal(mt.cpu().numpy(), m1) and np.array_equal(wt.cpu().numpy(), w1)
