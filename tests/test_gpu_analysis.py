"""GPU: quantizedsae_b200.analysis (sparse active lists + qsae_activation_counts / qsae_coactivation /
qsae_sq_error_accumulate through the C ABI) against the outputs of the reference's own analysis functions
(tests/golden/analysis_*.npz). Counts, co-activation, masks and per-feature token lists bit exact; MSE 2e-5 rel."""
import numpy as np
import pytest
import torch

import quantizedsae_b200 as Q
from quantizedsae_b200 import _lib as L
from quantizedsae_b200 import analysis as AN
from quantizedsae_b200.inference import framework as FW
from tests import analysis_common as AC
from tests.golden import cases

pytestmark = pytest.mark.gpu


def build_wrapper(kind, cfg, inp, dev):
    T = torch.from_numpy
    if kind == "b_sae":
        m = Q.BinarySAE(cfg["D"], cfg["H"], cfg["gamma"], cfg["n_bits"])
        m.load_state_dict({"encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]),
                           "decoder.weight": T(inp["logits"]), "decoder.bias": T(inp["bd"])}, strict=True)
    elif kind == "baseline_sae":
        m = Q.BaselineSparseAutoencoder(cfg["D"], cfg["H"])
        m.load_state_dict({"encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]),
                           "decoder.weight": T(inp["Wd"]), "decoder.bias": T(inp["bd"])}, strict=True)
    elif kind == "q_sae":
        m = Q.QuantizedMatryoshkaSAE(cfg["D"], cfg["H"], 32, cfg["abs_range"], cfg["n_bits"], cfg["allow_bias"])
        m.load_state_dict({"encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]), "decoder.weight": T(inp["W"]),
                           "decoder.weight_mirror": T(inp["Wm"]), "decoder.bias": T(inp["bd"])}, strict=True)
    else:
        m = Q.ResidualQuantizedSAE(cfg["D"], cfg["H"], 32, cfg["abs_range"], cfg["n_bits"])
        m.load_state_dict({k: T(v) for k, v in cases.rqsae_state_dict(inp, cfg["n_bits"]).items()}, strict=True)
    return FW.SAEWrapper(FW.SAE_REGISTRY[kind], m, dev)


@pytest.mark.parametrize("name", list(AC.ANALYSIS_CASES))
def test_analysis_matches_reference(cuda_device, golden_dir, name):
    kind, cfg = AC.ANALYSIS_CASES[name]
    g = np.load(golden_dir / f"{name}.npz")
    inp = AC.inputs(kind, cfg)
    sae = build_wrapper(kind, cfg, inp, cuda_device)
    loader = [torch.from_numpy(b) for b in AC.batches(inp["x"])]
    tok = torch.from_numpy(g["token_ids"])
    launches0 = L.launch_count()
    st = AN.compute_activation_stats(sae, loader, token_ids=tok, tokens_per_context=AC.TOKENS_PER_CONTEXT, device="cuda")
    res = {"mse": AN.compute_reconstruction_error(sae, loader, device="cuda"),
           "mse_by_level": AN.compute_reconstruction_error_by_level(sae, loader, device="cuda").numpy(),
           "l0_by_level": AN.compute_l0_by_level(sae, loader, device="cuda").numpy(),
           "activation_counts": st["activation_counts"].numpy(), "coactivation": st["coactivation"].numpy(),
           "tokens_per_feature": st["tokens_per_feature"]}
    assert L.launch_count() > launches0
    AC.check_against_golden(res, g, cfg["H"], mse_rtol=2e-5)
    # the reference-shaped dense mask, and the one-pass variant
    mask = torch.cat([AN._activation_mask(sae, b) for b in loader]).numpy()
    assert np.array_equal(mask, AC.golden_mask(g, cfg["H"]))
    one = AN.analyze_dataset(sae, loader, token_ids=tok, tokens_per_context=AC.TOKENS_PER_CONTEXT, device="cuda")
    assert abs(one["mse_final"] - float(g["mse"])) <= 2e-5 * float(g["mse"])
    assert torch.equal(one["activation_counts"], st["activation_counts"]) and torch.equal(one["coactivation"], st["coactivation"])
    assert one["tokens_per_feature"] == st["tokens_per_feature"]


def test_coactivation_full_size_properties(cuda_device):
    """H = 32768, k = 32, 8192 tokens: the HBM-resident [H, H] int32 matrix (4.3 GB) against size-independent
    properties: symmetry on a sample, diagonal == activation counts, total == sum of (row activity)^2."""
    H, B, k = 32768, 8192, 32
    g = torch.Generator(device=cuda_device).manual_seed(3)
    idx = torch.stack([torch.randperm(H, device=cuda_device, generator=g)[:k] for _ in range(64)]).to(torch.int32)
    idx = idx.repeat(B // 64, 1).contiguous()
    idx[::7, -1] = -1                                      # some empty slots
    vals = torch.rand((B, k), device=cuda_device, generator=g) - 0.1   # ~10 % non-positive: inactive
    counts = torch.zeros(H, dtype=torch.int64, device=cuda_device)
    cooc = torch.zeros((H, H), dtype=torch.int32, device=cuda_device)
    L.activation_counts(idx, vals, counts)
    L.coactivation(idx, vals, cooc)
    active = (idx >= 0) & (vals > 0)
    assert torch.equal(torch.diagonal(cooc).to(torch.int64), counts)
    assert int(counts.sum()) == int(active.sum())
    assert int(cooc.sum(dtype=torch.int64)) == int((active.sum(1).to(torch.int64) ** 2).sum())
    rows = idx[0][active[0]].long()
    assert torch.equal(cooc[rows][:, rows], cooc[rows][:, rows].t())
