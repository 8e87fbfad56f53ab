"""Generate tests/golden/analysis_*.npz: outputs of the UNMODIFIED reference analysis functions
(scripts/analysis/dynamic_analysis.py: compute_reconstruction_error, compute_reconstruction_error_by_level,
compute_l0_by_level, compute_activation_stats) on the seeded cases of tests/golden/cases.py, run through the
reference's own SAEWrapper (src/quantized_sae/inference/framework.py). Authoring container only:

    python tests/golden/make_golden_analysis.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402
from tests.golden import cases  # noqa: E402

OUT = Path(__file__).resolve().parent
T = torch.from_numpy
TOKENS_PER_CONTEXT = 5


def build_reference_model(ref, kind, cfg):
    c = ref.classes
    if kind == "b_sae":
        inp = cases.bsae_inputs(cfg)
        m = c.BinarySAE(cfg["D"], cfg["H"], cfg["gamma"], cfg["n_bits"])
        m.load_state_dict({"encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]),
                           "decoder.weight": T(inp["logits"]), "decoder.bias": T(inp["bd"])}, strict=True)
    elif kind == "baseline_sae":
        inp = cases.baseline_inputs(cfg)
        m = c.BaselineSparseAutoencoder(cfg["D"], cfg["H"])
        m.load_state_dict({"encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]),
                           "decoder.weight": T(inp["Wd"]), "decoder.bias": T(inp["bd"])}, strict=True)
    elif kind == "q_sae":
        inp = cases.qsae_inputs(cfg)
        m = c.QuantizedMatryoshkaSAE(cfg["D"], cfg["H"], 32, cfg["abs_range"], cfg["n_bits"], cfg["allow_bias"])
        m.load_state_dict({"encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]), "decoder.weight": T(inp["W"]),
                           "decoder.weight_mirror": T(inp["Wm"]), "decoder.bias": T(inp["bd"])}, strict=True)
    else:
        inp = cases.rqsae_inputs(cfg)
        m = c.ResidualQuantizedSAE(cfg["D"], cfg["H"], 32, cfg["abs_range"], cfg["n_bits"])
        m.load_state_dict({k: T(v) for k, v in cases.rqsae_state_dict(inp, cfg["n_bits"]).items()}, strict=True)
    return m, inp


def token_table(B, seed):
    n_ctx = -(-B // TOKENS_PER_CONTEXT)
    return np.random.default_rng(seed).integers(0, 50000, size=(n_ctx, TOKENS_PER_CONTEXT)).astype(np.int64)


def batches(x):
    """Two uneven batches, like a DataLoader with drop_last=False."""
    cut = x.shape[0] // 2 + 3
    return [x[:cut], x[cut:]]


ANALYSIS_CASES = {
    "analysis_bsae_polar_d64_h2048": ("b_sae", cases.BSAE_CASES["bsae_polar_d64_h2048"]),
    "analysis_bsae_soft_d64_h2048": ("b_sae", cases.BSAE_CASES["bsae_soft_d64_h2048"]),
    "analysis_baseline_d64_h2048": ("baseline_sae", cases.BASELINE_CASES["baseline_d64_h2048"]),
    "analysis_qsae_d64_h2048": ("q_sae", cases.QSAE_CASES["qsae_d64_h2048"]),
    "analysis_rqsae_d64_h2048": ("rq_sae", cases.RQSAE_CASES["rqsae_d64_h2048"]),
}


def main():
    ref = ref_shim.load_analysis()
    da, fw = ref.dynamic_analysis, ref.framework
    for name, (kind, cfg) in ANALYSIS_CASES.items():
        m, inp = build_reference_model(ref, kind, cfg)
        sae = fw.SAEWrapper(fw.SAE_REGISTRY[kind], m, "cpu")
        loader = [T(b) for b in batches(inp["x"])]
        tok = token_table(inp["x"].shape[0], cfg["seed"])
        mse = da.compute_reconstruction_error(sae, loader, device="cpu")
        mse_lv = da.compute_reconstruction_error_by_level(sae, loader, device="cpu")
        l0_lv = da.compute_l0_by_level(sae, loader, device="cpu")
        st = da.compute_activation_stats(sae, loader, token_ids=T(tok), tokens_per_context=TOKENS_PER_CONTEXT, device="cpu")
        mask = torch.cat([da._activation_mask(sae, b) for b in loader]).numpy()
        co = st["coactivation"].numpy()
        ci, cj = np.nonzero(co)
        tpf = st["tokens_per_feature"]
        if len(ci) > 200000:     # keep the fixture small: the full matrix is mask^T mask, pinned here by invariants
            keep = (ci * 131 + cj * 71) % 97 == 0
            ci_s, cj_s = ci[keep], cj[keep]
        else:
            ci_s, cj_s = ci, cj
        np.savez_compressed(
            OUT / f"{name}.npz", input_sha=cases.checksum(inp), mse=np.float64(mse), mse_by_level=mse_lv.numpy(),
            l0_by_level=l0_lv.numpy(), activation_counts=st["activation_counts"].numpy(),
            cooc_i=ci_s.astype(np.int32), cooc_j=cj_s.astype(np.int32), cooc_v=co[ci_s, cj_s].astype(np.int32),
            cooc_complete=np.bool_(len(ci_s) == len(ci)), cooc_nnz=np.int64(len(ci)), cooc_sum=np.int64(co.sum(dtype=np.int64)),
            cooc_trace=np.int64(np.trace(co, dtype=np.int64)),
            cooc_weighted=np.int64((co[ci, cj].astype(np.int64) * ((ci * 131 + cj * 71) % 1009)).sum()),
            mask=np.packbits(mask, axis=1), tpf_len=np.array([len(t) for t in tpf], dtype=np.int64),
            tpf_flat=np.array([t for ts in tpf for t in ts], dtype=np.int64), token_ids=tok)
        print(name, "mse", mse, "l0", l0_lv.numpy(), "nnz cooc", len(ci))


if __name__ == "__main__":
    main()
