"""Shared by the CPU (oracle) and GPU (product) analysis tests: cases, inputs, golden accessors."""
import numpy as np

from tests.golden import cases

TOKENS_PER_CONTEXT = 5

ANALYSIS_CASES = {
    "analysis_bsae_polar_d64_h2048": ("b_sae", cases.BSAE_CASES["bsae_polar_d64_h2048"]),
    "analysis_bsae_soft_d64_h2048": ("b_sae", cases.BSAE_CASES["bsae_soft_d64_h2048"]),
    "analysis_baseline_d64_h2048": ("baseline_sae", cases.BASELINE_CASES["baseline_d64_h2048"]),
    "analysis_qsae_d64_h2048": ("q_sae", cases.QSAE_CASES["qsae_d64_h2048"]),
    "analysis_rqsae_d64_h2048": ("rq_sae", cases.RQSAE_CASES["rqsae_d64_h2048"]),
}


def inputs(kind, cfg):
    if kind == "b_sae":
        return cases.bsae_inputs(cfg)
    if kind == "baseline_sae":
        return cases.baseline_inputs(cfg)
    if kind == "q_sae":
        return cases.qsae_inputs(cfg)
    inp = cases.rqsae_inputs(cfg)
    inp["stages"] = [(inp[f"We{i}"], inp[f"be{i}"], inp[f"W{i}"], inp[f"Wm{i}"], inp[f"bd{i}"]) for i in range(cfg["n_bits"])]
    return inp


def batches(x):
    cut = x.shape[0] // 2 + 3
    return [x[:cut], x[cut:]]


def golden_mask(g, H):
    return np.unpackbits(g["mask"], axis=1)[:, :H].astype(bool)


def golden_tokens_per_feature(g):
    out, pos = [], 0
    flat = g["tpf_flat"]
    for n in g["tpf_len"].tolist():
        out.append(flat[pos:pos + n].tolist())
        pos += n
    return out


def check_against_golden(res, g, H, *, mse_rtol):
    """res: dict(mse, mse_by_level, l0_by_level, activation_counts, coactivation [H,H], tokens_per_feature)."""
    assert abs(res["mse"] - float(g["mse"])) <= mse_rtol * abs(float(g["mse"]))
    np.testing.assert_allclose(np.asarray(res["mse_by_level"], dtype=np.float64), g["mse_by_level"], rtol=mse_rtol)
    np.testing.assert_array_equal(np.asarray(res["l0_by_level"], dtype=np.float64), g["l0_by_level"])       # exact counts / tokens
    np.testing.assert_array_equal(np.asarray(res["activation_counts"]), g["activation_counts"])
    co = np.asarray(res["coactivation"]).astype(np.int64)
    m = golden_mask(g, H).astype(np.int64)
    np.testing.assert_array_equal(co, m.T @ m)                                                               # A^T A of the reference mask
    ci, cj = np.nonzero(co)
    assert len(ci) == int(g["cooc_nnz"]) and co.sum() == int(g["cooc_sum"]) and np.trace(co) == int(g["cooc_trace"])
    assert int((co[ci, cj] * ((ci * 131 + cj * 71) % 1009)).sum()) == int(g["cooc_weighted"])
    np.testing.assert_array_equal(co[g["cooc_i"], g["cooc_j"]], g["cooc_v"])
    assert res["tokens_per_feature"] == golden_tokens_per_feature(g)
