"""Seeded input generators shared by the golden-fixture generator and the parity tests.

Inputs are regenerated from (case name -> config + numpy PCG64 seed) so the committed fixtures
only need to hold the *reference outputs* plus a checksum of the inputs they were made from.
"""
from __future__ import annotations

import hashlib

import numpy as np

F32 = np.float32


def round_bf16(a: np.ndarray) -> np.ndarray:
    """Round float32 to the nearest bf16-representable float32 (ties to even)."""
    u = np.ascontiguousarray(a, dtype=F32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(F32).reshape(a.shape)


def xavier_uniform(rng, out_f: int, in_f: int) -> np.ndarray:
    a = np.sqrt(6.0 / (in_f + out_f))
    return rng.uniform(-a, a, size=(out_f, in_f)).astype(F32)


def away_from_zero(w: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    """Push |w| < eps out to +-eps: keeps thresholded bits out of the float32 sigmoid band."""
    w = w.copy()
    small = np.abs(w) < eps
    w[small] = np.where(w[small] >= 0, eps, -eps).astype(F32)
    return w


# name -> config. "polar": logits are +-110 (sigmoid exactly 0/1); "bf16": x, We bf16-representable
BSAE_CASES = {
    "bsae_soft_d64_h2048":   dict(D=64,  H=2048, n_bits=4, gamma=4.0, B=32, polar=False, bf16=False, seed=11),
    "bsae_polar_d64_h2048":  dict(D=64,  H=2048, n_bits=4, gamma=4.0, B=32, polar=True,  bf16=True,  seed=12),
    "bsae_soft_d32_h1024_8b": dict(D=32, H=1024, n_bits=8, gamma=4.0, B=16, polar=False, bf16=False, seed=13),
    "bsae_polar_d512_h4096": dict(D=512, H=4096, n_bits=4, gamma=4.0, B=48, polar=True,  bf16=True,  seed=14),
    "bsae_soft_d512_h4096":  dict(D=512, H=4096, n_bits=4, gamma=1.5, B=24, polar=False, bf16=False, seed=15),
    "bsae_polar_d256_h8192_2b": dict(D=256, H=8192, n_bits=2, gamma=2.0, B=40, polar=True, bf16=True, seed=16),
    # the headline shape itself (BASELINE config 1: 512 -> 32768, n_bits = 4, gamma = 4.0, reference-default k = 65);
    # "big": the fixture stores a sha256 of the [H, D] integer dictionary instead of the array
    "bsae_polar_d512_h32768": dict(D=512, H=32768, n_bits=4, gamma=4.0, B=16, polar=True, bf16=True, seed=17, big=True),
}

BASELINE_CASES = {
    "baseline_d64_h2048":  dict(D=64,  H=2048, B=32, bf16=False, seed=21),
    "baseline_d512_h4096": dict(D=512, H=4096, B=40, bf16=True,  seed=22),
    "baseline_d512_h32768": dict(D=512, H=32768, B=16, bf16=True, seed=23),        # BASELINE config 2 shape
}

TSAE_CASES = {
    "tsae_d64_h2048":  dict(D=64,  H=2048, B=32, bf16=False, seed=31),
    "tsae_d512_h4096": dict(D=512, H=4096, B=24, bf16=True,  seed=32),
    "tsae_d512_h32768": dict(D=512, H=32768, B=8, bf16=True, seed=33),             # BASELINE config 3 shape
}

QSAE_CASES = {
    "qsae_d64_h2048":  dict(D=64,  H=2048, n_bits=4, abs_range=4.0, B=32, enc_bias=-0.5, bf16=False, allow_bias=True,  seed=41),
    "qsae_d512_h4096": dict(D=512, H=4096, n_bits=4, abs_range=1.5, B=24, enc_bias=-0.543, bf16=True, allow_bias=True, seed=42),
    "qsae_d64_h1024_dense_nobias": dict(D=64, H=1024, n_bits=3, abs_range=4.0, B=16, enc_bias=0.0, bf16=False, allow_bias=False, seed=43),
    # BASELINE config 4 shape (bias -0.543: mean L0 ~ 34)
    "qsae_d512_h32768": dict(D=512, H=32768, n_bits=4, abs_range=4.0, B=16, enc_bias=-0.543, bf16=True, allow_bias=True, seed=44),
}


# rq_sae (next row): cascade of n_bits one-bit q_saes on doubled residuals
RQSAE_CASES = {
    "rqsae_d64_h2048":  dict(D=64,  H=2048, n_bits=4, abs_range=4.0, B=32, enc_bias=-0.3, seed=51),
    "rqsae_d512_h4096": dict(D=512, H=4096, n_bits=3, abs_range=2.0, B=24, enc_bias=-0.5, seed=52),
}


def bsae_inputs(cfg: dict) -> dict:
    """Synthetic weights/inputs shaped like SURVEY.md 8(d) config 1 (scaled down)."""
    rng = np.random.default_rng(cfg["seed"])
    D, H, nb, B = cfg["D"], cfg["H"], cfg["n_bits"], cfg["B"]
    We = xavier_uniform(rng, H, D)                      # sae/binary.py:86
    be = (0.01 * rng.standard_normal(H)).astype(F32)    # non-zero so the bias path is exercised
    x = rng.standard_normal((B, D)).astype(F32)
    if cfg["bf16"]:
        We, x = round_bf16(We), round_bf16(x)
    if cfg["polar"]:
        logits = np.where(rng.random((H, D * nb)) < 0.5, 110.0, -110.0).astype(F32)
    else:
        logits = away_from_zero((rng.standard_normal((H, D * nb)) * 1.5).astype(F32))
    bd = rng.standard_normal(D).astype(F32)
    return dict(x=x, We=We, be=be, logits=logits, bd=bd)


def baseline_inputs(cfg: dict) -> dict:
    rng = np.random.default_rng(cfg["seed"])
    D, H, B = cfg["D"], cfg["H"], cfg["B"]
    We = xavier_uniform(rng, H, D)
    be = (0.01 * rng.standard_normal(H)).astype(F32)
    Wd = xavier_uniform(rng, D, H)                      # decoder.weight [D, H] (baseline.py:12)
    bd = rng.standard_normal(D).astype(F32)
    x = rng.standard_normal((B, D)).astype(F32)
    if cfg["bf16"]:
        We, x, Wd = round_bf16(We), round_bf16(x), round_bf16(Wd)
    return dict(x=x, We=We, be=be, Wd=Wd, bd=bd)


def tsae_inputs(cfg: dict) -> dict:
    rng = np.random.default_rng(cfg["seed"])
    D, H, B = cfg["D"], cfg["H"], cfg["B"]
    We = xavier_uniform(rng, H, D)
    be = (0.01 * rng.standard_normal(H)).astype(F32)
    # N(0, 0.4824^2): P(|w| >= 0.5) = 0.30, the RigL target density;
    # kaiming init would ternarise to all zeros.
    Wd = (0.4824 * rng.standard_normal((D, H))).astype(F32)
    x = rng.standard_normal((B, D)).astype(F32)
    if cfg["bf16"]:
        We, x = round_bf16(We), round_bf16(x)
    return dict(x=x, We=We, be=be, Wd=Wd)


def qsae_inputs(cfg: dict) -> dict:
    rng = np.random.default_rng(cfg["seed"])
    D, H, B = cfg["D"], cfg["H"], cfg["B"]
    We = xavier_uniform(rng, H, D)
    be = (cfg["enc_bias"] + 0.01 * rng.standard_normal(H)).astype(F32)
    W = away_from_zero(xavier_uniform(rng, H, D))       # quantized_matryoshka.py:43-44
    Wm = away_from_zero(xavier_uniform(rng, H, D))
    bd = rng.standard_normal(D).astype(F32)
    x = rng.standard_normal((B, D)).astype(F32)
    if cfg["bf16"]:
        We, x = round_bf16(We), round_bf16(x)
    return dict(x=x, We=We, be=be, W=W, Wm=Wm, bd=bd)


def rqsae_inputs(cfg: dict) -> dict:
    """Per-stage weights of ResidualQuantizedSAE (sae/residual_quantized.py:13-49): stage i is a
    QuantizedMatryoshkaSAE(input_dim, sizes[i], n_bits=1)."""
    rng = np.random.default_rng(cfg["seed"])
    D, H, B, nb = cfg["D"], cfg["H"], cfg["B"], cfg["n_bits"]
    sizes = [1 if i < 2 else 2 ** (i - 1) for i in range(nb)]
    f = H / sum(sizes)
    if sum(sizes) != H:
        sizes = [max(1, int(s * f)) for s in sizes]
        sizes[-1] = H - sum(sizes[:-1])
    out = dict(x=rng.standard_normal((B, D)).astype(F32))
    for i, hs in enumerate(sizes):
        out[f"We{i}"] = xavier_uniform(rng, hs, D)
        out[f"be{i}"] = (cfg["enc_bias"] + 0.01 * rng.standard_normal(hs)).astype(F32)
        out[f"W{i}"] = away_from_zero(xavier_uniform(rng, hs, D))
        out[f"Wm{i}"] = away_from_zero(xavier_uniform(rng, hs, D))
        out[f"bd{i}"] = rng.standard_normal(D).astype(F32)
    return out


def rqsae_state_dict(inp: dict, n_bits: int) -> dict:
    sd = {}
    for i in range(n_bits):
        sd[f"saes.{i}.encoder.0.weight"] = inp[f"We{i}"]
        sd[f"saes.{i}.encoder.0.bias"] = inp[f"be{i}"]
        sd[f"saes.{i}.decoder.weight"] = inp[f"W{i}"]
        sd[f"saes.{i}.decoder.weight_mirror"] = inp[f"Wm{i}"]
        sd[f"saes.{i}.decoder.bias"] = inp[f"bd{i}"]
    return sd


def int_weights_sha(int_w: np.ndarray) -> str:
    """sha256 of the [H, D] integer dictionary as int8, row-major (big fixtures store this instead of the array)."""
    return hashlib.sha256(np.ascontiguousarray(int_w, dtype=np.int8).tobytes()).hexdigest()


def int_weights_match(got: np.ndarray, g) -> bool:
    """got == the reference's quantized_int_weights() recorded in fixture g (array, or its sha256 for big cases)."""
    if "int_weights" in g.files:
        return bool(np.array_equal(got, g["int_weights"]))
    return int_weights_sha(got) == str(g["int_weights_sha"])


def checksum(arrays: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(arrays):
        h.update(k.encode())
        h.update(np.ascontiguousarray(arrays[k]).tobytes())
    return h.hexdigest()


def sparse_from_dense(latent: np.ndarray):
    """Dense [B,H] latent with a fixed nnz per row -> (vals, idx) ordered (value desc, index asc)."""
    B = latent.shape[0]
    nnz = (latent != 0).sum(1)
    k = int(nnz.max())
    vals = np.zeros((B, k), dtype=F32)
    idx = np.full((B, k), -1, dtype=np.int32)
    for b in range(B):
        j = np.nonzero(latent[b])[0]
        v = latent[b, j]
        o = np.lexsort((j, -v.astype(np.float64)))
        vals[b, : len(j)] = v[o]
        idx[b, : len(j)] = j[o]
    return vals, idx


# training-side fixtures: case name -> loss weight used by tests/golden/make_golden_train.py
TRAIN_BSAE = {"bsae_soft_d64_h2048": 0.3, "bsae_soft_d32_h1024_8b": 0.05, "bsae_soft_d512_h4096": 1.0}   # polarize_lambda
TRAIN_QSAE = {"qsae_d64_h2048": 1e-3, "qsae_d64_h1024_dense_nobias": 1e-3, "qsae_d512_h4096": 1e-3}      # sparsity_lambda

RIGL_CASES = {
    "rigl_d64_h2048": dict(D=64, H=2048, B=32, seed=61, sparsity=0.7, f_decay=0.3),
    "rigl_d32_h1024_ties": dict(D=32, H=1024, B=16, seed=62, sparsity=0.5, f_decay=0.1, quantise=True),
}


def rigl_inputs(cfg):
    rng = np.random.default_rng(cfg["seed"])
    D, H, B = cfg["D"], cfg["H"], cfg["B"]
    w = (0.4824 * rng.standard_normal((D, H))).astype(F32)
    if cfg.get("quantise"):
        # |w| on a 1/64 grid: many equal magnitudes, so the drop threshold removes whole tie groups (ternary.py:67-69)
        w = (np.round(w * 64) / 64).astype(F32)
    act = np.maximum(rng.standard_normal((B, H)), 0).astype(F32) + F32(1e-3)   # no exact-zero column means
    grad = rng.standard_normal((B, D)).astype(F32)
    return dict(w=w, act=act, grad=grad)


