"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the QuantizedSAE forward hot path.

This file is the parity oracle for the CUDA path. It is *not* part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it. The product package never does.

Every function cites the reference lines it restates (paths relative to
``/root/reference/src/quantized_sae/``). The restatement is pinned two ways
(``tests/test_oracle_golden.py``):

  * against the committed fixtures in ``tests/golden/*.npz``, produced by running the
    *unmodified* reference modules (``tests/golden/make_golden.py``), and
  * against the only known-answer vector the reference publishes
    (``README.md:100``: bits MSB-first [1,0,1,0], gamma=4 -> -6 -> -3.0).

The reference has no tests of its own, so those fixtures are the pin.

Conventions
-----------
``x``  [B, D] float32 activations; ``We`` [H, D], ``be`` [H] encoder; sparse latents are
returned as ``(values [B,k] float32, indices [B,k] int32)`` ordered by (value desc, index asc)
-- the reference's ``torch.topk`` order is unspecified for ties, so parity tests use tie-free
inputs and this oracle fixes the deterministic rule the CUDA path follows.
Reconstructions are accumulated in float64 from the sparse form, which agrees with the
reference's dense float32 matmul to ~1e-6 relative (checked in the golden tests).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# b_sae: two's-complement bit-plane dictionary          (sae/binary.py)
# --------------------------------------------------------------------------------------

def bit_coefficients(n_bits: int) -> np.ndarray:
    """[1, 2, 4, ..., -2^(n-1)] -- LSB first, sign bit last (sae/binary.py:28-29)."""
    c = (2.0 ** np.arange(n_bits)).astype(F32)
    c[-1] = -c[-1]
    return c


def sigmoid_f32(w: np.ndarray) -> np.ndarray:
    """float32 logistic, the literal expression torch evaluates: 1 / (1 + exp(-w))."""
    w = np.asarray(w, dtype=F32)
    with np.errstate(over="ignore"):
        return (F32(1.0) / (F32(1.0) + np.exp(-w, dtype=F32))).astype(F32)


def hard_bits(logits: np.ndarray) -> np.ndarray:
    """bit = sigmoid(w) > 0.5, strict (sae/binary.py:52).

    In exact arithmetic this is w > 0. In float32 the logistic rounds to exactly 0.5 for
    0 < w <~ 1.2e-7 (implementation dependent inside that band: torch-CPU, torch-CUDA and numpy
    use different expf), so parity inputs keep |w| >= 1e-6 and the band is tested separately.
    """
    return sigmoid_f32(logits) > F32(0.5)


def dequant_hard(logits: np.ndarray, n_bits: int) -> np.ndarray:
    """quantized_int_weights(): [H, D*n_bits] logits -> [H, D] int8 in [-2^(n-1), 2^(n-1)-1].

    Column d*n_bits + i of a row is bit i of output feature d (sae/binary.py:49-58, view at :53).
    """
    H = logits.shape[0]
    bits = hard_bits(logits).reshape(H, -1, n_bits).astype(np.int32)
    coef = bit_coefficients(n_bits).astype(np.int32)
    return (bits * coef).sum(-1).astype(np.int8)


def dequant_soft(logits: np.ndarray, n_bits: int) -> np.ndarray:
    """Soft effective weights used by the reference *forward* (sae/binary.py:26-35, :60-69).

    int_w[h,d] = sum_i sigmoid(L[h, d*n+i]) * c_i, float32, same summation order (i ascending).
    """
    H = logits.shape[0]
    p = sigmoid_f32(logits).reshape(H, -1, n_bits)
    coef = bit_coefficients(n_bits)
    acc = np.zeros(p.shape[:2], dtype=F32)
    for i in range(n_bits):
        acc = (acc + p[:, :, i] * coef[i]).astype(F32)
    return acc


def polarize_loss(logits: np.ndarray, n_bits: int) -> float:
    """mean(p * (1-p) * 2^i) with all-positive bit weights (sae/binary.py:41-42)."""
    H = logits.shape[0]
    p = sigmoid_f32(logits).reshape(H, -1, n_bits).astype(np.float64)
    w = 2.0 ** np.arange(n_bits)
    return float((p * (1.0 - p) * w).mean())


def pack_nibbles(int_w: np.ndarray) -> np.ndarray:
    """Packed-int4 dictionary layout used by the CUDA path: [H, D/2] uint8, feature 2j in the
    low nibble and 2j+1 in the high nibble of byte j, each nibble the two's-complement integer
    itself (bit i of the nibble == hard bit i)."""
    u = (int_w.astype(np.int16) & 0xF).astype(np.uint8)
    return (u[:, 0::2] | (u[:, 1::2] << 4)).astype(np.uint8)


def unpack_nibbles(packed: np.ndarray) -> np.ndarray:
    lo = (packed & 0xF).astype(np.int8)
    hi = (packed >> 4).astype(np.int8)
    lo = np.where(lo >= 8, lo - 16, lo)
    hi = np.where(hi >= 8, hi - 16, hi)
    out = np.empty((packed.shape[0], packed.shape[1] * 2), dtype=np.int8)
    out[:, 0::2] = lo
    out[:, 1::2] = hi
    return out


# --------------------------------------------------------------------------------------
# encoder + selection (all modules)
# --------------------------------------------------------------------------------------

def encode_pre(x: np.ndarray, We: np.ndarray, be: np.ndarray) -> np.ndarray:
    """z = x @ We^T + be in float32 (nn.Linear: sae/binary.py:82-84, baseline.py:8-10,
    ternary.py:95-98, quantized_matryoshka.py:206-209)."""
    return (x.astype(F32) @ We.astype(F32).T + be.astype(F32)).astype(F32)


def topk_rows(z: np.ndarray, k: int):
    """Per-row top-k of raw pre-activations (sae/binary.py:94, baseline.py:35).

    Deterministic order: value descending, index ascending among equal values."""
    B, H = z.shape
    if k > H:
        raise RuntimeError("selected index k out of range")  # torch.topk's error
    order = np.lexsort((np.broadcast_to(np.arange(H), z.shape), -z.astype(np.float64)), axis=1)
    idx = order[:, :k].astype(np.int32)
    vals = np.take_along_axis(z, idx.astype(np.int64), axis=1).astype(F32)
    return vals, idx


def densify(vals: np.ndarray, idx: np.ndarray, H: int) -> np.ndarray:
    """latent * mask (sae/binary.py:96-99) == zeros + scatter (baseline.py:38-39)."""
    out = np.zeros((vals.shape[0], H), dtype=F32)
    np.put_along_axis(out, idx.astype(np.int64), vals, axis=1)
    return out


def bsae_k(hidden_dim: int, k_frac: float = 0.002) -> int:
    """k = int(hidden_dim * self.k), self.k = 0.002 (sae/binary.py:80,94)."""
    return int(hidden_dim * k_frac)


# --------------------------------------------------------------------------------------
# decoders
# --------------------------------------------------------------------------------------

def decode_rows(vals: np.ndarray, idx: np.ndarray, rows: np.ndarray, scale: float,
                bias: np.ndarray | None) -> np.ndarray:
    """recon[b,:] = scale * sum_j vals[b,j] * rows[idx[b,j], :] + bias, float64 accumulate.

    ``rows`` is the [H, D] dictionary with features as rows: the b_sae integer / soft weights
    (sae/binary.py:38) or the transposed baseline decoder weight (baseline.py:29)."""
    g = rows[idx.astype(np.int64)].astype(np.float64)            # [B, k, D]
    acc = np.einsum("bk,bkd->bd", vals.astype(np.float64), g) * float(scale)
    if bias is not None:
        acc = acc + bias.astype(np.float64)
    return acc.astype(F32)


def bsae_forward(x, We, be, logits, dec_bias, *, n_bits: int, gamma: float, k: int,
                 mode: str = "soft"):
    """BinarySAE.forward restated sparsely (sae/binary.py:91-103).

    mode="soft": decoder uses sigmoid bits exactly like the reference forward (:26-38).
    mode="hard": decoder uses quantized_int_weights() (:49-58) -- equal to the reference forward
                 when the logits are polarised (sigmoid in {0.0, 1.0} exactly, e.g. +-110).
    Returns (values, indices, recon, polarize_loss)."""
    z = encode_pre(x, We, be)
    vals, idx = topk_rows(z, k)
    qstep = gamma / (2 ** (n_bits - 1))                           # sae/binary.py:18
    rows = dequant_soft(logits, n_bits) if mode == "soft" else dequant_hard(logits, n_bits).astype(F32)
    recon = decode_rows(vals, idx, rows, qstep, dec_bias)
    return vals, idx, recon, polarize_loss(logits, n_bits)


def baseline_forward(x, We, be, Wd, bd, *, k: int = 32):
    """BaselineSparseAutoencoder.forward (sae/baseline.py:17-40): Linear (no ReLU) -> top-k ->
    scatter -> Linear. ``Wd`` is decoder.weight [D, H] (features are columns)."""
    z = encode_pre(x, We, be)
    vals, idx = topk_rows(z, k)
    recon = decode_rows(vals, idx, np.ascontiguousarray(Wd.T), 1.0, bd)
    return vals, idx, recon


def ternarize(Wd: np.ndarray, threshold: float = 0.5) -> np.ndarray:
    """sign(w) * (|w| >= 0.5) in {-1,0,+1} (sae/ternary.py:46-49); the forward value of the STE
    expression at :51-52 equals this matrix exactly, the ``mask`` buffer does not enter it."""
    return (np.sign(Wd) * (np.abs(Wd) >= F32(threshold))).astype(np.int8)


def tsae_forward(x, We, be, Wd):
    """TernarySparseAutoencoder.forward (sae/ternary.py:116-122): h = relu(Linear(x)) dense,
    recon = h @ T^T with T = ternarize(decoder.weight [D,H]); no decoder bias."""
    h = np.maximum(encode_pre(x, We, be), F32(0))
    T = ternarize(Wd).astype(np.float64)
    recon = (h.astype(np.float64) @ T.T).astype(F32)
    return h, recon


def tsae_topk_activation(h: np.ndarray, k: int):
    """Dormant sparse mode (sae/ternary.py:102-114): top-k then clamp non-positive to 0."""
    vals, idx = topk_rows(h, k)
    return np.where(vals > 0, vals, F32(0)).astype(F32), idx


# --------------------------------------------------------------------------------------
# q_sae: quantized Matryoshka decoder                   (sae/quantized_matryoshka.py)
# --------------------------------------------------------------------------------------

def matryoshka_level_sizes(hidden_dim: int, n_bits: int) -> list[int]:
    """nested_dictionary_size (sae/quantized_matryoshka.py:25-38): base [1,1,2,4,...] scaled to
    hidden_dim with int() truncation, remainder folded into the last level."""
    sizes = [1 if i < 2 else 2 ** (i - 1) for i in range(n_bits)]
    total = sum(sizes)
    if total != hidden_dim:
        f = hidden_dim / total
        sizes = [max(1, int(s * f)) for s in sizes]
        sizes[-1] = hidden_dim - sum(sizes[:-1])
    return sizes


def sign_pm1(w: np.ndarray) -> np.ndarray:
    """+1 where sigmoid(w) >= 0.5 else -1 (sae/quantized_matryoshka.py:67-80)."""
    return np.where(sigmoid_f32(w) >= F32(0.5), 1, -1).astype(np.int8)


def qsae_dictionary(W: np.ndarray, Wm: np.ndarray, *, n_bits: int, abs_range: float):
    """T = S + S_mirror in {-2,0,2} and the per-row scale
    scale[h] = 2^(n_bits-l-2) * quant_step / (||T[h,:]||_2 + 1e-8)   (:82-91), float32."""
    H = W.shape[0]
    T = (sign_pm1(W).astype(np.int8) + sign_pm1(Wm).astype(np.int8)).astype(np.int8)
    norms = np.sqrt((T.astype(F32) ** 2).sum(1, dtype=F32)).astype(F32)
    qstep = abs_range / (2 ** (n_bits - 1))                        # :20
    sizes = matryoshka_level_sizes(H, n_bits)
    level = np.repeat(np.arange(n_bits), sizes)
    factor = (2.0 ** (n_bits - level - 2) * qstep).astype(F32)
    scale = (factor / (norms + F32(1e-8))).astype(F32)
    return T, scale, level, sizes


def qsae_forward(x, We, be, W, Wm, bd, *, n_bits: int, abs_range: float, allow_bias: bool = True):
    """QuantizedMatryoshkaSAE.forward (sae/quantized_matryoshka.py:217-220 -> :47-143).

    active a[b,h] = sigmoid(z) > 0.5 (:99); result[i] = cumulative reconstruction through level i
    (bias added once at level 0, :123-124); latent_group[i] = mean_b sum_{h in level i} a[b,h].
    Returns (latent_group [n_bits] float32, result [n_bits, B, D] float32, active [B,H] bool)."""
    z = encode_pre(x, We, be)
    act = sigmoid_f32(z) > F32(0.5)
    T, scale, level, sizes = qsae_dictionary(W, Wm, n_bits=n_bits, abs_range=abs_range)
    rowsT = T.astype(np.float64) * scale.astype(np.float64)[:, None]
    B = x.shape[0]
    D = W.shape[1]
    result = np.zeros((n_bits, B, D), dtype=np.float64)
    groups = np.zeros(n_bits, dtype=np.float64)
    acc = np.zeros((B, D), dtype=np.float64)
    start = 0
    for i, s in enumerate(sizes):
        a = act[:, start:start + s].astype(np.float64)
        acc = acc + a @ rowsT[start:start + s]
        if i == 0 and allow_bias:
            acc = acc + bd.astype(np.float64)
        groups[i] = a.sum(1).mean()
        result[i] = acc
        start += s
    return groups.astype(F32), result.astype(F32), act


def rqsae_forward(x, stages, *, abs_range: float):
    """ResidualQuantizedSAE.forward (sae/residual_quantized.py:53-69): stage t is a one-bit
    QuantizedMatryoshkaSAE on the residual r_t (bias only in stage 0, :46); r_{t+1} = 2 (r_t - recon_t).
    ``stages``: list of (We, be, W, Wm, bd). Returns (latent_group [T], recon [T, B, D])."""
    r = x.astype(F32)
    groups, recons = [], []
    for t, (We, be, W, Wm, bd) in enumerate(stages):
        g, res, _ = qsae_forward(r, We, be, W, Wm, bd, n_bits=1, abs_range=abs_range, allow_bias=(t == 0))
        groups.append(g[-1])
        recons.append(res[-1])
        r = ((r - res[-1]) * F32(2)).astype(F32)
    return np.array(groups, dtype=F32), np.stack(recons).astype(F32)


# --------------------------------------------------------------------------------------
# dense float32 port of BinarySAE.forward, op-for-op (CPU baseline timing only)
# --------------------------------------------------------------------------------------

def bsae_forward_dense_port_torch(x, We, be, logits, dec_bias, *, n_bits: int, gamma: float, k: int):
    """The reference's own op sequence on torch CPU tensors (sae/binary.py:91-103 + :24-47):
    addmm -> topk -> zeros_like/scatter_/mul -> sigmoid over all bit logits -> weighted bit sum ->
    dense [B,H]x[H,D] matmul. Used as the timed CPU baseline ("kind": "port") on the GPU box,
    where /root/reference does not exist. Takes and returns torch tensors."""
    import torch

    with torch.no_grad():
        latent = torch.addmm(be, x, We.t())
        _, top_idx = latent.topk(k, dim=1)
        keep = torch.zeros_like(latent)
        keep.scatter_(1, top_idx, 1.0)
        sparse_latent = latent * keep
        p = torch.sigmoid(logits)
        coef = 2.0 ** torch.arange(n_bits, dtype=p.dtype)
        coef[-1] = -coef[-1]
        p3 = p.view(logits.shape[0], -1, n_bits)
        int_w = (p3 * coef).sum(-1)
        recon = (gamma / 2 ** (n_bits - 1)) * sparse_latent.matmul(int_w) + dec_bias
        pol = (p3 * (1 - p3) * (2.0 ** torch.arange(n_bits, dtype=p.dtype))).mean()
    return sparse_latent, recon, pol
