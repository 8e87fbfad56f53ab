"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (via oracle/ref_shim).

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Each fixture holds the reference outputs for one seeded case of tests/golden/cases.py plus the
sha256 of the inputs it was generated from. Inputs are regenerated from the seed at test time.
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


def _np(t):
    return t.detach().cpu().numpy()


def make_bsae(ref, name, cfg):
    inp = cases.bsae_inputs(cfg)
    m = ref.BinarySAE(cfg["D"], cfg["H"], cfg["gamma"], cfg["n_bits"])
    sd = {
        "encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]),
        "decoder.weight": T(inp["logits"]), "decoder.bias": T(inp["bd"]),
    }
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        latent, recon, pol = m(T(inp["x"]))
        intw = m.decoder.quantized_int_weights()
        softw = m.decoder.quantized_int_weights_continuous()
    vals, idx = cases.sparse_from_dense(_np(latent))
    k = int(cfg["H"] * m.k)
    assert vals.shape[1] == k, (vals.shape, k)
    iw = _np(intw).astype(np.int8)
    dictionary = dict(int_weights_sha=cases.int_weights_sha(iw)) if cfg.get("big") else dict(int_weights=iw)
    np.savez_compressed(
        OUT / f"{name}.npz", input_sha=cases.checksum(inp), k=k,
        latent_vals=vals, latent_idx=idx, recon=_np(recon), polarize=np.float64(pol.item()),
        soft_weights_row0=_np(softw)[:4].copy(), **dictionary,
        state_keys=np.array(sorted(m.state_dict().keys())),
        state_shapes=np.array([str(tuple(v.shape)) for _, v in sorted(m.state_dict().items())]),
    )


def make_baseline(ref, name, cfg):
    inp = cases.baseline_inputs(cfg)
    m = ref.BaselineSparseAutoencoder(cfg["D"], cfg["H"])
    m.load_state_dict({
        "encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]),
        "decoder.weight": T(inp["Wd"]), "decoder.bias": T(inp["bd"]),
    }, strict=True)
    m.eval()
    with torch.no_grad():
        h, recon = m(T(inp["x"]))
    vals, idx = cases.sparse_from_dense(_np(h))
    np.savez_compressed(
        OUT / f"{name}.npz", input_sha=cases.checksum(inp), k=m.topk,
        latent_vals=vals, latent_idx=idx, recon=_np(recon),
        state_keys=np.array(sorted(m.state_dict().keys())),
        state_shapes=np.array([str(tuple(v.shape)) for _, v in sorted(m.state_dict().items())]),
    )


def make_tsae(ref, name, cfg):
    inp = cases.tsae_inputs(cfg)
    m = ref.TernarySparseAutoencoder(cfg["D"], cfg["H"])
    sd = m.state_dict()
    sd["encoder.0.weight"] = T(inp["We"])
    sd["encoder.0.bias"] = T(inp["be"])
    sd["decoder.weight"] = T(inp["Wd"])
    m.load_state_dict(sd, strict=True)
    m.eval()
    keys_before = sorted(m.state_dict().keys())
    with torch.no_grad():
        h, recon = m(T(inp["x"]))
    np.savez_compressed(
        OUT / f"{name}.npz", input_sha=cases.checksum(inp),
        h=_np(h), recon=_np(recon),
        state_keys=np.array(keys_before),
        state_shapes=np.array([str(tuple(m.state_dict()[k].shape)) for k in keys_before]),
        state_keys_after_forward=np.array(sorted(m.state_dict().keys())),
    )


def make_qsae(ref, name, cfg):
    inp = cases.qsae_inputs(cfg)
    m = ref.QuantizedMatryoshkaSAE(cfg["D"], cfg["H"], 32, cfg["abs_range"], cfg["n_bits"],
                                   cfg["allow_bias"])
    m.load_state_dict({
        "encoder.0.weight": T(inp["We"]), "encoder.0.bias": T(inp["be"]),
        "decoder.weight": T(inp["W"]), "decoder.weight_mirror": T(inp["Wm"]),
        "decoder.bias": T(inp["bd"]),
    }, strict=True)
    m.eval()
    with torch.no_grad():
        groups, result = m(T(inp["x"]))
        latent = m.encode(T(inp["x"]))
    np.savez_compressed(
        OUT / f"{name}.npz", input_sha=cases.checksum(inp),
        latent_group=np.array([g.item() for g in groups], dtype=np.float64),
        result=np.stack([_np(r) for r in result]),
        active=np.packbits(_np(latent > 0.5), axis=1),
        level_sizes=np.array(m.decoder.nested_dictionary_size),
        state_keys=np.array(sorted(m.state_dict().keys())),
        state_shapes=np.array([str(tuple(v.shape)) for _, v in sorted(m.state_dict().items())]),
    )


def make_rqsae(ref, name, cfg):
    inp = cases.rqsae_inputs(cfg)
    m = ref.ResidualQuantizedSAE(cfg["D"], cfg["H"], 32, cfg["abs_range"], cfg["n_bits"])
    m.load_state_dict({k: T(v) for k, v in cases.rqsae_state_dict(inp, cfg["n_bits"]).items()}, strict=True)
    m.eval()
    with torch.no_grad():
        groups, recons = m(T(inp["x"]))
    np.savez_compressed(
        OUT / f"{name}.npz", input_sha=cases.checksum(inp),
        latent_group=np.array([g.item() for g in groups], dtype=np.float64),
        recon=np.stack([_np(r) for r in recons]),
        stage_sizes=np.array(m.sae_hidden_dims),
        state_keys=np.array(sorted(m.state_dict().keys())),
        state_shapes=np.array([str(tuple(v.shape)) for _, v in sorted(m.state_dict().items())]),
    )


def make_misc(ref):
    """Known answers that are not tied to a seeded case."""
    # README.md:100 -- MSB-first [1,0,1,0] == storage order (LSB-first) [0,1,0,1] -> -6 -> -3.0
    dec = ref.binary_decoder(1, 1, gamma=4.0, n_bits=4)
    with torch.no_grad():
        dec.weight.copy_(torch.tensor([[-110.0, 110.0, -110.0, 110.0]]))
    readme_int = float(dec.quantized_int_weights()[0, 0])
    # all 16 nibbles, one output feature each
    dec16 = ref.binary_decoder(1, 16, gamma=4.0, n_bits=4)
    pat = np.array([[(v >> i) & 1 for i in range(4)] for v in range(16)], dtype=np.float32)
    with torch.no_grad():
        dec16.weight.copy_(T(np.where(pat.reshape(1, 64) > 0, 110.0, -110.0).astype(np.float32)))
    nibble_ints = _np(dec16.quantized_int_weights())[0]
    # level sizes of the Matryoshka decoder for a few shapes
    shapes = [(32768, 4), (32768, 8), (2 ** 20, 4), (2048, 4), (32768, 1), (1024, 3), (4096, 4)]
    sizes = {}
    for H, nb in shapes:
        # out_features=1 keeps the weights tiny; only the size arithmetic (:25-38) matters here
        dd = ref.QuantizedMatryoshkaDecoder(H, 1, abs_range=4, n_bits=nb)
        sizes[f"{H}_{nb}"] = np.array(dd.nested_dictionary_size)
    np.savez_compressed(
        OUT / "misc.npz", readme_int=readme_int, readme_value=readme_int * (4.0 / 8),
        nibble_ints=nibble_ints.astype(np.int8),
        **{f"sizes_{k}": v for k, v in sizes.items()},
    )


def main():
    torch.manual_seed(0)
    ref = ref_shim.load()
    if "--only-rqsae" in sys.argv:       # added after the first fixture set was committed
        for name, cfg in cases.RQSAE_CASES.items():
            make_rqsae(ref, name, cfg)
        return
    if "--only-headline-shapes" in sys.argv:   # the H = 32768 cases, added later still
        make_bsae(ref, "bsae_polar_d512_h32768", cases.BSAE_CASES["bsae_polar_d512_h32768"])
        make_baseline(ref, "baseline_d512_h32768", cases.BASELINE_CASES["baseline_d512_h32768"])
        make_tsae(ref, "tsae_d512_h32768", cases.TSAE_CASES["tsae_d512_h32768"])
        make_qsae(ref, "qsae_d512_h32768", cases.QSAE_CASES["qsae_d512_h32768"])
        for p in sorted(OUT.glob("*h32768.npz")):
            print(f"{p.name:40s} {p.stat().st_size/1024:8.1f} KiB")
        return
    for name, cfg in cases.RQSAE_CASES.items():
        make_rqsae(ref, name, cfg)
    for name, cfg in cases.BSAE_CASES.items():
        make_bsae(ref, name, cfg)
    for name, cfg in cases.BASELINE_CASES.items():
        make_baseline(ref, name, cfg)
    for name, cfg in cases.TSAE_CASES.items():
        make_tsae(ref, name, cfg)
    for name, cfg in cases.QSAE_CASES.items():
        make_qsae(ref, name, cfg)
    make_misc(ref)
    for p in sorted(OUT.glob("*.npz")):
        print(f"{p.name:40s} {p.stat().st_size/1024:8.1f} KiB")


if __name__ == "__main__":
    main()
