"""Inference registry / wrapper with the reference's surface (inference/framework.py:65-359):

    sae = load_sae("b_sae", device="cuda")          # registry lookup, checkpoint restore, .eval()
    out = sae(batch)                                # dict: latent / reconstruction / aux ... per variant
    recon = sae.reconstruct(batch)
    for r in sae.reconstruct_loader(loader): ...
    dictionary = sae.decoder_dictionary()           # b_sae: quantization_step * quantized_int_weights()

What differs: the models are the B200 modules of `quantizedsae_b200.sae` (CUDA only -- a wrapper on a
CPU device can hold and export weights, but its forward raises instead of falling back); checkpoint
locations are not hard-wired to the source tree (`checkpoint_root`, or the QSAE_CHECKPOINT_ROOT
environment variable, or an explicit `checkpoint_path`); entries can be added with `register_sae`.
Registry keyword arguments are the reference's shipped configurations (:165-220).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from dataclasses import dataclass, field, replace
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, Iterator, Optional, Union

import torch
import torch.nn as nn

from ..sae import (BaselineSparseAutoencoder, BinarySAE, QuantizedMatryoshkaSAE, ResidualQuantizedSAE,
                   TernarySparseAutoencoder)

DeviceLike = Union[torch.device, str]


def _cpu(t: torch.Tensor) -> torch.Tensor:
    return t.detach().cpu().clone()


def _as_tensor(batch: Any) -> torch.Tensor:
    """DataLoader batches arrive as tensors or as 1-tuples/lists of tensors (:47-57)."""
    if isinstance(batch, (list, tuple)):
        if not batch:
            raise ValueError("Received an empty batch; cannot infer tensor input.")
        batch = batch[0]
    if not isinstance(batch, torch.Tensor):
        raise TypeError(f"Expected batch to be a torch.Tensor, received {type(batch)} instead.")
    return batch


# ---- per-variant output adapters (:76-111) and dictionary exports (:114-162) --------------------
def _out_binary(model, x):
    latent, recon, pol = model(x)
    return {"latent": latent, "reconstruction": recon, "aux": {"polarize_loss": pol}}


def _out_levels(model, x):
    groups, levels = model(x)
    return {"latent_groups": groups, "reconstruction_levels": levels, "reconstruction": levels[-1]}


def _out_pair(model, x):
    latent, recon = model(x)
    return {"latent": latent, "reconstruction": recon}


def _dict_binary(model, options):
    dec = model.decoder
    with torch.no_grad():
        weight = dec.quantization_step * dec.quantized_int_weights().to(torch.float32)
    return {"weight": _cpu(weight), "bias": _cpu(dec.bias)}


def _dict_quantized(model, options):
    dec = model.decoder
    w, wm = _cpu(dec.weight), _cpu(dec.weight_mirror)
    return {"weight": w, "weight_mirror": wm, "effective_weight": w + wm, "bias": _cpu(dec.bias)}


def _dict_residual(model, options):
    out: Dict[str, torch.Tensor] = {}
    for level, sae in enumerate(model.saes):
        w, wm = _cpu(sae.decoder.weight), _cpu(sae.decoder.weight_mirror)
        out[f"level_{level}_weight"] = w
        out[f"level_{level}_weight_mirror"] = wm
        out[f"level_{level}_effective_weight"] = w + wm
        if getattr(sae.decoder, "bias", None) is not None:
            out[f"level_{level}_bias"] = _cpu(sae.decoder.bias)
    return out


def _dict_linear(model, options):
    return {"weight": _cpu(model.decoder.weight), "bias": _cpu(model.decoder.bias)}


def _dict_ternary(model, options):
    dec = model.decoder
    out = {"weight": _cpu(dec.weight), "mask": _cpu(dec.mask)}
    if dec.weight.is_cuda:
        out["hard_weight"] = _cpu(dec.hard_weights())
    return out


@dataclass(frozen=True)
class SAERegistryEntry:
    name: str
    constructor: Callable[..., nn.Module]
    checkpoint_path: Path            # relative paths are resolved against the checkpoint root
    checkpoint_format: str           # "torch" | "safetensors"
    kwargs: Dict[str, Any] = field(default_factory=dict)
    forward_adapter: Callable[[nn.Module, torch.Tensor], Dict[str, Any]] = _out_pair
    decoder_getter: Callable[[nn.Module, Dict[str, Any]], Dict[str, torch.Tensor]] = _dict_linear


SAE_REGISTRY: Dict[str, SAERegistryEntry] = {}


def register_sae(entry: SAERegistryEntry) -> None:
    SAE_REGISTRY[entry.name] = entry


register_sae(SAERegistryEntry("b_sae", BinarySAE, Path("Trained_SAEs/b_sae_32768_4_bits.pth"), "torch",
                              {"input_dim": 512, "hidden_dim": 32768, "gamma": 1.5, "n_bits": 4},
                              _out_binary, _dict_binary))
register_sae(SAERegistryEntry("q_sae", QuantizedMatryoshkaSAE, Path("Trained_SAEs/q_sae_32768_4_bits.pth"), "torch",
                              {"input_dim": 512, "hidden_dim": 32768, "top_k": 32, "abs_range": 1.5, "n_bits": 4,
                               "allow_bias": True}, _out_levels, _dict_quantized))
register_sae(SAERegistryEntry("rq_sae", ResidualQuantizedSAE, Path("Trained_SAEs/rq_sae_32768_4_bits.pth"), "torch",
                              {"input_dim": 512, "hidden_dim": 32768, "top_k": 32, "abs_range": 1.5, "n_bits": 4},
                              _out_levels, _dict_residual))
register_sae(SAERegistryEntry("baseline_sae", BaselineSparseAutoencoder, Path("SAEs/baseline_sae_32768.pth"), "torch",
                              {"input_dim": 512, "hidden_dim": 32768}, _out_pair, _dict_linear))
# not in the reference's registry (its t_sae checkpoints are loaded ad hoc, utils/inspector.py:32-33)
register_sae(SAERegistryEntry("t_sae", TernarySparseAutoencoder, Path("SAEs/t_sae_32768.pth"), "torch",
                              {"input_dim": 512, "hidden_dim": 32768}, _out_pair, _dict_ternary))


def _checkpoint_root(root: Optional[Union[str, Path]]) -> Path:
    if root is not None:
        return Path(root)
    return Path(os.environ.get("QSAE_CHECKPOINT_ROOT", "."))


_ELEUTHER_KEYS = {"encoder.weight", "encoder.bias", "W_dec", "b_dec"}


def remap_eleuther_keys(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """EleutherAI sae-pythia checkpoints -> baseline_sae layout (:253-271): encoder.weight ->
    encoder.0.weight, W_dec [H, D] -> decoder.weight [D, H], b_dec -> decoder.bias."""
    if "encoder.0.weight" in state or not _ELEUTHER_KEYS.issubset(state.keys()):
        return state
    return OrderedDict({"encoder.0.weight": state["encoder.weight"], "encoder.0.bias": state["encoder.bias"],
                        "decoder.weight": state["W_dec"].t().contiguous(), "decoder.bias": state["b_dec"]})


def load_state_dict_file(path: Union[str, Path], checkpoint_format: str = "torch") -> Dict[str, torch.Tensor]:
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"Checkpoint not found: {path}")
    if checkpoint_format == "torch":
        state = torch.load(path, map_location="cpu", weights_only=True)
        for key in ("state_dict", "model_state_dict", "model"):       # tolerant unwrapping, as the reference's
            if isinstance(state, dict) and key in state and isinstance(state[key], dict):   # eval script does
                state = state[key]
        return state
    if checkpoint_format == "safetensors":
        from safetensors.torch import load_file

        return remap_eleuther_keys(load_file(str(path)))
    raise ValueError(f"Unsupported checkpoint format '{checkpoint_format}'.")


class SAEWrapper:
    """Uniform inference interface over the SAE variants (:280-337)."""

    def __init__(self, entry: SAERegistryEntry, model: nn.Module, device: Optional[DeviceLike]) -> None:
        self._entry = entry
        self.model = model
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.model.to(self.device)
        self.model.eval()

    def to(self, device: DeviceLike) -> "SAEWrapper":
        self.device = torch.device(device)
        self.model.to(self.device)
        return self

    def eval(self) -> "SAEWrapper":
        self.model.eval()
        return self

    @torch.no_grad()
    def __call__(self, batch: Any) -> Dict[str, Any]:
        x = _as_tensor(batch).to(self.device)
        return self._entry.forward_adapter(self.model, x)

    @torch.no_grad()
    def reconstruct(self, batch: Any) -> torch.Tensor:
        return self(batch)["reconstruction"]

    @torch.no_grad()
    def reconstruct_loader(self, dataloader: Iterable[Any], *, return_details: bool = False
                           ) -> Iterator[Union[torch.Tensor, Dict[str, Any]]]:
        for batch in dataloader:
            out = self(batch)
            yield out if return_details else out["reconstruction"]

    def decoder_dictionary(self, **options: Any) -> Dict[str, torch.Tensor]:
        return self._entry.decoder_getter(self.model, options)


def available_saes(checkpoint_root: Optional[Union[str, Path]] = None) -> Dict[str, Path]:
    root = _checkpoint_root(checkpoint_root)
    return {name: (e.checkpoint_path if e.checkpoint_path.is_absolute() else root / e.checkpoint_path)
            for name, e in SAE_REGISTRY.items()}


def load_sae(name: str, *, device: Optional[DeviceLike] = None, strict: bool = True,
             checkpoint_path: Optional[Union[str, Path]] = None, checkpoint_root: Optional[Union[str, Path]] = None,
             **kwargs_override: Any) -> SAEWrapper:
    """Instantiate a registered variant, restore its weights and wrap it for inference (:340-359)."""
    if name not in SAE_REGISTRY:
        raise KeyError(f"Unknown SAE '{name}'. Available: {list(SAE_REGISTRY)}")
    entry = SAE_REGISTRY[name]
    if kwargs_override:
        entry = replace(entry, kwargs={**entry.kwargs, **kwargs_override})
    path = Path(checkpoint_path) if checkpoint_path is not None else available_saes(checkpoint_root)[name]
    fmt = "safetensors" if path.suffix == ".safetensors" else entry.checkpoint_format
    state = load_state_dict_file(path, fmt)
    model = entry.constructor(**entry.kwargs)
    model.load_state_dict(state, strict=strict)
    return SAEWrapper(entry, model, device)
