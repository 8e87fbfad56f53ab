"""Inference registry mirroring the reference's `quantized_sae.inference` package."""
from .framework import (SAE_REGISTRY, SAERegistryEntry, SAEWrapper, available_saes, load_sae, load_state_dict_file,
                        register_sae)

__all__ = ["SAE_REGISTRY", "SAERegistryEntry", "SAEWrapper", "available_saes", "load_sae", "load_state_dict_file",
           "register_sae"]
