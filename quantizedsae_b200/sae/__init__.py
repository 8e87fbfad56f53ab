"""nn.Module mirror of the reference's `quantized_sae.sae` package (same class names)."""
from .base import SparseAutoencoder
from .baseline import BaselineSparseAutoencoder
from .binary import BinarySAE, binary_decoder
from .quantized_matryoshka import QuantizedMatryoshkaDecoder, QuantizedMatryoshkaSAE
from .residual_quantized import ResidualQuantizedSAE
from .ternary import STEWeights, TernarySparseAutoencoder

__all__ = ["SparseAutoencoder", "BaselineSparseAutoencoder", "BinarySAE", "binary_decoder",
           "QuantizedMatryoshkaDecoder", "QuantizedMatryoshkaSAE", "STEWeights", "TernarySparseAutoencoder", "ResidualQuantizedSAE"]
