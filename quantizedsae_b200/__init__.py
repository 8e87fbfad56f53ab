"""quantizedsae_b200 -- B200 (sm_100a) forward hot path of ASSERT-KTH/QuantizedSAE.

Python mirrors the reference's nn.Module API (same constructors, state_dict layout and forward
return structure); all arithmetic runs in libqsae_b200.so (hand-written CUDA behind the C ABI of
include/qsae_b200.h). Importing this package does not need a GPU; running a forward does.
"""
from .sae import (BaselineSparseAutoencoder, BinarySAE, QuantizedMatryoshkaDecoder, QuantizedMatryoshkaSAE,
                  ResidualQuantizedSAE, SparseAutoencoder, STEWeights, TernarySparseAutoencoder, binary_decoder)
from .sparse import SparseLatents

__version__ = "0.1.0"
__all__ = ["BaselineSparseAutoencoder", "BinarySAE", "SparseAutoencoder", "binary_decoder", "SparseLatents",
           "QuantizedMatryoshkaDecoder", "QuantizedMatryoshkaSAE", "STEWeights", "TernarySparseAutoencoder", "ResidualQuantizedSAE"]
