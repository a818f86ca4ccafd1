"""Model geometry of the EfficientSpeech acoustic path.

Mirrors the constructor arguments of the reference modules (layers/networks.py:18-30,
:264-270, :310-333; CLI defaults utils/tools.py:354-389) and derives every dimension the
kernels need.  The three named variants are the same code with different arguments
(README.md:173-196).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

# len(text.symbols.symbols) + 1 (layers/networks.py:32, text/symbols.py:22-31): 152 symbols
# plus the padding row 0.
N_SYMBOLS = 153

# preprocessed_data/LJSpeech/stats.json[:2] (read at model.py:127-130): (min, max) of the
# normalised pitch / energy; they become the bucketize bins (layers/networks.py:109-122).
LJSPEECH_PITCH_STATS = (-2.9170793047299672, 11.391254536985771)
LJSPEECH_ENERGY_STATS = (-1.431044578552246, 8.184337615966797)


@dataclass(frozen=True)
class ESConfig:
    depth: int = 2              # encoder pyramid levels; only 2 is functional upstream
    reduction: int = 4
    head: int = 1
    embed_dim: int = 128
    kernel_size: int = 3        # encoder merge-conv / fuse upsample kernel
    expansion: int = 1          # MixFFN hidden multiplier
    n_blocks: int = 2           # decoder blocks
    block_depth: int = 2        # depthwise-separable layers per decoder block
    decoder_kernel_size: int = 5
    n_mel: int = 80
    n_symbols: int = N_SYMBOLS
    pitch_stats: Tuple[float, float] = LJSPEECH_PITCH_STATS
    energy_stats: Tuple[float, float] = LJSPEECH_ENERGY_STATS

    # ---- derived (layers/networks.py:22-30) ----
    @property
    def dim(self) -> int:                       # "d": fuse / predictor width
        return self.embed_dim // self.reduction

    @property
    def enc_dims_in(self) -> List[int]:
        return [self.embed_dim] + [self.dim * 2 ** i for i in range(self.depth - 1)]

    @property
    def enc_dims(self) -> List[int]:
        return [self.dim * 2 ** i for i in range(self.depth)]

    @property
    def enc_heads(self) -> List[int]:
        return [self.head * (i + 1) for i in range(self.depth)]

    @property
    def enc_kernels(self) -> List[int]:
        return [self.kernel_size - (2 if i > 0 else 0) for i in range(self.depth)]

    @property
    def enc_strides(self) -> List[int]:
        return [1] + [2] * (self.depth - 1)

    @property
    def dx4(self) -> int:                       # decoder input width (4 feature groups)
        return 4 * self.dim

    @property
    def dx2(self) -> int:                       # decoder hidden width (layers/networks.py:269)
        return min(4 * self.dim, 256)

    @property
    def n_dec_layers(self) -> int:
        return self.n_blocks * self.block_depth

    def validate(self) -> None:
        if self.depth != 2:
            raise ValueError("only encoder depth 2 is supported (the reference's Fuse cannot "
                             "reach the input length for depth > 2)")
        if self.kernel_size not in (3, 5):
            raise ValueError("encoder kernel_size must be 3 or 5")
        if self.decoder_kernel_size % 2 != 1 or self.decoder_kernel_size > 9:
            raise ValueError("decoder_kernel_size must be odd and <= 9")
        if self.embed_dim % self.reduction:
            raise ValueError("embed_dim must be divisible by reduction")
        if self.dim % 32 or self.dim > 128:
            raise ValueError("dim = embed_dim // reduction must be 32, 64, 96 or 128")


VARIANTS = {
    # README.md:173-196 / utils/tools.py:354-389
    "tiny": ESConfig(),
    "small": ESConfig(n_blocks=3, reduction=2),
    "base": ESConfig(head=2, reduction=1, expansion=2, kernel_size=5, n_blocks=3, block_depth=3),
}


def variant(name: str) -> ESConfig:
    return VARIANTS[name]
