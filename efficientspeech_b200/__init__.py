"""efficientspeech_b200 -- Blackwell-native (sm_100a) EfficientSpeech acoustic forward path.

Drop-in for the reference's ``layers`` package (layers/__init__.py:1):

    from efficientspeech_b200 import PhonemeEncoder, MelDecoder, Phoneme2Mel

The compute lives in ``libes_b200.so`` (hand-written CUDA behind the C ABI of
``include/es_b200.h``); this package is the thin host mirror of the reference's module API.
"""
from .config import ESConfig, VARIANTS, variant, LJSPEECH_PITCH_STATS, LJSPEECH_ENERGY_STATS  # noqa: F401
from .modules import (AcousticDecoder, Encoder, FeatureUpsampler, Fuse, MelDecoder, MixFFN,  # noqa: F401
                      Phoneme2Mel, PhonemeEncoder, SelfAttention, mel_to_half)

from .collate import collate, collate_flat  # noqa: F401
from . import hifigan  # noqa: F401  (drop-in for the reference's hifigan package)
from . import training  # noqa: F401
from . import text  # noqa: F401
from .pipeline import synthesize  # noqa: F401

__version__ = "0.2.0"


def build_model(cfg_or_name="tiny") -> "Phoneme2Mel":
    """Construct Phoneme2Mel(PhonemeEncoder, MelDecoder) exactly as model.py:132-147 does."""
    cfg = variant(cfg_or_name) if isinstance(cfg_or_name, str) else cfg_or_name
    enc = PhonemeEncoder(pitch_stats=cfg.pitch_stats, energy_stats=cfg.energy_stats, depth=cfg.depth,
                         reduction=cfg.reduction, head=cfg.head, embed_dim=cfg.embed_dim,
                         kernel_size=cfg.kernel_size, expansion=cfg.expansion)
    dec = MelDecoder(dim=cfg.embed_dim // cfg.reduction, kernel_size=cfg.decoder_kernel_size,
                     n_mel_channels=cfg.n_mel, n_blocks=cfg.n_blocks, block_depth=cfg.block_depth)
    return Phoneme2Mel(enc, dec)


def load_numpy_state(model, state) -> None:
    """load_state_dict(strict=True) from a {name: numpy array} dict in the reference layout."""
    import torch
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in state.items()}, strict=True)
