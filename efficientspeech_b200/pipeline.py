"""Text -> waveform with every stage of this repository chained on the device: the flow of the reference's
``synthesize.py:66-110`` / ``demo.py`` + ``EfficientSpeech.predict_step`` (``model.py:159-164``).

    ids (CPU strings, text.py) -> collate (es_collate) -> Phoneme2Mel (free running) -> HiFi-GAN -> wav

The mel never leaves the GPU: the vocoder reads the acoustic model's [B, T, 80] output through strides.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import text as _text
from .collate import collate

__all__ = ["synthesize"]

HOP_LENGTH = 256          # config/LJSpeech/preprocess.yaml:16 -- samples per mel frame (= the vocoder's total upsampling)


def synthesize(texts: Sequence[str], lexicon: Dict[str, List[str]], g2p: Optional[Callable], phoneme2mel, vocoder,
               preprocess_config: dict, device="cuda") -> Dict[str, torch.Tensor]:
    """Returns ``wav`` [B, L] fp32 in (-1, 1) (rows in the ORDER OF ``texts``), ``wav_len`` [B] (= mel_len * hop: samples
    past it are the vocoder's response to zero-padded frames), ``mel`` [B, T, 80], ``mel_len`` and ``duration``."""
    items = [{"phoneme": _text.text2phoneme(lexicon, g2p, t, preprocess_config)} for t in texts]
    if any(len(it["phoneme"]) < 2 for it in items):
        raise ValueError("every text must yield at least two phonemes")
    batch = collate(items, device)                              # sorted by length, padded, masks -- on the device
    with torch.no_grad():
        mel, mel_len, duration = phoneme2mel({"phoneme": batch["phoneme"], "phoneme_mask": batch["phoneme_mask"]}, train=False)
        wav = vocoder(mel.transpose(1, 2)).squeeze(1)           # model.py:160-161; the transpose stays a view
    inv = torch.empty_like(batch["perm"])
    inv[batch["perm"].long()] = torch.arange(len(items), device=inv.device, dtype=inv.dtype)
    idx = inv.long()                                            # back to the caller's order
    return {"wav": wav[idx], "wav_len": (mel_len * HOP_LENGTH)[idx], "mel": mel[idx], "mel_len": mel_len[idx],
            "duration": duration[idx], "phoneme_len": batch["phoneme_len"][idx]}
