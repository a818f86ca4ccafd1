"""Seeded synthetic batches in the reference's batch-dict layout (datamodule.py:29-76).

Used by bench.py (the headline workload) and by the parity tests; numpy PCG64 so the same
seed gives the same tensors on every machine.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .config import ESConfig


def make_batch(cfg: ESConfig, B: int, N: int, seed: int = 0, ragged: bool = False,
               fixed_duration: Optional[int] = 6, min_len: Optional[int] = None,
               max_dur: int = 11) -> Dict[str, np.ndarray]:
    """phoneme ~ U{1..n_symbols-2} int32 [B,N]; mask True = padding (utils/tools.py:43-51).

    ``ragged=False``: every utterance has N phonemes and (if ``fixed_duration``) the same
    duration per phoneme (the headline workload: durations all 6 -> T = 6N).
    ``ragged=True``: lengths ~ U{min_len..N}, durations ~ U{0..max_dur} (zeros included), pads
    zeroed, batch sorted by decreasing length like the reference collate_fn (datamodule.py:53).
    pitch/energy ~ N(0,1) clipped into the stats range (teacher-forcing targets).
    """
    rng = np.random.default_rng(np.random.SeedSequence([0xBA7C4, seed, B, N, int(ragged)]))
    if ragged:
        lo = min_len if min_len is not None else max(1, N // 2)
        lens = np.sort(rng.integers(lo, N + 1, size=B))[::-1].copy()
        lens[0] = N
    else:
        lens = np.full(B, N)
    ids = rng.integers(1, cfg.n_symbols - 1, size=(B, N)).astype(np.int32)
    mask = np.arange(N)[None, :] >= lens[:, None]
    ids[mask] = 0
    if ragged or fixed_duration is None:
        dur = rng.integers(0, max_dur + 1, size=(B, N)).astype(np.int32)
    else:
        dur = np.full((B, N), fixed_duration, dtype=np.int32)
    dur[mask] = 0
    pitch = np.clip(rng.standard_normal((B, N)), cfg.pitch_stats[0], cfg.pitch_stats[1]).astype(np.float32)
    energy = np.clip(rng.standard_normal((B, N)), cfg.energy_stats[0], cfg.energy_stats[1]).astype(np.float32)
    pitch[mask] = 0
    energy[mask] = 0
    return {"phoneme": ids, "phoneme_mask": mask, "phoneme_len": lens.astype(np.int32),
            "pitch": pitch, "energy": energy, "duration": dur,
            "mel_len": dur.sum(axis=1).astype(np.int32)}
