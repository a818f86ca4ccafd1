"""Batch collation on the device: the reference's ``LJSpeechDataModule.collate_fn`` (datamodule.py:29-76) and
``get_mask_from_lengths`` (utils/tools.py:43-51) for the acoustic model's inputs, as two CUDA kernels (es_collate.cu).

``collate(items, device)`` takes the same list of per-utterance dicts the reference's dataset yields (numpy arrays:
``phoneme`` int, optional ``pitch`` / ``energy`` float and ``duration`` int) and returns the reference's batch dict with
every tensor on the device: ``phoneme`` int32 [B,N], ``phoneme_len``, ``phoneme_mask`` bool (True = padding), and for
training batches ``pitch``, ``energy``, ``duration``, ``mel_len``; plus ``perm`` (row r = items[perm[r]]).  The ragged
arrays cross PCIe once, concatenated; sorting, padding and masks happen on the GPU.  ``text`` / ``mel`` targets are not
inputs of the acoustic forward path and stay with the caller.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _cabi

__all__ = ["collate", "collate_flat"]


def collate_flat(offsets: torch.Tensor, phoneme: torch.Tensor, n_max: int, pitch=None, energy=None, duration=None) -> Dict[str, torch.Tensor]:
    """Device-resident ragged arrays (CSR ``offsets`` [B+1] int32) -> padded, length-sorted batch (es_collate)."""
    if not offsets.is_cuda:
        raise RuntimeError("efficientspeech_b200.collate: inputs must be CUDA tensors (no CPU fallback)")
    dev = offsets.device
    B = offsets.numel() - 1
    i32, f32 = dict(dtype=torch.int32, device=dev), dict(dtype=torch.float32, device=dev)
    out = {"perm": torch.empty(B, **i32), "phoneme": torch.empty(B, n_max, **i32),
           "phoneme_mask": torch.empty(B, n_max, dtype=torch.uint8, device=dev), "phoneme_len": torch.empty(B, **i32)}
    if pitch is not None:
        out["pitch"] = torch.empty(B, n_max, **f32)
    if energy is not None:
        out["energy"] = torch.empty(B, n_max, **f32)
    if duration is not None:
        out["duration"] = torch.empty(B, n_max, **i32)
        out["mel_len"] = torch.empty(B, **i32)

    def ptr(t):
        return None if t is None else t.data_ptr()

    with torch.cuda.device(dev):
        _cabi.check(_cabi.load().es_collate(
            torch.cuda.current_stream(dev).cuda_stream, B, int(n_max), offsets.data_ptr(), phoneme.data_ptr(), ptr(pitch),
            ptr(energy), ptr(duration), out["perm"].data_ptr(), out["phoneme"].data_ptr(), out["phoneme_mask"].data_ptr(),
            out["phoneme_len"].data_ptr(), ptr(out.get("pitch")), ptr(out.get("energy")), ptr(out.get("duration")),
            ptr(out.get("mel_len"))))
    out["phoneme_mask"] = out["phoneme_mask"].view(torch.bool)
    return out


def collate(items: Sequence[Dict[str, np.ndarray]], device) -> Dict[str, torch.Tensor]:
    """The reference's collate_fn for the acoustic inputs, on `device`."""
    device = torch.device(device)
    lens = np.array([len(it["phoneme"]) for it in items], dtype=np.int64)
    offsets = np.zeros(len(items) + 1, dtype=np.int32)
    np.cumsum(lens, out=offsets[1:])

    def flat(key, dtype):
        if any(key not in it for it in items):
            return None
        a = np.concatenate([np.asarray(it[key]).astype(dtype, copy=False) for it in items])
        return torch.from_numpy(a).pin_memory().to(device, non_blocking=True)

    return collate_flat(torch.from_numpy(offsets).to(device), flat("phoneme", np.int32), int(lens.max()),
                        flat("pitch", np.float32), flat("energy", np.float32), flat("duration", np.int32))
