"""Multi-GPU plumbing: utterances are independent (only LayerNorm, no batch statistics;
SURVEY.md section 8e), so the batch shards contiguously across ranks with NO data-path
collective.  The only communication is one broadcast of the weights at start-up."""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist


def broadcast_weights(model: torch.nn.Module, src: int = 0) -> None:
    """Single flat broadcast (NCCL over NVLink on GPUs, gloo in the CPU tests) of every
    parameter in state-dict order; in-place copies bump the parameter versions so the packed
    weight image is rebuilt on the next forward."""
    params = [p for p in model.parameters()]
    if not params:
        return
    flat = torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params])
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n


def shard_bounds(batch_size: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of the batch owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(batch_size, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, object], rank: int, world: int) -> Dict[str, object]:
    """Slice every per-utterance tensor along dim 0.  The phoneme dimension is NOT re-padded:
    the unmasked attention makes each utterance depend on its own padding to the GLOBAL N_max
    (SURVEY.md H3), so shards keep the full batch's phoneme length to stay bit-comparable with
    the unsharded run, and teacher-forced shards keep the global frame count (``max_mel_len``) for the same reason
    (SURVEY.md H4).  Free-running shards derive T from their own predicted durations, as the reference would.
    The shard carries ``global_batch_size``: the reference drops the phoneme mask for a one-utterance batch
    (networks.py:338), and a shard of size 1 must keep following the GLOBAL batch's policy."""
    B = batch["phoneme"].shape[0]
    lo, hi = shard_bounds(B, rank, world)
    out = {}
    for k, v in batch.items():
        out[k] = v[lo:hi] if hasattr(v, "shape") and len(v.shape) > 0 and v.shape[0] == B else v
    out["global_batch_size"] = int(batch.get("global_batch_size", B))
    if "mel_len" in batch and "max_mel_len" not in batch:
        # teacher-forced: padded frames flow through the decoder, so the frames next to an utterance's end
        # depend on whether pad-row frames or the convolution's zero padding follow them -- shards keep the global T
        out["max_mel_len"] = int(max(batch["mel_len"]))
    return out
