"""Training-step pieces around the network (BASELINE configs[4]): the reference's loss
(model.py:167-217), its optimizer (torch.optim.AdamW, model.py:279-283) as one kernel over flat buffers, its warm-up /
cosine schedule (model.py:77-101) and the data-parallel gradient all-reduce (train.py:66-76 delegates it to Lightning's
DDP).  The backward of the network is NOT built: ``loss`` also returns the gradients with respect to the predictions --
the seeds a backward would start from -- and ``FusedAdamW`` takes gradients from wherever the caller computes them.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

import torch

from . import _cabi

__all__ = ["loss", "FusedAdamW", "lr_lambda", "allreduce_flat"]


def _ws(device):
    n = _cabi.load().es_loss_workspace_bytes()
    return torch.empty(n + 256, dtype=torch.uint8, device=device)


def loss(y_hat: Dict[str, torch.Tensor], y: Dict[str, torch.Tensor], x: Dict[str, torch.Tensor], with_grads: bool = False):
    """``EfficientSpeech.loss`` (model.py:167-209): returns (mel_loss, pitch_loss, energy_loss, duration_loss) as device
    scalars; with ``with_grads`` also a dict ``total`` (10 mel + 2 pitch + 2 energy + duration, model.py:215) and the
    gradients of the total with respect to ``mel`` / ``pitch`` / ``energy`` / ``duration`` predictions."""
    mel_pred = y_hat["mel"].contiguous()
    if not mel_pred.is_cuda:
        raise RuntimeError("efficientspeech_b200.training.loss: tensors must be on a CUDA device (no CPU fallback)")
    dev = mel_pred.device
    B, T, C = mel_pred.shape
    N = x["pitch"].shape[-1]
    f32 = dict(dtype=torch.float32, device=dev)

    def pred(k):
        return y_hat[k][:, :N].reshape(B, N).to(**f32).contiguous()               # model.py:190-191 slice + squeeze

    pp, ep, dp = pred("pitch"), pred("energy"), pred("duration")
    mel_t = y["mel"].to(**f32).contiguous()
    if tuple(mel_t.shape) != (B, T, C):
        raise RuntimeError(f"mel target {tuple(mel_t.shape)} does not match the prediction {(B, T, C)}")
    mel_len = x["mel_len"].to(device=dev, dtype=torch.int32).contiguous()
    pm = x.get("phoneme_mask")
    pm = None if pm is None else pm.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8)
    out = torch.empty(5, **f32)
    grads = {}
    if with_grads:
        grads = {"mel": torch.empty_like(mel_pred), "pitch": torch.empty(B, N, **f32), "energy": torch.empty(B, N, **f32),
                 "duration": torch.empty(B, N, **f32)}
    ws = _ws(dev)

    def ptr(t):
        return None if t is None else t.data_ptr()

    with torch.cuda.device(dev):
        _cabi.check(_cabi.load().es_loss(
            torch.cuda.current_stream(dev).cuda_stream, B, N, T, C, mel_pred.data_ptr(), mel_t.data_ptr(), mel_len.data_ptr(),
            pp.data_ptr(), ep.data_ptr(), dp.data_ptr(), x["pitch"].to(**f32).contiguous().data_ptr(),
            x["energy"].to(**f32).contiguous().data_ptr(), x["duration"].to(device=dev, dtype=torch.int32).contiguous().data_ptr(),
            ptr(pm), out.data_ptr(), ptr(grads.get("mel")), ptr(grads.get("pitch")), ptr(grads.get("energy")),
            ptr(grads.get("duration")), ws.data_ptr(), ws.numel()))
    losses = (out[1], out[2], out[3], out[4])
    if with_grads:
        return losses, {"total": out[0], **grads}
    return losses


def lr_lambda(current_step: int, warmup_steps: int, total_steps: int, min_lr: float = 0.0) -> float:
    """model.py:90-98 (get_lr_scheduler): linear warm-up, then cosine decay."""
    if current_step < warmup_steps:
        return float(current_step) / float(max(1, warmup_steps))
    progress = float(current_step - warmup_steps) / float(max(1, total_steps - warmup_steps))
    return max(min_lr, 0.5 * (1.0 + math.cos(math.pi * progress)))


def allreduce_flat(flat_grad: torch.Tensor, average: bool = True) -> torch.Tensor:
    """ONE collective for the whole gradient (NCCL over NVLink on GPUs, gloo in the CPU tests): what DDP's bucketed
    all-reduce (train.py:66-76, strategy "ddp") amounts to for a 0.27-4 M parameter model."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        if average:
            flat_grad.div_(dist.get_world_size())
    return flat_grad


class FusedAdamW:
    """torch.optim.AdamW(params, lr, weight_decay) (model.py:280) over ONE flat fp32 buffer: the parameters are
    re-pointed at views of it, so ``step`` is a single kernel (es_adamw_step) and the gradient all-reduce a single
    collective.  ``step(lr_scale)`` takes the scheduler's multiplier (lr_lambda)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        self.params = [p for p in params if p.requires_grad]
        if not self.params or not self.params[0].is_cuda:
            raise RuntimeError("FusedAdamW needs CUDA parameters (no CPU fallback)")
        dev = self.params[0].device
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                p.grad = self.flat_grad[off:off + k].view_as(p)
                off += k
        self.t = 0

    def zero_grad(self):
        self.flat_grad.zero_()

    def step(self, lr_scale: float = 1.0, allreduce: bool = False):
        if allreduce:
            allreduce_flat(self.flat_grad)
        self.t += 1
        lr = self.lr * lr_scale
        b1, b2 = self.betas
        step_size = lr / (1.0 - b1 ** self.t)
        bc2_sqrt = math.sqrt(1.0 - b2 ** self.t)
        dev = self.flat.device
        with torch.cuda.device(dev):
            _cabi.check(_cabi.load().es_adamw_step(torch.cuda.current_stream(dev).cuda_stream, self.flat.numel(), self.flat.data_ptr(),
                                                   self.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                                   lr, b1, b2, self.eps, self.weight_decay, step_size, bc2_sqrt))
        for p in self.params:                     # the kernel wrote through raw pointers: tell torch the tensors changed, the
            torch.autograd.graph.increment_version(p)   # forward re-packs its weight images when a version moves
        return lr
