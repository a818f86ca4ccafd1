"""Training-step pieces around the network (BASELINE configs[4]): the reference's loss
(model.py:167-217), its optimizer (torch.optim.AdamW, model.py:279-283) as one kernel over flat buffers, its warm-up /
cosine schedule (model.py:77-101) and the data-parallel gradient all-reduce (train.py:66-76 delegates it to Lightning's
DDP), plus the differentiable forward of the network itself: ``forward_train`` restates ``Phoneme2Mel.forward(x,
train=True)`` (layers/networks.py:336-434) operator by operator over ``train_ops`` (this repository's kernels with
their adjoints; torch.autograd only keeps the tape), and ``TrainStep`` chains forward, loss, backward, all-reduce and
AdamW into the step ``EfficientSpeech.training_step`` + Lightning perform (model.py:211-217).

The inference kernels (one fused tcgen05 kernel per layer, nothing kept) are a different code path; the training path
keeps every intermediate and is a first correct version built from generic fp32 kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

import torch

from . import _cabi
from . import train_ops as ops

__all__ = ["loss", "FusedAdamW", "lr_lambda", "allreduce_flat", "forward_train", "TrainStep"]


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
        # every parameter starts on a 16-byte boundary of the flat buffer (kernels read weights with 128-bit loads when
        # they can); the padding elements are zero and stay zero (zero gradient, zero moments)
        n = sum(-(-p.numel() // 4) * 4 for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
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
                off += -(-k // 4) * 4
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


# ------------------------------------------------------------------------------------------------------------------
# differentiable forward (teacher-forced), layers/networks.py operator by operator
# ------------------------------------------------------------------------------------------------------------------
def _pool_mask(mask: torch.Tensor, pool: int) -> torch.Tensor:
    """layers/blocks.py:52-58: a pooled position is padding when ANY of its `pool` inputs is (tail padded with True)."""
    if pool <= 1:
        return mask
    B, n = mask.shape
    pad = (-n) % pool
    if pad:
        mask = torch.cat([mask, torch.ones(B, pad, dtype=torch.bool, device=mask.device)], dim=1)
    return mask.view(B, -1, pool).any(dim=-1)


def _acoustic_decoder(dec, fused):
    """AcousticDecoder.forward (layers/networks.py:151-165).  Note what the reference does: the scalar head reads the
    ReLU'd conv2 output, NOT norm2's; norm2 only shapes the duration decoder's feature output."""
    c1, c2 = dec.conv1[0], dec.conv2[0]
    y = ops.act(ops.conv1d(fused, c1.weight, c1.bias, 1, 1), ops.ACT_RELU)
    y = ops.act(ops.layernorm(y, dec.norm1.weight, dec.norm1.bias), ops.ACT_RELU)
    y = ops.act(ops.conv1d(y, c2.weight, c2.bias, 1, 1), ops.ACT_RELU)
    feats = ops.layernorm(y, dec.norm2.weight, dec.norm2.bias) if dec.duration else None
    out = ops.linear(y, dec.linear.weight, dec.linear.bias)
    if dec.duration:
        out = ops.act(out, ops.ACT_RELU)
    return out, feats


def forward_train(model, x: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``Phoneme2Mel.forward(x, train=True)`` with a tape: returns the reference's dict (``mel`` [B,T,80], ``pitch`` /
    ``energy`` / ``duration`` [B,N,1], ``mel_len`` [B] int32, ``features`` [B,T,4d], ``masks`` [B,T,4d] bool or None).
    ``model`` is this package's ``Phoneme2Mel`` (its parameters are the leaves).  ``x["max_mel_len"]`` (python int)
    spares the one host sync of networks.py:344."""
    pe, md = model.encoder, model.decoder
    enc = pe.encoder
    phoneme = x["phoneme"]
    if not phoneme.is_cuda:
        raise RuntimeError("efficientspeech_b200.training.forward_train: tensors must be on a CUDA device (no CPU fallback)")
    dev = phoneme.device
    B, N = phoneme.shape
    gB = int(x.get("global_batch_size", B))
    pmask = x["phoneme_mask"].to(device=dev, dtype=torch.bool) if gB > 1 else None            # networks.py:338

    # ---- Encoder.forward (networks.py:52-87)
    h = ops.embedding(phoneme.to(torch.int32), enc.embed.weight, padding_idx=0)
    feats = []
    for conv3, conv1, attn, ffn, norm1, norm2 in enc.attn_blocks:
        h = ops.conv1d(h, conv3.weight, None, conv3.stride[0], conv3.padding[0])
        h = ops.linear(h, conv1.weight.squeeze(-1), None)
        n_l, C = h.shape[1], h.shape[2]
        mask_l = None
        if pmask is not None:
            mask_l = _pool_mask(pmask, int(round(N / n_l)))[:, :n_l].contiguous()
        H = attn.num_heads
        qkv = ops.linear(h, attn.qkv.weight, attn.qkv.bias)
        a = ops.attention_core(qkv, H, C, float((C // H) ** -0.5))
        y = ops.linear(a, attn.proj.weight, attn.proj.bias)
        h = ops.layernorm(ops.add(y, h), norm1.weight, norm1.bias)
        if mask_l is not None:
            h = ops.mask_rows(h, mask_l)
        f = ops.linear(h, ffn.mlp1.weight, ffn.mlp1.bias)
        f = ops.act(ops.conv1d(f, ffn.conv.weight, ffn.conv.bias, 1, 1), ops.ACT_GELU)
        f = ops.linear(f, ffn.mlp2.weight, ffn.mlp2.bias)
        h = ops.layernorm(ops.add(f, h), norm2.weight, norm2.bias)
        if mask_l is not None:
            h = ops.mask_rows(h, mask_l)
        feats.append(h)

    # ---- Fuse.forward (networks.py:190-219): project, upsample back to N positions, concatenate, fuse
    outs = []
    for feat, (mlp, up) in zip(feats, pe.fuse.mlps):
        z = ops.linear(feat, mlp.weight, mlp.bias)
        if isinstance(up, torch.nn.ConvTranspose1d):
            z = ops.conv_transpose1d(z, up.weight, up.bias, up.stride[0], N)
        outs.append(z)
    fused = ops.linear(ops.concat_channels(outs) if len(outs) > 1 else outs[0], pe.fuse.fuse.weight, pe.fuse.fuse.bias)
    if pmask is not None:
        fused = ops.mask_rows(fused, pmask)

    # ---- predictors, teacher-forced embeddings (networks.py:346-371)
    f32 = dict(device=dev, dtype=torch.float32)
    pitch_pred, _ = _acoustic_decoder(pe.pitch_decoder, fused)
    pf = ops.embedding(ops.bucketize(x["pitch"].to(**f32), pe.pitch_decoder.pitch_bins), pe.pitch_decoder.pitch_embedding.weight)
    energy_pred, _ = _acoustic_decoder(pe.energy_decoder, fused)
    ef = ops.embedding(ops.bucketize(x["energy"].to(**f32), pe.energy_decoder.energy_bins), pe.energy_decoder.energy_embedding.weight)
    dur_pred, df = _acoustic_decoder(pe.duration_decoder, fused)
    if pmask is not None:
        pf, ef, df = ops.mask_rows(pf, pmask), ops.mask_rows(ef, pmask), ops.mask_rows(df, pmask)
    fused4 = ops.concat_channels([fused, pf, ef, df])

    # ---- FeatureUpsampler with the TARGET durations (networks.py:380-389)
    dur = x["duration"].to(device=dev, dtype=torch.int32)
    if pmask is not None:
        dur = dur.masked_fill(pmask, 0)
    cum = torch.cumsum(dur.clamp(min=0), dim=1, dtype=torch.int32).contiguous()
    T = x.get("max_mel_len")
    T = int(x["mel_len"].max().item()) if T is None else int(T)
    features = ops.expand_rows(fused4, cum, T)
    mel_len_pred = cum[:, -1].contiguous()
    frame_mask = None
    if pmask is not None:
        frame_mask = torch.arange(T, device=dev, dtype=torch.int32)[None, :] >= mel_len_pred[:, None]

    # ---- MelDecoder.forward (networks.py:290-304)
    skip = ops.layernorm(ops.act(ops.linear(features, md.proj[0].weight, md.proj[0].bias), ops.ACT_TANH),
                         md.proj[2].weight, md.proj[2].bias)
    for convs, skip_norm in md.blocks:
        z = skip
        for conv, norm in convs:
            z = ops.dwconv1d(z, conv[0].weight, conv[0].bias)
            z = ops.act(ops.linear(z, conv[1].weight.squeeze(-1), conv[1].bias), ops.ACT_TANH)
            z = ops.layernorm(z, norm.weight, norm.bias)
        skip = ops.layernorm(ops.add(z, skip), skip_norm.weight, skip_norm.bias)
    mel = ops.linear(skip, md.mel_linear.weight, md.mel_linear.bias)
    if frame_mask is not None and gB > 1:
        mel = ops.mask_rows(mel, frame_mask)                                                   # networks.py:424-427
    # the reference returns the [B,T,4d] bool tensor; the same values as a broadcast view (None for a single utterance)
    masks = None if frame_mask is None else frame_mask.unsqueeze(-1).expand(B, T, fused4.shape[-1])
    return {"pitch": pitch_pred, "energy": energy_pred, "duration": dur_pred, "mel_len": mel_len_pred, "features": features,
            "masks": masks, "mel": mel}


class TrainStep:
    """One optimisation step of the reference's training loop (model.py:211-217 + torch.optim.AdamW + the warm-up /
    cosine schedule, DDP's gradient average when torch.distributed is initialised): ``losses = step(x, y)``.

    ``forward_train`` -> ``loss`` (which also produces the gradient seeds) -> backward through the tape -> ONE flat
    all-reduce -> ONE AdamW kernel.  Returns the five device scalars (total, mel, pitch, energy, duration); nothing
    synchronises the host.  The tensor-core GEMMs bound every mbarrier wait and raise a device flag instead of hanging;
    ``Phoneme2Mel.check_async_errors()`` (one stream sync) reads it -- call it wherever the loop synchronises anyway
    (logging a loss value, an epoch end)."""

    def __init__(self, model, lr: float = 1e-3, weight_decay: float = 1e-6, warmup_steps: int = 50, total_steps: int = 5000,
                 use_graphs: bool = False):
        self.model = model
        self.use_graphs = bool(use_graphs)
        self._graphs, self._pool = {}, None
        # norm2 of the pitch / energy predictors never reaches an output (layers/networks.py:160-161: the head reads the
        # tensor BEFORE norm2, only the duration predictor's features use it): torch leaves their .grad None and AdamW
        # skips them, weight decay included -- so they stay out of the flat buffer here
        unused = ("pitch_decoder.norm2.", "energy_decoder.norm2.")
        self.opt = FusedAdamW([p for n, p in model.named_parameters() if not any(u in n for u in unused)],
                              lr=lr, weight_decay=weight_decay)
        self.warmup_steps, self.total_steps = int(warmup_steps), int(total_steps)
        self.n = 0

    def forward_backward(self, x: Dict[str, torch.Tensor], y: Dict[str, torch.Tensor]):
        """Zero the flat gradient, run forward, loss and backward; returns (total, mel, pitch, energy, duration) device
        scalars.  No host synchronisation when ``x["max_mel_len"]`` is given, so the whole call can be graph-captured."""
        self.opt.zero_grad()
        pred = forward_train(self.model, x)
        losses, g = loss(pred, y, x, with_grads=True)
        B, N = x["phoneme"].shape
        torch.autograd.backward(
            [pred["mel"], pred["pitch"], pred["energy"], pred["duration"]],
            [g["mel"], g["pitch"].view(B, N, 1), g["energy"].view(B, N, 1), g["duration"].view(B, N, 1)])
        return (g["total"],) + tuple(losses)

    def optimizer_step(self):
        import torch.distributed as dist
        ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        # LambdaLR evaluates lr_lambda(0) at construction, so optimiser step k (0-based) runs at lr_lambda(k)
        self.opt.step(lr_scale=lr_lambda(self.n, self.warmup_steps, self.total_steps), allreduce=ddp)
        self.n += 1

    def __call__(self, x: Dict[str, torch.Tensor], y: Dict[str, torch.Tensor]):
        if self.use_graphs:
            return self._graphed(x, y)
        out = self.forward_backward(x, y)
        self.optimizer_step()
        return out

    # ---- CUDA-graph replay: a step is ~450 launches of 2-40 us kernels, and enqueueing them from Python takes longer
    # than running them.  forward + loss + backward are captured once per batch geometry (B, N, T) and replayed; the
    # all-reduce and the AdamW kernel stay outside the graph (the learning rate changes every step and NCCL keeps its own
    # stream semantics).  Callers bucket T (pad the mel target, pass the bucket as x["max_mel_len"]) to bound the number
    # of graphs; all graphs share one memory pool (replays are sequential; outputs are copied out of the pool at once).
    def _graphed(self, x, y):
        if "max_mel_len" not in x:
            raise RuntimeError("TrainStep(use_graphs=True) needs x['max_mel_len'] (python int): a captured step cannot "
                               "read it back from the device")
        key = (tuple(x["phoneme"].shape), int(x["max_mel_len"]), int(x.get("global_batch_size", x["phoneme"].shape[0])))
        g = self._graphs.get(key)
        if g is None:
            g = self._capture(x, y)
            self._graphs[key] = g
        sx, sy, graph, out = g
        for k, v in x.items():
            if torch.is_tensor(v):
                sx[k].copy_(v, non_blocking=True)
        sy["mel"].copy_(y["mel"], non_blocking=True)
        graph.replay()
        # the graphs share one memory pool: a temporary of a graph captured EARLIER may live where this graph's output
        # was allocated later, so the five scalars are copied out before any other graph replays
        res = torch.stack(out)
        self.optimizer_step()
        return tuple(res.unbind(0))

    def _capture(self, x, y):
        dev = x["phoneme"].device
        sx = {k: (v.detach().to(dev).clone() if torch.is_tensor(v) else v) for k, v in x.items()}
        sy = {"mel": y["mel"].detach().to(dev).clone()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # warm-up off the capture: lazy allocations, function attributes
            for _ in range(2):
                self.forward_backward(sx, sy)
        torch.cuda.current_stream(dev).wait_stream(side)
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, pool=self._pool):
            out = self.forward_backward(sx, sy)
        return sx, sy, graph, out
