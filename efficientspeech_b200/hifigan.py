"""Host-side mirror of the reference's ``hifigan`` package (hifigan/__init__.py, hifigan/models.py:84-134), backed by
``es_hifigan_forward`` (csrc/es_hifigan.cu).

``Generator(h)`` keeps the reference's constructor (``h``: the AttrDict of hifigan/LJ_V2/config.json), sub-module names
and weight-norm parametrisation, so ``model.py:23-48``'s ``get_hifigan`` works unchanged:

    vocoder = hifigan.Generator(config); vocoder.load_state_dict(ckpt["generator"]); vocoder.eval()
    vocoder.remove_weight_norm()

The torch sub-modules only HOLD parameters; ``forward(mel [B, 80, T]) -> wav [B, 1, 256 T]`` packs the effective conv
weights once per weight version and runs the CUDA kernels on the current stream.  A mel that is the transposed view of
the acoustic model's [B, T, 80] output is consumed in place (strides), no copy.  ResBlock1 configurations only.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch
from torch import nn
from torch.nn import Conv1d, ConvTranspose1d
from torch.nn.utils import remove_weight_norm, weight_norm

from . import _cabi

__all__ = ["AttrDict", "Generator", "ResBlock1", "LRELU_SLOPE"]

LRELU_SLOPE = 0.1


class AttrDict(dict):
    """hifigan/__init__.py:4-7."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


def _init_weights(m, mean=0.0, std=0.01):
    if m.__class__.__name__.find("Conv") != -1:
        m.weight.data.normal_(mean, std)


class ResBlock1(nn.Module):
    """Parameters of hifigan/models.py:18-43 (forward :45-52 runs inside es_hifigan_forward)."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.h = h
        self.convs1 = nn.ModuleList([
            weight_norm(Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=get_padding(kernel_size, d)))
            for d in dilation])
        self.convs1.apply(_init_weights)
        self.convs2 = nn.ModuleList([
            weight_norm(Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=get_padding(kernel_size, 1)))
            for _ in dilation])
        self.convs2.apply(_init_weights)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("ResBlock1 only holds parameters; call Generator, which runs the fused CUDA kernels")

    def remove_weight_norm(self):
        for l in self.convs1:
            remove_weight_norm(l)
        for l in self.convs2:
            remove_weight_norm(l)


def _effective_weight(m: nn.Module) -> torch.Tensor:
    """The conv weight the reference would use: g * v / ||v|| per slice of dim 0 while weight norm is attached
    (torch._weight_norm), the plain weight after remove_weight_norm()."""
    if hasattr(m, "weight_g"):
        v, g = m.weight_v.detach().float(), m.weight_g.detach().float()
        norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (v.dim() - 1))
        return v * (g / norm)
    return m.weight.detach().float()


class Generator(nn.Module):
    """HiFi-GAN generator -- drop-in for hifigan/models.py:84-134."""

    def __init__(self, h):
        super().__init__()
        if str(h.resblock) != "1":
            raise ValueError("only ResBlock1 generators (\"resblock\": \"1\") are implemented")
        self.h = h
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        self.conv_pre = weight_norm(Conv1d(80, h.upsample_initial_channel, 7, 1, padding=3))
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            self.ups.append(weight_norm(ConvTranspose1d(h.upsample_initial_channel // (2 ** i),
                                                        h.upsample_initial_channel // (2 ** (i + 1)), k, u, padding=(k - u) // 2)))
        self.resblocks = nn.ModuleList()
        ch = h.upsample_initial_channel
        for i in range(len(self.ups)):
            ch = h.upsample_initial_channel // (2 ** (i + 1))
            for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(ResBlock1(h, ch, k, d))
        self.conv_post = weight_norm(Conv1d(ch, 1, 7, 1, padding=3))
        self.ups.apply(_init_weights)
        self.conv_post.apply(_init_weights)
        self._key = None
        self._flat: Optional[torch.Tensor] = None
        self._handle = C.c_void_p(None)
        self._workspace: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if self._handle.value:
                _cabi.load().es_hifigan_destroy(self._handle)
        except Exception:
            pass

    def remove_weight_norm(self):
        for l in self.ups:
            remove_weight_norm(l)
        for l in self.resblocks:
            l.remove_weight_norm()
        remove_weight_norm(self.conv_pre)
        remove_weight_norm(self.conv_post)

    @property
    def total_upsampling(self) -> int:
        return int(np.prod(self.h.upsample_rates))

    def _convs(self) -> List[nn.Module]:
        out = [self.conv_pre] + list(self.ups)
        for rb in self.resblocks:
            out += list(rb.convs1) + list(rb.convs2)
        return out + [self.conv_post]

    def _ensure(self, device: torch.device) -> C.c_void_p:
        params = list(self.parameters())
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in params)
        if key == self._key:
            return self._handle
        lib = _cabi.load()
        h = self.h
        chunks, offs, off = [], {}, 0
        for idx, m in enumerate(self._convs()):
            for name, t in (("w", _effective_weight(m)), ("b", m.bias.detach().float())):
                a = t.reshape(-1).cpu()
                offs[(idx, name)] = off
                off += (a.numel() + 63) // 64 * 64                     # 256-byte aligned slots
                chunks.append((a, offs[(idx, name)]))
        flat = torch.zeros(off, dtype=torch.float32)
        for a, o in chunks:
            flat[o:o + a.numel()] = a
        self._flat = flat.to(device)
        base = self._flat.data_ptr()

        def cw(idx):
            return _cabi.es_hg_conv_w_t(base + 4 * offs[(idx, "w")], base + 4 * offs[(idx, "b")])

        W = _cabi.es_hifigan_weights_t()
        W.conv_pre = cw(0)
        idx = 1
        for i in range(self.num_upsamples):
            W.ups[i] = cw(idx)
            idx += 1
        for i in range(self.num_upsamples):
            for j in range(self.num_kernels):
                for d in range(3):
                    W.res[i][j].convs1[d] = cw(idx + d)
                    W.res[i][j].convs2[d] = cw(idx + 3 + d)
                idx += 6
        W.conv_post = cw(idx)
        cfg = _cabi.es_hifigan_config_t()
        cfg.n_mel, cfg.initial_channel = 80, int(h.upsample_initial_channel)
        cfg.n_up, cfg.n_res = self.num_upsamples, self.num_kernels
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            cfg.up_rate[i], cfg.up_kernel[i] = int(u), int(k)
        for j, (k, d) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
            if len(d) != 3:
                raise ValueError("ResBlock1 takes three dilations per kernel size")
            cfg.res_kernel[j] = int(k)
            for q in range(3):
                cfg.res_dilation[j][q] = int(d[q])
        if self._handle.value:
            lib.es_hifigan_destroy(self._handle)
            self._handle = C.c_void_p(None)
        hd = C.c_void_p(None)
        _cabi.check(lib.es_hifigan_create(C.byref(cfg), C.byref(W), C.byref(hd)))
        self._handle, self._key = hd, key
        return hd

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("efficientspeech_b200.hifigan: mel must be a CUDA tensor -- the vocoder runs only as "
                               "sm_100a kernels, there is no CPU fallback")
        if x.dim() != 3 or x.shape[1] != 80:
            raise RuntimeError(f"mel must be [B, 80, T], got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            x = x.float()
        B, _, T = x.shape
        dev = x.device
        lib = _cabi.load()
        with torch.cuda.device(dev):
            hd = self._ensure(dev)
            need = lib.es_hifigan_workspace_bytes(hd, B, T)
            if self._workspace is None or self._workspace.device != dev or self._workspace.numel() < need:
                self._workspace = torch.empty(int(need) + 1024, dtype=torch.uint8, device=dev)
            wav = torch.empty(B, 1, T * self.total_upsampling, dtype=torch.float32, device=dev)
            sb, sc, st = x.stride()
            _cabi.check(lib.es_hifigan_forward(hd, torch.cuda.current_stream(dev).cuda_stream, B, T, x.data_ptr(), sb, sc, st,
                                               wav.data_ptr(), self._workspace.data_ptr(), self._workspace.numel()))
        return wav
