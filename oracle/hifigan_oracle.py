"""CPU oracle for the HiFi-GAN generator (mel -> waveform), the step after the acoustic path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as es_oracle.py).  Plain-numpy restatement of
hifigan/models.py (jik876/hifi-gan as vendored by the reference), every function citing the lines it follows.

Parity pin: the reference's own ``hifigan.Generator`` with the checkpoint that ships in its tree
(hifigan/LJ_V2/generator_v2) runs in the build container; ``oracle/make_golden.py`` dumps its output for a seeded mel to
``tests/golden/hifigan_v2_t6.npz`` and ``tests/test_hifigan_cpu.py`` checks this restatement against the fixture
everywhere and against the live reference where it is staged (oracle/_ref).
"""
from __future__ import annotations

import json
import os

import numpy as np

LRELU_SLOPE = 0.1                                             # hifigan/models.py:16


def leaky_relu(x, slope):
    return np.where(x > 0, x, x * np.asarray(slope, x.dtype))


def conv1d(x, w, b, dilation=1, padding=0):
    """nn.Conv1d, stride 1: x [B,Cin,L], w [Cout,Cin,K] -> [B,Cout,L + 2 pad - dil (K-1)]."""
    B, Cin, L = x.shape
    Cout, _, K = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding)))
    Lo = L + 2 * padding - dilation * (K - 1)
    y = np.zeros((B, Cout, Lo), x.dtype)
    for j in range(K):
        y += np.einsum("oc,bcl->bol", w[:, :, j], xp[:, :, j * dilation:j * dilation + Lo])
    return y + b[None, :, None]


def conv_transpose1d(x, w, b, stride, padding):
    """nn.ConvTranspose1d: x [B,Cin,L], w [Cin,Cout,K] -> [B,Cout,(L-1) stride - 2 pad + K]."""
    B, Cin, L = x.shape
    _, Cout, K = w.shape
    full = np.zeros((B, Cout, (L - 1) * stride + K), x.dtype)
    for j in range(K):
        full[:, :, j:j + (L - 1) * stride + 1:stride] += np.einsum("co,bcl->bol", w[:, :, j], x)
    Lo = (L - 1) * stride - 2 * padding + K
    return full[:, :, padding:padding + Lo] + b[None, :, None]


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)       # hifigan/models.py:14-15


def effective_weight(state, prefix):
    """weight_norm(dim=0): w = g * v / ||v|| per slice of dim 0 (torch.nn.utils.weight_norm); plain weight after
    remove_weight_norm (hifigan/models.py:129-134, model.py:44)."""
    if prefix + ".weight" in state:
        return np.asarray(state[prefix + ".weight"], np.float32)
    v = np.asarray(state[prefix + ".weight_v"], np.float32)
    g = np.asarray(state[prefix + ".weight_g"], np.float32)
    norm = np.sqrt((v.reshape(v.shape[0], -1).astype(np.float64) ** 2).sum(1)).astype(np.float32)
    return v * (g.reshape(-1) / norm).reshape([-1] + [1] * (v.ndim - 1))


def resblock1(x, state, pre, kernel, dilations):
    """ResBlock1.forward, hifigan/models.py:45-52."""
    for d, dil in enumerate(dilations):
        xt = leaky_relu(x, LRELU_SLOPE)
        xt = conv1d(xt, effective_weight(state, f"{pre}.convs1.{d}"), state[f"{pre}.convs1.{d}.bias"], dil, get_padding(kernel, dil))
        xt = leaky_relu(xt, LRELU_SLOPE)
        xt = conv1d(xt, effective_weight(state, f"{pre}.convs2.{d}"), state[f"{pre}.convs2.{d}.bias"], 1, get_padding(kernel, 1))
        x = xt + x
    return x


def generator(mel, state, cfg):
    """Generator.forward, hifigan/models.py:111-127.  mel [B,80,T] float32 -> wav [B,1,T*prod(rates)]."""
    state = {k: np.asarray(v) for k, v in state.items()}
    x = conv1d(mel.astype(np.float32), effective_weight(state, "conv_pre"), state["conv_pre.bias"], 1, 3)        # :112
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = leaky_relu(x, LRELU_SLOPE)                                                                            # :114
        x = conv_transpose1d(x, effective_weight(state, f"ups.{i}"), state[f"ups.{i}.bias"], u, (k - u) // 2)     # :115
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            r = resblock1(x, state, f"resblocks.{i * nk + j}", rk, rd)                                            # :117-121
            xs = r if xs is None else xs + r
        x = xs / np.float32(nk)                                                                                   # :122
    x = leaky_relu(x, 0.01)                                                                                       # :123 (F.leaky_relu default)
    x = conv1d(x, effective_weight(state, "conv_post"), state["conv_post.bias"], 1, 3)                           # :124
    return np.tanh(x)                                                                                             # :125


def load_reference_checkpoint(ref_dir):
    """(config dict, state dict of numpy arrays) of <ref_dir>/hifigan/LJ_V2 (torch needed only to unpickle)."""
    import torch
    with open(os.path.join(ref_dir, "hifigan", "LJ_V2", "config.json")) as f:
        cfg = json.load(f)
    ck = torch.load(os.path.join(ref_dir, "hifigan", "LJ_V2", "generator_v2"), map_location="cpu")
    return cfg, {k: v.numpy() for k, v in ck["generator"].items()}
