"""Parameter layout of the acoustic model = the reference's state-dict layout.

``param_shapes(cfg)`` reproduces, name for name and shape for shape, what
``layers.Phoneme2Mel(PhonemeEncoder(...), MelDecoder(...)).state_dict()`` holds
(layers/networks.py:32-47, :98-122, :176-187, :272-288) so that
checkpoints move both ways with ``load_state_dict(strict=True)``.

``init_state_dict(cfg, seed)`` is a deterministic, platform-independent (numpy PCG64)
initialisation used for synthetic-weight benchmarks and parity fixtures; real use loads a
reference checkpoint instead.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

from .config import ESConfig

# kind tags: "w" weight (fan_in given), "b" bias of the preceding weight, "ln_w"/"ln_b"
# LayerNorm affine, "embed" phoneme table, "table" variance embedding, "bins" bucket edges.
_Spec = "OrderedDict[str, Tuple[Tuple[int, ...], str, int]]"


def _spec(cfg: ESConfig):
    cfg.validate()
    s = OrderedDict()

    def w(name, shape, fan_in, bias=True):
        s[name + ".weight"] = (tuple(shape), "w", fan_in)
        if bias:
            # (the one ConvTranspose1d, weight (in, out, k), has in == out, so shape[0] holds)
            s[name + ".bias"] = ((shape[0],), "b", fan_in)

    def ln(name, c):
        s[name + ".weight"] = ((c,), "ln_w", 0)
        s[name + ".bias"] = ((c,), "ln_b", 0)

    e = "encoder.encoder."
    s[e + "embed.weight"] = ((cfg.n_symbols, cfg.embed_dim), "embed", 0)
    for i, (cin, c, h, k) in enumerate(zip(cfg.enc_dims_in, cfg.enc_dims, cfg.enc_heads, cfg.enc_kernels)):
        p = e + f"attn_blocks.{i}."
        hc = c * cfg.expansion
        w(p + "0", (cin, cin, k), cin * k, bias=False)      # dense merge conv, no bias
        w(p + "1", (c, cin, 1), cin, bias=False)            # 1x1 projection, no bias
        w(p + "2.qkv", (3 * h * c, c), c, bias=False)       # row order [q|k|v][head][c]
        w(p + "2.proj", (c, h * c), h * c)
        w(p + "3.mlp1", (hc, c), c)
        w(p + "3.conv", (hc, hc, 3), hc * 3)
        w(p + "3.mlp2", (c, hc), hc)
        ln(p + "4", c)
        ln(p + "5", c)
    d = cfg.dim
    f = "encoder.fuse."
    for i, c in enumerate(cfg.enc_dims):
        w(f + f"mlps.{i}.0", (d, c), c)
        if i > 0:                                           # ConvTranspose1d weight is (in, out, k)
            w(f + f"mlps.{i}.1", (d, d, cfg.kernel_size), d * cfg.kernel_size)
    w(f + "fuse", (d, d * cfg.depth), d * cfg.depth)
    for which in ("pitch", "energy", "duration"):
        p = f"encoder.{which}_decoder."
        if which != "duration":
            s[p + f"{which}_bins"] = ((d - 1,), "bins", 0)
        w(p + "conv1.0", (d, d, 3), d * 3)
        ln(p + "norm1", d)
        w(p + "conv2.0", (d, d, 3), d * 3)
        ln(p + "norm2", d)
        w(p + "linear", (1, d), d)
        if which != "duration":
            s[p + f"{which}_embedding.weight"] = ((d, d), "table", 0)
    q = "decoder."
    w(q + "proj.0", (cfg.dx2, cfg.dx4), cfg.dx4)
    ln(q + "proj.2", cfg.dx2)
    for b in range(cfg.n_blocks):
        for l in range(cfg.block_depth):
            p = q + f"blocks.{b}.0.{l}."
            w(p + "0.0", (cfg.dx2, 1, cfg.decoder_kernel_size), cfg.decoder_kernel_size)  # depthwise
            w(p + "0.1", (cfg.dx2, cfg.dx2, 1), cfg.dx2)                                  # pointwise
            ln(p + "1", cfg.dx2)
        ln(q + f"blocks.{b}.1", cfg.dx2)
    w(q + "mel_linear", (cfg.n_mel, cfg.dx2), cfg.dx2)
    return s


def param_shapes(cfg: ESConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    return OrderedDict((k, v[0]) for k, v in _spec(cfg).items())


def variance_bins(stats: Tuple[float, float], d: int) -> np.ndarray:
    """torch.linspace(min, max, d-1) in fp32, exactly as layers/networks.py:111,120."""
    import torch
    return torch.linspace(float(stats[0]), float(stats[1]), d - 1, dtype=torch.float32).numpy().copy()


def init_state_dict(cfg: ESConfig, seed: int = 0, duration_bias: float = 2.5) -> Dict[str, np.ndarray]:
    """Deterministic synthetic weights (fp32 numpy).

    Linear/conv weights and biases ~ U(+-1/sqrt(fan_in)) (torch's default bound), LayerNorm
    affine parameters are perturbed away from (1, 0) so that they are actually exercised,
    the embedding padding row is zero (padding_idx=0), and the duration head bias is raised
    so that free-running inference produces non-empty utterances (with torch's default
    init every duration rounds to 0).
    """
    rng = np.random.default_rng(np.random.SeedSequence([0xE5B200, seed]))
    sd: Dict[str, np.ndarray] = {}
    for name, (shape, kind, fan_in) in _spec(cfg).items():
        if kind == "bins":
            v = variance_bins(cfg.pitch_stats if "pitch" in name else cfg.energy_stats, cfg.dim)
        elif kind == "embed":
            v = rng.standard_normal(shape)
            v[0] = 0.0
        elif kind == "table":
            v = rng.standard_normal(shape)
        elif kind == "ln_w":
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif kind == "ln_b":
            v = 0.1 * rng.standard_normal(shape)
        else:
            bound = 1.0 / np.sqrt(fan_in)
            v = rng.uniform(-bound, bound, size=shape)
        sd[name] = np.ascontiguousarray(v, dtype=np.float32)
    sd["encoder.duration_decoder.linear.bias"] += np.float32(duration_bias)
    return sd


def state_checksum(sd: Dict[str, np.ndarray]) -> str:
    """Order-independent content hash (pins fixtures to the weights they were made with)."""
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k]).tobytes())
    return h.hexdigest()[:16]
