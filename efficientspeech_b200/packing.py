"""Fold the reference state dict into the kernel-ready weight image (include/es_b200.h).

Everything here is host-side algebra done once per weight update, in float64, then rounded
to fp32:

* encoder block 0: embedding (networks.py:54) + dense merge conv (:65) + 1x1 (:66) collapse
  into ``k`` gather tables ``Tab[tau] = E (W1x1 Wk[:,:,tau])^T`` of shape [153, d];
* encoder block 1: merge conv (k-2, stride 2) and 1x1 fold into one strided conv;
* MixFFN: ``mlp1`` folds into the dense k=3 conv (blocks.py:23-25).  The conv zero-pads
  ``mlp1``'s OUTPUT, so the folded bias is split per tap (``ffn1_tapb``) and added only for
  taps that read inside the sequence;
* Fuse (networks.py:189-219) is linear end to end: Linear -> ConvTranspose1d -> concat ->
  Linear becomes ``A0`` (level 0), per-tap ``G_tau``/``g_tau`` (level 1) and a constant;
* all dense weights are stored K-major ``[taps][K][Nout_padded_to_32]`` so a warp reads
  consecutive output channels.

The result is one flat fp32 buffer plus an ``offsets`` table (element offsets); modules.py
uploads the buffer and points the ``es_weights_t`` fields into it.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .config import ESConfig

ALIGN = 64          # floats (256 B)


def _pad_cols(a: np.ndarray) -> np.ndarray:
    """Pad the last dim to a multiple of 32 with zeros."""
    n = a.shape[-1]
    npad = (n + 31) // 32 * 32
    if npad == n:
        return a
    out = np.zeros(a.shape[:-1] + (npad,), dtype=a.dtype)
    out[..., :n] = a
    return out


def fold_encoder(sd: Dict[str, np.ndarray], cfg: ESConfig) -> Dict[str, np.ndarray]:
    f64 = lambda k: np.asarray(sd[k], dtype=np.float64)  # noqa: E731
    out: Dict[str, np.ndarray] = {}
    e = "encoder.encoder."
    E = f64(e + "embed.weight")
    for i in range(2):
        p = e + f"attn_blocks.{i}."
        Wk = f64(p + "0.weight")                   # [Cin, Cin, k]
        W1 = f64(p + "1.weight")[:, :, 0]          # [C, Cin]
        k = Wk.shape[2]
        weff = np.stack([W1 @ Wk[:, :, t] for t in range(k)])        # [k][C][Cin]
        if i == 0:
            out["enc0.merge_w"] = np.stack([E @ weff[t].T for t in range(k)])        # [k][153][C]
        else:
            out["enc1.merge_w"] = _pad_cols(np.stack([weff[t].T for t in range(k)]))  # [k][Cin][C]
        out[f"enc{i}.qkv_w"] = _pad_cols(f64(p + "2.qkv.weight").T[None])
        out[f"enc{i}.proj_w"] = _pad_cols(f64(p + "2.proj.weight").T[None])
        out[f"enc{i}.proj_b"] = f64(p + "2.proj.bias")
        out[f"enc{i}.ln1_g"] = f64(p + "4.weight")
        out[f"enc{i}.ln1_b"] = f64(p + "4.bias")
        Wm1, bm1 = f64(p + "3.mlp1.weight"), f64(p + "3.mlp1.bias")  # [hC, C]
        Wc, bc = f64(p + "3.conv.weight"), f64(p + "3.conv.bias")    # [hC, hC, 3]
        out[f"enc{i}.ffn1_w"] = _pad_cols(np.stack([(Wc[:, :, t] @ Wm1).T for t in range(3)]))   # [3][C][hC]
        out[f"enc{i}.ffn1_tapb"] = _pad_cols(np.stack([Wc[:, :, t] @ bm1 for t in range(3)]))    # [3][hC]
        out[f"enc{i}.ffn1_b"] = bc
        out[f"enc{i}.ffn2_w"] = _pad_cols(f64(p + "3.mlp2.weight").T[None])
        out[f"enc{i}.ffn2_b"] = f64(p + "3.mlp2.bias")
        out[f"enc{i}.ln2_g"] = f64(p + "5.weight")
        out[f"enc{i}.ln2_b"] = f64(p + "5.bias")
    # ---- Fuse
    f = "encoder.fuse."
    d = cfg.dim
    Wf, bf = f64(f + "fuse.weight"), f64(f + "fuse.bias")
    Wf0, Wf1 = Wf[:, :d], Wf[:, d:]
    Wm0, bm0 = f64(f + "mlps.0.0.weight"), f64(f + "mlps.0.0.bias")
    Wm1, bm1 = f64(f + "mlps.1.0.weight"), f64(f + "mlps.1.0.bias")
    Wct, bct = f64(f + "mlps.1.1.weight"), f64(f + "mlps.1.1.bias")  # [in, out, k]
    k = Wct.shape[2]
    out["fuse_a0"] = (Wf0 @ Wm0).T                                    # [d(K)][d(out)]
    out["fuse_g"] = np.stack([(Wf1 @ Wct[:, :, t].T @ Wm1).T for t in range(k)])   # [k][2d][d]
    out["fuse_gb"] = np.stack([Wf1 @ Wct[:, :, t].T @ bm1 for t in range(k)])      # [k][d]
    out["fuse_c"] = Wf0 @ bm0 + Wf1 @ bct + bf
    # ---- predictors
    for which in ("pitch", "energy", "duration"):
        p = f"encoder.{which}_decoder."
        for c in ("conv1", "conv2"):
            W = f64(p + c + ".0.weight")                              # [d, d, 3]
            out[f"{which}.{c}_w"] = np.stack([W[:, :, t].T for t in range(3)])
            out[f"{which}.{c}_b"] = f64(p + c + ".0.bias")
        for n in ("1", "2"):
            out[f"{which}.ln{n}_g"] = f64(p + f"norm{n}.weight")
            out[f"{which}.ln{n}_b"] = f64(p + f"norm{n}.bias")
        out[f"{which}.lin_w"] = f64(p + "linear.weight")[0]
        out[f"{which}.lin_b"] = f64(p + "linear.bias")
        if which != "duration":
            out[f"{which}.bins"] = np.asarray(sd[p + f"{which}_bins"], dtype=np.float32)   # exact copy
            out[f"{which}.table"] = f64(p + f"{which}_embedding.weight")
    return out


def fold_decoder(sd: Dict[str, np.ndarray], cfg: ESConfig) -> Dict[str, np.ndarray]:
    f64 = lambda k: np.asarray(sd[k], dtype=np.float64)  # noqa: E731
    out: Dict[str, np.ndarray] = {}
    q = "decoder."
    out["dproj_w"] = f64(q + "proj.0.weight").T[None]                 # [1][dx4][dx2]
    out["dproj_b"] = f64(q + "proj.0.bias")
    out["dproj_ln_g"] = f64(q + "proj.2.weight")
    out["dproj_ln_b"] = f64(q + "proj.2.bias")
    layer = 0
    for b in range(cfg.n_blocks):
        for l in range(cfg.block_depth):
            p = q + f"blocks.{b}.0.{l}."
            out[f"dec{layer}.dw_w"] = f64(p + "0.0.weight")[:, 0, :].T.copy()      # [k][dx2]
            out[f"dec{layer}.dw_b"] = f64(p + "0.0.bias")
            out[f"dec{layer}.pw_w"] = f64(p + "0.1.weight")[:, :, 0].T[None].copy()  # [1][K][N]
            out[f"dec{layer}.pw_b"] = f64(p + "0.1.bias")
            out[f"dec{layer}.ln_g"] = f64(p + "1.weight")
            out[f"dec{layer}.ln_b"] = f64(p + "1.bias")
            layer += 1
        out[f"blk{b}.ln_g"] = f64(q + f"blocks.{b}.1.weight")
        out[f"blk{b}.ln_b"] = f64(q + f"blocks.{b}.1.bias")
    out["mel_w"] = _pad_cols(f64(q + "mel_linear.weight").T[None])   # [1][dx2][96]
    out["mel_b"] = _pad_cols(f64(q + "mel_linear.bias"))
    return out


def split_fp16(w: np.ndarray) -> np.ndarray:
    """hi = fp16(w), lo = fp16(w - hi): the two-term fp16 image used by the tcgen05 kernels.

    Returned as a float32-typed container of the raw halves ([2, ...] fp16 viewed as fp32 words)
    so that it can live in the same flat buffer.  hi + lo reproduces w to ~2^-22 relative.
    """
    w32 = np.asarray(w, dtype=np.float32)
    hi = w32.astype(np.float16)
    lo = (w32 - hi.astype(np.float32)).astype(np.float16)
    both = np.stack([hi, lo]).reshape(-1)
    assert both.size % 2 == 0
    return both.view(np.float32)


def canon_split_fp16(w_nk: np.ndarray) -> np.ndarray:
    """[N][K] fp32 weight -> split-fp16 image in the UMMA canonical K-major no-swizzle order
    ``[2 (hi, lo)][K/8][N][8]`` (8x8 core matrices of 128 contiguous bytes; es_umma.cuh), so the
    kernel's weight load is a straight bulk copy.  Returned as raw fp32 words for the flat buffer."""
    w32 = np.ascontiguousarray(w_nk, dtype=np.float32)
    n, k = w32.shape
    assert k % 8 == 0
    hi = w32.astype(np.float16)
    lo = (w32 - hi.astype(np.float32)).astype(np.float16)
    planes = np.stack([hi, lo]).reshape(2, n, k // 8, 8).transpose(0, 2, 1, 3)
    return np.ascontiguousarray(planes).reshape(-1).view(np.float32)


def canon_split_chunks(w_nk: np.ndarray, kc: int = 32) -> np.ndarray:
    """[N][K] fp32 weight -> K-chunked split-fp16 image ``[K/kc][2 (hi, lo)][kc/8][N][8]`` for the
    wide-decoder kernel (es_umma_dec256.cu): every chunk is one contiguous bulk copy and is, by
    itself, a UMMA canonical K-major no-swizzle B operand of kc columns."""
    w32 = np.ascontiguousarray(w_nk, dtype=np.float32)
    n, k = w32.shape
    assert k % kc == 0 and kc % 8 == 0
    hi = w32.astype(np.float16)
    lo = (w32 - hi.astype(np.float32)).astype(np.float16)
    planes = np.stack([hi, lo]).reshape(2, n, k // kc, kc // 8, 8)          # [pl][n][chunk][panel][8]
    img = planes.transpose(2, 0, 3, 1, 4)                                   # [chunk][pl][panel][n][8]
    return np.ascontiguousarray(img).reshape(-1).view(np.float32)


def canon_split_taps(w_tkn: np.ndarray, n: int) -> np.ndarray:
    """Folded conv weight [taps][K][N_padded] -> per-tap canonical split-fp16 images
    ``[taps][2 (hi, lo)][K/8][N][8]`` for the tcgen05 row GEMM (es_umma_enc.cu)."""
    taps = w_tkn.shape[0]
    return np.concatenate([canon_split_fp16(np.ascontiguousarray(w_tkn[t].T[:n])) for t in range(taps)])


def canon_split_units(w_tkn: np.ndarray, n: int, nt: int, kc: int = 32) -> np.ndarray:
    """Folded conv weight [taps][K][N_padded] -> streamed units for es_umma_wide.cu, in the order the
    kernel consumes them: ``[n/nt][K/kc][taps][2 (hi, lo)][kc/8][nt][8]`` halves -- each unit is the
    canonical K-major image of one tap's [nt][kc] slice (== canon_split_fp16 of that slice)."""
    taps, k, _ = w_tkn.shape
    assert n % nt == 0 and k % kc == 0
    out = []
    for y in range(n // nt):
        for c in range(k // kc):
            for t in range(taps):
                sl = np.ascontiguousarray(w_tkn[t, c * kc:(c + 1) * kc, y * nt:(y + 1) * nt].T)   # [nt][kc]
                out.append(canon_split_fp16(sl))
    return np.concatenate(out)


def pair_stride2_taps(w_tkn: np.ndarray) -> np.ndarray:
    """A 3-tap stride-2 conv [3][K][N] as a 3-tap stride-1 "same" conv over paired rows X2[t] = [x[2t] | x[2t+1]]:
    y[t] = W0 x[2t-1] + W1 x[2t] + W2 x[2t+1] = [0 | W0] X2[t-1] + [W1 | W2] X2[t]  ->  [3][2K][N], third tap zero."""
    assert w_tkn.shape[0] == 3
    _, k, n = w_tkn.shape
    out = np.zeros((3, 2 * k, n), dtype=w_tkn.dtype)
    out[0, k:] = w_tkn[0]
    out[1, :k] = w_tkn[1]
    out[1, k:] = w_tkn[2]
    return out


def pack(folded: Dict[str, np.ndarray]) -> Tuple[np.ndarray, Dict[str, int]]:
    """Concatenate the folded arrays (fp32, 256-byte aligned) -> (flat buffer, element offsets)."""
    offsets: Dict[str, int] = {}
    total = 0
    for k, v in folded.items():
        offsets[k] = total
        total += (int(v.size) + ALIGN - 1) // ALIGN * ALIGN
    flat = np.zeros(max(total, ALIGN), dtype=np.float32)
    for k, v in folded.items():
        flat[offsets[k]:offsets[k] + v.size] = np.asarray(v).astype(np.float32, copy=False).reshape(-1)
    return flat, offsets
