"""Numpy model of what the CUDA kernels compute FROM THE FOLDED WEIGHT IMAGE (test-only).

It follows the kernels' formulation (gather tables, boundary-aware tap bias, parity taps of
the transposed conv, upper_bound length regulator) rather than the reference's op sequence,
so comparing it with the oracle on CPU validates efficientspeech_b200/packing.py and the
index arithmetic of the kernels without a GPU.
"""
import numpy as np

from oracle.es_oracle import gelu_erf, layer_norm, softmax_last


def conv_taps(x, w, taps, stride, pad, n_out, tap_bias=None):
    """x [B,n,K]; w [taps][K][Np] -> [B,n_out,Np] (+ per-tap bias where the tap is in range)."""
    B, n, K = x.shape
    y = np.zeros((B, n_out, w.shape[2]), dtype=np.float64)
    for t in range(n_out):
        for tau in range(taps):
            ti = t * stride + tau - pad
            if 0 <= ti < n:
                y[:, t] += x[:, ti] @ w[tau]
                if tap_bias is not None:
                    y[:, t] += tap_bias[tau]
    return y


def encoder_model(F, cfg, ids, mask):
    B, N = ids.shape
    d = cfg.dim
    tab = F["enc0.merge_w"]
    k0 = tab.shape[0]
    x = np.zeros((B, N, d))
    for t in range(N):
        for tau in range(k0):
            ti = t + tau - k0 // 2
            if 0 <= ti < N:
                x[:, t] += tab[tau][ids[:, ti]]
    feats = []
    n = N
    m = mask
    for i in range(2):
        C, H, hC = cfg.enc_dims[i], cfg.enc_heads[i], cfg.enc_dims[i] * cfg.expansion
        if i == 1:
            k1 = cfg.enc_kernels[1]
            n1 = (N + 2 * (k1 // 2) - k1) // 2 + 1
            x = conv_taps(x, F["enc1.merge_w"], k1, 2, k1 // 2, n1)[..., :C]
            n = n1
            if mask is not None:
                pool = int(np.rint(np.float32(N / n1)))
                mp = np.ones((B, n1 * pool), dtype=bool)
                mp[:, :N] = mask
                m = mp.reshape(B, n1, pool).max(axis=2)
        qkv = x @ F[f"enc{i}.qkv_w"][0][:, :3 * H * C]
        q = qkv[..., 0:H * C].reshape(B, n, H, C)
        kk = qkv[..., H * C:2 * H * C].reshape(B, n, H, C)
        v = qkv[..., 2 * H * C:].reshape(B, n, H, C)
        s = np.einsum("bqhc,bkhc->bhqk", q, kk) / np.sqrt(C // H)
        o = np.einsum("bhqk,bkhc->bqhc", softmax_last(s), v).reshape(B, n, H * C)
        y = o @ F[f"enc{i}.proj_w"][0][:, :C] + F[f"enc{i}.proj_b"]
        x1 = layer_norm(y + x, F[f"enc{i}.ln1_g"], F[f"enc{i}.ln1_b"])
        if m is not None:
            x1 = np.where(m[..., None], 0.0, x1)
        h = conv_taps(x1, F[f"enc{i}.ffn1_w"], 3, 1, 1, n, tap_bias=F[f"enc{i}.ffn1_tapb"])[..., :hC]
        h = gelu_erf(h + F[f"enc{i}.ffn1_b"])
        y = h @ F[f"enc{i}.ffn2_w"][0][:, :C] + F[f"enc{i}.ffn2_b"]
        x = layer_norm(y + x1, F[f"enc{i}.ln2_g"], F[f"enc{i}.ln2_b"])
        if m is not None:
            x = np.where(m[..., None], 0.0, x)
        feats.append(x)
    # fuse
    f0, f1 = feats
    n1 = f1.shape[1]
    k = F["fuse_g"].shape[0]
    fused = np.zeros((B, N, d))
    for t in range(N):
        acc = F["fuse_c"] + f0[:, t] @ F["fuse_a0"]
        for tau in range(t & 1, k, 2):
            j = (t - tau) >> 1
            if t - tau < 0 or j >= n1:
                continue
            acc = acc + f1[:, j] @ F["fuse_g"][tau] + F["fuse_gb"][tau]
        fused[:, t] = acc
    if mask is not None:
        fused = np.where(mask[..., None], 0.0, fused)
    return feats, fused


def predictor_model(F, which, fused):
    B, N, d = fused.shape
    y = conv_taps(fused, F[f"{which}.conv1_w"], 3, 1, 1, N) + F[f"{which}.conv1_b"]
    y = np.maximum(layer_norm(np.maximum(y, 0), F[f"{which}.ln1_g"], F[f"{which}.ln1_b"]), 0)
    y = np.maximum(conv_taps(y, F[f"{which}.conv2_w"], 3, 1, 1, N) + F[f"{which}.conv2_b"], 0)
    pred = y @ F[f"{which}.lin_w"] + F[f"{which}.lin_b"][0]
    feat = layer_norm(y, F[f"{which}.ln2_g"], F[f"{which}.ln2_b"])
    if which == "duration":
        pred = np.maximum(pred, 0)
    return pred, feat


def length_regulator_model(cum, T):
    """src[b,t] = upper_bound(cum[b], t) for t < cum[b,-1], else -1."""
    B, N = cum.shape
    src = np.full((B, T), -1, dtype=np.int32)
    for b in range(B):
        L = min(int(cum[b, -1]), T)
        src[b, :L] = np.searchsorted(cum[b], np.arange(L), side="right")
    return src
