"""CPU validation of the weight folding (packing.py) against the oracle, via tests/kernel_model.py."""
import os

import numpy as np
import pytest

from efficientspeech_b200 import packing
from helpers import golden_files, load_golden
from kernel_model import encoder_model, length_regulator_model, predictor_model
from oracle import es_oracle


@pytest.mark.parametrize("path", golden_files()[::3] + golden_files()[2::3], ids=lambda p: os.path.basename(p)[:-4])
def test_folded_encoder_matches_oracle(path):
    vname, cfg, sd, batch, g = load_golden(path)
    F = packing.fold_encoder(sd, cfg)
    o = es_oracle.phoneme2mel(batch, sd, train=True, dtype=np.float64)
    mask = batch["phoneme_mask"] if batch["phoneme"].shape[0] > 1 else None
    feats, fused = encoder_model(F, cfg, batch["phoneme"], mask)
    for a, b in zip(feats, o["_enc_features"]):
        assert np.abs(a - b).max() < 1e-9
    assert np.abs(fused - o["_fused"]).max() < 1e-9
    for which in ("pitch", "energy", "duration"):
        pred, feat = predictor_model(F, which, fused)
        assert np.abs(pred - o[which][..., 0]).max() < 1e-9
    d = cfg.dim
    assert np.abs(feat * (1 if mask is None else ~mask[..., None]) - o["_fused4"][..., 3 * d:]).max() < 1e-9


def test_length_regulator_model_matches_oracle():
    rng = np.random.default_rng(5)
    for _ in range(30):
        B, N = int(rng.integers(1, 5)), int(rng.integers(1, 50))
        dur = rng.integers(0, 7, size=(B, N)).astype(np.int32)
        feats = np.zeros((B, N, 1), np.float32)
        _, _, ml, src = es_oracle.feature_upsampler(feats, np.zeros((B, N, 1), bool), dur)
        T = int(ml.max())
        got = length_regulator_model(np.cumsum(dur, axis=1), T)
        assert np.array_equal(got, src)


def test_pack_layout_and_split_fp16():
    vname, cfg, sd, batch, g = load_golden(golden_files()[0])
    F = dict(packing.fold_encoder(sd, cfg))
    F.update(packing.fold_decoder(sd, cfg))
    flat, off = packing.pack(F)
    assert flat.dtype == np.float32
    for k, v in F.items():
        assert off[k] % packing.ALIGN == 0
        assert np.array_equal(flat[off[k]:off[k] + v.size], np.asarray(v, np.float32).reshape(-1))
    assert F["mel_w"].shape == (1, cfg.dx2, 96) and F["mel_b"].shape == (96,)
    w = F["dec0.pw_w"][0].T.astype(np.float32)                    # [N][K]
    n, k = w.shape
    img = packing.canon_split_fp16(w).view(np.float16).reshape(2, k // 8, n, 8)
    # canonical K-major no-swizzle order: element (row n, col kk) lives at [kk // 8][n][kk % 8]
    back = img.transpose(0, 2, 1, 3).reshape(2, n, k).astype(np.float64)
    assert np.abs(back[0] + back[1] - w).max() < 2e-7
    assert np.array_equal(back[0].astype(np.float16), w.astype(np.float16))


def test_chunked_split_image_is_per_chunk_canonical():
    """es_umma_dec256.cu streams the weight in 32-column K chunks; every chunk must by itself be
    the canonical K-major image of that column slice (so chunk c == canon_split_fp16(w[:, c*32:(c+1)*32]))."""
    rng = np.random.default_rng(5)
    for n, k in ((256, 256), (80, 256), (256, 512)):
        w = rng.standard_normal((n, k)).astype(np.float32)
        img = packing.canon_split_chunks(w).view(np.float16).reshape(k // 32, -1)
        for c in range(k // 32):
            ref = packing.canon_split_fp16(w[:, c * 32:(c + 1) * 32]).view(np.float16)
            assert np.array_equal(img[c], ref)


def test_streamed_units_image_order():
    """es_umma_wide.cu consumes [n/nt][K/32][taps] units, each the canonical image of one tap's
    [nt][32] slice."""
    rng = np.random.default_rng(6)
    taps, k, n, npad, nt = 3, 128, 256, 256, 128
    w = rng.standard_normal((taps, k, npad)).astype(np.float32)
    img = packing.canon_split_units(w, n, nt).view(np.float16).reshape(n // nt, k // 32, taps, -1)
    for y in range(n // nt):
        for c in range(k // 32):
            for t in range(taps):
                ref = packing.canon_split_fp16(np.ascontiguousarray(w[t, c * 32:(c + 1) * 32, y * nt:(y + 1) * nt].T))
                assert np.array_equal(img[y, c, t], ref.view(np.float16))


def test_pair_stride2_taps_is_the_same_convolution():
    """The paired-row form of a 3-tap stride-2 conv (base, encoder block 1 merge): a stride-1 'same' conv over
    X2[t] = [x[2t] | x[2t+1]] with K doubled must give the strided conv's output."""
    import numpy as np
    from efficientspeech_b200 import packing
    rng = np.random.default_rng(5)
    n, K, N = 12, 6, 4
    x = rng.standard_normal((n, K))
    w = rng.standard_normal((3, K, N))
    xp = np.concatenate([np.zeros((1, K)), x, np.zeros((1, K))])
    want = np.stack([sum(xp[2 * t + tau] @ w[tau] for tau in range(3)) for t in range(n // 2)])      # pad 1, stride 2
    w2 = packing.pair_stride2_taps(w)
    x2 = x.reshape(n // 2, 2 * K)
    x2p = np.concatenate([np.zeros((1, 2 * K)), x2, np.zeros((1, 2 * K))])
    got = np.stack([sum(x2p[t + tau] @ w2[tau] for tau in range(3)) for t in range(n // 2)])           # pad 1, stride 1
    assert np.abs(got - want).max() < 1e-12
