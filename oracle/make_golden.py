"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py

Each fixture holds: variant name, weight seed + checksum (weights are regenerated from the
seed by ``efficientspeech_b200.params.init_state_dict``, not stored), the input batch, and
the reference's outputs for the teacher-forced (``train=True``) and free-running
(``train=False``) calls of ``layers.Phoneme2Mel.forward`` (layers/networks.py:415-434),
plus the length regulator's integer source map obtained by pushing index-coded features
through the reference ``FeatureUpsampler`` (layers/networks.py:228-258).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from efficientspeech_b200.config import VARIANTS  # noqa: E402
from efficientspeech_b200.params import init_state_dict, state_checksum  # noqa: E402
from efficientspeech_b200.synthetic import make_batch  # noqa: E402
from oracle.ref_shim import build_reference_model, import_reference_layers, run_reference  # noqa: E402

CASES = [  # (tag, B, N, ragged, weight seed)
    ("b3n24", 3, 24, True, 11),
    ("b1n33", 1, 33, False, 12),
    ("b2n17", 2, 17, True, 13),
]


def reference_src_map(duration, max_mel_len):
    """Reference FeatureUpsampler on features whose value is the phoneme index."""
    import torch
    layers = import_reference_layers()
    up = layers.networks.FeatureUpsampler()
    B, N = duration.shape
    feats = torch.arange(N, dtype=torch.float32).view(1, N, 1).expand(B, N, 1).contiguous()
    masks = torch.zeros(B, N, 1, dtype=torch.bool)
    f, m, mel_len = up(feats, masks, torch.from_numpy(duration.copy()), max_mel_len=max_mel_len)
    src = f[..., 0].numpy().astype(np.int32)
    src[m[..., 0].numpy()] = -1
    return src, mel_len.numpy()


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for vname, cfg in VARIANTS.items():
        for tag, B, N, ragged, wseed in CASES:
            sd = init_state_dict(cfg, seed=wseed)
            ref = build_reference_model(cfg, sd)
            batch = make_batch(cfg, B, N, seed=wseed, ragged=ragged, fixed_duration=None, max_dur=7)
            tf = run_reference(ref, batch, train=True)
            fr = run_reference(ref, batch, train=False)
            src, ml = reference_src_map(batch["duration"], int(batch["mel_len"].max()))
            assert np.array_equal(ml, tf["mel_len"])
            rec = {"variant": vname, "weight_seed": wseed, "weight_checksum": state_checksum(sd)}
            for k, v in batch.items():
                rec["in_" + k] = v
            for k in ("mel", "mel_len", "pitch", "energy", "duration"):
                rec["tf_" + k] = tf[k]
            rec["tf_src"] = src
            for k in ("mel", "mel_len", "duration"):
                rec["fr_" + k] = fr[k]
            path = os.path.join(out_dir, f"{vname}_{tag}.npz")
            np.savez_compressed(path, **rec)
            print(path, os.path.getsize(path) // 1024, "KB  T_tf", tf["mel"].shape[1], "T_fr", fr["mel"].shape[1])


def make_collate_fixture():
    """tests/golden/collate_b7.npz: outputs of the reference's own collate_fn source (datamodule.py:29-76, extracted with
    oracle.ref_shim.reference_functions -- the module itself needs lightning) on seeded ragged items of distinct lengths."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_collate_cpu as t
    from oracle import es_oracle
    items = t.ragged_items(123, 7)
    for k, it in enumerate(items):
        for key in it:
            it[key] = np.resize(it[key], 4 + 5 * k)
    ref = t.reference_collate(items)
    o = es_oracle.collate(items)
    out = {"B": 7}
    for i, it in enumerate(items):
        for k, v in it.items():
            out[f"{k}_{i}"] = v
    for k in ("phoneme", "phoneme_len", "phoneme_mask", "pitch", "energy", "duration", "mel_len"):
        assert np.array_equal(np.asarray(ref[k]), o[k]), k
        out["out_" + k] = np.asarray(ref[k])
    out["out_perm"] = o["perm"]
    path = os.path.join(ROOT, "tests", "golden", "collate_b7.npz")
    np.savez_compressed(path, **out)
    print(path)


HIFIGAN_V2 = {"resblock": "1", "upsample_rates": [8, 8, 2, 2], "upsample_kernel_sizes": [16, 16, 4, 4],
              "upsample_initial_channel": 128, "resblock_kernel_sizes": [3, 7, 11],
              "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]]}


def hifigan_seeded_state(cfg, seed):
    """Plain (weight-norm removed) generator weights from a numpy seed: N(0, 1/sqrt(fan_in)) weights, small biases."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(name, shape, fan_in):
        sd[name + ".weight"] = (rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)
        sd[name + ".bias"] = (0.1 * rng.standard_normal(shape[1] if name.startswith("ups") else shape[0])).astype(np.float32)

    c0 = cfg["upsample_initial_channel"]
    conv("conv_pre", (c0, 80, 7), 80 * 7)
    ch = c0
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        conv(f"ups.{i}", (ch, ch // 2, k), ch * k // u)
        ch //= 2
        for j, rk in enumerate(cfg["resblock_kernel_sizes"]):
            for d in range(3):
                conv(f"resblocks.{i * len(cfg['resblock_kernel_sizes']) + j}.convs1.{d}", (ch, ch, rk), ch * rk)
                conv(f"resblocks.{i * len(cfg['resblock_kernel_sizes']) + j}.convs2.{d}", (ch, ch, rk), ch * rk)
    conv("conv_post", (1, ch, 7), ch * 7)
    return sd


def make_hifigan_fixture():
    """tests/golden/hifigan_v2_t6.npz: output of the reference's hifigan.Generator (hifigan/models.py:84-127, V2 geometry)
    for a seeded mel, once with seeded plain weights (self-contained fixture) and once with the checkpoint of the tree
    (hifigan/LJ_V2/generator_v2; the test needs oracle/_ref for that half)."""
    import warnings
    import torch
    from oracle import hifigan_oracle as ho
    from oracle import ref_shim
    warnings.filterwarnings("ignore")
    sys.path.insert(0, ref_shim.REF_DIR)
    import hifigan as ref_hg
    rng = np.random.default_rng(5)
    mel = (rng.standard_normal((2, 80, 6)) * 1.5 - 4).astype(np.float32)
    out = {"mel": mel, "weight_seed": 11}
    g = ref_hg.Generator(ref_hg.AttrDict(HIFIGAN_V2)).eval()
    g.remove_weight_norm()
    sd = hifigan_seeded_state(HIFIGAN_V2, 11)
    g.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    with torch.no_grad():
        out["wav_seeded"] = g(torch.from_numpy(mel)).numpy()
    cfg, ck = ho.load_reference_checkpoint(ref_shim.REF_DIR)
    g2 = ref_hg.Generator(ref_hg.AttrDict(cfg)).eval()
    g2.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()}, strict=True)
    g2.remove_weight_norm()
    with torch.no_grad():
        out["wav_checkpoint"] = g2(torch.from_numpy(mel)).numpy()
    assert np.abs(ho.generator(mel, sd, HIFIGAN_V2) - out["wav_seeded"]).max() < 2e-5
    assert np.abs(ho.generator(mel, ck, cfg) - out["wav_checkpoint"]).max() < 2e-5
    path = os.path.join(ROOT, "tests", "golden", "hifigan_v2_t6.npz")
    np.savez_compressed(path, **out)
    print(path)


if __name__ == "__main__":
    main()
    make_collate_fixture()
    make_hifigan_fixture()
