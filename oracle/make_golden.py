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


if __name__ == "__main__":
    main()
    make_collate_fixture()
