"""The oracle's restatement of the reference collate_fn (datamodule.py:29-76) against the reference's own source, run
where the reference tree is mounted, and against a committed golden fixture everywhere."""
import os

import numpy as np
import pytest

from oracle import es_oracle, ref_shim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collate_b7.npz")


def ragged_items(seed, B, lo=3, hi=40, train=True):
    rng = np.random.default_rng(seed)
    items = []
    for _ in range(B):
        n = int(rng.integers(lo, hi + 1))
        it = {"phoneme": rng.integers(1, 152, size=n).astype(np.int64)}
        if train:
            it["pitch"] = rng.standard_normal(n).astype(np.float32)
            it["energy"] = rng.standard_normal(n).astype(np.float32)
            it["duration"] = rng.integers(0, 12, size=n).astype(np.int64)
        items.append(it)
    return items


def reference_collate(items):
    """Run the reference's collate_fn source (extracted, not imported: datamodule.py needs lightning)."""
    fns, ns = ref_shim.reference_functions("utils/tools.py", ["pad_1D", "pad_2D", "get_mask_from_lengths"])
    got, ns2 = ref_shim.reference_functions("datamodule.py", ["collate_fn"])
    ns2.update({k: v for k, v in ns.items() if k in fns})
    x = [dict(it, text="") for it in items]
    y = [{"mel": np.zeros((int(np.sum(it.get("duration", [1]))), 80), np.float32)} for it in items]
    bx, _ = got["collate_fn"](None, list(zip(x, y)))
    return {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in bx.items()}


@pytest.mark.skipif(not (os.path.isfile(os.path.join(ref_shim._MOUNTED, "datamodule.py")) or
                         os.path.isfile(os.path.join(ref_shim.REF_DIR, "datamodule.py"))), reason="reference sources not available")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_collate_matches_reference_source(seed):
    items = ragged_items(seed, 9)
    lens = [len(it["phoneme"]) for it in items]
    if len(set(lens)) != len(lens):                       # tie order is unspecified upstream: make lengths distinct
        for k, it in enumerate(items):
            n = 3 + 4 * k
            for key in it:
                it[key] = np.resize(it[key], n)
    ref = reference_collate(items)
    o = es_oracle.collate(items)
    for k in ("phoneme", "phoneme_len", "phoneme_mask", "pitch", "energy", "duration", "mel_len"):
        assert np.array_equal(np.asarray(ref[k]), o[k]), k
    assert ref["phoneme"].dtype == np.int32 and ref["mel_len"].dtype == np.int32


def test_oracle_collate_golden():
    z = np.load(GOLDEN)
    B = int(z["B"])
    items = [{k: z[f"{k}_{i}"] for k in ("phoneme", "pitch", "energy", "duration")} for i in range(B)]
    o = es_oracle.collate(items)
    for k in ("perm", "phoneme", "phoneme_len", "phoneme_mask", "pitch", "energy", "duration", "mel_len"):
        assert np.array_equal(z["out_" + k], o[k]), k


def test_stable_ties_and_inference_only_items():
    items = [{"phoneme": np.arange(1, 6)}, {"phoneme": np.arange(1, 9)}, {"phoneme": np.arange(2, 7)}, {"phoneme": np.arange(1, 9)}]
    o = es_oracle.collate(items)
    assert o["perm"].tolist() == [1, 3, 0, 2]             # equal lengths keep their input order
    assert "pitch" not in o and "mel_len" not in o
    assert o["phoneme_mask"].sum() == 6
