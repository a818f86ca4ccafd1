"""The numpy oracle reproduces the reference's own outputs (fixtures made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from helpers import golden_files, load_golden
from oracle import es_oracle


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_fixture(path):
    vname, cfg, sd, batch, g = load_golden(path)
    tf = es_oracle.phoneme2mel(batch, sd, train=True)
    assert np.array_equal(tf["mel_len"], g["tf_mel_len"]) and tf["mel_len"].dtype == np.int32
    assert np.array_equal(tf["_src"], g["tf_src"])            # length-regulator index map: bit-exact
    assert tf["mel"].shape == g["tf_mel"].shape
    assert np.abs(tf["mel"] - g["tf_mel"]).max() < 2e-5
    for k in ("pitch", "energy", "duration"):
        assert np.abs(tf[k] - g["tf_" + k]).max() < 1e-5
    fr = es_oracle.phoneme2mel(batch, sd, train=False)
    assert np.array_equal(fr["mel_len"], g["fr_mel_len"])
    assert fr["mel"].shape == g["fr_mel"].shape
    assert np.abs(fr["mel"] - g["fr_mel"]).max() < 2e-5
    assert np.abs(fr["duration"] - g["fr_duration"]).max() < 1e-5


def test_fixtures_exist():
    assert len(golden_files()) == 9


def test_oracle_fp64_close_to_fp32():
    vname, cfg, sd, batch, g = load_golden(golden_files()[0])
    a = es_oracle.phoneme2mel(batch, sd, train=True, dtype=np.float32)
    b = es_oracle.phoneme2mel(batch, sd, train=True, dtype=np.float64)
    assert np.abs(a["mel"] - b["mel"]).max() < 2e-5


def test_length_regulator_properties():
    rng = np.random.default_rng(0)
    for _ in range(20):
        B, N, C = int(rng.integers(1, 5)), int(rng.integers(1, 40)), 4
        dur = rng.integers(0, 6, size=(B, N))
        feats = rng.standard_normal((B, N, C)).astype(np.float32)
        masks = np.zeros((B, N, C), dtype=bool)
        f, m, ml, src = es_oracle.feature_upsampler(feats, masks, dur)
        assert np.array_equal(ml, dur.sum(1))
        for b in range(B):
            cs = np.cumsum(dur[b])
            want = np.searchsorted(cs, np.arange(ml[b]), side="right")   # SURVEY A.7 closed form
            assert np.array_equal(src[b, :ml[b]], want)
            assert (src[b, ml[b]:] == -1).all() and m[b, ml[b]:].all() and not m[b, :ml[b]].any()
            assert np.array_equal(f[b, :ml[b]], feats[b, want])


def test_reference_pin_when_reference_present():
    from oracle.ref_shim import reference_available
    if not reference_available():
        pytest.skip("reference tree not mounted (GPU box)")
    from oracle.ref_shim import build_reference_model, run_reference
    vname, cfg, sd, batch, g = load_golden(golden_files()[0])
    r = run_reference(build_reference_model(cfg, sd), batch, train=True)
    assert np.array_equal(r["mel"], g["tf_mel"])
