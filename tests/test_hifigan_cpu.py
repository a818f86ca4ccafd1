"""HiFi-GAN row, CPU side: the numpy oracle (oracle/hifigan_oracle.py) against the reference's outputs -- the committed
fixture everywhere, the live reference hifigan.Generator where oracle/_ref is staged -- and the drop-in surface of
efficientspeech_b200.hifigan (state-dict names, weight-norm round trip, loud failure without CUDA)."""
import importlib.util
import os
import sys
import warnings

import numpy as np
import pytest
import torch

import efficientspeech_b200 as es
from oracle import hifigan_oracle as ho
from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "hifigan_v2_t6.npz")
TOL_WAV = 1e-4     # max-abs on a waveform in (-1, 1): fp32 summation-order noise through ~80 convolutions stays below 3e-5


def _make_golden_module():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "oracle", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


def staged_checkpoint():
    p = os.path.join(ref_shim.REF_DIR, "hifigan", "LJ_V2", "generator_v2")
    return p if os.path.isfile(p) else None


def test_oracle_matches_reference_fixture_seeded_weights():
    z = np.load(GOLDEN)
    mg = _make_golden_module()
    sd = mg.hifigan_seeded_state(mg.HIFIGAN_V2, int(z["weight_seed"]))
    got = ho.generator(z["mel"], sd, mg.HIFIGAN_V2)
    assert got.shape == z["wav_seeded"].shape == (2, 1, 6 * 256)
    assert np.abs(got - z["wav_seeded"]).max() <= TOL_WAV


@pytest.mark.skipif(staged_checkpoint() is None, reason="reference checkpoint not staged (oracle/build_ref.py)")
def test_oracle_matches_reference_fixture_and_live_reference_with_the_checkpoint():
    z = np.load(GOLDEN)
    cfg, ck = ho.load_reference_checkpoint(ref_shim.REF_DIR)
    got = ho.generator(z["mel"], ck, cfg)
    assert np.abs(got - z["wav_checkpoint"]).max() <= TOL_WAV
    warnings.filterwarnings("ignore")
    sys.path.insert(0, ref_shim.REF_DIR)
    import hifigan as ref_hg
    g = ref_hg.Generator(ref_hg.AttrDict(cfg)).eval()
    g.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()}, strict=True)
    g.remove_weight_norm()
    mel = (np.random.default_rng(1).standard_normal((1, 80, 9)) - 3).astype(np.float32)
    with torch.no_grad():
        want = g(torch.from_numpy(mel)).numpy()
    assert np.abs(ho.generator(mel, ck, cfg) - want).max() <= TOL_WAV
    # our holder takes the same checkpoint, strict, before and after remove_weight_norm (model.py:41-44)
    G = es.hifigan.Generator(es.hifigan.AttrDict(cfg))
    assert list(G.state_dict()) == list(ref_hg.Generator(ref_hg.AttrDict(cfg)).state_dict())
    G.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()}, strict=True)
    w_before = es.hifigan._effective_weight(G.conv_pre).numpy()
    G.eval()
    G.remove_weight_norm()
    assert "conv_pre.weight" in G.state_dict() and "conv_pre.weight_g" not in G.state_dict()
    assert np.abs(G.conv_pre.weight.detach().numpy() - w_before).max() <= 1e-6
    assert np.abs(w_before - ho.effective_weight(ck, "conv_pre")).max() <= 1e-6


def test_generator_has_no_cpu_fallback_and_validates():
    mg = _make_golden_module()
    G = es.hifigan.Generator(es.hifigan.AttrDict(mg.HIFIGAN_V2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G(torch.zeros(1, 80, 4))
    with pytest.raises(ValueError):
        es.hifigan.Generator(es.hifigan.AttrDict(dict(mg.HIFIGAN_V2, resblock="2")))
    assert G.total_upsampling == 256
