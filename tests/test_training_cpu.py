"""Training-step pieces, CPU side: the schedule against the reference's own get_lr_scheduler source (model.py:77-101),
and the oracle used by the GPU tests (the reference's `loss` method source, model.py:167-209, extracted and run under
torch on CPU) against a hand computation."""
import math
import os

import numpy as np
import pytest
import torch

from efficientspeech_b200 import training
from oracle import ref_shim

HAVE_REF = os.path.isfile(os.path.join(ref_shim._MOUNTED, "model.py")) or os.path.isfile(os.path.join(ref_shim.REF_DIR, "model.py"))


def reference_loss_fn():
    """EfficientSpeech.loss as written upstream (the module itself needs lightning: the method source is compiled alone)."""
    fns, ns = ref_shim.reference_functions("model.py", ["loss"])
    ns["nn"] = torch.nn
    return fns["loss"]


def synthetic_step(seed, B=3, N=11, T=40, C=80):
    rng = np.random.default_rng(seed)
    plen = np.sort(rng.integers(3, N + 1, size=B))[::-1].copy()
    plen[0] = N
    pmask = np.arange(N)[None, :] >= plen[:, None]
    mel_len = rng.integers(5, T + 1, size=B).astype(np.int32)
    mel_len[0] = T
    x = {"phoneme_mask": torch.from_numpy(pmask), "mel_mask": torch.from_numpy(np.arange(T)[None, :] >= mel_len[:, None]),
         "mel_len": torch.from_numpy(mel_len), "pitch": torch.from_numpy(rng.standard_normal((B, N)).astype(np.float32)),
         "energy": torch.from_numpy(rng.standard_normal((B, N)).astype(np.float32)),
         "duration": torch.from_numpy(rng.integers(0, 9, size=(B, N)).astype(np.int32))}
    y = {"mel": torch.from_numpy(rng.standard_normal((B, T, C)).astype(np.float32))}
    y_hat = {"mel": torch.from_numpy(rng.standard_normal((B, T, C)).astype(np.float32)),
             "pitch": torch.from_numpy(rng.standard_normal((B, N, 1)).astype(np.float32)),
             "energy": torch.from_numpy(rng.standard_normal((B, N, 1)).astype(np.float32)),
             "duration": torch.from_numpy(np.abs(rng.standard_normal((B, N, 1)) * 4).astype(np.float32))}
    return y_hat, y, x


@pytest.mark.skipif(not HAVE_REF, reason="reference model.py not available (mounted tree or oracle/_ref)")
def test_lr_lambda_matches_reference_source():
    import types
    fns, ns = ref_shim.reference_functions("model.py", ["get_lr_scheduler"])
    ns["math"] = math
    captured = {}
    ns["LambdaLR"] = lambda opt, fn: captured.setdefault("fn", fn)
    fns["get_lr_scheduler"](None, 50, 5000, min_lr=0)
    for step in (0, 1, 49, 50, 51, 2500, 4999, 5000):
        assert training.lr_lambda(step, 50, 5000, 0.0) == captured["fn"](step)


@pytest.mark.skipif(not HAVE_REF, reason="reference model.py not available (mounted tree or oracle/_ref)")
def test_reference_loss_source_runs_and_matches_a_hand_computation():
    y_hat, y, x = synthetic_step(0)
    mel, pitch, energy, dur = reference_loss_fn()(None, y_hat, y, x)
    valid = ~x["mel_mask"].numpy()
    d = np.abs(y_hat["mel"].numpy() - y["mel"].numpy())[valid]
    assert abs(float(mel) - d.mean()) < 1e-6
    pv = ~x["phoneme_mask"].numpy()
    assert abs(float(pitch) - ((y_hat["pitch"].numpy()[..., 0] - x["pitch"].numpy())[pv] ** 2).mean()) < 1e-6
    dd = np.log(y_hat["duration"].numpy()[..., 0][pv] + 1) - np.log(x["duration"].numpy()[pv].astype(np.float32) + 1)
    assert abs(float(dur) - (dd ** 2).mean()) < 1e-6
