"""Host-side logic of the training path that needs no GPU: the pooled padding mask against the reference's own
expression, the no-CPU-fallback rule of every differentiable operator, the synthetic training batches of bench.py."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from efficientspeech_b200 import train_ops as ops
from efficientspeech_b200 import training

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,pool", [(128, 2), (129, 2), (11, 2), (10, 4), (13, 4), (7, 1)])
def test_pool_mask_matches_reference_expression(n, pool):
    """layers/blocks.py:52-58: pad the mask with True up to a multiple of `pool`, then a max over groups of `pool`."""
    from einops import reduce
    g = torch.Generator().manual_seed(n * 10 + pool)
    lens = torch.randint(1, n + 1, (5,), generator=g)
    mask = torch.arange(n)[None, :] >= lens[:, None]
    want = mask
    if pool > 1:
        mod = n % pool
        m = F.pad(mask, [0, pool - mod], value=True) if mod > 0 else mask
        want = reduce(m, "b (n p) -> b n", "max", p=pool)
    got = training._pool_mask(mask, pool)
    assert got.dtype == torch.bool and torch.equal(got, want)


def test_operators_refuse_cpu_tensors():
    x, w, b = torch.randn(2, 3, 8), torch.randn(4, 8), torch.randn(4)
    for call in (lambda: ops.linear(x, w, b),
                 lambda: ops.conv1d(x, torch.randn(4, 8, 3), None, 1, 1),
                 lambda: ops.dwconv1d(x, torch.randn(8, 1, 5), torch.randn(8)),
                 lambda: ops.layernorm(x, torch.ones(8), torch.zeros(8)),
                 lambda: ops.act(x, ops.ACT_TANH),
                 lambda: ops.add(x, x),
                 lambda: ops.embedding(torch.zeros(3, dtype=torch.long), torch.randn(5, 8))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    from types import SimpleNamespace
    model = SimpleNamespace(encoder=SimpleNamespace(encoder=None), decoder=None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        training.forward_train(model, {"phoneme": torch.zeros(2, 3, dtype=torch.long)})


def test_train_step_graph_mode_needs_the_length_bound():
    step = training.TrainStep.__new__(training.TrainStep)
    step.use_graphs, step._graphs, step._pool = True, {}, None
    with pytest.raises(RuntimeError, match="max_mel_len"):
        step({"phoneme": torch.zeros(2, 3, dtype=torch.long)}, {"mel": torch.zeros(2, 4, 80)})


def test_bench_training_batches_are_consistent():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from efficientspeech_b200.config import VARIANTS
    cfg = VARIANTS["tiny"]
    for b in bench.train_batches(cfg, 6, 32, 2, seed0=3):
        T = int(b["mel_len"].max())
        assert b["mel"].shape == (6, T, cfg.n_mel) and b["mel"].dtype == np.float32
        assert (b["duration"].sum(axis=1) == b["mel_len"]).all()
        assert (b["duration"][b["phoneme_mask"]] == 0).all() and (b["phoneme"][b["phoneme_mask"]] == 0).all()
        assert b["phoneme_len"].max() == 32 and (np.diff(b["phoneme_len"]) <= 0).all()      # sorted like collate_fn
