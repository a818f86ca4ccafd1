"""Training-step pieces on the GPU: es_loss against the reference's own loss source (values) and its autograd (gradient
seeds); es_adamw_step against torch.optim.AdamW run on the CPU for several steps with the reference's schedule."""
import os
import sys

import numpy as np
import pytest
import torch

import efficientspeech_b200 as es
from efficientspeech_b200 import training

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_training_cpu import HAVE_REF, reference_loss_fn, synthetic_step  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.skipif(not HAVE_REF, reason="reference model.py not staged (oracle/build_ref.py)")
@pytest.mark.parametrize("seed,B,N,T", [(0, 3, 11, 40), (1, 16, 128, 768), (2, 1, 7, 9)])
def test_loss_and_gradient_seeds_match_reference(seed, B, N, T):
    y_hat, y, x = synthetic_step(seed, B, N, T)
    for k in ("mel", "pitch", "energy", "duration"):
        y_hat[k].requires_grad_(True)
    ref = reference_loss_fn()(None, y_hat, y, x)
    total = 10. * ref[0] + 2. * ref[1] + 2. * ref[2] + ref[3]                     # model.py:215
    total.backward()
    to = lambda d: {k: (v.detach().to(DEV) if torch.is_tensor(v) else v) for k, v in d.items()}
    got, g = training.loss(to(y_hat), to(y), to(x), with_grads=True)
    for a, b in zip(got, ref):
        assert abs(float(a) - float(b)) <= 1e-5 * max(1.0, abs(float(b)))
    assert abs(float(g["total"]) - float(total)) <= 1e-5 * abs(float(total))
    assert np.abs(g["mel"].cpu().numpy() - y_hat["mel"].grad.numpy()).max() <= 1e-7
    for k in ("pitch", "energy", "duration"):
        want = y_hat[k].grad.numpy()[..., 0]
        assert np.abs(g[k].cpu().numpy() - want).max() <= 1e-6 * max(1.0, np.abs(want).max()), k
    # deterministic: same bits on a second call
    got2, _ = training.loss(to(y_hat), to(y), to(x), with_grads=True)
    assert all(float(a) == float(b) for a, b in zip(got, got2))


def test_fused_adamw_tracks_torch_adamw():
    torch.manual_seed(0)
    shapes = [(153, 128), (32, 128, 3), (32,), (1, 32), (80, 128)]
    ref_params = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    params = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref_params]
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-3, weight_decay=1e-6)         # model.py:280
    sched = torch.optim.lr_scheduler.LambdaLR(ref_opt, lambda s: training.lr_lambda(s, 50, 5000))
    opt = training.FusedAdamW(params, lr=1e-3, weight_decay=1e-6)
    versions = [p._version for p in params]
    for step in range(60):
        for rp, p in zip(ref_params, params):
            g = torch.randn(rp.shape) * (1.0 + step)
            rp.grad = g.clone()
            p.grad.copy_(g)
        scale = training.lr_lambda(step, 50, 5000)
        ref_opt.step()
        sched.step()
        opt.step(lr_scale=scale)
    assert all(p._version > v for p, v in zip(params, versions))
    for rp, p in zip(ref_params, params):
        a, b = p.detach().cpu().numpy(), rp.detach().numpy()
        assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(b).max())
