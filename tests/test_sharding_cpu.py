"""world_size-2 gloo test of the N>1 plumbing (weight broadcast + batch sharding), CPU only."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import efficientspeech_b200 as es
from efficientspeech_b200.params import init_state_dict, state_checksum
from efficientspeech_b200.sharding import broadcast_weights, shard_batch, shard_bounds
from efficientspeech_b200.synthetic import make_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = es.VARIANTS["tiny"]
    model = es.build_model("tiny")
    es.load_numpy_state(model, init_state_dict(cfg, seed=10 + rank))      # ranks start different
    v0 = [p._version for p in model.parameters()]
    broadcast_weights(model, src=0)
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    out[rank] = (state_checksum(sd), all(p._version > a for p, a in zip(model.parameters(), v0)))
    dist.destroy_process_group()


def test_weight_broadcast_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    want = state_checksum(init_state_dict(es.VARIANTS["tiny"], seed=10))
    assert out[0][0] == want and out[1][0] == want
    assert out[1][1]                     # versions bumped -> packed image rebuilt on next forward


def _allreduce_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from efficientspeech_b200.training import allreduce_flat
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    allreduce_flat(g)                                      # the ONE collective of the training step: flat gradient, averaged
    out[rank] = g.tolist()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_allreduce_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    want = (torch.arange(10, dtype=torch.float32) * 1.5).tolist()
    assert out[0] == want and out[1] == want


def test_shards_partition_the_batch():
    cfg = es.VARIANTS["tiny"]
    batch = make_batch(cfg, 13, 16, seed=0, ragged=True, fixed_duration=None)
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            lo, hi = shard_bounds(13, r, world)
            sb = shard_batch(batch, r, world)
            assert sb["phoneme"].shape == (hi - lo, 16)          # global N kept
            seen.extend(range(lo, hi))
            assert np.array_equal(sb["duration"], batch["duration"][lo:hi])
        assert seen == list(range(13))
