"""A few eager training steps (for ncu): python tools/run_train_step.py [B] [N] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import efficientspeech_b200 as es
from efficientspeech_b200 import training
from efficientspeech_b200.config import VARIANTS
from efficientspeech_b200.synthetic import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = "cuda:0"
cfg = VARIANTS["tiny"]
m = es.build_model("tiny").to(dev)
step = training.TrainStep(m)
b = make_batch(cfg, B, N, seed=0, ragged=True, fixed_duration=None, max_dur=12)
T = int(b["mel_len"].max())
x = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items()}
x["max_mel_len"] = T
y = {"mel": torch.randn(B, T, cfg.n_mel, device=dev)}
for _ in range(steps):
    out = step(x, y)
torch.cuda.synchronize()
print("loss", float(out[0]), "T", T)
