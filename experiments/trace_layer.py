"""Per-role timeline of the 128-channel decoder layer kernel (es_umma_layer.cu, CTA 0) from in-kernel clock64 stamps.
Needs a -DES_LAYER_TRACE build:  ES_B200_LIB=.../libes_TRACE.so ES_TRACE_LAUNCH=<layer 0..3> python tools/trace_layer.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import efficientspeech_b200 as es  # noqa: E402
from efficientspeech_b200 import _cabi  # noqa: E402
from efficientspeech_b200.params import init_state_dict  # noqa: E402

cfg = es.VARIANTS["tiny"]
model = es.build_model("tiny")
es.load_numpy_state(model, init_state_dict(cfg, 0))
model = model.cuda().eval()
B, T = 256, 768
feats = torch.randn(B, T, 128, device="cuda")
lib = _cabi.load()
with torch.no_grad():
    for _ in range(3):
        model.decoder(feats)
    torch.cuda.synchronize()
    buf = torch.zeros(6 * 16 * 8, dtype=torch.int64, device="cuda")
    lib.es_debug_set_trace(buf.data_ptr())
    model.decoder(feats)
    torch.cuda.synchronize()
    lib.es_debug_set_trace(None)
tr = buf.cpu().numpy().reshape(6, 16, 8)
names = {0: ["start", "aready", "tfree", "committed"],
         1: ["start", "p0_loads", "p0_afree", "p0_stored", "p1_loads", "p1_afree", "p1_stored", "arrived"],
         2: ["start", "p0_loads", "p0_afree", "p0_stored", "p1_loads", "p1_afree", "p1_stored", "arrived"],
         3: ["start", "mma_done", "chunk0", "chunk1", "released"],
         4: ["start", "mma_done", "chunk0", "chunk1", "released"],
         5: ["start", "mma_done", "chunk0", "chunk1", "released"]}
t0 = tr[tr > 0].min()
print("layer launch picked (ES_TRACE_LAUNCH):", os.environ.get("ES_TRACE_LAUNCH", "0"))
for role, rn in [(0, "issue"), (1, "producer w0"), (2, "producer w7"), (3, "epilogue slot0 half0"), (4, "epilogue slot0 half1"), (5, "epilogue slot1 half0")]:
    print(f"== {rn}: events {names[role]}")
    for it in range(0, 12):
        row = tr[role, it, :len(names[role])]
        if row[0] == 0:
            continue
        print(f"  it {it:2d} start {row[0] - t0:8d}  deltas {[int(x) for x in np.diff(row)]}  total {int(row[-1] - row[0])}")
