"""Per-role timeline of the tcgen05 decoder-layer kernel (CTA 0), from in-kernel clock64 stamps.
Run under gpurun: python tools/trace_decoder.py [block_end:0|1]"""
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
    buf = torch.zeros(4 * 32 * 8, dtype=torch.int64, device="cuda")
    # the decoder forward launches proj, 4 layers, mel: keep the stamps of the LAST traced launch of
    # interest by enabling the trace for exactly one forward and reading what the final kernels left
    which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    lib.es_debug_set_trace(buf.data_ptr())
    model.decoder(feats)
    torch.cuda.synchronize()
    lib.es_debug_set_trace(None)
tr = buf.cpu().numpy().reshape(4, 32, 8)
names = {0: ["start", "aready", "tfree", "mma_issued", "x_issued"],
         1: ["start", "x_landed", "pass0_done", "a_free", "stored", "arrived"],
         2: ["start", "mma_done", "tmem_ld", "ln1", "skip_ln2", "stored"],
         3: ["start", "mma_done", "tmem_ld", "ln1", "skip_ln2", "stored"]}
t0 = tr[tr > 0].min()
print("launch picked (ES_TRACE_LAUNCH; 0=proj, 1..4=layers, 5=mel):", os.environ.get("ES_TRACE_LAUNCH", "0"))
for role, rn in [(0, "issue"), (1, "producer w8"), (2, "epilogue g0"), (3, "epilogue g1")]:
    print(f"== {rn}: events {names[role]}")
    for it in range(2, 12):
        row = tr[role, it, :len(names[role])]
        if row[0] == 0:
            continue
        rel = row - t0
        d = np.diff(row)
        print(f"  it {it:2d} start {rel[0]:8d}  deltas {list(d)}  total {row[-1] - row[0]}")
