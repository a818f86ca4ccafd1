"""Timeline of the fused phoneme kernel (CTA 0, thread 0) from in-kernel clock64 stamps.
Event codes: 0 utterance start, 4/5 weight wait begin/end, 1 GEMM (with weights) done, 3 attention GEMM done,
2 phase sync passed, 6 row reduction barrier passed, 9 utterance end.  Run under gpurun."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import efficientspeech_b200 as es  # noqa: E402
from efficientspeech_b200 import _cabi  # noqa: E402
from efficientspeech_b200.params import init_state_dict  # noqa: E402
from efficientspeech_b200.synthetic import make_batch  # noqa: E402

cfg = es.VARIANTS["tiny"]
model = es.build_model("tiny")
es.load_numpy_state(model, init_state_dict(cfg, 0))
model = model.cuda().eval()
batch = make_batch(cfg, 256, 128, seed=1, ragged=False, fixed_duration=6)
x = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
x["max_mel_len"] = 768
lib = _cabi.load()
model.encoder.materialize_features = False
with torch.no_grad():
    for _ in range(3):
        model.encoder(x, train=True)
    torch.cuda.synchronize()
    buf = torch.zeros(2 * 256, dtype=torch.int64, device="cuda")
    lib.es_debug_set_phoneme_trace(buf.data_ptr())
    model.encoder(x, train=True)
    torch.cuda.synchronize()
    lib.es_debug_set_phoneme_trace(None)
tr = buf.cpu().numpy().reshape(2, 128, 2)
names = {0: "start", 1: "gemm_done", 2: "phase_sync", 3: "attn_gemm_done", 4: "w_wait_begin", 5: "w_wait_end", 6: "row_reduce", 9: "end"}
for u in range(2):
    ev = [(int(t), int(c)) for t, c in tr[u] if t > 0]
    if not ev:
        continue
    t0 = ev[0][0]
    print(f"== utterance {u}: {len(ev)} events, total {ev[-1][0] - t0} cycles")
    prev = t0
    acc = {}
    for t, c in ev:
        acc[names.get(c, str(c))] = acc.get(names.get(c, str(c)), 0) + (t - prev)
        print(f"  {t - t0:8d}  +{t - prev:6d}  {names.get(c, c)}")
        prev = t
    print("  time attributed to the interval ENDING at each event kind:", acc)
