"""Time the training step (forward_train + loss + backward + AdamW) on synthetic LJSpeech-shaped batches and list the
kernels by time share.  python tools/bench_train.py [B] [N] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import efficientspeech_b200 as es
from efficientspeech_b200 import training
from efficientspeech_b200.config import VARIANTS
from efficientspeech_b200.synthetic import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = "cuda:0"
cfg = VARIANTS["tiny"]
m = es.build_model("tiny").to(dev)
step = training.TrainStep(m)
batches = []
for i in range(4):
    b = make_batch(cfg, B, N, seed=i, ragged=True, fixed_duration=None, max_dur=12)
    T = int(b["mel_len"].max())
    x = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items()}
    x["max_mel_len"] = T
    y = {"mel": torch.randn(B, T, cfg.n_mel, device=dev)}
    batches.append((x, y, int(b["mel_len"].sum())))
for i in range(8):          # every batch shape twice: the caching allocator has to have seen all of them
    out = step(*batches[i % 4][:2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
frames = 0
for i in range(steps):
    out = step(*batches[i % 4][:2])
    frames += batches[i % 4][2]
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = e0.elapsed_time(e1) / steps
print(f"B={B} N={N} T~{batches[0][0]['max_mel_len']}: {ms:.2f} ms/step (device), {wall / steps * 1e3:.2f} ms wall, "
      f"{frames / (ms * steps) * 1e3 / 1e6:.2f} M frames/s, loss {float(out[0]):.4f}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
import os
if os.environ.get("ES_TRAIN_TC") == "0":
    from efficientspeech_b200 import train_ops
    train_ops.set_tensor_core(False)
    for i in range(3):
        out = step(*batches[i % 4][:2])
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        out = step(*batches[i % 4][:2])
    e1.record()
    torch.cuda.synchronize()
    print(f"SIMT GEMMs: {e0.elapsed_time(e1) / steps:.2f} ms/step")
    train_ops.set_tensor_core(True)
    for i in range(4):
        out = step(*batches[i % 4][:2])
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        out = step(*batches[i % 4][:2])
    e1.record()
    torch.cuda.synchronize()
    print(f"tensor-core GEMMs again: {e0.elapsed_time(e1) / steps:.2f} ms/step")
gstep = training.TrainStep(m, use_graphs=True)
gstep.opt = step.opt            # same flat buffers
for i in range(8):
    out = gstep(*batches[i % 4][:2])
torch.cuda.synchronize()
e0.record()
for i in range(steps):
    out = gstep(*batches[i % 4][:2])
e1.record()
torch.cuda.synchronize()
print(f"graph replay (forward+loss+backward), eager AdamW: {e0.elapsed_time(e1) / steps:.2f} ms/step, loss {float(out[0]):.4f}")
# host-only cost of a step: enqueue without waiting (the queue is deep enough for one step)
torch.cuda.synchronize()
t0 = time.perf_counter()
step(*batches[0][:2])
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue time of one step: {(t1 - t0) * 1e3:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(*batches[0][:2])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

# idle time on the device in front of each kernel, grouped by kernel name
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
gaps, prev_end = {}, None
for e in evs:
    if prev_end is not None:
        g = max(0.0, e.time_range.start - prev_end)
        k = e.name[:70]
        a = gaps.setdefault(k, [0.0, 0, 0.0])
        a[0] += g; a[1] += 1; a[2] += e.time_range.end - e.time_range.start
    prev_end = max(prev_end or 0, e.time_range.end)
print("idle gap in front of kernel (us total, launches, kernel us total):")
for k, a in sorted(gaps.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {a[0]:9.1f} {a[1]:5d} {a[2]:9.1f}  {k}")
print("span", (evs[-1].time_range.end - evs[0].time_range.start) / 1e3, "ms")
