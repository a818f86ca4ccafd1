"""HiFi-GAN sub-record of bench.py on its own (B x T from argv), for quick A/B runs."""
import json
import sys
import types

sys.argv = [sys.argv[0]] + sys.argv[1:]
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 768
a = types.SimpleNamespace(gpus=1)
ctx = bench.Ctx(a)
r = bench.measure_hifigan(ctx, B, T, 3)
print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "achieved_tflops", "audio_rtf")}), r["roofline"]["frac"])
