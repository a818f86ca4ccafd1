#!/bin/bash
# One GPU-box call: parity suite, then the tiny bench under the decoder-gather modes given as arguments (A/B).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -22 gpurun_out/pytest_gpu.log
for mode in "$@"; do
  ES_DEC_GATHER_MODE=$mode timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_tiny_g$mode.json 2> gpurun_out/bench_tiny_g$mode.err
  echo "mode $mode rc=$?"
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_tiny_g$mode.json"))
    print($mode, round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), "ms; roofline", round(j["roofline"]["frac"], 3), j["kernel_ms_per_step"])
except Exception as e:
    print("no bench line", e)
PY
done
