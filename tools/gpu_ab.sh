#!/bin/bash
# One GPU-box call: parity suite, then the tiny bench with the environment settings given as arguments (A/B),
# e.g.  tools/gpu_ab.sh "ES_FUSED_PHONEME=1" "ES_FUSED_PHONEME=0"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -40 gpurun_out/pytest_gpu.log
i=0
for setting in "$@"; do
  env $setting timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_tiny_ab$i.json 2> gpurun_out/bench_tiny_ab$i.err
  echo "[$setting] rc=$?"; tail -3 gpurun_out/bench_tiny_ab$i.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_tiny_ab$i.json"))
    print(round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), "ms; e2e", round(j["e2e"]["value"] / 1e6, 1), "; roofline", round(j["roofline"]["frac"], 3), j["kernel_ms_per_step"])
except Exception as e:
    print("no bench line", e)
PY
  i=$((i+1))
done
