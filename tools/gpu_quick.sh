#!/bin/bash
# Quick GPU iteration: the GPU parity suite (optionally a -k selection in $1), then tiny bench lines for each
# environment setting given in the remaining arguments ("" = defaults).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
sel="$1"; shift
if [ -n "$sel" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "$sel" > gpurun_out/pytest_gpu.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
fi
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
i=0
for setting in "$@"; do
  env $setting timeout 200 python bench.py --no-sub --no-cpu-baseline --long-steps 0 > gpurun_out/bench_q$i.json 2> gpurun_out/bench_q$i.err
  echo "[$setting] rc=$?"; tail -2 gpurun_out/bench_q$i.err
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/bench_q$i.json") if l.startswith("{")][-1])
    print(round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), "ms; e2e", round(j["e2e"]["value"] / 1e6, 1), "; roofline", round(j["roofline"]["frac"], 3),
          {k: round(x, 4) for k, x in j["kernel_ms_per_step"].items()}, [round(p["ms"], 4) for p in j["roofline"].get("per_launch_position", [])])
except Exception as e:
    print("no bench line", e)
PY
  i=$((i+1))
done
