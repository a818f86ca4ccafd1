#!/bin/bash
# Full GPU suite (all failures listed) + small / base bench lines.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for v in small base; do
  timeout 300 python bench.py --variant $v --no-cpu-baseline --no-sub > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/bench_$v.json") if l.startswith("{")][-1])
print("$v", round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), {k: round(x, 4) for k, x in j["kernel_ms_per_step"].items()})
PY
done
