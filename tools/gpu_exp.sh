#!/bin/bash
# Timing experiments: one tiny bench line per experiment library efficientspeech_b200/exp/libes_<NAME>.so given as arguments
# (results of these builds are numerically WRONG on purpose; only their kernel times are read).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for e in "$@"; do
  ES_B200_LIB=$PWD/efficientspeech_b200/exp/libes_$e.so timeout 150 python bench.py --no-sub --no-cpu-baseline --long-steps 0 > gpurun_out/bench_exp_$e.json 2> gpurun_out/bench_exp_$e.err
  echo "exp $e rc=$?"; tail -1 gpurun_out/bench_exp_$e.err
  python - <<PY
import json
try:
    x = json.loads([l for l in open("gpurun_out/bench_exp_$e.json") if l.startswith("{")][-1])
    print("$e", round(x["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in x["kernel_ms_per_step"].items()}, [round(p["ms"], 4) for p in x["roofline"]["per_launch_position"]])
except Exception as ex:
    print("no line", ex)
PY
done
