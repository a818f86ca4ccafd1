#!/bin/bash
# A/B of bench lines only (no test suite): each argument is an environment setting for one bench run,
# "variant:ENV=..." selects the variant (default tiny).
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  variant=tiny; setting="$spec"
  case "$spec" in *:*) variant="${spec%%:*}"; setting="${spec#*:}";; esac
  env $setting timeout 150 python bench.py --variant $variant --no-cpu-baseline > gpurun_out/bench_ab$i.json 2> gpurun_out/bench_ab$i.err
  echo "[$variant $setting] rc=$?"; tail -2 gpurun_out/bench_ab$i.err
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/bench_ab$i.json") if l.startswith("{")][-1])
    print(round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), "ms; e2e", round(j["e2e"]["value"] / 1e6, 1), "; roofline", round(j["roofline"]["frac"], 3), {k: round(x, 4) for k, x in j["kernel_ms_per_step"].items()})
except Exception as e:
    print("no bench line", e)
PY
  i=$((i+1))
done
