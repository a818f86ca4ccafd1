#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "hifigan or ragged or collate" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --long-steps 0 > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_tiny.err
python - <<'PY'
import json
def last(path):
    try:
        return json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as e:
        return None
j = last("gpurun_out/bench_tiny.json")
if j:
    print("tiny", round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), "ms; e2e", round(j["e2e"]["value"] / 1e6, 1), "; roofline", round(j["roofline"]["frac"], 3), {k: round(x, 4) for k, x in j["kernel_ms_per_step"].items()})
    for k, v in (j.get("configs") or {}).items():
        print(" ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "ms_per_utt_mean", "graph_replay_ms_per_utt", "achieved_tflops", "audio_rtf")})
PY
