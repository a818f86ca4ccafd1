#!/bin/bash
# Round-2 evidence in one 1-GPU call: parity suite, bench lines (tiny with sub-records + CPU baseline, reference arm),
# ncu launch list of the bench command, ncu --set full of the decoder-side kernels (traffic), D2H ceiling.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "bench tiny rc=$?"; tail -2 gpurun_out/bench_tiny.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
for v in small base; do
  timeout 300 python bench.py --variant $v --no-cpu-baseline --no-sub > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
done
timeout 120 python tools/d2h_ceiling.py > gpurun_out/d2h_ceiling_1gpu.json 2> gpurun_out/d2h_ceiling_1gpu.err; echo "d2h rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tiny.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub --long-steps 0 > gpurun_out/ncu_bench_tiny.log 2>&1; echo "ncu list rc=$?"
timeout 500 ncu --set full --import-source on --clock-control none -k regex:umma_dec_kernel --launch-skip 6 --launch-count 6 -o gpurun_out/prof_dec_r02 -f python tools/run_forward.py tiny 2 > gpurun_out/ncu_dec_r02.log 2>&1; echo "ncu dec rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:hg_conv_kernel --launch-skip 40 --launch-count 4 -o gpurun_out/prof_hifigan_r02 -f python tools/run_hifigan.py 4 192 > gpurun_out/ncu_hifigan_r02.log 2>&1; echo "ncu hifigan rc=$?"
python - <<'PY'
import json
def last(path):
    try:
        return json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as e:
        return None
for v in ("tiny", "small", "base"):
    j = last(f"gpurun_out/bench_{v}.json")
    if not j: print(v, "no line"); continue
    print(v, round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), "ms; e2e", round(j["e2e"]["value"] / 1e6, 1), "; roofline", round(j["roofline"]["frac"], 3),
          {k: round(x, 4) for k, x in j["kernel_ms_per_step"].items()}, "long", (j.get("long_run") or {}).get("value"))
    for k, s in (j.get("configs") or {}).items():
        print("  ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in s.items() if kk in ("value", "ms_per_step", "ms_per_utt_mean", "graph_replay_ms_per_utt", "achieved_tflops")})
    if j.get("cpu_baseline"): print("   cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"], j["cpu_baseline"]["kind"])
r = last("gpurun_out/bench_reference.json")
print("reference", r and (r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["kind"], {k: v.get("value") for k, v in (r.get("configs") or {}).items()}))
PY
