#!/bin/bash
# Round-end evidence in one GPU-box call: parity suite, bench lines (tiny / small / base / reference arm),
# ncu launch list of the bench command, ncu --set full captures of the decoder kernels and the phoneme kernel.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "bench tiny rc=$?"
for v in small base; do
  timeout 200 python bench.py --variant $v --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tiny.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_tiny.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:umma_dec_kernel --launch-skip 6 --launch-count 6 -o gpurun_out/prof_dec_final -f python tools/run_forward.py tiny 2 > gpurun_out/ncu_dec_final.log 2>&1; echo "ncu dec rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:umma_phoneme --launch-skip 1 --launch-count 1 -o gpurun_out/prof_phoneme_final -f python tools/run_forward.py tiny 2 > gpurun_out/ncu_phoneme_final.log 2>&1; echo "ncu phoneme rc=$?"
python - <<'PY'
import json
for v in ("tiny", "small", "base"):
    try:
        j = json.load(open(f"gpurun_out/bench_{v}.json"))
        print(v, round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), "ms; e2e", round(j["e2e"]["value"] / 1e6, 1),
              "; roofline", round(j["roofline"]["frac"], 3), {k: round(x, 4) for k, x in j["kernel_ms_per_step"].items()})
    except Exception as e:
        print(v, "no bench line", e)
print(open("gpurun_out/bench_reference.json").read()[:400])
PY
