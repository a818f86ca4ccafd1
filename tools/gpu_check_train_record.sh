export PYTHONPATH=$PWD; mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "bench tiny rc=$?"; tail -2 gpurun_out/bench_tiny.err
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_tiny.json") if l.startswith("{")][-1])
print(json.dumps(j["configs"]["tiny_train_b128_per_gpu"], indent=1)[:2500])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_base.csv python bench.py --variant base --steps 2 --warmup 1 --no-cpu-baseline --no-sub --long-steps 0 > gpurun_out/ncu_bench_base.log 2>&1; echo "base list rc=$?"
