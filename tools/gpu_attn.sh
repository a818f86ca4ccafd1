export PYTHONPATH=$PWD; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py -m gpu -q --timeout 200 > gpurun_out/pytest_umma.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_umma.log
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "base" > gpurun_out/pytest_base.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_base.log
timeout 300 python bench.py --variant base --no-cpu-baseline --no-sub > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; echo "bench base rc=$?"
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_base.json") if l.startswith("{")][-1])
print(round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), {k: round(x, 4) for k, x in j["kernel_ms_per_step"].items()})
PY
