#!/bin/bash
# Round-2 call 1: full GPU parity suite (incl. the BASELINE-shape tests), the bench line with sub-records, the
# reference arm, compute-sanitizer over a small slice of the suite, and timing experiments of the decoder layer kernel.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -30 >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "bench tiny rc=$?"; tail -3 gpurun_out/bench_tiny.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
for e in NOTANH NOSTORE NOEPI NODW NOSPLIT; do
  ES_B200_LIB=$PWD/efficientspeech_b200/exp/libes_$e.so timeout 120 python bench.py --no-sub --no-cpu-baseline --long-steps 0 > gpurun_out/bench_exp_$e.json 2> gpurun_out/bench_exp_$e.err; echo "exp $e rc=$?"
done
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x --timeout 400 -k "reference_fixture_parity and (tiny_b3n24 or small_b2n17 or base_b2n17)" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x --timeout 400 -k "reference_fixture_parity and tiny_b3n24" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/sanitizer_racecheck.log
python - <<'PY'
import json
def last(path):
    try:
        return json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as e:
        return None
j = last("gpurun_out/bench_tiny.json")
if j:
    print("tiny", round(j["value"] / 1e6, 1), "M frames/s", round(j["ms_per_step"], 4), "ms; e2e", round(j["e2e"]["value"] / 1e6, 1),
          "; roofline", round(j["roofline"]["frac"], 3), {k: round(x, 4) for k, x in j["kernel_ms_per_step"].items()})
    for k, v in (j.get("configs") or {}).items():
        print(" ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "ms_per_utt_mean", "graph_replay_ms_per_utt", "kernel_ms_per_step")},
              "roofline", (v.get("roofline") or {}).get("frac"), "e2e", (v.get("e2e") or {}).get("value"))
    print(" cpu", j.get("cpu_baseline"), j.get("host_placement"))
for e in ("NOTANH", "NOSTORE", "NOEPI", "NODW", "NOSPLIT"):
    x = last(f"gpurun_out/bench_exp_{e}.json")
    if x:
        print(e, round(x["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in x["kernel_ms_per_step"].items()}, [round(p["ms"], 4) for p in x["roofline"]["per_launch_position"]])
r = last("gpurun_out/bench_reference.json")
print("reference", r and (r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["kind"]))
PY
