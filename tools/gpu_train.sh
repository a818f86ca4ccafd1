#!/bin/bash
# Training-path iteration: the backward tests (all failures listed), then the step timing.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_train_backward.py tests/test_gpu_training.py -m gpu -q --timeout 300 > gpurun_out/pytest_train.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/pytest_train.log
timeout 300 python tools/bench_train.py ${1:-128} ${2:-128} > gpurun_out/bench_train.log 2>&1
echo "bench rc=$?"; head -45 gpurun_out/bench_train.log
