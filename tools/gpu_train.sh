#!/bin/bash
# Training-path iteration: the backward tests (all failures listed), then the step timing.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_train_backward.py tests/test_gpu_training.py -m gpu -q --timeout 300 > gpurun_out/pytest_train.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_train.log
ES_TRAIN_TC=0 timeout 300 python tools/bench_train.py ${1:-128} ${2:-128} > gpurun_out/bench_train.log 2>&1
echo "bench rc=$?"; grep -v "^-----\|autograd::engine\|Backward  \|^ *_[A-Z]" gpurun_out/bench_train.log | cut -c1-60,100-112,150-200 | head -60
