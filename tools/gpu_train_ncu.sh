#!/bin/bash
# ncu evidence for the training step: launch list of two eager steps, --set full of the tensor-core GEMM launches of the
# second step; and the launch list of the base variant's forward.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/run_train_step.py 128 128 2 > gpurun_out/ncu_train_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:t_gemm_tc_kernel --launch-skip 114 --launch-count 28 -o gpurun_out/prof_train_gemm_r02 -f python tools/run_train_step.py 128 128 2 > gpurun_out/ncu_train_gemm.log 2>&1; echo "full rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_base.csv python bench.py --variant base --steps 2 --warmup 1 --no-cpu-baseline --no-sub --long-steps 0 > gpurun_out/ncu_bench_base.log 2>&1; echo "base list rc=$?"
