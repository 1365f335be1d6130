#!/bin/bash
# one iteration on the training step: GEMM probe, training parity tests, the C5 step time, bench line, ncu launch list
mkdir -p gpurun_out
timeout 120 python scratch/gemm_shapes.py > gpurun_out/t_gemm_shapes.log 2>&1; echo "gemm_shapes rc=$?"; cat gpurun_out/t_gemm_shapes.log
timeout 600 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/t_pytest_train.log 2>&1; echo "pytest train rc=$?"
tail -5 gpurun_out/t_pytest_train.log
timeout 200 python scratch/train_step.py 8 > gpurun_out/t_train_step.log 2>&1; echo "train_step rc=$?"; tail -2 gpurun_out/t_train_step.log
timeout 300 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/t_train_g1.json 2> gpurun_out/t_train_g1.err; echo "bench rc=$?"
tail -n1 gpurun_out/t_train_g1.json | cut -c1-1200
if [ "$1" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/t_train_launches.csv python scratch/train_step.py 3 > gpurun_out/t_ncu_train.log 2>&1; echo "ncu rc=$?"
fi
