#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/train_step.py 6 1 > gpurun_out/c11_train.log 2>&1
timeout 300 python scratch/train_step.py 6 0 >> gpurun_out/c11_train.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_train_launches.csv python scratch/train_step.py 3 0 > gpurun_out/c11_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?"
cat gpurun_out/c11_train.log; tail -5 gpurun_out/c11_pytest.log
