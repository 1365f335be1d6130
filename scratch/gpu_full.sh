#!/bin/bash
# the whole GPU suite, then the training iteration
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f_pytest.log
bash scratch/gpu_train_iter.sh $1
