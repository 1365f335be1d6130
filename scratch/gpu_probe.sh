#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scratch/gemm_shapes.py > gpurun_out/t_gemm_shapes.log 2>&1; echo "gemm_shapes rc=$?"; cat gpurun_out/t_gemm_shapes.log
timeout 120 python scratch/gemm_prof.py > gpurun_out/t_gemm_prof.log 2>&1; echo "gemm_prof rc=$?"; cat gpurun_out/t_gemm_prof.log
timeout 200 python scratch/train_step.py 8 > gpurun_out/t_train_step.log 2>&1; echo "train_step rc=$?"; tail -2 gpurun_out/t_train_step.log
