#!/bin/bash
# ncu sections of the TMA-fed GEMM launches of the second C5 training step (after the plain run exits 0); the report is
# turned into CSV on the box (the .ncu-rep is too large to bring back)
mkdir -p gpurun_out
timeout 200 python scratch/train_step.py 2 > gpurun_out/t_train_step.log 2>&1 || { echo "plain run failed"; tail gpurun_out/t_train_step.log; exit 1; }
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --clock-control none -k regex:"sgemm_tma" --launch-skip 58 -c 58 -f -o /tmp/r2_train_gemm python scratch/train_step.py 2 > gpurun_out/ncu_train_full.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2_train_gemm.ncu-rep --page raw --csv > gpurun_out/r2_train_gemm_raw.csv 2> gpurun_out/ncu_train_csv.err; echo "csv rc=$?"
ls -la /tmp/*.ncu-rep gpurun_out/r2_train_gemm_raw.csv; tail -3 gpurun_out/ncu_train_full.log
