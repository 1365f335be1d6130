#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scratch/fit_throughput.py 8100 270 3 > gpurun_out/t_fit.log 2>&1; echo "fit rc=$?"; tail -12 gpurun_out/t_fit.log
