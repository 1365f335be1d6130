#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py -m gpu -q -x > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c18_pytest.log
for G in 16 128; do
  timeout 200 python bench.py --crystals $G --atoms 40 --steps 200 --warmup 5 --no-cpu-baseline --no-other-precision --no-e2e-trajectory > gpurun_out/c18_g${G}_plain.json 2> gpurun_out/c18_err.log
  timeout 200 python bench.py --crystals $G --atoms 40 --steps 200 --warmup 5 --no-cpu-baseline --no-other-precision --no-e2e-trajectory --cuda-graph > gpurun_out/c18_g${G}_graph.json 2>> gpurun_out/c18_err.log
done
timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-other-precision --no-e2e-trajectory > gpurun_out/c18_c2_plain.json 2>> gpurun_out/c18_err.log
timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-other-precision --no-e2e-trajectory --cuda-graph > gpurun_out/c18_c2_graph.json 2>> gpurun_out/c18_err.log
python scratch/show_bench.py gpurun_out/c18_g16_plain.json gpurun_out/c18_g16_graph.json gpurun_out/c18_g128_plain.json gpurun_out/c18_g128_graph.json gpurun_out/c18_c2_plain.json gpurun_out/c18_c2_graph.json | grep value; tail -3 gpurun_out/c18_err.log
