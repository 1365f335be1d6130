#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/ab_edge.py > gpurun_out/c2_ab_edge.log 2>&1; echo "rc=$?" >> gpurun_out/c2_ab_edge.log
timeout 300 python scratch/ab_edge.py 256 200 7.0 8 >> gpurun_out/c2_ab_edge.log 2>&1; echo "rc=$?" >> gpurun_out/c2_ab_edge.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc=$?" >> gpurun_out/c2_bench.err
cat gpurun_out/c2_ab_edge.log; tail -5 gpurun_out/c2_pytest.log; python scratch/show_bench.py gpurun_out/c2_bench.json 2>/dev/null | head -20
