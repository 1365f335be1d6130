#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/ab_mlp.py > gpurun_out/c13_ab_mlp.log 2>&1; echo "rc=$?" >> gpurun_out/c13_ab_mlp.log
timeout 300 python scratch/ab_mlp.py 3 7 >> gpurun_out/c13_ab_mlp.log 2>&1; echo "rc=$?" >> gpurun_out/c13_ab_mlp.log
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_longrows.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c13_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-precision > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
cat gpurun_out/c13_ab_mlp.log; tail -4 gpurun_out/c13_pytest.log; python scratch/show_bench.py gpurun_out/c13_bench.json
