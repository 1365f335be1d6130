#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_fullsize.py tests/test_gpu_longrows.py tests/test_gpu_step.py -m gpu -q -x > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?"
timeout 300 python bench.py --crystals 256 --atoms 200 --radius 7 --cap 8 --state teacher --t0 300 --steps 20 --no-cpu-baseline --no-other-precision > gpurun_out/c12_bench_c3_teacher.json 2> gpurun_out/c12_c3t.err
timeout 300 python bench.py --crystals 256 --atoms 200 --radius 7 --cap 8 --steps 20 --no-cpu-baseline --no-other-precision > gpurun_out/c12_bench_c3_sampler.json 2> gpurun_out/c12_c3s.err
tail -4 gpurun_out/c12_pytest.log; python scratch/show_bench.py gpurun_out/c12_bench_c3_teacher.json gpurun_out/c12_bench_c3_sampler.json
