set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -5 gpurun_out/pytest_quick.log
timeout 600 python bench.py --no-cpu-baseline --no-other-precision > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python scratch/show_bench.py gpurun_out/bench_quick.json || cat gpurun_out/bench_quick.json gpurun_out/bench_quick.err | tail -20
