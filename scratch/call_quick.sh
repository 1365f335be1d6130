mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_state.py tests/test_gpu_step.py tests/test_gpu_wrapper.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-precision 2>/dev/null > gpurun_out/ab_B.json
python scratch/show_bench.py gpurun_out/ab_B.json
