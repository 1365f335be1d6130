mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-precision 2>/dev/null > gpurun_out/ab_B.json
python scratch/show_bench.py gpurun_out/ab_B.json
