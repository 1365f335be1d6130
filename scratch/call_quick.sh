mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log
(cd scratch/ab_base && cp ../ab_message.py scratch_ab_message.py 2>/dev/null; mkdir -p scratch; cp ../ab_message.py scratch/ab_message.py; timeout 300 python scratch/ab_message.py 2>&1 | tail -2)
timeout 300 python scratch/ab_message.py 2>&1 | tail -2
