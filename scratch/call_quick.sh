mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -6 gpurun_out/pytest_quick.log
for v in A B A B; do
  if [ $v = A ]; then d=scratch/ab_base; else d=.; fi
  (cd $d && timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-precision 2>/dev/null) > gpurun_out/ab_$v.json
  echo "== $v"; python scratch/show_bench.py gpurun_out/ab_$v.json
done
