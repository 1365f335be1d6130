mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -5 gpurun_out/pytest_quick.log
for v in A B A B; do
  if [ $v = A ]; then d=scratch/ab_base; else d=.; fi
  (cd $d && timeout 300 python bench.py --crystals 256 --atoms 200 --radius 7 --steps 10 --state teacher --t0 300 --no-cpu-baseline --no-other-precision 2>/dev/null) > gpurun_out/ab_c3t_$v.json
  echo "== $v C3 teacher"; python scratch/show_bench.py gpurun_out/ab_c3t_$v.json
done
(cd . && timeout 300 python bench.py --crystals 256 --atoms 200 --radius 7 --steps 10 --no-cpu-baseline --no-other-precision 2>/dev/null) > gpurun_out/ab_c3_B.json
echo "== B C3 sampler"; python scratch/show_bench.py gpurun_out/ab_c3_B.json
(cd . && timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-other-precision 2>/dev/null) > gpurun_out/ab_B.json
echo "== B C2"; python scratch/show_bench.py gpurun_out/ab_B.json
