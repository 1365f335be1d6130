#!/bin/bash
# round 2, call 1: full GPU test suite, smoke, bench (both arms), compute-sanitizer passes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c1_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?" >> gpurun_out/c1_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 30 python scratch/sanitize_step.py > gpurun_out/c1_san_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/c1_san_$tool.log
done
tail -3 gpurun_out/c1_pytest.log; tail -2 gpurun_out/c1_smoke.log; tail -c 600 gpurun_out/c1_bench.json; for t in memcheck racecheck synccheck; do tail -3 gpurun_out/c1_san_$t.log; done
