mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log
timeout 600 python bench.py --cap 0 --state teacher --t0 500 --steps 5 --no-cpu-baseline --no-other-precision > gpurun_out/bench_c2u.json 2> gpurun_out/bench_c2u.err; echo "c2u rc=$?"
python scratch/show_bench.py gpurun_out/bench_c2u.json || tail -3 gpurun_out/bench_c2u.err
timeout 900 python bench.py --crystals 256 --atoms 200 --radius 7 --cap 0 --state teacher --t0 100 --steps 3 --no-cpu-baseline --no-other-precision > gpurun_out/bench_c3u.json 2> gpurun_out/bench_c3u.err; echo "c3u rc=$?"
python scratch/show_bench.py gpurun_out/bench_c3u.json || tail -3 gpurun_out/bench_c3u.err
