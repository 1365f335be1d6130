mkdir -p gpurun_out
timeout 600 python bench.py --trajectory --no-cpu-baseline --no-other-precision > gpurun_out/bench_traj.json 2> gpurun_out/bench_traj.err; echo "traj rc=$?"
python scratch/show_bench.py gpurun_out/bench_traj.json
timeout 600 python bench.py --crystals 256 --atoms 200 --radius 7 --no-cpu-baseline --no-other-precision > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
python scratch/show_bench.py gpurun_out/bench_c3.json
timeout 600 python bench.py --workload train > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc=$?"
tail -c 1500 gpurun_out/bench_train.json
