set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --trajectory --no-cpu-baseline --no-other-precision > gpurun_out/bench_traj.json 2> gpurun_out/bench_traj.err; echo "traj rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-precision > gpurun_out/b2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-precision > gpurun_out/ncu_launch.log 2>&1
echo "launch rc=$?"
timeout 120 python scratch/one_step.py 3 > gpurun_out/one_step.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"edge_kernels_tc|message_fiber_norm_fused|convnext_mlp_tc|graph_fill|graph_count|readout_pooled|node_embed" --launch-skip 30 -c 15 -f -o gpurun_out/full_step python scratch/one_step.py 3 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
