#!/bin/bash
# final validation of the tree: full GPU suite, smoke, both bench arms, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/f_smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
timeout 300 python scratch/one_step.py 3 > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python scratch/one_step.py 3 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"edge_kernels_tc|message_fiber_norm_fused|convnext_mlp_tc|readout_pooled|node_embed|graph_fill|graph_count" --launch-skip 30 -c 15 -f -o gpurun_out/r2_step python scratch/one_step.py 3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
tail -3 gpurun_out/f_pytest.log; tail -3 gpurun_out/f_smoke.log; python scratch/show_bench.py gpurun_out/f_bench.json
