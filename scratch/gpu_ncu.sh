#!/bin/bash
# ncu evidence of one C2 step (fp16 tensor path): launch list + --set full of the large kernels (after the plain run exits 0)
mkdir -p gpurun_out
timeout 300 python scratch/one_step.py 3 > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/ncu_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python scratch/one_step.py 3 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"edge_kernels_tc|message_fiber_norm_fused|convnext_mlp_tc|readout_pooled|node_embed|graph_fill|graph_count" --launch-skip 30 -c 15 -f -o gpurun_out/r2_step python scratch/one_step.py 3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_full.log
