mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"readout_pooled|node_embed" --launch-skip 2 -c 2 -f -o gpurun_out/readout python scratch/one_step.py 3 > gpurun_out/ncu_readout.log 2>&1
echo "rc=$?"
tail -3 gpurun_out/ncu_readout.log
