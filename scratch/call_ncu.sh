set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"message_fiber_norm_cached" --launch-skip 10 -c 1 -f -o gpurun_out/cached python scratch/one_step.py 3 > gpurun_out/ncu_cached.log 2>&1
echo "rc=$?"
tail -3 gpurun_out/ncu_cached.log
