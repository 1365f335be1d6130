mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scratch/generate_mg.py 2048 > gpurun_out/generate_g8.json 2> gpurun_out/generate_g8.err; echo "gen rc=$?"
cat gpurun_out/generate_g8.json; tail -3 gpurun_out/generate_g8.err
