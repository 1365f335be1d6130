mkdir -p gpurun_out
bash scratch/call_mgN.sh 4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 scratch/generate_mg.py 2048 > gpurun_out/generate_g4.json 2> gpurun_out/generate_g4.err; echo "gen rc=$?"
cat gpurun_out/generate_g4.json; tail -5 gpurun_out/generate_g4.err
