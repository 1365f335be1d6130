#!/bin/bash
# BASELINE configs[3] in full: 65 536 crystals x 40 atoms over 8 B200 through the public generate driver
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scratch/generate_mg.py 8192 > gpurun_out/c4_full_g8.json 2> gpurun_out/c4_full_g8.err; echo "rc=$?"
tail -2 gpurun_out/c4_full_g8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/c4_bench_g8.json 2> gpurun_out/c4_bench_g8.err; echo "rc=$?"
tail -c 400 gpurun_out/c4_bench_g8.json
