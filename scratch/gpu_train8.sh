#!/bin/bash
# C5 training step on 1, 2 and 8 GPUs of the same box (DDP: one all-reduce of the flat gradient per step)
mkdir -p gpurun_out
timeout 300 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/train_g1.json 2> gpurun_out/train_g1.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --workload train --gpus 8 --steps 20 --warmup 3 > gpurun_out/train_g8.json 2> gpurun_out/train_g8.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --workload train --gpus 2 --steps 20 --warmup 3 > gpurun_out/train_g2.json 2> gpurun_out/train_g2.err; echo "rc=$?"
tail -n1 gpurun_out/train_g1.json | cut -c1-1200; tail -n1 gpurun_out/train_g8.json | cut -c1-1200; tail -n1 gpurun_out/train_g2.json | cut -c1-1200
