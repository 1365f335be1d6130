#!/bin/bash
# final multi-GPU record: training step under DDP (1, 8 ranks) and the sampling bench at 8 ranks, both arms
mkdir -p gpurun_out
timeout 300 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/train_g1.json 2> gpurun_out/train_g1.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --workload train --gpus 8 --steps 20 --warmup 3 > gpurun_out/train_g8.json 2> gpurun_out/train_g8.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/ref_g8.json 2> gpurun_out/ref_g8.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "rc=$?"
for f in train_g1 train_g8 ref_g8 bench_g8 bench_g2; do tail -n1 gpurun_out/$f.json | cut -c1-300; done
