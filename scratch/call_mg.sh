mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "g2 rc=$?"
python scratch/show_bench.py gpurun_out/bench_g2.json || tail -5 gpurun_out/bench_g2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_g2.json 2> gpurun_out/bench_ref_g2.err; echo "ref g2 rc=$?"
cut -c1-200 gpurun_out/bench_ref_g2.json
