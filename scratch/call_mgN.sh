# multi-GPU weak-scaling line: bash scratch/call_mgN.sh N   (run under gpurun --gpus N)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "g$N rc=$?"
python scratch/show_bench.py gpurun_out/bench_g$N.json || tail -5 gpurun_out/bench_g$N.err
