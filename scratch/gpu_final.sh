#!/bin/bash
# final validation of the tree: full GPU suite, smoke, both bench arms, training bench + its ncu launch list
# (the ncu captures of the sampling step are scratch/gpu_ncu.sh, of the training GEMMs scratch/gpu_ncu_train.sh)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/f_smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/f_train.json 2> gpurun_out/f_train.err; echo "train rc=$?"
timeout 200 python scratch/train_step.py 3 > gpurun_out/t_train_step.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/t_train_launches.csv python scratch/train_step.py 3 > gpurun_out/t_ncu_train.log 2>&1; echo "ncu rc=$?"
timeout 120 python scratch/gemm_shapes.py > gpurun_out/t_gemm_shapes.log 2>&1; echo "gemm_shapes rc=$?"
timeout 120 python scratch/gemm_prof.py > gpurun_out/t_gemm_prof.log 2>&1; echo "gemm_prof rc=$?"
tail -3 gpurun_out/f_pytest.log; tail -3 gpurun_out/f_smoke.log; python scratch/show_bench.py gpurun_out/f_bench.json; tail -n1 gpurun_out/f_train.json | cut -c1-300; tail -2 gpurun_out/t_train_step.log
