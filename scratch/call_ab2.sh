# A = committed HEAD (scratch/ab_base), B = working tree; alternate on the same box
mkdir -p gpurun_out
for rep in 1 2; do
  for v in A B; do
    if [ $v = A ]; then d=scratch/ab_base; else d=.; fi
    (cd $d && timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-other-precision 2>/dev/null) > gpurun_out/ab_$v$rep.json
    echo "== $v$rep"; python scratch/show_bench.py gpurun_out/ab_$v$rep.json
  done
done
