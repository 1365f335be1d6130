mkdir -p gpurun_out
for x in 0 1 2 3 4 7; do echo "EXP=$x"; ARREAU_EXP=$x timeout 300 python scratch/ab_message.py 2>&1 | tail -2; done
