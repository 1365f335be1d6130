#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/ab_message.py > gpurun_out/c7_ab_msg.log 2>&1; echo "rc=$?" >> gpurun_out/c7_ab_msg.log
timeout 300 python scratch/ab_message.py 256 200 7.0 8 >> gpurun_out/c7_ab_msg.log 2>&1; echo "rc=$?" >> gpurun_out/c7_ab_msg.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
cat gpurun_out/c7_ab_msg.log; tail -15 gpurun_out/c7_pytest.log
