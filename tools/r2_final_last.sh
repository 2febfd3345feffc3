#!/bin/bash
# Last GPU call of round 2: full GPU suite on the final tree, then the ncu captures behind profiles/r2_traffic.json
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/gpu_tests_last.log
bash tools/ncu_traffic.sh damped > gpurun_out/ncu_traffic_last.log 2>&1
cat gpurun_out/gpu_tests_last.log; tail -2 gpurun_out/ncu_traffic_last.log
