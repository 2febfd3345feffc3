#!/bin/bash
# One GPU call: contact-scene test at 1M tets, parity of the barrier-free kernels, timings with spare warps skipping the colour loop.
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
{
  echo "== tests: contact scene + full size"; timeout 400 python -m pytest tests/test_gpu_trajectory.py -x -q -k "full_size" 2>&1 | tail -4
  echo "== tests: parity, dataflow schedule"; XF_TEST_SCHEDULES=4 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -3
  echo "== 55 plain"; $B
  echo "== 55 plain again"; $B
  echo "== 55 chain"; $B --grouping chains
  echo "== 40 plain"; $B --cells 40 --substeps-per-step 20
  echo "== 28 plain"; $B --cells 28 --substeps-per-step 20
  echo "== 16 plain"; $B --cells 16 --substeps-per-step 20
  echo "== 55 plain block160"; XF_DATAFLOW_BLOCK=160 $B
} > gpurun_out/e1_check.log 2>&1
grep -o '^== .*\|"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"grid": [^]]*\]\|rror.*\|[0-9]* passed.*\|[0-9]* failed.*' gpurun_out/e1_check.log | cut -c1-160
