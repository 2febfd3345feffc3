#!/bin/bash
# Round 2: packed (f32x2) element arithmetic: full GPU suite, then bench (plain and chained), small-mesh latency regime.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/gpu_tests_packed.log
B="timeout 300 python bench.py --no-cpu-baseline"
{
  echo "== 55 plain"; $B --extras damped,native_rate,batch
  echo "== 55 chains"; $B --no-extras --grouping chains
  echo "== 16 plain"; $B --no-extras --cells 16 --substeps-per-step 200
  echo "== 40 plain"; $B --no-extras --cells 40 --substeps-per-step 100
  echo "== 110 plain"; $B --no-extras --cells 110 --substeps-per-step 20 --steps 5
} > gpurun_out/bench_packed.log 2>&1
cat gpurun_out/gpu_tests_packed.log
grep -o '^== .*\|"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"frames_per_s": [0-9.]*\|rror.*' gpurun_out/bench_packed.log | cut -c1-160
