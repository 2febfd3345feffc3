#!/bin/bash
mkdir -p gpurun_out
B="timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
{
  echo "== tests"; XF_TEST_SCHEDULES=4 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
  echo "== pair"; $B --schedule dataflow
  echo "== no pair"; XF_NO_PAIRS=1 $B --schedule dataflow
  echo "== pair cells 16"; $B --schedule dataflow --cells 16 --substeps-per-step 20
  echo "== pair cells 28"; $B --schedule dataflow --cells 28 --substeps-per-step 20
  echo "== pair cells 70"; $B --schedule dataflow --cells 70 --substeps-per-step 20
  echo "== pair cells 110"; $B --schedule dataflow --cells 110 --substeps-per-step 20
  echo "== pair mixedsel"; $B --schedule dataflow --energy mixedsel
} > gpurun_out/df_sweep.log 2>&1
grep -o '^== .*\|"ms_per_step": [0-9.]*\|[0-9]* passed\|[0-9]* failed\|rror: .*' gpurun_out/df_sweep.log | cut -c1-150 | head -80
