#!/bin/bash
# Round-1 measurement sweep of the barrier-free schedule (run under gpurun).
mkdir -p gpurun_out
B="timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
{
  echo "== tests"; XF_TEST_SCHEDULES=4 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
  echo "== dataflow prefetch"; $B --schedule dataflow
  echo "== dataflow no prefetch"; XF_DATAFLOW_NO_PREFETCH=1 $B --schedule dataflow
  echo "== dataflow fast"; $B --schedule dataflow --precision fast
  echo "== dataflow esleep 100"; XF_DATAFLOW_ESLEEP_NS=100 $B --schedule dataflow
  echo "== dataflow cells 110"; $B --schedule dataflow --cells 110 --substeps-per-step 20
  echo "== dataflow cells 40"; $B --schedule dataflow --cells 40 --substeps-per-step 20
  echo "== dataflow cells 28"; $B --schedule dataflow --cells 28 --substeps-per-step 20
  echo "== dataflow mixedsel"; $B --schedule dataflow --energy mixedsel
} > gpurun_out/df_sweep.log 2>&1
grep -o '^== .*\|"ms_per_step": [0-9.]*\|[0-9]* passed\|[0-9]* failed\|rror: .*' gpurun_out/df_sweep.log | cut -c1-150 | head -80
