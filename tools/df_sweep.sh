#!/bin/bash
mkdir -p gpurun_out
B="timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
{
  echo "== dataflow"; $B --schedule dataflow
  echo "== dataflow sentinel"; XF_DATAFLOW_SENTINEL=1 $B --schedule dataflow
  echo "== dataflow sentinel esleep 100"; XF_DATAFLOW_SENTINEL=1 XF_DATAFLOW_ESLEEP_NS=100 $B --schedule dataflow
  echo "== sentinel cells 16"; XF_DATAFLOW_SENTINEL=1 $B --schedule dataflow --cells 16 --substeps-per-step 20
  echo "== sentinel cells 110"; XF_DATAFLOW_SENTINEL=1 $B --schedule dataflow --cells 110 --substeps-per-step 20
  echo "== tests sentinel"; XF_DATAFLOW_SENTINEL=1 XF_TEST_SCHEDULES=4 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bit_identical" 2>&1 | tail -2
} > gpurun_out/df_sweep.log 2>&1
grep -o '^== .*\|"ms_per_step": [0-9.]*\|[0-9]* passed\|[0-9]* failed\|rror: .*' gpurun_out/df_sweep.log | cut -c1-150 | head -80
