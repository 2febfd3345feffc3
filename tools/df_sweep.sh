#!/bin/bash
# Round-1 measurement sweep of the barrier-free schedule (run under gpurun).
mkdir -p gpurun_out
B="timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
{
for mode in 0; do
  echo "== mode $mode tests"
  XF_DATAFLOW_MODE=$mode XF_TEST_SCHEDULES=4 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu  2>&1 | tail -15
  echo "== mode $mode dataflow"; XF_DATAFLOW_MODE=$mode $B --schedule dataflow
  echo "== mode $mode dataflow esleep 200"; XF_DATAFLOW_MODE=$mode XF_DATAFLOW_ESLEEP_NS=200 XF_DATAFLOW_SLEEP_NS=200 $B --schedule dataflow
  echo "== mode $mode dataflow cells 110"; XF_DATAFLOW_MODE=$mode $B --schedule dataflow --cells 110 --substeps-per-step 20
  echo "== mode $mode dataflow cells 28"; $B --schedule dataflow --cells 28 --substeps-per-step 20
  echo "== dataflow cells 16"; $B --schedule dataflow --cells 16 --substeps-per-step 20
  echo "== dataflow cells 70"; $B --schedule dataflow --cells 70 --substeps-per-step 20
done
} > gpurun_out/df_sweep.log 2>&1
grep -o '^== .*\|"ms_per_step": [0-9.]*\|[0-9]* passed\|[0-9]* failed\|rror: .*' gpurun_out/df_sweep.log | cut -c1-150 | head -80
