#!/bin/bash
# Round-1 measurement sweep of the barrier-free schedule and the device vertex numbering (run under gpurun).
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
{
XF_TEST_SCHEDULES=4,2 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_trajectory.py -x -q -m gpu -k "full_size" 2>&1 | tail -5
echo "== persistent, no renumber"; XF_NO_RENUMBER=1 $B --schedule persistent
echo "== persistent, renumber"; $B --schedule persistent
echo "== dataflow, no renumber"; XF_NO_RENUMBER=1 $B --schedule dataflow
echo "== dataflow, renumber"; $B --schedule dataflow
for es in 100 300; do
  echo "== dataflow, renumber, esleep $es"; XF_DATAFLOW_ESLEEP_NS=$es XF_DATAFLOW_SLEEP_NS=$es $B --schedule dataflow
done
for cells in 16 28 40 70 110; do
  echo "== dataflow cells $cells"; $B --schedule dataflow --cells $cells --substeps-per-step 20
  echo "== persistent cells $cells"; $B --schedule persistent --cells $cells --substeps-per-step 20
done
} > gpurun_out/df_sweep.log 2>&1
grep -o '^== .*\|"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|passed\|failed\|error' gpurun_out/df_sweep.log | grep -v '"value"' | head -80
