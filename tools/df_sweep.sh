#!/bin/bash
# Round-1 measurement sweep of the barrier-free schedule (run under gpurun).
mkdir -p gpurun_out
{
XF_TEST_SCHEDULES=4 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_trajectory.py -x -q -m gpu -k "full_size" 2>&1 | tail -5
for sched in persistent dataflow; do
  timeout 300 python bench.py --schedule $sched --steps 5 --warmup 3 --no-cpu-baseline
done
for ns in 50 200 1000; do
  XF_DATAFLOW_SLEEP_NS=$ns timeout 300 python bench.py --schedule dataflow --steps 5 --warmup 3 --no-cpu-baseline
done
for cells in 16 28 40 70 110; do
  timeout 300 python bench.py --schedule dataflow --cells $cells --substeps-per-step 20 --steps 5 --warmup 3 --no-cpu-baseline
done
} > gpurun_out/df_sweep.log 2>&1
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|passed\|failed\|error' gpurun_out/df_sweep.log | head -80
