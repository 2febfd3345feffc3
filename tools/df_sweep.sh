#!/bin/bash
# Single-GPU measurement sweep behind profiles/r1_results.md (run under gpurun): schedules, mesh sizes, precision, grouping.
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
{
  for sched in dataflow persistent bricks per_color; do echo "== $sched"; $B --schedule $sched; done
  echo "== dataflow fast"; $B --schedule dataflow --precision fast
  echo "== dataflow clusters"; XF_CLUSTERS=1 $B --schedule dataflow
  echo "== dataflow no L1 prefetch"; XF_DATAFLOW_NO_PREFETCH=1 $B --schedule dataflow
  for cells in 16 28 40 70 110; do
    echo "== dataflow cells $cells"; $B --schedule dataflow --cells $cells --substeps-per-step 20
    echo "== persistent cells $cells"; $B --schedule persistent --cells $cells --substeps-per-step 20
  done
  echo "== batch boxL"; timeout 200 python tools/batch_bench.py --shape boxL
  echo "== batch beamL"; timeout 200 python tools/batch_bench.py --shape beamL
} > gpurun_out/df_sweep.log 2>&1
grep -o '^== .*\|"element_substeps_per_s": [0-9.e+]*\|"ms_per_step": [0-9.]*\|rror: .*' gpurun_out/df_sweep.log | cut -c1-150
