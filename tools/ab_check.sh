#!/bin/bash
# One GPU call: A/B of (i) warps per CTA of the barrier-free kernels (only as many as a colour needs vs 256 threads),
# (ii) numbering of the lattice colour classes (ring vs type-major), (iii) chained vs plain; goldens and parity first.
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
{
  echo "== tests: goldens on the GPU"; timeout 300 python -m pytest tests/test_golden.py -x -q -m gpu 2>&1 | tail -3
  echo "== tests: parity, dataflow schedule"; XF_TEST_SCHEDULES=4 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
  echo "== 55 plain ring auto-block"; $B --grouping elements
  echo "== 55 plain ring block256"; XF_DATAFLOW_BLOCK=256 $B --grouping elements
  echo "== 55 plain type auto-block"; $B --grouping elements --hint-order type
  echo "== 55 plain type block256"; XF_DATAFLOW_BLOCK=256 $B --grouping elements --hint-order type
  echo "== 55 chain ring auto-block"; $B --grouping chains
  echo "== 55 plain ring block192"; XF_DATAFLOW_BLOCK=192 $B --grouping elements
  echo "== 55 plain ring auto-block sleep0"; XF_DATAFLOW_SLEEP_NS=0 $B --grouping elements
  echo "== 55 plain ring auto-block sleep2000"; XF_DATAFLOW_SLEEP_NS=2000 $B --grouping elements
  for cells in 16 40 70; do
    echo "== $cells plain ring auto-block"; $B --grouping elements --cells $cells --substeps-per-step 20
    echo "== $cells plain ring block256"; XF_DATAFLOW_BLOCK=256 $B --grouping elements --cells $cells --substeps-per-step 20
  done
} > gpurun_out/ab_check.log 2>&1
grep -o '^== .*\|"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"grid": [^]]*\]\|rror.*\|[0-9]* passed.*\|[0-9]* failed.*' gpurun_out/ab_check.log | cut -c1-160
