#!/bin/bash
# One GPU call: parity of the chained sweep (XF_GROUPING_CHAINS), chained vs plain timings, launch list + one full ncu
# capture of k_substeps_chain.  Every step has its own timeout; results land in gpurun_out/.
mkdir -p gpurun_out
B="timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
{
  echo "== tests: chained"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "chained" 2>&1 | tail -5
  echo "== bench 55 chains"; $B --grouping chains
  echo "== bench 55 elements"; $B --grouping elements
  echo "== tests: dataflow schedule with XF_CHAIN=1"; XF_CHAIN=1 XF_TEST_SCHEDULES=4 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
  for cells in 16 28 40; do
    echo "== bench $cells chains"; $B --grouping chains --cells $cells --substeps-per-step 20
    echo "== bench $cells elements"; $B --grouping elements --cells $cells --substeps-per-step 20
  done
  echo "== bench 55 chains no-prefetch"; XF_DATAFLOW_NO_PREFETCH=1 $B --grouping chains
  echo "== bench 55 chains mixedsel"; $B --grouping chains --energy mixedsel
  echo "== bench 55 elements mixedsel"; $B --grouping elements --energy mixedsel
} > gpurun_out/chain_check.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_chain.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --grouping chains > gpurun_out/ncu_chain_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_substeps_chain -s 3 -c 1 -o gpurun_out/prof_chain_r1 -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --grouping chains > gpurun_out/ncu_chain_full.log 2>&1
timeout 300 python bench.py --grouping chains > gpurun_out/bench_chain_full.json 2> gpurun_out/bench_chain_full.err
grep -o '^== .*\|"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|rror.*\|[0-9]* passed.*\|[0-9]* failed.*' gpurun_out/chain_check.log | cut -c1-160
