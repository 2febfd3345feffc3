#!/bin/bash
# Stall statistics of the barrier-free cross-GPU schedule at 20M tets: gap between peer store and local store 0 vs 400 ns
N=${1:-4}
mkdir -p gpurun_out
T="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu --dims 150 150 --substeps 4 --check 0 --schedule dataflow --time-substeps 50 --time-calls 8"
{
  for rep in 1 2 3; do
    for gap in 0 400; do
      echo "== rep $rep gap $gap"; XF_PART_STORE_GAP_NS=$gap $T 2>&1 | grep -E "PART_RESULT|XfError:" | head -1 | cut -c1-330
    done
  done
} > gpurun_out/part_gap_n$N.log 2>&1
cat gpurun_out/part_gap_n$N.log
