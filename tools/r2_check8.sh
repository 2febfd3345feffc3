#!/bin/bash
# Debug: graph partition on 4 GPUs at growing sizes (flag protocol), full error output
N=${1:-4}
mkdir -p gpurun_out
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu"
{
  for d in 30 55 88; do
    echo "== $d^3 graph auto"; $T --dims $d $d --substeps 4 --check 0 --schedule auto --partition graph --time-substeps 10 2>&1 | grep -E "PART_RESULT|rror|peers|xf " | head -6
  done
  echo "== 88^3 graph persistent"; $T --dims 88 88 --substeps 4 --check 0 --schedule persistent --partition graph --time-substeps 10 2>&1 | grep -E "PART_RESULT|rror" | head -4
  echo "== 110^3 slabs per_color"; $T --dims 110 110 --substeps 4 --check 0 --schedule per_color --time-substeps 10 2>&1 | grep -E "PART_RESULT|rror" | head -4
} > gpurun_out/part_graph_debug_n$N.log 2>&1
cut -c1-420 gpurun_out/part_graph_debug_n$N.log
