#!/bin/bash
# Debug: graph partition at scale (2 GPUs), flag protocol vs barrier-free
N=${1:-2}
mkdir -p gpurun_out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu"
{
  echo "== 1M slabs per_color"; $T --dims 55 55 --substeps 4 --check 0 --schedule per_color --time-substeps 20 2>&1 | grep -E "PART_RESULT|rror" | head -4
  echo "== 1M graph per_color"; $T --dims 55 55 --substeps 4 --check 0 --schedule per_color --partition graph --time-substeps 20 2>&1 | grep -E "PART_RESULT|rror" | head -4
  echo "== 1M graph dataflow"; $T --dims 55 55 --substeps 4 --check 0 --schedule dataflow --partition graph --time-substeps 50 2>&1 | grep -E "PART_RESULT|rror" | head -4
  echo "== 4M graph per_color"; $T --dims 88 88 --substeps 4 --check 0 --schedule per_color --partition graph --time-substeps 20 2>&1 | grep -E "PART_RESULT|rror" | head -4
  echo "== 4M graph persistent"; $T --dims 88 88 --substeps 4 --check 0 --schedule persistent --partition graph --time-substeps 20 2>&1 | grep -E "PART_RESULT|rror" | head -4
  echo "== 384k graph per_color parity"; $T --dims 40 40 --substeps 8 --schedule per_color --partition graph 2>&1 | grep -E "PART_RESULT|rror" | head -4
} > gpurun_out/part_graph_debug_n$N.log 2>&1
cut -c1-420 gpurun_out/part_graph_debug_n$N.log
