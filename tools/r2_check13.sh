#!/bin/bash
# Debug: 20M tets on 4 GPUs, barrier-free: energy 4 vs 7, one long call vs several back-to-back calls, bench leg
N=${1:-4}
mkdir -p gpurun_out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu --dims 150 150 --substeps 4 --check 0 --schedule dataflow"
{
  echo "== energy 4, 1 x 100"; $T --energy 4 --time-substeps 100 2>&1 | grep -E "PART_RESULT|XfError:" | head -2
  echo "== energy 4, 6 x 50"; $T --energy 4 --time-substeps 50 --time-calls 6 2>&1 | grep -E "PART_RESULT|XfError:" | head -2
  echo "== energy 7, 6 x 50"; $T --energy 7 --time-substeps 50 --time-calls 6 2>&1 | grep -E "PART_RESULT|XfError:" | head -2
  echo "== bench partitioned leg"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --extras partitioned 2>&1 | grep -E "XfError:|^\{" | cut -c1-200 | head -3
} > gpurun_out/part_dbg13_n$N.log 2>&1
cut -c1-400 gpurun_out/part_dbg13_n$N.log
