#!/bin/bash
# Partitioned-mesh checks and timings (run under gpurun --gpus N).
N=${1:-2}
mkdir -p gpurun_out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tools/part_worker.py --mode gpu"
{
  echo "== small parity dataflow"; $T --dims 12 6 --substeps 15 --schedule dataflow 2>&1 | grep -E "PART_RESULT|rror" | tail -3
  echo "== small parity dataflow nohint"; $T --dims 12 6 --substeps 15 --schedule dataflow --no-hint --energy 5 2>&1 | grep -E "PART_RESULT|rror" | tail -3
  echo "== 384k parity+timing dataflow"; $T --dims 40 40 --substeps 8 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|rror" | tail -3
  echo "== 1M timing dataflow"; $T --dims 55 55 --substeps 4 --check 0 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|rror" | tail -3
  echo "== 8M timing dataflow"; $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 60 2>&1 | grep -E "PART_RESULT|rror" | tail -3
  echo "== 8M timing dataflow, 2x iface warps"; XF_PART_IFACE_WARPS=400 $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 60 2>&1 | grep -E "PART_RESULT|rror" | tail -3
  if [ "$N" -ge 4 ]; then
  echo "== 20M timing dataflow"; $T --dims 150 150 --substeps 4 --check 0 --schedule dataflow --time-substeps 60 2>&1 | grep -E "PART_RESULT|rror" | tail -3
  fi
} > gpurun_out/part_sweep_$N.log 2>&1
cat gpurun_out/part_sweep_$N.log | cut -c1-400
