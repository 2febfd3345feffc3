#!/bin/bash
# One mesh on N GPUs (run under gpurun --gpus N): parity against the oracle at 384k tets, timings up to 20M tets.
N=${1:-2}
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu"
{
  echo "== 384k parity+timing dataflow"; $T --dims 40 40 --substeps 8 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|Error" | head -3
  echo "== 1M timing dataflow"; $T --dims 55 55 --substeps 4 --check 0 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|Error" | head -3
  echo "== 1M timing per_color"; $T --dims 55 55 --substeps 4 --check 0 --schedule per_color --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -3
  echo "== 8M timing dataflow"; $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -3
  if [ "$N" -ge 4 ]; then
    echo "== 20M timing dataflow"; $T --dims 150 150 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -3
  fi
} > gpurun_out/part_sweep_$N.log 2>&1
cut -c1-400 gpurun_out/part_sweep_$N.log
