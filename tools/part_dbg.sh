#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tools/part_worker.py --mode gpu"
{
  echo "== 384k parity+timing dataflow"; $T --dims 40 40 --substeps 8 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|Error" | head -3
  echo "== 8M timing dataflow"; $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -3
  echo "== 20M timing dataflow"; $T --dims 150 150 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -3
} > gpurun_out/part_dbg_$N.log 2>&1
cut -c1-400 gpurun_out/part_dbg_$N.log
