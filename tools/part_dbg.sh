#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tools/part_worker.py --mode gpu"
{
  for w in 48 96 400; do
  echo "== 8M timing dataflow iface warps $w"; XF_PART_IFACE_WARPS=$w $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -3
  done
} > gpurun_out/part_dbg_$N.log 2>&1
grep -o '^== .*\|"us_per_substep": [0-9.]*\|Error.*' gpurun_out/part_dbg_$N.log
