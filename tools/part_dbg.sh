#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tools/part_worker.py --mode gpu"
{
  echo "== 384k parity+timing dataflow, 4 interface SMs"; XF_PART_IFACE_SMS=4 $T --dims 40 40 --substeps 8 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|Error" | head -3
  for k in 0 3 6 12; do
  echo "== 8M timing dataflow iface SMs $k"; XF_PART_IFACE_SMS=$k $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -3
  done
} > gpurun_out/part_dbg_$N.log 2>&1
grep -o '^== .*\|"ok": [a-z]*\|"us_per_substep": [0-9.]*\|Error.*' gpurun_out/part_dbg_$N.log
