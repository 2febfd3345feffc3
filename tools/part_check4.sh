#!/bin/bash
# Four GPUs (gpurun --gpus 4): 4-rank bench line (independent scenes) and one 8M-tet mesh partitioned over 4 GPUs.
mkdir -p gpurun_out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu"
{
  echo "== bench --gpus 4"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -E '^\{|Error' | cut -c1-900
  echo "== 384k parity+timing dataflow, 4 GPUs"; $T --dims 40 40 --substeps 8 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|Error" | head -3
  echo "== 8M timing dataflow, 4 GPUs"; $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -3
} > gpurun_out/part_check4.log 2>&1
cut -c1-600 gpurun_out/part_check4.log
