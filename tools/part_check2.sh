#!/bin/bash
# Two GPUs (gpurun --gpus 2): partitioned-mesh GPU test, parity + timing of one mesh on 2 GPUs, 2-rank bench line.
mkdir -p gpurun_out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu"
{
  echo "== tests: partition (2 GPUs)"; timeout 300 python -m pytest tests/test_partition.py -q -m gpu 2>&1 | tail -3
  echo "== 384k parity+timing dataflow"; $T --dims 40 40 --substeps 8 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|Error" | head -3
  echo "== 8M timing dataflow"; $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -3
  echo "== bench --gpus 2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -E '^\{|Error' | cut -c1-700
} > gpurun_out/part_check2.log 2>&1
cut -c1-500 gpurun_out/part_check2.log
