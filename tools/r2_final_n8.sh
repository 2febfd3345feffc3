#!/bin/bash
# Round 2 final 8-GPU run (gpurun --gpus 8): bench.py --gpus 8 with every extra leg, raw logs of one mesh on 8 GPUs.
N=8
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 \
   > gpurun_out/bench_final_n$N.json 2> gpurun_out/bench_final_n$N.err; echo "bench rc=$?" > gpurun_out/final_n$N.log
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu"
{
  echo "== 384k parity+timing, slabs, barrier-free"; $T --dims 40 40 --substeps 8 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|XfError:" | head -3
  echo "== 8M timing, slabs, barrier-free"; $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|XfError:" | head -3
  echo "== 20M timing, slabs, barrier-free, 4 x 50"; $T --dims 150 150 --substeps 4 --check 0 --schedule dataflow --time-substeps 50 --time-calls 4 2>&1 | grep -E "PART_RESULT|XfError:" | head -3
  echo "== 384k parity, graph partition, web default damping (flag protocol)"; $T --dims 40 40 --substeps 8 --schedule auto --partition graph --damping 0.005 --rayleigh 3 2>&1 | grep -E "PART_RESULT|XfError:" | head -3
} > gpurun_out/part_final_n$N.log 2>&1
cat gpurun_out/final_n$N.log; cut -c1-400 gpurun_out/part_final_n$N.log; grep -E '^\{' gpurun_out/bench_final_n$N.json | cut -c1-200; grep -E "XfError|rror:" gpurun_out/bench_final_n$N.err | head -3
