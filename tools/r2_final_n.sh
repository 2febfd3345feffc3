#!/bin/bash
# Round 2 final multi-GPU run (gpurun --gpus N): partition + hardening tests, bench.py --gpus N (both arms at N), raw logs.
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_partition.py tests/test_gpu_hardening.py tests/test_batch_sharding.py -q -m gpu 2>&1 | tail -5 > gpurun_out/gpu_tests_final_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 \
   > gpurun_out/bench_final_n$N.json 2> gpurun_out/bench_final_n$N.err; echo "bench rc=$?" >> gpurun_out/gpu_tests_final_n$N.log
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu"
{
  echo "== 384k parity+timing, slabs, barrier-free"; $T --dims 40 40 --substeps 8 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|rror" | head -3
  echo "== 8M timing, slabs, barrier-free"; $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|rror" | head -3
  echo "== 20M timing, slabs, barrier-free"; $T --dims 150 150 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|rror" | head -3
  echo "== 8M timing, graph partition, flag protocol (auto)"; $T --dims 110 110 --substeps 4 --check 0 --schedule auto --partition graph --time-substeps 10 2>&1 | grep -E "PART_RESULT|rror" | head -3
  echo "== 384k parity, graph partition, web default damping"; $T --dims 40 40 --substeps 8 --schedule auto --partition graph --damping 0.005 --rayleigh 3 2>&1 | grep -E "PART_RESULT|rror" | head -3
} > gpurun_out/part_final_n$N.log 2>&1
cat gpurun_out/gpu_tests_final_n$N.log; cut -c1-400 gpurun_out/part_final_n$N.log; grep -E '^\{' gpurun_out/bench_final_n$N.json | cut -c1-200; tail -2 gpurun_out/bench_final_n$N.err
