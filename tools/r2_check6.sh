#!/bin/bash
# Round 2, N GPUs (gpurun --gpus N): partition tests (slabs + graph partition, 2 and 4 ranks), IFACE_WARPS sweep of the partitioned
# barrier-free kernel, bench.py --gpus N with the extras.
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_partition.py tests/test_gpu_hardening.py -q -m gpu 2>&1 | tail -5 > gpurun_out/gpu_tests_n$N.log
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu"
{
  for w in 4 8 16 32 64; do
    echo "== 8M tets, XF_PART_IFACE_WARPS=$w"; XF_PART_IFACE_WARPS=$w $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|Error" | head -2
  done
  echo "== 8M tets, graph partition (flag protocol)"; $T --dims 110 110 --substeps 4 --check 0 --schedule auto --partition graph --time-substeps 20 2>&1 | grep -E "PART_RESULT|Error" | head -2
} > gpurun_out/part_iface_sweep_n$N.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 \
   > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench rc=$?" >> gpurun_out/gpu_tests_n$N.log
cat gpurun_out/gpu_tests_n$N.log; cut -c1-330 gpurun_out/part_iface_sweep_n$N.log; grep -E '^\{' gpurun_out/bench_r2_n$N.json | cut -c1-200; tail -3 gpurun_out/bench_r2_n$N.err
