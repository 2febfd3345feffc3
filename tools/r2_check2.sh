#!/bin/bash
# Round 2, N GPUs (gpurun --gpus N): two-GPU tests (partition parity, torn records over NVLink), bench.py --gpus N with the extras.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_partition.py tests/test_gpu_hardening.py -q -m gpu 2>&1 | tail -5 > gpurun_out/gpu_tests_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 \
   > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench rc=$?" >> gpurun_out/gpu_tests_n$N.log
cat gpurun_out/gpu_tests_n$N.log; grep -E '^\{' gpurun_out/bench_r2_n$N.json | cut -c1-300; tail -5 gpurun_out/bench_r2_n$N.err
