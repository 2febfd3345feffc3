#!/bin/bash
# A/B: precomputed alpha plane (DeviceScene::eAlpha) in the barrier-free kernels; parity first
mkdir -p gpurun_out
XF_TEST_SCHEDULES=4 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hardening.py tests/test_frame_driver.py tests/test_golden.py -q -m gpu -x 2>&1 | tail -4 > gpurun_out/gpu_tests_alpha.log
B="timeout 300 python bench.py --no-cpu-baseline --extras native_rate,damped"
{
  echo "== alpha plane"; $B
  echo "== no alpha plane"; XF_NO_ALPHA_PLANE=1 $B
  echo "== alpha plane again"; $B
  echo "== alpha plane, chains"; $B --grouping chains
} > gpurun_out/bench_alpha.log 2>&1
cat gpurun_out/gpu_tests_alpha.log
grep -o '^== .*\|"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"frames_per_s": [0-9.]*\|rror.*' gpurun_out/bench_alpha.log | cut -c1-160
