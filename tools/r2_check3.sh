#!/bin/bash
# Round 2: damping sweeps / volume passes on the barrier-free schedule: parity subset, then the damped bench leg.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hardening.py tests/test_gpu_trajectory.py tests/test_frame_driver.py -q -m gpu -x -k "damp or volume or wrap or stalled or frames or trajectory" 2>&1 | tail -15 > gpurun_out/gpu_tests_damp.log
timeout 600 python bench.py --no-cpu-baseline --extras damped,native_rate > gpurun_out/bench_r2_damped.json 2> gpurun_out/bench_r2_damped.err
cat gpurun_out/gpu_tests_damp.log; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2_damped.json"))
print(d["value"], d["extra"]["damped"], d["extra"]["native_rate"]["frames_per_s"])
PY
tail -3 gpurun_out/bench_r2_damped.err
