#!/bin/bash
# Round-2 first check: smoke, GPU suite, default bench, L2 bandwidth probe (copy/read/write, 256-bit accesses).
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python - > gpurun_out/l2_probe.log 2>&1 <<'PY'
from __graft_entry__ import load_package
xf = load_package()
for mode in (0, 1):
    for mb in (8, 16, 32, 48, 64, 96, 256, 1024):
        for bps in (4, 8, 16):
            print("mode", mode, "MB", mb, "blocks/SM", bps, "GB/s %.1f" % xf.l2_bandwidth(0, mode, mb, passes=200 if mb <= 96 else 20, reps=5, blocks_per_sm=bps), flush=True)
PY
tail -3 gpurun_out/smoke.log; cat gpurun_out/gpu_tests.log; cut -c1-900 gpurun_out/bench_default.json; tail -30 gpurun_out/l2_probe.log
