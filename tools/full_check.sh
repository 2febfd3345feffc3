#!/bin/bash
# Full GPU check: smoke, test suite, default bench (both arms), ncu launch list + one full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_dataflow.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_substeps_dataflow -s 3 -c 1 -o gpurun_out/prof_dataflow_r1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/smoke.log; cat gpurun_out/gpu_tests.log; cut -c1-700 gpurun_out/bench_default.json; cut -c1-400 gpurun_out/bench_ref.json
