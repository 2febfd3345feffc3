#!/bin/bash
# Round 2 final check on one GPU: smoke, full GPU suite, bench (both arms), ncu launch list + full captures (headline, damped).
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/gpu_tests_final.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_b.log 2>&1
bash tools/ncu_traffic.sh damped > gpurun_out/ncu_traffic.log 2>&1
tail -3 gpurun_out/smoke.log; cat gpurun_out/gpu_tests_final.log; cut -c1-600 gpurun_out/bench_final_n1.json; tail -3 gpurun_out/bench_final_n1.err; cut -c1-300 gpurun_out/bench_final_ref.json
