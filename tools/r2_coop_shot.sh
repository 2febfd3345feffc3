#!/bin/bash
# Last GPU seconds of round 2: the four-lanes-per-element probe (parity, lone-warp latency, full-chip throughput), its GPU test,
# then - if time is left - a bench line of the final tree (headline + damped + native rate legs).
mkdir -p gpurun_out
timeout 70 python tools/coop_probe.py > gpurun_out/r2_coop_probe.json 2> gpurun_out/r2_coop_probe.err; echo "probe rc=$?"
timeout 60 python -m pytest tests/test_gpu_coop.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_gpu_tests_coop.log; cat gpurun_out/r2_gpu_tests_coop.log
timeout 100 python bench.py --steps 20 --warmup 5 --extras damped,native_rate > gpurun_out/r2_bench_last_n1.json 2> gpurun_out/r2_bench_last_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_coop_probe.json"))
    for r in d["runs"]:
        print(r["energy"], r["poisson"], r["parity"])
        for t in r["timing"]:
            print("  ", t["warps_per_sm"], t["lone_warp_cycles_per_solve"], t["element_solves_per_s"], t["mismatched_doubles"])
except Exception as e:
    print("probe output unreadable:", e)
PY
cut -c1-400 gpurun_out/r2_bench_last_n1.json; tail -2 gpurun_out/r2_coop_probe.err gpurun_out/r2_bench_last_n1.err
