#!/bin/bash
# Round 2: batch kernel with O/V in global memory + general kernel without register prefetch: tests and bench legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -q -m gpu -x -k "batch or damp or volume" 2>&1 | tail -4 > gpurun_out/gpu_tests_batch.log
timeout 600 python bench.py --no-cpu-baseline --extras damped,batch > gpurun_out/bench_r2_batch.json 2> gpurun_out/bench_r2_batch.err
cat gpurun_out/gpu_tests_batch.log; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2_batch.json"))
print("headline", d["value"], "damped", d["extra"]["damped"]["value"], d["extra"]["damped"]["ms_per_step"])
for k, v in d["extra"]["batch"].items():
    print(k, v["value"], v["ms_per_step"], "e2e", v["e2e"]["value"], "smem", v["smem_bytes"], "block", v["block_threads"], "group", v["group_threads"])
PY
tail -3 gpurun_out/bench_r2_batch.err
