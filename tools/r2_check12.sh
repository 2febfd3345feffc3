#!/bin/bash
# parity of the final element arithmetic + bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/gpu_tests_final2.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err
cat gpurun_out/gpu_tests_final2.log; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final2_n1.json"))
e = d["extra"]
print("value %.4e ms %.4f e2e %.4e frac %.4f frac_l2 %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["frac_l2"]))
print("damped %.4e native fps %.2f boxL %.4e beamL %.4e part %.4e %s" % (e["damped"]["value"], e["native_rate"]["frames_per_s"], e["batch"]["boxL"]["value"], e["batch"]["beamL"]["value"], e["partitioned"]["value"], e["partitioned"]["parity_ok"]))
PY
tail -2 gpurun_out/bench_final2_n1.err
