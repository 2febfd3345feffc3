#!/bin/bash
# Final GPU call of round 2: the four-lane probe with the scalar one-thread baseline added, its GPU test, smoke() on the final tree.
mkdir -p gpurun_out
timeout 50 python tools/coop_probe.py > gpurun_out/r2_coop_probe.json 2> gpurun_out/r2_coop_probe.err; echo "probe rc=$?"
timeout 30 python -m pytest tests/test_gpu_coop.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_gpu_tests_coop.log; cat gpurun_out/r2_gpu_tests_coop.log
timeout 40 python __graft_entry__.py smoke > gpurun_out/r2_smoke_last.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke_last.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_coop_probe.json"))
    for r in d["runs"]:
        print(r["energy"], r["poisson"], [(p["variant"], p["iterations"], p["mismatched_doubles"]) for p in r["parity"]])
        for t in r["timing"]:
            print("  v%d w%2d" % (t["variant"], t["warps_per_sm"]), t["lone_warp_cycles_per_solve"], t["element_solves_per_s"])
except Exception as e:
    print("probe output unreadable:", e)
PY
tail -3 gpurun_out/r2_coop_probe.err
