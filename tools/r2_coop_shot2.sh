#!/bin/bash
# Second (last) GPU call for the four-lane probe: both variants (vertex 3 broadcast / gathered by every lane, branch-free stores and
# selects), then an ncu capture of the two lone-warp launches (one-thread, four-lane) for the stall breakdown.
mkdir -p gpurun_out
timeout 60 python tools/coop_probe.py > gpurun_out/r2_coop_probe.json 2> gpurun_out/r2_coop_probe.err; echo "probe rc=$?"
timeout 40 python -m pytest tests/test_gpu_coop.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_gpu_tests_coop.log; cat gpurun_out/r2_gpu_tests_coop.log
timeout 70 ncu --set full --import-source on --clock-control none -k regex:k_probe --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2_prof_coop \
    python tools/coop_probe.py --one --variant 1 > gpurun_out/r2_ncu_coop.log 2>&1; echo "ncu rc=$?"
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
tail -3 gpurun_out/r2_coop_probe.err; tail -4 gpurun_out/r2_ncu_coop.log
