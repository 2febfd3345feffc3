#!/bin/bash
# Regenerates profiles/r2_traffic.json's inputs for the CURRENT build (run on the GPU box through gpurun):
# one `ncu --set full` capture of the dominant kernel of bench.py's headline step (+ the damped variant with "damped"),
# plus the hash of the kernel sources the capture belongs to.  Post-process here with
#   python tools/ncu_traffic.py gpurun_out/prof_headline.ncu-rep [gpurun_out/prof_damped.ncu-rep]
# bench.py refuses a capture whose source hash differs from the sources it is run with (roofline.traffic = null, loudly).
mkdir -p gpurun_out
python - > gpurun_out/traffic_source_hash.txt <<'PY'
import bench
print(bench.source_hash())
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_substeps_ -s 3 -c 1 -o gpurun_out/prof_headline -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_headline.log 2>&1
if [ "$1" = "damped" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_substeps_ -s 3 -c 1 -o gpurun_out/prof_damped -f \
      python bench.py --damped --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_damped.log 2>&1
fi
tail -2 gpurun_out/ncu_headline.log; cat gpurun_out/traffic_source_hash.txt
