#!/bin/bash
# Stage time against elements per colour around the 2-warps-per-scheduler boundary (37 888 elements per colour on 148 SMs).
mkdir -p gpurun_out
{
  for cells in 48 50 52 53 54 55 56 58 60 63; do
    echo "== cells $cells"; timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --cells $cells --substeps-per-step 50
  done
} > gpurun_out/size_scan.log 2>&1
grep -o '^== .*\|"tets": [0-9]*\|"ms_per_step": [0-9.]*\|"grid": [^]]*\]\|rror.*' gpurun_out/size_scan.log | cut -c1-160
