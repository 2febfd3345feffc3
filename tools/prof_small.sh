#!/bin/bash
# ncu source-level stall sampling of the chain-bound regime (24k tets: one warp per SM) for the pair and the single-thread kernels
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_substeps_pair -s 3 -c 1 -o gpurun_out/prof_pair_small -f python bench.py --cells 16 --substeps-per-step 20 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pair_small.log 2>&1
XF_NO_PAIRS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_substeps_dataflow -s 3 -c 1 -o gpurun_out/prof_df_small -f python bench.py --cells 16 --substeps-per-step 20 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_df_small.log 2>&1
ls -la gpurun_out/*.ncu-rep
