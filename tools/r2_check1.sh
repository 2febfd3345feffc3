#!/bin/bash
# Round 2: the new bench line (extras) on one GPU, the reference arm, and the batch kernel's instruction counts (issue figure).
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err
for shape in boxL beamL; do
  timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_batch -s 3 -c 1 --csv \
    --log-file gpurun_out/batch_inst_$shape.csv python tools/batch_bench.py --shape $shape --steps 1 > gpurun_out/batch_inst_$shape.log 2>&1
done
cut -c1-3000 gpurun_out/bench_r2_n1.json; tail -5 gpurun_out/bench_r2_n1.err; cut -c1-600 gpurun_out/bench_r2_ref.json; tail -3 gpurun_out/batch_inst_boxL.csv
