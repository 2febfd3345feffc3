#!/bin/bash
# A/B of the release fence between peer store and local store in k_part_dataflow (XF_PART_RELEASE), slabs and graph partition
N=${1:-4}
mkdir -p gpurun_out
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/part_worker.py --mode gpu"
{
  for r in 0 1; do
    echo "== 110^3 slabs dataflow release=$r"; XF_PART_RELEASE=$r $T --dims 110 110 --substeps 4 --check 0 --schedule dataflow --time-substeps 100 2>&1 | grep -E "PART_RESULT|rror" | head -3
    echo "== 110^3 graph dataflow release=$r"; XF_PART_RELEASE=$r $T --dims 110 110 --substeps 4 --check 0 --schedule auto --partition graph --time-substeps 100 2>&1 | grep -E "PART_RESULT|rror" | head -3
    echo "== 55^3 slabs dataflow release=$r"; XF_PART_RELEASE=$r $T --dims 55 55 --substeps 4 --check 0 --schedule dataflow --time-substeps 200 2>&1 | grep -E "PART_RESULT|rror" | head -3
  done
  echo "== 40^3 graph parity release=1"; XF_PART_RELEASE=1 $T --dims 40 40 --substeps 8 --schedule auto --partition graph 2>&1 | grep -E "PART_RESULT|rror" | head -3
} > gpurun_out/part_release_ab_n$N.log 2>&1
cut -c1-420 gpurun_out/part_release_ab_n$N.log
