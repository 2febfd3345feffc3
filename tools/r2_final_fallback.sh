#!/bin/bash
# The bench's fallback for a stalled barrier-free partition leg, exercised with a fake stall on 2 GPUs (8M tets to keep it short)
mkdir -p gpurun_out
XF_BENCH_FAKE_STALL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 \
  --no-cpu-baseline --extras partitioned --part-cells 110 > gpurun_out/bench_fallback_n2.json 2> gpurun_out/bench_fallback_n2.err; echo "rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/bench_fallback_n2.json"):
    if l.startswith("{"):
        p = json.loads(l)["extra"]["partitioned"]
        print({k: p.get(k) for k in ("parity_ok", "schedule", "kernel", "stalls", "us_per_substep", "value", "gpu_launches")})
PY
grep -E "rror" gpurun_out/bench_fallback_n2.err | head -3
