"""Measurement helper: throughput of the batched-scene path (BASELINE config 3) on N GPUs, one process per GPU.
Scenes are sharded over the ranks (xf.shard_scenes) with no data-path collective; torch.distributed only carries the barrier and the
max-over-ranks of the device time.  One GPU: `python tools/batch_bench.py`; N GPUs: `python -m torch.distributed.run --nnodes=1
--nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/batch_bench.py`."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from __graft_entry__ import load_package
xf = load_package()
ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=4096, help="scenes of the whole job (all ranks)")
ap.add_argument("--shape", choices=["beamL", "boxL"], default="boxL")
ap.add_argument("--precision", choices=["exact", "fast"], default="exact")
ap.add_argument("--substeps", type=int, default=50)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
first, count = xf.shard_scenes(a.scenes, world, rank)
dims = (8, 2) if a.shape == "beamL" else (8, 8)
nodes, idx, hint = xf.GenerateTetBlock(*dims)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
b = xf.GeoBatchCuda(nodes, idx, count, device=local_rank, precision=xf.PRECISION_EXACT if a.precision == "exact" else xf.PRECISION_FAST,
                    color_hint=hint, stream=stream.cuda_stream)
arr = (xf.Settings * count)()
for k in range(count):
    s = first + k  # global scene index: per-scene gravity / compliance do not depend on the sharding
    arr[k] = xf.make_settings(energy=7, poisson=0.5, gravity=(0.0, -0.4905 * (1 + 0.1 * (s % 7))), compliance=1.0 + 0.25 * (s % 4))
dt = np.float32(1 / 3000)
for _ in range(3):
    b.Substep(arr, dt, a.substeps)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
for s0, s1 in ev:
    s0.record(); b.Substep(arr, dt, a.substeps); s1.record()
torch.cuda.synchronize()
ms = torch.tensor([sum(x.elapsed_time(y) for x, y in ev) / a.steps], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms.item())
X, V, w = b.get_state(0, 1)
if rank == 0:
    print(json.dumps({"workload": "%d x %s (%d tets each), yeohskinfast nu=0.5, per-scene gravity/compliance" % (a.scenes, a.shape, b.nT),
                      "n_gpus": world, "scenes_rank0": count, "precision": a.precision, "ms_per_step": ms, "substeps_per_step": a.substeps,
                      "element_substeps_per_s": a.scenes * b.nT * a.substeps / (ms * 1e-3), "finite": bool(np.isfinite(X).all()), **b.info()}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
