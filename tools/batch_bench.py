"""Measurement helper: throughput of the batched-scene path (BASELINE config 3) on one GPU."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from __graft_entry__ import load_package
xf = load_package()
ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=4096)
ap.add_argument("--shape", choices=["beamL", "boxL"], default="boxL")
ap.add_argument("--precision", choices=["exact", "fast"], default="exact")
ap.add_argument("--substeps", type=int, default=50)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dims = (8, 2) if a.shape == "beamL" else (8, 8)
nodes, idx, hint = xf.GenerateTetBlock(*dims)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
b = xf.GeoBatchCuda(nodes, idx, a.scenes, precision=xf.PRECISION_EXACT if a.precision == "exact" else xf.PRECISION_FAST, color_hint=hint,
                    stream=stream.cuda_stream)
arr = (xf.Settings * a.scenes)()
for s in range(a.scenes):
    arr[s] = xf.make_settings(energy=7, poisson=0.5, gravity=(0.0, -0.4905 * (1 + 0.1 * (s % 7))), compliance=1.0 + 0.25 * (s % 4))
dt = np.float32(1 / 3000)
for _ in range(3):
    b.Substep(arr, dt, a.substeps)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
for s0, s1 in ev:
    s0.record(); b.Substep(arr, dt, a.substeps); s1.record()
torch.cuda.synchronize()
ms = sum(x.elapsed_time(y) for x, y in ev) / a.steps
X, V, w = b.get_state(0, 1)
print(json.dumps({"workload": "%d x %s (%d tets each), yeohskinfast nu=0.5, per-scene gravity/compliance" % (a.scenes, a.shape, b.nT),
                  "precision": a.precision, "ms_per_step": ms, "substeps_per_step": a.substeps,
                  "element_substeps_per_s": a.scenes * b.nT * a.substeps / (ms * 1e-3), "finite": bool(np.isfinite(X).all()), **b.info()}))
