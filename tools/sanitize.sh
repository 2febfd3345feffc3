#!/bin/bash
# compute-sanitizer over the small end-to-end paths (hardening): memcheck on smoke() (every schedule, bit-exact against the
# oracle under the tool), racecheck on the batched kernel (the one that keeps state in shared memory between __syncwarp / __syncthreads)
# and on the per-colour schedule, synccheck on the batched kernel.
mkdir -p gpurun_out
cat > /tmp/sanitize_batch.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from __graft_entry__ import load_package
xf = load_package()
for dims in ((4, 2), (5, 3)):
    nodes, idx, hint = xf.GenerateTetBlock(*dims, wonkiness=0.2)
    b = xf.GeoBatchCuda(nodes, idx, 6, color_hint=hint)
    st = xf.make_settings(energy=7, poisson=0.5, damping=0.005, rayleigh=3, pbd_damping=0.03)
    st.volumeAndTimeCorrectedPbdDamping = 1e-6; st.amortizedVolumeAndTimeCorrectedPbdDamping = 7e-6
    b.Substep(st, np.float32(1 / 3000), 9)
    X, V, w = b.get_state()
    print("batch", dims, b.info(), bool(np.isfinite(X).all()))
    g = xf.GeoLinear3dCuda(nodes, idx, schedule=xf.SCHEDULE_LAUNCH_PER_COLOR, color_hint=hint)
    g.Substep(st, np.float32(1 / 3000), 3)
    print("per_color", dims, bool(np.isfinite(g.get_state()[0]).all()))
PY
{
  echo "== memcheck: smoke()"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python __graft_entry__.py smoke 2>&1 | tail -12; echo "rc=$?"
  echo "== racecheck: batch + per-colour"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python /tmp/sanitize_batch.py 2>&1 | tail -8; echo "rc=$?"
  echo "== synccheck: batch"; timeout 900 compute-sanitizer --tool synccheck --error-exitcode 3 python /tmp/sanitize_batch.py 2>&1 | tail -6; echo "rc=$?"
} > gpurun_out/sanitizer.log 2>&1
cat gpurun_out/sanitizer.log
