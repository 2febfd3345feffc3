"""TEST INFRASTRUCTURE - worker for tests/test_batch_sharding.py (torchrun, gloo): BASELINE config 3's sharding of a batch of independent scenes over
ranks, emulated on the CPU.  Every rank steps ITS scenes (xf.shard_scenes) with the C oracle as the stepper - test infrastructure,
like the partition emulation - using per-scene settings derived from the global scene index, and rank 0 checks the gathered
per-scene digests against the unsharded run: same scenes, each exactly once, same bits whatever the world size."""
import argparse, hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as dist
from __graft_entry__ import load_package
from oracle import bindings as ob
xf = load_package()
ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=11)
ap.add_argument("--substeps", type=int, default=5)
a = ap.parse_args()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
nodes, idx, hint = xf.GenerateTetBlock(4, 2)
order = xf.GeoLinear3dCuda(nodes, idx, device=-1, color_hint=hint).get_order()
dt = np.float32(1 / 3000)


def digest(s):
    o = ob.OracleScene(nodes, idx)
    o.set_order(order)
    st = ob.make_settings(energy=7, poisson=0.5, gravity=(0.0, -0.4905 * (1 + 0.1 * (s % 7))), compliance=1.0 + 0.25 * (s % 4))
    o.substep(st, dt, a.substeps)
    X, V, w = o.get_state()
    return hashlib.sha256(X.tobytes() + V.tobytes()).hexdigest()


first, count = xf.shard_scenes(a.scenes, world, rank)
mine = {first + k: digest(first + k) for k in range(count)}
gathered = [None] * world
dist.all_gather_object(gathered, (first, count, mine))
if rank == 0:
    msg, seen = "", {}
    for r, (f, c, d) in enumerate(gathered):
        if sorted(d) != list(range(f, f + c)):
            msg = "rank %d holds %s, expected [%d, %d)" % (r, sorted(d), f, f + c)
        for s, h in d.items():
            if s in seen:
                msg = "scene %d stepped twice" % s
            seen[s] = h
    if not msg and sorted(seen) != list(range(a.scenes)):
        msg = "scenes missing: %s" % sorted(set(range(a.scenes)) - set(seen))
    if not msg:
        bad = [s for s in range(a.scenes) if seen[s] != digest(s)]
        msg = "scenes differ from the unsharded run: %s" % bad if bad else ""
    print("SHARD_RESULT " + json.dumps({"ok": not msg, "msg": msg, "world": world, "counts": [g[1] for g in gathered],
                                        "distinct": len(set(seen.values()))}))
dist.barrier()
dist.destroy_process_group()
