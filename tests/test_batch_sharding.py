"""BASELINE config 3, N > 1: a batch of independent scenes shards over ranks with no data-path collective.  CPU only:
`xf.shard_scenes` as a function, and a gloo run (world sizes 2 and 3) in which every rank steps its scenes - the C oracle stands in
for the device, as in the partition emulation - and the gathered results are compared with the unsharded run."""
import json
import os
import subprocess
import sys

import pytest

from __graft_entry__ import ROOT, build, load_package

build()
xf = load_package()
WORKER = os.path.join(ROOT, "tests", "batch_shard_worker.py")


@pytest.mark.parametrize("n", [0, 1, 7, 8, 4096, 4099])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_scenes_partitions_the_batch(n, world):
    shards = [xf.shard_scenes(n, world, r) for r in range(world)]
    assert shards[0][0] == 0 and sum(c for _, c in shards) == n
    for (f0, c0), (f1, c1) in zip(shards, shards[1:]):
        assert f1 == f0 + c0 and 0 <= c0 - c1 <= 1          # contiguous, ragged by at most one, larger shards first
    with pytest.raises(ValueError):
        xf.shard_scenes(n, world, world)


@pytest.mark.parametrize("world,scenes", [(2, 11), (3, 10)])
def test_sharded_batch_matches_unsharded_run(world, scenes):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + world), WORKER, "--scenes", str(scenes)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    lines = [l for l in p.stdout.splitlines() if l.startswith("SHARD_RESULT ")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-4000:]
    out = json.loads(lines[-1][len("SHARD_RESULT "):])
    assert out["ok"], out["msg"]
    assert sum(out["counts"]) == scenes and max(out["counts"]) - min(out["counts"]) <= 1
    assert out["distinct"] > 1                               # per-scene settings really differ
