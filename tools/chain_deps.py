#!/usr/bin/env python
"""Host-only analysis behind DESIGN.md section 9 ("what comes next"): for every colour of a MeshGen lattice with the ring-ordered
hint, how many records an element of the chained sweep gathers from L2, how many of those were written in the immediately preceding
stage, and how far away (in 32-element chunks = warps) their writers sit.  No GPU needed.  usage: python tools/chain_deps.py [cells]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

xf = load_package()
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
nodes, idx, hint = xf.GenerateTetBlock(cells, cells)
g = xf.GeoLinear3dCuda(nodes, idx, device=-1, grouping=xf.GROUPING_CHAINS, color_hint=hint)
order, colors = g.get_order(), g.get_colors()
info, permille = g.chain_info()
tets = idx.reshape(-1, 5)[:, 1:].astype(np.int64)
sizes = np.bincount(colors, minlength=g.nColors)
start = np.concatenate([[0], np.cumsum(sizes)])
last_pos = np.full(g.nV, -1, dtype=np.int64)  # position inside its colour of the last writer of every vertex
last_col = np.full(g.nV, -1, dtype=np.int64)
print("%d tets, %d colours, %d per mille of corner uses kept in the thread" % (len(tets), g.nColors, permille))
for c in range(g.nColors):
    T = tets[order[start[c]:start[c + 1]]]
    pos = np.arange(sizes[c])
    words = info[start[c]:start[c + 1]].astype(np.int64)
    if c > 0:
        gathered = prev_stage = same_warp = 0
        offsets = set()
        for n in range(4):
            first = ((words >> (5 * n)) & 8) != 0
            v = T[first, n]
            written = last_col[v] >= 0
            d = last_pos[v][written] - pos[first][written]
            prev = last_col[v][written] == c - 1
            same = (last_pos[v][written] // 32) == (pos[first][written] // 32)
            gathered += int(first.sum())
            prev_stage += int(prev.sum())
            same_warp += int((prev & same).sum())
            offsets |= set(int(x) for x in np.unique(np.round(d[prev] / 32.0).astype(int)))
        print("colour %2d: %.2f gathered records per element, %.2f written in the previous stage, %.2f by a lane of the same warp; "
              "writer's chunk offset %s" % (c, gathered / sizes[c], prev_stage / sizes[c], same_warp / sizes[c], sorted(offsets)[:8]))
    for n in range(4):
        last_pos[T[:, n]] = pos
        last_col[T[:, n]] = c
