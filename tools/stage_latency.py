import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
xf = load_package()
for mode, name in ((0, "element solve, lone warp (cycles)"), (1, "record hand-off SM -> L2 -> polling SM (cycles, one way)"), (2, "dependent 256-bit L2 load (cycles)")):
    vals = [xf.stage_latency(0, mode, 4000) for _ in range(3)]
    print(name, ["%.0f" % v for v in vals], flush=True)
