"""Measurement helper: cost of the bare grid barrier for a few grid shapes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
xf = load_package()
L = xf.lib()
L.xf_debug_barrier_us.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_float)]
names = {0: "fence+atomicAdd+ld.acquire spin+fence", 1: "red.release + ld.acquire spin", 2: "red.release + ld.relaxed spin + fence",
         3: "variant 2 + nanosleep(40)", 4: "cooperative_groups grid.sync", 5: "tree, groups of 8", 6: "tree, groups of 16", 7: "tree, groups of 32"}
for variant in (0, 2, 4, 5, 6, 7):
    for bps, th in [(1, 128), (1, 256), (2, 256)]:
        us = C.c_float()
        rc = L.xf_debug_barrier_us(0, variant, bps, th, 2000, C.byref(us))
        print("variant %d (%s) blocks/SM=%d threads=%d rc=%d  %.3f us/barrier" % (variant, names[variant], bps, th, rc, us.value))
