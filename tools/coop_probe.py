#!/usr/bin/env python
"""Warp-cooperative element solve (four lanes per element, xf_element_coop.cuh) against the one-thread solve of the stepping
kernels: bit parity, lone-warp latency (cycles per solve) and full-chip throughput (element solves per second) on real element
records of a wonky MeshGen block in a deformed state.  Output: one JSON document (profiles/r2_coop_probe.json).
usage (GPU box): python tools/coop_probe.py > gpurun_out/r2_coop_probe.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

xf = load_package()


def main():
    one = "--one" in sys.argv  # a single call (for ncu): YeohSkinFast, nu = 0.5, 2000 iterations, 8 warps per SM, variant from --variant N
    nodes, idx, hint = xf.GenerateTetBlock(10, 10, wonkiness=0.3)   # 6000 tets
    geo = xf.GeoLinear3dCuda(nodes, idx, device=-1, color_hint=hint)  # host-only scene: element constants only
    el = geo.get_elements()
    X0, _, w = geo.get_state()
    rng = np.random.default_rng(7)
    h = float(np.abs(np.diff(np.unique(np.round(X0[:, 0], 6)))).min())
    X = X0 + rng.uniform(-0.12 * h, 0.12 * h, size=X0.shape)
    w = np.where(w > 0, w, np.float32(1.0)).astype(np.float32)
    consts, Xg, wg = xf.gathered_elements(el, X, w)
    if one:
        variant = int(sys.argv[sys.argv.index("--variant") + 1]) if "--variant" in sys.argv else 0
        r = xf.coop_element_probe(consts, Xg, wg, xf.substep_constants(1.0, 0.5, 1.0 / 3000.0), energy=xf.Energy_YeohSkinFast, iterations=2000,
                                  warps_per_sm=8, variant=variant)
        print(json.dumps({k: v for k, v in r.items() if not k.startswith("x_")}))
        return
    out = {"what": "xf_debug_coop_element: four lanes per element vs one thread per element, same gathered elements",
           "elements": int(consts.shape[0]), "mesh": "GenerateTetBlock(10, 10, wonkiness 0.3), positions perturbed by +-0.12 h", "runs": []}
    for energy, name in ((xf.Energy_YeohSkinFast, "yeohskinfast"), (xf.Energy_MixedSel, "mixedsel")):
        for poisson in (0.5, 0.45):
            p4 = xf.substep_constants(1.0, poisson, 1.0 / 3000.0)
            run = {"energy": name, "poisson": poisson, "parity": [], "timing": []}
            # parity: variant 0 (lane 3 broadcasts vertex 3) on a short and a longer chain of solves; variant 1 (every lane gathered
            # vertex 3 itself) on one solve - the probe gathers once, so its chains are not comparable (xf_probe_coop.cu)
            # variant 2: the one-thread side runs the scalar arithmetic (no two-wide instructions)
            for variant, its in ((0, (1, 20)), (1, (1,)), (2, (1, 20))):
                for it in its:
                    r = xf.coop_element_probe(consts, Xg, wg, p4, energy=energy, iterations=it, warps_per_sm=1, variant=variant)
                    run["parity"].append({"variant": variant, "iterations": it, "mismatched_doubles": r["mismatched"],
                                          "compared_doubles": r["compared"], "moved": bool(not np.array_equal(r["x_single"], Xg)),
                                          "finite": bool(np.isfinite(r["x_coop"]).all())})
            if poisson == 0.5:
                for variant in (0, 1, 2):
                    for wps in ((1, 2, 4, 8, 12, 16) if variant < 2 else (4, 8, 12, 16)):
                        r = xf.coop_element_probe(consts, Xg, wg, p4, energy=energy, iterations=2000, warps_per_sm=wps, variant=variant)
                        run["timing"].append({"variant": variant, "warps_per_sm": wps, "iterations": 2000,
                                              "one_thread_arithmetic": "scalar" if variant & 2 else "two-wide (what the kernels run)",
                                              "lone_warp_cycles_per_solve": {"one_thread": r["cycles_single"], "four_lane": r["cycles_coop"]},
                                              "element_solves_per_s": {"one_thread": r["solves_per_s_single"], "four_lane": r["solves_per_s_coop"]},
                                              "sm_count": r["sm_count"], "clock_khz": r["clock_khz"]})
            out["runs"].append(run)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
