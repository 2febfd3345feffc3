#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE
(oracle/_ref/libxpbd_ref_strict.so = /root/reference/XPBDFEM/*.cpp behind oracle/ref_harness.cpp,
g++ -O2 -ffp-contract=off).  The reference ships no golden vectors of its own, so these
are the pins.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

Scenes are the reference's own small shapes as Demo::UpdateSettings builds them for Element_T4
(Demo.cpp:289-318): Beam L (8x2x2 hexes, 192 tets), Box L (8x8x8, 3072 tets), Armadillo (1189 tets).
Each file stores the mesh, the element order used, the Settings bytes, and X,V (fp64) after
1, 10 and 100 calls of Geo3d::Substep at dt = 1/3000.
Orders: "lcg" = the reference's own tOrder (Geo.cpp:759-769); "colour" = the CUDA schedule's equivalent
serial order as computed by the product's host code at generation time (stored, so the fixture stays
valid as a pin of the *reference* even if the colouring heuristic changes later).
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DT = np.float32(1.0 / 3000.0)
SNAPSHOTS = (1, 10, 100)
ENERGIES = {"mixed": 3, "mixedsel": 4, "yeohskin": 5, "yeohskinfast": 7}
NUS = (0.45, 0.495, 0.4999, 0.5)


def make_ref(scene):
    if scene == "beamL":
        return ob.RefScene.block(8, 2, wonkiness=0.0), dict(density=1.0, auto_resize=False)
    if scene == "beamL_wonky":
        return ob.RefScene.block(8, 2, wonkiness=0.3), dict(density=1.0, auto_resize=False)
    if scene == "boxL":
        return ob.RefScene.block(8, 8, wonkiness=0.0), dict(density=1.0, auto_resize=False)
    if scene == "armadillo":
        return ob.RefScene.armadillo(2.0), dict(density=2.0, auto_resize=True)
    raise ValueError(scene)


def colour_order(xf, nodes, idx, meta, lattice_dims):
    hint = None
    if lattice_dims is not None:
        _, _, hint = xf.GenerateTetBlock(*lattice_dims)
    g = xf.GeoLinear3dCuda(nodes, idx, device=-1, color_hint=hint, **meta)  # host-only: colouring, no device
    return g.get_order()


def run_case(scene, order_kind, energy_name, sim, nu, xf):
    ref, meta = make_ref(scene)
    nodes, idx = ref.get_mesh()
    if order_kind == "colour":
        dims = {"beamL": (8, 2), "beamL_wonky": (8, 2), "boxL": (8, 8)}.get(scene)
        order = colour_order(xf, nodes, idx, meta, dims)
        ref.set_order(order)
    else:
        order = ref.get_order()
    st = ob.make_settings(energy=ENERGIES[energy_name], simultaneous=sim, poisson=nu)
    if scene == "armadillo":  # the native defaults, Demo.cpp:20-35 (minus in-constraint damping)
        st.compliance = 3.2
        st.gravity[1] = -0.602
        st.flags &= ~ob.Settings_LockLeft
    out = dict(nodes=nodes, idx=idx, order=order.astype(np.uint32), settings=np.frombuffer(bytes(st), dtype=np.uint8).copy(),
               dt=np.float32(DT), density=np.float32(meta["density"]), auto_resize=np.uint8(meta["auto_resize"]),
               snapshots=np.array(SNAPSHOTS, dtype=np.uint32))
    done = 0
    for snap in SNAPSHOTS:
        ref.substep(st, DT, snap - done)
        st.tickId += snap - done
        done = snap
        X, V, w = ref.get_state()
        out["X_%d" % snap] = X
        out["V_%d" % snap] = V
    out["w_final"] = w
    out["volume_final"] = np.float32(ref.volume())
    return out


def main():
    if not ob.have_ref("strict"):
        raise SystemExit("oracle/_ref/libxpbd_ref_strict.so missing: run `make -C oracle ref` where /root/reference exists")
    xf = load_package()
    os.makedirs(OUT, exist_ok=True)
    cases = []
    for energy, sim, nu in itertools.product(ENERGIES, (True, False), NUS):
        cases.append(("beamL", "colour", energy, sim, nu))
    for energy, sim in itertools.product(ENERGIES, (True, False)):
        cases.append(("beamL_wonky", "lcg", energy, sim, 0.5))
    for energy in ENERGIES:
        cases.append(("boxL", "colour", energy, True, 0.5))
    for energy, order in itertools.product(("mixedsel", "yeohskinfast"), ("lcg", "colour")):
        cases.append(("armadillo", order, energy, True, 0.5))
    total = 0
    for scene, order, energy, sim, nu in cases:
        name = "%s_%s_%s_%s_nu%s.npz" % (scene, order, energy, "sim" if sim else "ser", str(nu).replace(".", "p"))
        data = run_case(scene, order, energy, sim, nu, xf)
        path = os.path.join(OUT, name)
        np.savez_compressed(path, **data)
        total += os.path.getsize(path)
    print("wrote %d fixtures, %.1f KiB" % (len(cases), total / 1024.0))


if __name__ == "__main__":
    main()
