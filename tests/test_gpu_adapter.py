"""Drop-in demonstration (GPU): the reference's own `Geo` virtual interface (Geo.h:15-35) implemented by the CUDA
library (xpbd-fem_b200/host/GeoLinear3dCuda.h, compiled against the UNMODIFIED reference headers into
oracle/_ref/libxpbd_ref_adapter.so) behaves bit-identically to the reference's GeoLinear3d when both are driven
through the same virtual calls, one Substep per call, the way Sim::Update does (Demo.cpp:86-88)."""
import numpy as np
import pytest

from __graft_entry__ import build, load_package
from oracle import bindings as ob

build()
xf = load_package()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ob.have_ref("adapter"), reason="oracle/_ref/libxpbd_ref_adapter.so not built")]
DT = np.float32(1.0 / 3000.0)


def test_adapter_matches_reference_geo_through_virtual_interface():
    nodes, idx, hint = xf.GenerateTetBlock(8, 8, wonkiness=0.15)
    cuda_geo = ob.AdapterScene(nodes, idx, color_hint=hint)
    ref_geo = ob.RefScene.mesh(nodes, idx)
    ref_geo.set_order(cuda_geo.get_order())
    assert (cuda_geo.nV, cuda_geo.nT) == (ref_geo.nV, ref_geo.nT)
    rot = np.array([0.0, -1.0, 0.0, 1.0, 0.0, 0.0, 0.01, 0.02, 1.0], dtype=np.float32)  # Sim::FinishAddingBlocks, Demo.cpp:157-161
    cuda_geo.transform(rot)
    ref_geo.transform(rot)
    assert cuda_geo.volume() == ref_geo.volume()
    st = ob.make_settings(energy=ob.Energy_YeohSkinFast, poisson=0.5, damping=0.004, rayleigh=ob.Rayleigh_PostAmortized, pbd_damping=0.03)
    st.volumeAndTimeCorrectedPbdDamping = 1e-6
    st.amortizedVolumeAndTimeCorrectedPbdDamping = 7e-6
    manip = ob.Manipulator()
    manip.pos[:] = (0.0, 0.0, 0.3)
    manip.manipPlaneNormal[:] = (0.0, 0.0, 1.0)
    manip.pick0[:] = (0.01, 0.0, 0.0)
    manip.pickDirTarget[:] = (0.02, 0.05, -1.0)
    manip.picked = 1
    manip.pickedPointIdx = 100
    cuda_geo.substep(st, DT, 40, manip=manip)
    ref_geo.substep(st, DT, 40, manip=manip)
    Xc, Vc, wc = cuda_geo.get_state()
    Xr, Vr, wr = ref_geo.get_state()
    assert np.array_equal(Xc, Xr) and np.array_equal(Vc, Vr) and np.array_equal(wc, wr)
    assert cuda_geo.volume() == ref_geo.volume()


def test_adapter_pick_selects_the_reference_vertex():
    """Geo::Pick (Geo.cpp:366-385) through both implementations: same vertex, point and distance for rays that hit the body, graze
    it, start inside it, point away from it (nothing in front: outputs untouched), and after the left lock has zeroed the inverse
    masses of the locked vertices (they must not be grabbed)."""
    nodes, idx, hint = xf.GenerateTetBlock(6, 4, wonkiness=0.1)
    cuda_geo = ob.AdapterScene(nodes, idx, color_hint=hint)
    ref_geo = ob.RefScene.mesh(nodes, idx)
    ref_geo.set_order(cuda_geo.get_order())
    st = ob.make_settings(energy=ob.Energy_MixedSel, poisson=0.45, lock_left=True)
    rng = np.random.default_rng(5)
    P = nodes.reshape(-1, 3)
    lo, hi = P.min(axis=0), P.max(axis=0)
    rays = []
    for _ in range(40):
        target = lo + (hi - lo) * rng.random(3)
        origin = target + np.array([0.0, 0.0, 0.4]) + 0.05 * rng.standard_normal(3)
        d = target - origin
        rays.append((origin, d / np.linalg.norm(d)))
    rays.append((0.5 * (lo + hi), np.array([0.0, 0.0, -1.0])))                       # starts inside
    rays.append((hi + 0.3, np.array([1.0, 0.0, 0.0])))                               # everything behind the origin
    rays.append((np.array([lo[0], lo[1], 0.5]), np.array([0.0, 0.0, -1.0])))         # aims at a locked corner
    for phase in range(2):
        for origin, d in rays:
            a = cuda_geo.pick(origin, d)
            r = ref_geo.pick(origin, d)
            assert a[0] == r[0] and a[1] == r[1], (phase, origin, d, a, r)
            assert np.array_equal(a[2], r[2]) and a[3] == r[3]
        # two substeps: the first zeroes w of the locked vertices (Geo.cpp:320), changing which vertices can be picked
        cuda_geo.substep(st, DT, 2)
        ref_geo.substep(st, DT, 2)
