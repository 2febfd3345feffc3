"""Pins the plain-C restatement (oracle/xpbd_oracle.c) against the UNMODIFIED reference
(oracle/_ref/libxpbd_ref_strict.so = /root/reference/XPBDFEM/*.cpp behind oracle/ref_harness.cpp).
Everything here is bit-exact: both sides are built with -ffp-contract=off.  CPU only."""
import itertools

import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = pytest.mark.skipif(not ob.have_ref("strict"), reason="oracle/_ref not built (no /root/reference here)")
DT = np.float32(1.0 / 3000.0)


def pair(width, height, wonk=0.0, pattern=0, density=1.0):
    r = ob.RefScene.block(width, height, wonkiness=wonk, pattern=pattern, density=density)
    nodes, idx = r.get_mesh()
    return r, ob.OracleScene(nodes, idx, density), nodes, idx


def assert_same_state(r, o):
    Xr, Vr, wr = r.get_state()
    Xo, Vo, wo = o.get_state()
    assert np.array_equal(Xr, Xo)
    assert np.array_equal(Vr, Vo)
    assert np.array_equal(wr, wo)


@pytest.mark.parametrize("wonk,pattern", [(0.0, 0), (0.3, 0), (0.0, 1), (0.45, 1)])
def test_mesh_generation_matches_reference(wonk, pattern):
    r = ob.RefScene.block(5, 3, wonkiness=wonk, pattern=pattern)
    nodes, idx = r.get_mesh()
    n2, i2 = ob.generate_tet_block(5, 3, wonkiness=wonk, pattern=pattern)
    assert np.array_equal(nodes, n2)
    assert np.array_equal(idx, i2)


@pytest.mark.parametrize("wonk,pattern,density", [(0.0, 0, 1.0), (0.3, 0, 1.0), (0.3, 1, 2.5)])
def test_init_matches_reference(wonk, pattern, density):
    r, o, _, _ = pair(4, 3, wonk, pattern, density)
    er, eo = r.get_elements(), o.get_elements()
    for k in er:
        assert np.array_equal(er[k], eo[k]), k
    assert np.array_equal(r.get_order(), o.get_order())
    assert np.array_equal(r.get_rest()[2], o.get_rest()[2])
    assert_same_state(r, o)
    assert r.volume() == o.volume()


def test_armadillo_autoresize_init_and_run():
    r = ob.RefScene.armadillo()
    nodes, idx = r.get_mesh()
    o = ob.OracleScene(nodes, idx, density=2.0, auto_resize=True)
    assert (r.nV, r.nT) == (456, 1189) == (o.nV, o.nT)
    er, eo = r.get_elements(), o.get_elements()
    for k in er:
        assert np.array_equal(er[k], eo[k]), k
    assert np.array_equal(r.get_rest()[0], o.get_rest()[0])
    st = ob.make_settings(energy=ob.Energy_YeohSkinFast, poisson=0.5, compliance=3.2, gravity=(0.0, -0.602), lock_left=False)
    r.substep(st, DT, 30)
    o.substep(st, DT, 30)
    assert_same_state(r, o)


@pytest.mark.parametrize("energy,sim,nu", list(itertools.product([3, 4, 5, 7], [True, False], [0.45, 0.495, 0.4999, 0.5])))
def test_substep_bit_exact(energy, sim, nu):
    r, o, _, _ = pair(6, 3, 0.3)
    st = ob.make_settings(energy=energy, simultaneous=sim, poisson=nu)
    r.substep(st, DT, 40)
    o.substep(st, DT, 40)
    assert_same_state(r, o)
    assert r.volume() == o.volume()


def test_unknown_energy_id_takes_the_default_label():
    """Energies outside SURVEY §8 (Pixar*, YeohSkinSel, Cube*, V*) are out of scope; unknown ids take the
    reference's `default:` label = MixedSelective (Fem.cpp:884-887), which is restated."""
    r, o, _, _ = pair(3, 2, 0.2)
    st = ob.make_settings(energy=31)
    r.substep(st, DT, 5)
    o.substep(st, DT, 5)
    assert_same_state(r, o)


@pytest.mark.parametrize("rayleigh,sim,energy", list(itertools.product([0, 1, 2, 3], [True, False], [3, 4, 5, 7])))
def test_damping_variants_bit_exact(rayleigh, sim, energy):
    r, o, _, _ = pair(5, 2, 0.25)
    st = ob.make_settings(energy=energy, simultaneous=sim, poisson=0.495, damping=0.005, rayleigh=rayleigh, pbd_damping=0.03,
                          drag_tc=0.0007)
    st.volumeAndTimeCorrectedPbdDamping = 1e-6
    st.amortizedVolumeAndTimeCorrectedPbdDamping = 7e-6
    r.substep(st, DT, 24)
    o.substep(st, DT, 24)
    assert_same_state(r, o)


def test_volume_passes_lock_right_and_manipulator():
    r, o, _, _ = pair(6, 2, 0.2)
    st = ob.make_settings(energy=ob.Energy_MixedSel, poisson=0.5, lock_right=True, volume_passes=2)
    # animate the right lock: scale + shear in the transform (Demo.cpp:72-80 builds a rotation*scale)
    st.lockedRightTransform3d[0] = 0.9
    st.lockedRightTransform3d[1] = 0.1
    st.lockedRightTransform3d[4] = -0.1
    manip = ob.Manipulator()
    manip.pos[:] = (0.0, 0.0, 0.3)
    manip.manipPlaneNormal[:] = (0.0, 0.0, 1.0)
    manip.pick0[:] = (0.01, 0.0, 0.0)
    manip.pickDirTarget[:] = (0.02, 0.05, -1.0)
    manip.picked = 1
    manip.pickedPointIdx = r.nV // 2
    r.substep(st, DT, 20, manip=manip)
    o.substep(st, DT, 20, manip=manip)
    assert_same_state(r, o)


def test_transform_and_volume():
    r, o, _, _ = pair(4, 2, 0.1)
    m = np.array([0.0, -1.0, 0.0, 1.0, 0.0, 0.0, 0.013, -0.02, 1.0], dtype=np.float32)  # rot90 + offset (Demo.cpp:157-161)
    r.transform(m)
    o.transform(m)
    assert_same_state(r, o)
    assert np.array_equal(r.get_origin(), o.get_origin())
    assert r.volume() == o.volume()
    st = ob.make_settings(lock_right=True)
    r.substep(st, DT, 10)
    o.substep(st, DT, 10)
    assert_same_state(r, o)


def test_extended_substep_equals_reference_substep_when_extensions_off():
    a = ob.RefScene.block(6, 3, wonkiness=0.3)
    b = ob.RefScene.block(6, 3, wonkiness=0.3)
    st = ob.make_settings(energy=ob.Energy_YeohSkin, poisson=0.5, damping=0.004, rayleigh=ob.Rayleigh_PostAmortized)
    a.substep(st, DT, 25)
    b.substep(st, DT, 25, ext=True)
    assert_same_state(a, b)


def test_ground_and_handles_extensions_match_harness():
    """x1/x2 are not in the reference; the harness splices them around the reference's own Constrain()."""
    r, o, _, _ = pair(4, 4, 0.2)
    st = ob.make_settings(poisson=0.5, lock_left=False, gravity=(0.0, -9.81))
    y0 = float(r.get_state()[0][:, 1].min()) - 1e-4
    idxs = np.array([3, r.nV - 1], dtype=np.uint32)
    tg = np.array([[0.0, 0.1, 0.0], [0.05, 0.05, 0.05]], dtype=np.float32)
    for s in (r, o):
        s.set_ground(True, y0, 0.25)
        s.set_handles(idxs, tg)
    r.substep(st, DT, 60, ext=True)
    o.substep(st, DT, 60)
    assert_same_state(r, o)
    assert r.get_state()[0][:, 1].min() >= y0
