"""GPU parity tests (run on the B200 box with -m gpu).  The CUDA path is driven through the C ABI and
compared with the CPU oracle (oracle/xpbd_oracle.c, itself pinned bit-for-bit to the unmodified reference)
running the schedule's equivalent serial order.

XF_PRECISION_EXACT: bit-exact (positions, velocities, inverse masses, volume) - the bar is equality.
XF_PRECISION_FAST : FMA contraction; the bar is north_star's 1e-5 x bounding box per substep."""
import itertools
import os

import numpy as np
import pytest

from __graft_entry__ import build, load_package
from oracle import bindings as ob

build()
xf = load_package()
pytestmark = pytest.mark.gpu
DT = np.float32(1.0 / 3000.0)
SCHEDULES = [xf.SCHEDULE_DATAFLOW, xf.SCHEDULE_PERSISTENT, xf.SCHEDULE_LAUNCH_PER_COLOR]
if os.environ.get("XF_TEST_SCHEDULES"):  # development aid: restrict the schedule axis, e.g. XF_TEST_SCHEDULES=4
    SCHEDULES = [int(x) for x in os.environ["XF_TEST_SCHEDULES"].split(",")]


def settings_pair(**kw):
    return xf.make_settings(**kw), ob.make_settings(**kw)


def make_pair(width=6, height=3, wonk=0.3, pattern=0, use_hint=True, precision=None, schedule=xf.SCHEDULE_PERSISTENT, density=1.0,
              grouping=xf.GROUPING_AUTO):
    nodes, idx, hint = xf.GenerateTetBlock(width, height, wonkiness=wonk, pattern=pattern)
    geo = xf.GeoLinear3dCuda(nodes, idx, density=density, precision=xf.PRECISION_EXACT if precision is None else precision,
                             schedule=schedule, color_hint=hint if (use_hint and pattern == 0) else None, grouping=grouping)
    orc = ob.OracleScene(nodes, idx, density)
    orc.set_order(geo.get_order())
    return geo, orc


def assert_bit_exact(geo, orc):
    X, V, w = geo.get_state()
    Xo, Vo, wo = orc.get_state()
    if not np.array_equal(X, Xo):
        bbox = Xo.max() - Xo.min()
        raise AssertionError("X differs: max |dX|/bbox = %.3e" % (np.abs(X - Xo).max() / bbox))
    assert np.array_equal(V, Vo)
    assert np.array_equal(w, wo)


def rel_err(geo, orc):
    X = geo.get_state()[0]
    Xo = orc.get_state()[0]
    return np.abs(X - Xo).max() / (Xo.max(0) - Xo.min(0)).max()


@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("energy,sim,nu", list(itertools.product([3, 4, 5, 7], [True, False], [0.45, 0.495, 0.4999, 0.5])))
def test_exact_substeps_bit_identical(energy, sim, nu, schedule):
    geo, orc = make_pair(schedule=schedule)
    st, ost = settings_pair(energy=energy, simultaneous=sim, poisson=nu)
    for n in (1, 9, 30):  # several calls: exercises the call-boundary post/predict split
        geo.Substep(st, DT, n)
        orc.substep(ost, DT, n)
        assert_bit_exact(geo, orc)
    assert geo.CalculateVolume() == orc.volume()


@pytest.mark.parametrize("schedule", [xf.SCHEDULE_DATAFLOW, xf.SCHEDULE_PERSISTENT, xf.SCHEDULE_LAUNCH_PER_COLOR])
@pytest.mark.parametrize("energy,sim,nu,pattern", [(7, True, 0.5, 0), (4, False, 0.4999, 0), (5, True, 0.495, 1), (3, False, 0.5, 0)])
def test_exact_clustered_coloring_bit_identical(energy, sim, nu, pattern, schedule):
    """XF_GROUPING_CLUSTERS: one thread solves the six tets of a cell back to back on the barrier-free schedule
    (k_substeps_cluster); for the other schedules the same serial order is an ordinary 48-colouring."""
    geo, orc = make_pair(7, 4, 0.25, pattern=pattern, schedule=schedule, grouping=xf.GROUPING_CLUSTERS)
    assert geo.nColors % 6 == 0
    st, ost = settings_pair(energy=energy, simultaneous=sim, poisson=nu)
    for n in (1, 2, 25):
        geo.Substep(st, DT, n)
        orc.substep(ost, DT, n)
        assert_bit_exact(geo, orc)
    assert geo.CalculateVolume() == orc.volume()


@pytest.mark.parametrize("energy,sim,nu,pattern,dims", [
    (7, True, 0.5, 0, (7, 4)), (4, True, 0.5, 0, (6, 3)), (4, False, 0.4999, 0, (7, 4)), (5, True, 0.495, 1, (6, 4)),
    (3, False, 0.5, 0, (5, 5)), (5, False, 0.45, 0, (8, 2)), (7, False, 0.5, 1, (4, 4)), (3, True, 0.4999, 0, (12, 1))])
def test_exact_chained_sweep_bit_identical(energy, sim, nu, pattern, dims):
    """XF_GROUPING_CHAINS on the barrier-free schedule (k_substeps_chain): the records an element shares with the thread's
    next element stay in private shared-memory slots.  Same colouring and serial order as XF_GROUPING_ELEMENTS, same bits;
    several calls (the slots are empty at every call boundary), ground plane, handles and the manipulator on top."""
    geo, orc = make_pair(dims[0], dims[1], 0.25, pattern=pattern, schedule=xf.SCHEDULE_DATAFLOW, grouping=xf.GROUPING_CHAINS)
    assert geo.info()["chainedPermille"] >= 500
    st, ost = settings_pair(energy=energy, simultaneous=sim, poisson=nu)
    y0 = float(orc.get_state()[0][:, 1].min()) - 2e-5
    idxs = np.array([1, geo.nV - 2], dtype=np.uint32)
    tg = np.array([[0.0, 0.08, 0.0], [0.04, 0.05, 0.03]], dtype=np.float32)
    for s in (geo, orc):
        s.set_ground(True, y0, 0.1)
        s.set_handles(idxs, tg)
    mg, mo = xf.Manipulator(), ob.Manipulator()
    for m in (mg, mo):
        m.pos[:] = (0.0, 0.0, 0.3)
        m.manipPlaneNormal[:] = (0.0, 0.0, 1.0)
        m.pick0[:] = (0.01, 0.0, 0.0)
        m.pickDirTarget[:] = (0.02, 0.05, -1.0)
        m.picked = 1
        m.pickedPointIdx = geo.nV // 3
    for n in (1, 2, 40):
        geo.Substep(st, DT, n, manip=mg)
        orc.substep(ost, DT, n, manip=mo)
        assert_bit_exact(geo, orc)
    assert geo.CalculateVolume() == orc.volume()


def test_chained_sweep_matches_plain_dataflow_at_100k_tets():
    """Many warps per SM, several CTAs per colour: the chained and the plain barrier-free kernels must agree bit for bit
    (they share the colouring), and the chained one must have run (info)."""
    nodes, idx, hint = xf.GenerateTetBlock(26, 26, wonkiness=0.15)
    st = xf.make_settings(energy=xf.Energy_YeohSkinFast, poisson=0.5)
    out = []
    for grouping in (xf.GROUPING_CHAINS, xf.GROUPING_ELEMENTS):
        geo = xf.GeoLinear3dCuda(nodes, idx, schedule=xf.SCHEDULE_DATAFLOW, color_hint=hint, grouping=grouping)
        assert (geo.info()["chainedPermille"] == 625) == (grouping == xf.GROUPING_CHAINS)
        geo.Substep(st, DT, 30)
        geo.Substep(st, DT, 7)
        out.append(geo.get_state())
        geo.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)
    assert np.isfinite(out[0][0]).all()


@pytest.mark.parametrize("schedule", SCHEDULES)
def test_exact_against_unmodified_reference(schedule):
    if not ob.have_ref():
        pytest.skip("oracle/_ref not present")
    nodes, idx, hint = xf.GenerateTetBlock(8, 8, wonkiness=0.25)
    geo = xf.GeoLinear3dCuda(nodes, idx, schedule=schedule, color_hint=hint)
    ref = ob.RefScene.mesh(nodes, idx)
    ref.set_order(geo.get_order())
    st, ost = settings_pair(energy=xf.Energy_YeohSkinFast, poisson=0.5)
    geo.Substep(st, DT, 100)
    ref.substep(ost, DT, 100)
    assert_bit_exact(geo, ref)
    assert geo.CalculateVolume() == ref.volume()


@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("rayleigh,sim,energy", list(itertools.product([0, 1, 2, 3], [True, False], [4, 5, 7])))
def test_exact_damping_variants(rayleigh, sim, energy, schedule):
    geo, orc = make_pair(5, 2, 0.25, schedule=schedule)
    kw = dict(energy=energy, simultaneous=sim, poisson=0.495, damping=0.005, rayleigh=rayleigh, pbd_damping=0.03, drag_tc=0.0007)
    st, ost = settings_pair(**kw)
    for s in (st, ost):
        s.volumeAndTimeCorrectedPbdDamping = 1e-6
        s.amortizedVolumeAndTimeCorrectedPbdDamping = 7e-6
    for n in (3, 21):
        geo.Substep(st, DT, n)
        orc.substep(ost, DT, n)
        st.tickId += n   # Sim::Update advances tickId per substep (Demo.cpp:81, 89)
        ost.tickId += n
        assert_bit_exact(geo, orc)
    if schedule == xf.SCHEDULE_DATAFLOW:
        # post-solve damping sweeps stay on the barrier-free schedule; only in-constraint damping (Paper / Limit) needs grid barriers
        assert geo.info()["lastKernel"] == ("k_substeps_dataflow_general" if rayleigh >= 2 else "k_substeps_persistent")


@pytest.mark.parametrize("rayleigh", [2, 3])
@pytest.mark.parametrize("pbd,volume_passes", [(0.03, 0), (0.0, 0), (0.03, 2)])
def test_exact_damped_default_scene_barrier_free_24k_tets(rayleigh, pbd, volume_passes):
    """The web demo's default damping (wasm/ui.js:76-88: Rayleigh_PostAmortized 0.005, pbdDamping 0.03, drag) at 24 576 tets, over more
    than one amortisation period and across launches, on k_substeps_dataflow_general: V records versioned by write counts."""
    nodes, idx, hint = xf.GenerateTetBlock(16, 16, wonkiness=0.2)
    geo = xf.GeoLinear3dCuda(nodes, idx, schedule=xf.SCHEDULE_DATAFLOW, color_hint=hint)
    orc = ob.OracleScene(nodes, idx)
    orc.set_order(geo.get_order())
    kw = dict(energy=xf.Energy_MixedSel, simultaneous=True, poisson=0.495, damping=0.005, rayleigh=rayleigh, pbd_damping=pbd, drag_tc=0.0007,
              volume_passes=volume_passes)
    st, ost = settings_pair(**kw)
    for s in (st, ost):
        s.volumeAndTimeCorrectedPbdDamping = 1e-6
        s.amortizedVolumeAndTimeCorrectedPbdDamping = 7e-6
    for n in (1, 11, 13):
        geo.Substep(st, DT, n)
        orc.substep(ost, DT, n)
        st.tickId += n
        ost.tickId += n
        assert_bit_exact(geo, orc)
    assert geo.info()["lastKernel"] == "k_substeps_dataflow_general"
    # and back to the plain kernel in the same scene: the tags of the two kernels do not collide
    st2, ost2 = settings_pair(energy=xf.Energy_MixedSel, poisson=0.495)
    geo.Substep(st2, DT, 5)
    orc.substep(ost2, DT, 5)
    assert_bit_exact(geo, orc)
    assert geo.info()["lastKernel"] == "k_substeps_dataflow"


@pytest.mark.parametrize("schedule", SCHEDULES)
def test_exact_volume_passes_lock_right_manipulator(schedule):
    geo, orc = make_pair(6, 2, 0.2, schedule=schedule)
    st, ost = settings_pair(energy=xf.Energy_MixedSel, poisson=0.5, lock_right=True, volume_passes=2)
    for s in (st, ost):
        s.lockedRightTransform3d[0] = 0.9
        s.lockedRightTransform3d[1] = 0.1
        s.lockedRightTransform3d[4] = -0.1
    mg, mo = xf.Manipulator(), ob.Manipulator()
    for m in (mg, mo):
        m.pos[:] = (0.0, 0.0, 0.3)
        m.manipPlaneNormal[:] = (0.0, 0.0, 1.0)
        m.pick0[:] = (0.01, 0.0, 0.0)
        m.pickDirTarget[:] = (0.02, 0.05, -1.0)
        m.picked = 1
        m.pickedPointIdx = geo.nV // 2
    geo.Substep(st, DT, 20, manip=mg)
    orc.substep(ost, DT, 20, manip=mo)
    assert_bit_exact(geo, orc)
    if schedule == xf.SCHEDULE_DATAFLOW:
        assert geo.info()["lastKernel"] == "k_substeps_dataflow_general"


@pytest.mark.parametrize("schedule", SCHEDULES)
def test_exact_ground_and_handles(schedule):
    geo, orc = make_pair(4, 4, 0.2, schedule=schedule)
    st, ost = settings_pair(poisson=0.5, lock_left=False, gravity=(0.0, -9.81))
    y0 = float(orc.get_state()[0][:, 1].min()) - 1e-4
    idxs = np.array([3, geo.nV - 1], dtype=np.uint32)
    tg = np.array([[0.0, 0.1, 0.0], [0.05, 0.05, 0.05]], dtype=np.float32)
    for s in (geo, orc):
        s.set_ground(True, y0, 0.25)
        s.set_handles(idxs, tg)
    geo.Substep(st, DT, 60)
    orc.substep(ost, DT, 60)
    assert_bit_exact(geo, orc)
    assert geo.get_state()[0][:, 1].min() >= y0


def test_exact_transform_then_run():
    geo, orc = make_pair(4, 2, 0.1)
    m = np.array([0.0, -1.0, 0.0, 1.0, 0.0, 0.0, 0.013, -0.02, 1.0], dtype=np.float32)
    geo.Transform(m)
    orc.transform(m)
    assert_bit_exact(geo, orc)
    assert np.array_equal(geo.get_origin(), orc.get_origin())
    assert geo.CalculateVolume() == orc.volume()
    st, ost = settings_pair(lock_right=True)
    geo.Substep(st, DT, 10)
    orc.substep(ost, DT, 10)
    assert_bit_exact(geo, orc)


@pytest.mark.parametrize("pattern", [0, 1])
def test_exact_generic_coloring_and_mirrored_pattern(pattern):
    geo, orc = make_pair(7, 4, 0.3, pattern=pattern, use_hint=False)
    st, ost = settings_pair(energy=xf.Energy_Mixed, poisson=0.5)
    geo.Substep(st, DT, 25)
    orc.substep(ost, DT, 25)
    assert_bit_exact(geo, orc)


def test_exact_armadillo():
    if not ob.have_ref():
        pytest.skip("needs the reference's embedded Armadillo tables")
    ref = ob.RefScene.armadillo()
    nodes, idx = ref.get_mesh()
    geo = xf.GeoLinear3dCuda(nodes, idx, density=2.0, auto_resize=True)
    ref.set_order(geo.get_order())
    st, ost = settings_pair(energy=xf.Energy_YeohSkinFast, poisson=0.5, compliance=3.2, gravity=(0.0, -0.602), lock_left=False)
    geo.Substep(st, DT, 50)
    ref.substep(ost, DT, 50)
    assert_bit_exact(geo, ref)


def test_state_roundtrip_and_teacher_forcing():
    geo, orc = make_pair(5, 3, 0.2)
    st, ost = settings_pair(energy=xf.Energy_YeohSkin, poisson=0.4999)
    orc.substep(ost, DT, 40)
    X, V, w = orc.get_state()
    geo.set_state(X, V, w)
    X2, V2, w2 = geo.get_state()
    assert np.array_equal(X, X2) and np.array_equal(V, V2) and np.array_equal(w, w2)
    geo.Substep(st, DT, 1)
    orc.substep(ost, DT, 1)
    assert_bit_exact(geo, orc)


@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("energy,sim,nu", list(itertools.product([3, 4, 5, 7], [True, False], [0.45, 0.495, 0.4999])))
def test_fast_precision_within_1e5_bbox_per_substep(energy, sim, nu, schedule):
    """north_star: per-substep positions within 1e-5 x bbox.  Teacher-forced: both sides restart every substep
    from the oracle's state, at three points of a trajectory."""
    geo, orc = make_pair(precision=xf.PRECISION_FAST, schedule=schedule)
    st, ost = settings_pair(energy=energy, simultaneous=sim, poisson=nu)
    worst = 0.0
    for warm in (0, 20, 60):
        orc.substep(ost, DT, warm)
        X, V, w = orc.get_state()
        geo.set_state(X, V, w)
        geo.Substep(st, DT, 1)
        orc.substep(ost, DT, 1)
        worst = max(worst, rel_err(geo, orc))
    assert worst < 1e-5, worst


def test_fast_precision_free_running_100_substeps():
    geo, orc = make_pair(8, 4, 0.2, precision=xf.PRECISION_FAST)
    st, ost = settings_pair(energy=xf.Energy_MixedSel, poisson=0.495)
    geo.Substep(st, DT, 100)
    orc.substep(ost, DT, 100)
    assert rel_err(geo, orc) < 1e-5


def test_unsupported_energy_is_an_error_not_a_fallback():
    geo, _ = make_pair(3, 2, 0.0)
    st = xf.make_settings(energy=0)  # Energy_Pixar
    with pytest.raises(xf.XfError) as e:
        geo.Substep(st, DT, 1)
    assert e.value.status == xf.XF_ERR_UNSUPPORTED


def test_device_stats_match_oracle():
    geo, orc = make_pair(6, 4, 0.2)
    st, ost = settings_pair(energy=xf.Energy_MixedSel, poisson=0.495)
    geo.Substep(st, DT, 50)
    orc.substep(ost, DT, 50)
    s = geo.stats(st)
    ke, pe, ed, ev = orc.energy(ost)
    assert s["nonfinite"] == 0
    assert abs(s["kinetic"] - ke) <= 1e-9 * max(1.0, abs(ke)) + 1e-12 * abs(ke) + 1e-18
    assert abs(s["gravitational"] - pe) <= 1e-9 * abs(pe) + 1e-18
    assert abs(s["deviatoric"] - ed) <= 1e-6 * abs(ed) + 1e-12
    assert abs(s["volumetric"] - ev) <= 1e-6 * abs(ev) + 1e-12
    assert abs(s["volume"] - orc.volume()) <= 1e-5 * orc.volume()


@pytest.mark.parametrize("dims", [(1, 1), (12, 1), (2, 1)])
def test_exact_degenerate_shapes(dims):
    """Shape_Single (one hex = 6 tets) and Shape_Line (12x1x1), Demo.cpp:290-291: fewer elements than threads in a warp,
    colours with a single element, vertices of low valence."""
    geo, orc = make_pair(dims[0], dims[1], 0.1)
    assert geo.nT == 6 * dims[0] * dims[1] * dims[1]
    st, ost = settings_pair(energy=xf.Energy_YeohSkin, simultaneous=False, poisson=0.5, lock_left=False)
    geo.Substep(st, DT, 25)
    orc.substep(ost, DT, 25)
    assert_bit_exact(geo, orc)


def test_zero_substeps_and_empty_mesh():
    geo, orc = make_pair(3, 2, 0.0)
    geo.Substep(xf.make_settings(), DT, 0)  # n == 0 is a no-op
    assert_bit_exact(geo, orc)
    nodes, idx, _ = xf.GenerateTetBlock(2, 2)
    with pytest.raises(xf.XfError) as e:
        xf.GeoLinear3dCuda(nodes, idx[:0])
    assert e.value.status == xf.XF_ERR_INVALID


def test_async_state_transfers_roundtrip():
    """xf_set_state_async / xf_get_state_async (the per-frame e2e path of bench.py) against the synchronous getters."""
    geo, orc = make_pair(5, 3, 0.2)
    st, ost = settings_pair(energy=xf.Energy_MixedSel, poisson=0.495)
    orc.substep(ost, DT, 15)
    Xo, Vo, wo = orc.get_state()
    geo.set_state(w=wo)
    hX, hV = np.ascontiguousarray(Xo), np.ascontiguousarray(Vo)
    geo.set_state_async(hX.ctypes.data, hV.ctypes.data)
    geo.Substep(st, DT, 4)
    oX, oV = np.empty_like(hX), np.empty_like(hV)
    geo.get_state_async(oX.ctypes.data, oV.ctypes.data)
    geo.Sync()
    orc.substep(ost, DT, 4)
    Xr, Vr, _ = orc.get_state()
    assert np.array_equal(oX, Xr) and np.array_equal(oV, Vr)
