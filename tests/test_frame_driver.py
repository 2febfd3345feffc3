"""Frame driver: xf_frame_update restates Sim::Update (Demo.cpp:37-103).

CPU: the bookkeeping (substep count per frame, residual time, time-corrected damping / drag constants, animated
right-lock transform, manipulator ray lerp, tickId) is compared bit for bit with the reference's own Sim over a
sequence of irregular frames - no device needed (scene = NULL steps nothing).
GPU: the same frames actually stepped on the device vs the reference's Sim::Update driving its GeoLinear3d."""
import ctypes as C

import numpy as np
import pytest

from __graft_entry__ import build, load_package
from oracle import bindings as ob

build()
xf = load_package()
pytestmark = pytest.mark.skipif(not ob.have_ref("strict"), reason="oracle/_ref not built (no /root/reference here)")

ROTATE90, ROTATE_LOCK = 1 << 24, 1 << 28
FRAMES = [(1 / 60, 1 / 60), (1 / 55, 1 / 60), (0.031, 1 / 60), (1 / 144, 1 / 120), (0.25, 1 / 30), (1 / 60, 1 / 60), (0.0009, 1 / 60)]
AUTO_FIELDS = ["areaAndTimeCorrectedPbdDamping", "volumeAndTimeCorrectedPbdDamping", "amortizedAreaAndTimeCorrectedPbdDamping",
               "amortizedVolumeAndTimeCorrectedPbdDamping", "timeCorrectedDrag", "tickId"]


def make_pair_settings(extra_flags=0, **kw):
    a, b = xf.make_settings(**kw), ob.make_settings(**kw)
    for s in (a, b):
        s.flags |= extra_flags
        s.substepsPerSecond = 3000.0
        s.drag = 0.002
        s.pbdDamping = 0.03
        s.leftRightSeparation = 1.0
    return a, b


def bits(x):
    return np.frombuffer(bytes(x), dtype=np.uint8)


def picked_manip(cls, idx):
    m = cls()
    m.pos[:] = (0.0, 0.0, 0.3)
    m.manipPlaneNormal[:] = (0.0, 0.0, 1.0)
    m.pick0[:] = (0.01, 0.0, 0.0)
    m.pickDir[:] = (0.02, 0.05, -1.0)
    m.pickDirOld[:] = (0.0, 0.0, -1.0)
    m.picked = 1
    m.pickedPointIdx = idx
    return m


@pytest.mark.parametrize("case", ["plain", "lock_right_rotating", "picked"])
def test_frame_bookkeeping_matches_reference_sim(case):
    nodes, idx, _ = xf.GenerateTetBlock(3, 2)
    flags = {"plain": 0, "lock_right_rotating": ROTATE_LOCK | ROTATE90, "picked": ROTATE90}[case]
    kw = dict(energy=4, poisson=0.495, damping=0.005, rayleigh=3, pbd_damping=0.03, lock_right=(case == "lock_right_rotating"))
    sx, so = make_pair_settings(flags, **kw)
    sim = ob.RefSim(nodes, idx, so)
    state = xf.new_frame_state()
    mx = picked_manip(xf.Manipulator, 5) if case == "picked" else None
    mo = picked_manip(ob.Manipulator, 5) if case == "picked" else None
    for k, (dt, med) in enumerate(FRAMES):
        if case == "lock_right_rotating":
            sx.leftRightSeparation = so.leftRightSeparation = 1.0 - 0.05 * k
        n_ref = sim.update(so, np.float32(dt), np.float32(med), manip=mo)
        n = C.c_uint32()
        rc = xf.lib().xf_frame_update(None, C.byref(sx), C.byref(mx) if mx is not None else None, np.float32(dt), np.float32(med), C.byref(state),
                                      C.byref(n))
        assert rc == 0
        assert n.value == n_ref, "frame %d" % k
        for f in AUTO_FIELDS:
            assert getattr(sx, f) == getattr(so, f), (k, f)
        if case == "lock_right_rotating" and n_ref:
            assert list(sx.lockedRightTransform) == list(so.lockedRightTransform)
            assert [sx.lockedRightTransform3d[i] for i in (0, 1, 2, 4, 5, 6, 8, 9, 10)] == [so.lockedRightTransform3d[i] for i in (0, 1, 2, 4, 5, 6, 8, 9, 10)]
        if mx is not None and n_ref:
            assert list(mx.pickDirTarget) == list(mo.pickDirTarget)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["plain", "lock_right_rotating", "picked"])
def test_frames_stepped_on_device_match_reference_sim(case):
    nodes, idx, hint = xf.GenerateTetBlock(6, 3, wonkiness=0.2)
    flags = {"plain": ROTATE90, "lock_right_rotating": ROTATE_LOCK, "picked": 0}[case]
    kw = dict(energy=7, poisson=0.5, damping=0.004, rayleigh=3, pbd_damping=0.03, lock_right=(case == "lock_right_rotating"))
    sx, so = make_pair_settings(flags, **kw)
    sim = ob.RefSim(nodes, idx, so)
    geo = xf.GeoLinear3dCuda(nodes, idx, color_hint=hint)
    sim.set_order(geo.get_order())
    if flags & ROTATE90:  # Sim::FinishAddingBlocks, Demo.cpp:157-161: rot = [(0,-1),(1,0)], zero offset for a single geo
        geo.Transform(np.array([0.0, -1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0], dtype=np.float32))
    else:
        geo.Transform(np.array([1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0], dtype=np.float32))
    state = xf.new_frame_state()
    mx = picked_manip(xf.Manipulator, 40) if case == "picked" else None
    mo = picked_manip(ob.Manipulator, 40) if case == "picked" else None
    stepped = False
    for k, (dt, med) in enumerate(FRAMES):
        if case == "lock_right_rotating":
            sx.leftRightSeparation = so.leftRightSeparation = 1.0 - 0.03 * k
        n_ref = sim.update(so, np.float32(dt), np.float32(med), manip=mo)
        before = geo.info()["kernelLaunches"]
        n = geo.FrameUpdate(sx, np.float32(dt), np.float32(med), state, manip=mx)
        assert n == n_ref
        # one launch per frame, also while dragging or animating the lock (per-substep rows on the device, xf_substep_varying);
        # the first stepped frame also evaluates the per-element alpha plane for its (compliance, nu, dt) once
        extra = 1 if (n and not stepped) else 0
        stepped = stepped or n > 0
        assert geo.info()["kernelLaunches"] - before == (1 if n else 0) + extra, "frame %d took %d launches for %d substeps" % (
            k, geo.info()["kernelLaunches"] - before, n)
        Xg, Vg, wg = geo.get_state()
        Xr, Vr, wr = sim.get_state()
        assert np.array_equal(Xg, Xr), "frame %d: max |dX| %.3e" % (k, np.abs(Xg - Xr).max())
        assert np.array_equal(Vg, Vr) and np.array_equal(wg, wr)


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [xf.SCHEDULE_DATAFLOW, xf.SCHEDULE_PERSISTENT, xf.SCHEDULE_LAUNCH_PER_COLOR])
def test_substep_varying_equals_single_substeps(schedule):
    """xf_substep_varying (one launch, per-substep lock transform + manipulator ray) against the same substeps issued one by one
    on the oracle, with damping sweeps (k_substeps_dataflow_general) and without."""
    nodes, idx, hint = xf.GenerateTetBlock(6, 3, wonkiness=0.2)
    for damped in (False, True):
        geo = xf.GeoLinear3dCuda(nodes, idx, color_hint=hint, schedule=schedule)
        orc = ob.OracleScene(nodes, idx)
        orc.set_order(geo.get_order())
        kw = dict(energy=7, poisson=0.5, lock_right=True)
        if damped:
            kw.update(damping=0.004, rayleigh=3, pbd_damping=0.03)
        sx, so = xf.make_settings(**kw), ob.make_settings(**kw)
        for s in (sx, so):
            s.volumeAndTimeCorrectedPbdDamping = 1e-6
            s.amortizedVolumeAndTimeCorrectedPbdDamping = 7e-6
        mx, mo = picked_manip(xf.Manipulator, 40), picked_manip(ob.Manipulator, 40)
        n = 13
        rng = np.random.default_rng(5)
        lock = np.zeros((n, 12), dtype=np.float32)
        dirs = np.zeros((n, 3), dtype=np.float32)
        for k in range(n):
            c, s_ = np.float32(np.cos(0.01 * k)), np.float32(np.sin(0.01 * k))
            lock[k, [0, 1, 4, 5, 10]] = (c * np.float32(0.95), s_ * np.float32(0.95), -s_, c, 1.0)
            dirs[k] = (0.02 * k / n, 0.05 * k / n, -1.0)
        dirs += rng.normal(scale=1e-3, size=dirs.shape).astype(np.float32)
        for k in range(n):
            so.lockedRightTransform3d[:] = lock[k].tolist()
            mo.pickDirTarget[:] = dirs[k].tolist()
            orc.substep(so, np.float32(1 / 3000), 1, manip=mo)
            so.tickId += 1
        geo.SubstepVarying(sx, np.float32(1 / 3000), lock, dirs, manip=mx)
        Xg, Vg, wg = geo.get_state()
        Xo, Vo, wo = orc.get_state()
        assert np.array_equal(Xg, Xo) and np.array_equal(Vg, Vo) and np.array_equal(wg, wo), "damped=%s" % damped
        geo.close()
