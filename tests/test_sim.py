"""Sim layer: xf_sim_* restates Sim::AddBlock / FinishAddingBlocks / SetGeoOffset / Update (Demo.cpp:37-186) for
several Geos, and xf_block_from_settings / xf_sim_add_block_from_settings the block table of Demo::UpdateSettings (Demo.cpp:289-318).

CPU: for every shape x pattern the block built from a Settings block is compared with what the reference's own
Demo::UpdateSettings builds (rest positions and element indices, bit for bit) - host-only scenes, no device.
GPU: a Sim with two blocks (one of them dragged by the manipulator, right side locked and animated) stepped over irregular frames
against the reference's Sim, bit for bit, one launch per geo and frame."""
import numpy as np
import pytest

from __graft_entry__ import build, load_package
from oracle import bindings as ob

build()
xf = load_package()
pytestmark = pytest.mark.skipif(not ob.have_ref("strict"), reason="oracle/_ref not built (no /root/reference here)")

SHAPES = {"Single": 0, "Line": 1, "BeamL": 2, "BeamM": 3, "BeamH": 4, "BeamL1x2": 5, "BeamL2x1": 6, "BeamL4x1": 7, "BeamL8x1": 8, "BoxL": 9,
          "BoxM": 10, "BoxH": 11}
SHAPE_BIT, PATTERN_BIT = 12, 16
ROTATE90, ROTATE_LOCK = 1 << 24, 1 << 28
FRAMES = [(1 / 60, 1 / 60), (1 / 55, 1 / 60), (0.031, 1 / 60), (1 / 144, 1 / 120), (0.25, 1 / 30), (1 / 60, 1 / 60), (0.0009, 1 / 60)]


def shaped_settings(mod, shape, pattern=0, wonk=0.0, extra=0, **kw):
    s = mod.make_settings(**kw)
    s.flags |= (shape << SHAPE_BIT) | (pattern << PATTERN_BIT) | extra
    s.wonkiness = wonk
    return s


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("pattern,wonk", [(0, 0.0), (1, 0.3)])
def test_block_from_settings_matches_demo_update_settings(shape, pattern, wonk):
    if shape in ("BoxH", "BeamH") and pattern == 1:
        pytest.skip("large blocks once are enough")
    sx, so = (shaped_settings(m, SHAPES[shape], pattern, wonk, energy=4) for m in (xf, ob))
    ref = ob.RefMultiSim(so, from_settings=True)
    assert ref.geo_count() == 1
    X0r, idxr = ref.get_mesh(0)
    sim = xf.SimCuda(device=-1)  # host-only: mesh generation, element init and colouring without a device
    sim.AddBlockFromSettings(sx)
    geo = sim.geo(0)
    X0, O, flags = geo.get_rest()
    el = geo.get_elements()
    assert (geo.nV, geo.nT) == ref.sizes(0)
    assert np.array_equal(X0, X0r)
    assert np.array_equal(el["idx"], idxr)
    w, h, scx, scy, pat = xf.block_from_settings(sx)
    assert geo.nT == 6 * w * h * h and pat == pattern
    sim.close()
    ref.close()


def test_block_from_settings_rejects_what_is_out_of_scope():
    s = shaped_settings(xf, 12, energy=4)  # Shape_Armadillo: the caller's asset
    with pytest.raises(xf.XfError) as e:
        xf.block_from_settings(s)
    assert e.value.status == xf.XF_ERR_UNSUPPORTED
    s = shaped_settings(xf, SHAPES["BoxL"], energy=4)
    s.flags = (s.flags & ~15) | 7  # Element_H8
    with pytest.raises(xf.XfError) as e:
        xf.block_from_settings(s)
    assert e.value.status == xf.XF_ERR_UNSUPPORTED


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


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["two_blocks_picked", "three_blocks_rotating_lock", "from_settings_offset"])
def test_sim_with_several_geos_matches_reference_sim(case):
    extra = {"two_blocks_picked": ROTATE90, "three_blocks_rotating_lock": ROTATE_LOCK, "from_settings_offset": 0}[case]
    kw = dict(energy=7, poisson=0.5, damping=0.004, rayleigh=3, pbd_damping=0.03, lock_right=(case == "three_blocks_rotating_lock"))
    sx, so = (shaped_settings(m, SHAPES["BeamL"], 0, 0.2, extra, **kw) for m in (xf, ob))
    for s in (sx, so):
        s.substepsPerSecond = 3000.0
        s.drag = 0.002
    sim = xf.SimCuda(device=0)
    if case == "from_settings_offset":
        ref = ob.RefMultiSim(so, from_settings=True)
        sim.AddBlockFromSettings(sx)
    else:
        ref = ob.RefMultiSim(so)
        blocks = [xf.GenerateTetBlock(6, 3, wonkiness=0.2), xf.GenerateTetBlock(4, 4, wonkiness=0.1)]
        if case == "three_blocks_rotating_lock":
            blocks.append(xf.GenerateTetBlock(5, 2))
        for nodes, idx, hint in blocks:
            ref.add_block(nodes, idx)
            sim.AddBlock(nodes, idx, color_hint=hint)
        ref.finish()
    sim.FinishAddingBlocks(sx)
    if case == "from_settings_offset":
        ref.set_geo_offset(0.0, 4.65 * 0.2 / 31.0)
        sim.SetGeoOffset(0.0, 4.65 * 0.2 / 31.0)
    n_geo = sim.geo_count()
    assert n_geo == ref.geo_count()
    geos = [sim.geo(i) for i in range(n_geo)]
    for i, g in enumerate(geos):
        ref.set_order(i, g.get_order())
        assert sim.volume0(i) == ref.volume0(i)
        assert np.array_equal(g.get_state()[0], ref.get_state(i)[0])
    picked = 1 if case == "two_blocks_picked" else -1
    mx = picked_manip(xf.Manipulator, 17) if picked >= 0 else None
    mo = picked_manip(ob.Manipulator, 17) if picked >= 0 else None
    stepped = False
    for k, (dt, med) in enumerate(FRAMES):
        if case == "three_blocks_rotating_lock":
            sx.leftRightSeparation = so.leftRightSeparation = 1.0 - 0.03 * k
        n_ref = ref.update(so, np.float32(dt), np.float32(med), manip=mo, picked_geo=picked)
        before = [g.info()["kernelLaunches"] for g in geos]
        n = sim.Update(sx, np.float32(dt), np.float32(med), manip=mx, picked_geo=picked)
        assert n == n_ref
        extra = 1 if (n and not stepped) else 0  # the first stepped frame also evaluates the per-element alpha plane once
        stepped = stepped or n > 0
        for i, g in enumerate(geos):
            assert g.info()["kernelLaunches"] - before[i] == (1 if n else 0) + extra
            Xg, Vg, wg = g.get_state()
            Xr, Vr, wr = ref.get_state(i)
            assert np.array_equal(Xg, Xr), "frame %d geo %d: max |dX| %.3e" % (k, i, np.abs(Xg - Xr).max())
            assert np.array_equal(Vg, Vr) and np.array_equal(wg, wr)
    sim.close()
    ref.close()
