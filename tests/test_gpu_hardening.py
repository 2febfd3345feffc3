"""GPU tests of what the barrier-free schedules rely on and how they fail (VERDICT r1 "weak" 1 and 6):
  * a 32-byte-aligned 256-bit access is never torn, between SMs of one GPU and across NVLink peer stores;
  * the 24-bit stage tags wrap around without a hiccup (a wrap needs ~13 000 launches in normal use);
  * a schedule that cannot make progress is REPORTED (XF_ERR_CUDA from the next sync / getter) - the device does not hang,
    the CUDA context survives, and other scenes of the process keep working."""
import numpy as np
import pytest

from __graft_entry__ import build, load_package
from oracle import bindings as ob

build()
xf = load_package()
pytestmark = pytest.mark.gpu
DT = np.float32(1.0 / 3000.0)


def test_256bit_records_are_never_torn_between_sms():
    reads, torn = xf.torn_records(device=0, remote_device=-1, n_records=1 << 16, rounds=16000)
    assert reads >= 1_000_000_000
    assert torn == 0, "%d of %d record reads saw words of two different stores" % (torn, reads)


@pytest.mark.skipif(xf.device_count() < 2, reason="needs two GPUs with peer access")
def test_256bit_records_are_never_torn_across_nvlink():
    reads, torn = xf.torn_records(device=0, remote_device=1, n_records=1 << 16, rounds=16000)
    assert reads >= 1_000_000_000
    assert torn == 0, "%d of %d record reads saw words of two different peer stores" % (torn, reads)


@pytest.mark.parametrize("grouping", [xf.GROUPING_AUTO, xf.GROUPING_CHAINS])
def test_stage_tags_wrap_around(grouping):
    """verBase is put three substeps below 2^24; ten launches of 4 substeps cross the wrap inside a launch and between launches."""
    nodes, idx, hint = xf.GenerateTetBlock(6, 5, wonkiness=0.2)
    geo = xf.GeoLinear3dCuda(nodes, idx, device=0, schedule=xf.SCHEDULE_DATAFLOW, color_hint=hint, grouping=grouping)
    stride = geo.nColors + 1
    geo.debug_knob(0, (1 << 24) - 3 * stride)
    st = xf.make_settings(energy=xf.Energy_YeohSkinFast, poisson=0.5)
    orc = ob.OracleScene(nodes, idx)
    orc.set_order(geo.get_order())
    ost = ob.make_settings(energy=ob.Energy_YeohSkinFast, poisson=0.5)
    for launch in range(10):
        geo.Substep(st, DT, 4)
        orc.substep(ost, DT, 4)
        X, V, w = geo.get_state()
        Xo, Vo, wo = orc.get_state()
        assert np.array_equal(X, Xo) and np.array_equal(V, Vo) and np.array_equal(w, wo), "launch %d" % launch
    geo.close()


def test_stage_tags_wrap_around_with_damping_sweeps_and_volume_passes():
    """The same wrap on k_substeps_dataflow_general (stride = 1 + colours * (1 + volume passes) + 1 tags per substep)."""
    nodes, idx, hint = xf.GenerateTetBlock(6, 5, wonkiness=0.2)
    geo = xf.GeoLinear3dCuda(nodes, idx, device=0, schedule=xf.SCHEDULE_DATAFLOW, color_hint=hint)
    stride = 1 + geo.nColors * 2 + 1
    geo.debug_knob(0, (1 << 24) - 3 * stride)
    kw = dict(energy=xf.Energy_YeohSkinFast, poisson=0.495, damping=0.005, rayleigh=3, pbd_damping=0.03, volume_passes=1)
    st, ost = xf.make_settings(**kw), ob.make_settings(**kw)
    for s in (st, ost):
        s.volumeAndTimeCorrectedPbdDamping = 1e-6
        s.amortizedVolumeAndTimeCorrectedPbdDamping = 7e-6
    orc = ob.OracleScene(nodes, idx)
    orc.set_order(geo.get_order())
    for launch in range(6):
        geo.Substep(st, DT, 4)
        orc.substep(ost, DT, 4)
        st.tickId += 4
        ost.tickId += 4
        X, V, w = geo.get_state()
        Xo, Vo, wo = orc.get_state()
        assert np.array_equal(X, Xo) and np.array_equal(V, Vo) and np.array_equal(w, wo), "launch %d" % launch
    assert geo.info()["lastKernel"] == "k_substeps_dataflow_general"
    geo.close()


def test_a_stalled_schedule_is_reported_and_the_context_survives():
    nodes, idx, hint = xf.GenerateTetBlock(6, 5)
    broken = xf.GeoLinear3dCuda(nodes, idx, device=0, schedule=xf.SCHEDULE_DATAFLOW, color_hint=hint)
    healthy = xf.GeoLinear3dCuda(nodes, idx, device=0, schedule=xf.SCHEDULE_DATAFLOW, color_hint=hint)
    st = xf.make_settings(energy=xf.Energy_MixedSel, poisson=0.45)
    broken.debug_knob(1, 1 << 13)   # give up after 8192 polls instead of 2^24
    broken.debug_knob(2, 251)       # vertex 0 now waits for a stage nobody ever writes
    broken.Substep(st, DT, 3)
    with pytest.raises(xf.XfError) as err:
        broken.Sync()
    assert err.value.status == xf.XF_ERR_CUDA and "stalled" in str(err.value)
    with pytest.raises(xf.XfError):
        broken.get_state()          # keeps failing: the state is undefined
    # the context is intact: another scene steps and matches the oracle
    healthy.Substep(st, DT, 6)
    X, V, w = healthy.get_state()
    orc = ob.OracleScene(nodes, idx)
    orc.set_order(healthy.get_order())
    orc.substep(ob.make_settings(energy=ob.Energy_MixedSel, poisson=0.45), DT, 6)
    Xo, Vo, wo = orc.get_state()
    assert np.array_equal(X, Xo) and np.array_equal(V, Vo)
    # a new state ends the report
    broken.set_state(X=Xo, V=Vo)
    broken.Sync()
    broken.close()
    healthy.close()
