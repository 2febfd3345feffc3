"""GPU tests of the batched-scene path (BASELINE config 3): every scene of a batch must be bit-identical
(XF_PRECISION_EXACT) to the CPU oracle stepping that scene alone with that scene's Settings."""
import ctypes as C

import numpy as np
import pytest

from __graft_entry__ import build, load_package
from oracle import bindings as ob

build()
xf = load_package()
pytestmark = pytest.mark.gpu
DT = np.float32(1.0 / 3000.0)


def jittered_settings(n, energy, sim, nu, **kw):
    """Per-scene gravity / compliance jittered deterministically from the scene index."""
    arr_x = (xf.Settings * n)()
    arr_o = []
    for s in range(n):
        g = (0.02 * ((s * 7) % 5 - 2), -0.4905 * (1.0 + 0.1 * ((s * 3) % 7)))
        comp = 1.0 * (1.0 + 0.25 * (s % 4))
        arr_x[s] = xf.make_settings(energy=energy, simultaneous=sim, poisson=nu, gravity=g, compliance=comp, **kw)
        arr_o.append(ob.make_settings(energy=energy, simultaneous=sim, poisson=nu, gravity=g, compliance=comp, **kw))
    return arr_x, arr_o


@pytest.mark.parametrize("dims,nscenes", [((8, 2), 37), ((4, 4), 9), ((8, 8), 5)])
@pytest.mark.parametrize("energy,sim,nu", [(4, True, 0.5), (7, True, 0.5), (3, False, 0.495), (5, True, 0.4999), (7, False, 0.45)])
def test_batch_scenes_bit_identical_to_oracle(dims, nscenes, energy, sim, nu):
    nodes, idx, hint = xf.GenerateTetBlock(*dims, wonkiness=0.2)
    batch = xf.GeoBatchCuda(nodes, idx, nscenes, color_hint=hint)
    sx, so = jittered_settings(nscenes, energy, sim, nu)
    order = batch.get_order()
    for n in (1, 12):
        batch.Substep(sx, DT, n)
    X, V, w = batch.get_state()
    for s in sorted(set([0, 1, nscenes // 2, nscenes - 1])):
        orc = ob.OracleScene(nodes, idx)
        orc.set_order(order)
        orc.substep(so[s], DT, 13)
        Xo, Vo, wo = orc.get_state()
        assert np.array_equal(X[s], Xo), "scene %d" % s
        assert np.array_equal(V[s], Vo)
        assert np.array_equal(w[s], wo)


def test_batch_damping_volume_passes_and_ground():
    nodes, idx, hint = xf.GenerateTetBlock(5, 2, wonkiness=0.25)
    n = 6
    batch = xf.GeoBatchCuda(nodes, idx, n, color_hint=hint)
    kw = dict(damping=0.005, rayleigh=xf.Rayleigh_PostAmortized, pbd_damping=0.03, drag_tc=0.0007, volume_passes=1, lock_left=False)
    sx, so = jittered_settings(n, 4, True, 0.495, **kw)
    for s in range(n):
        for a in (sx[s], so[s]):
            a.volumeAndTimeCorrectedPbdDamping = 1e-6
            a.amortizedVolumeAndTimeCorrectedPbdDamping = 7e-6
    y0 = float(nodes.reshape(-1, 3)[:, 1].min()) - 2e-4
    batch.set_ground(True, y0, 0.2)
    batch.Substep(sx, DT, 19)
    X, V, w = batch.get_state()
    order = batch.get_order()
    for s in (0, 3, 5):
        orc = ob.OracleScene(nodes, idx)
        orc.set_order(order)
        orc.set_ground(True, y0, 0.2)
        orc.substep(so[s], DT, 19)
        Xo, Vo, wo = orc.get_state()
        assert np.array_equal(X[s], Xo) and np.array_equal(V[s], Vo) and np.array_equal(w[s], wo)


def test_batch_set_state_roundtrip_and_shared_settings():
    nodes, idx, hint = xf.GenerateTetBlock(4, 2)
    batch = xf.GeoBatchCuda(nodes, idx, 4, color_hint=hint)
    X, V, w = batch.get_state()
    X[2] += 1e-3
    batch.set_state(2, X=X[2:3], V=V[2:3], w=w[2:3])
    X2, _, _ = batch.get_state()
    assert np.array_equal(X, X2)
    st = xf.make_settings(energy=7)
    batch.Substep(st, DT, 5)
    Xa, _, _ = batch.get_state()
    assert np.array_equal(Xa[0], Xa[1]) and not np.array_equal(Xa[0], Xa[2])


def test_batch_rejects_mixed_energies_and_oversized_scenes():
    nodes, idx, hint = xf.GenerateTetBlock(3, 2)
    batch = xf.GeoBatchCuda(nodes, idx, 2, color_hint=hint)
    arr = (xf.Settings * 2)()
    arr[0] = xf.make_settings(energy=4)
    arr[1] = xf.make_settings(energy=7)
    with pytest.raises(xf.XfError) as e:
        batch.Substep(arr, DT, 1)
    assert e.value.status == xf.XF_ERR_UNSUPPORTED
    big_nodes, big_idx, big_hint = xf.GenerateTetBlock(20, 20)
    with pytest.raises(xf.XfError) as e:
        xf.GeoBatchCuda(big_nodes, big_idx, 2, color_hint=big_hint)
    assert e.value.status == xf.XF_ERR_UNSUPPORTED
