"""The warp-cooperative element solve (four lanes per element, xpbd-fem_b200/csrc/xf_element_coop.cuh) checked WITHOUT a GPU:
the same header is compiled for the host (tests/coop_emu/coop_emu.cpp: four threads play the four lanes, shuffles and the quad's
shared-memory row go through a mailbox) and one Gauss-Seidel sweep of it is compared bit for bit with one Constrain() of the
reference (oracle/_ref, else the C oracle) in the same element order.  On the GPU, xf_debug_coop_element compares the device
instantiation of the same header with the one-thread solve (tests/test_gpu_coop.py)."""
import ctypes as C

import numpy as np
import pytest

from __graft_entry__ import build, load_package
from oracle import bindings as ob

build()
xf = load_package()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("energy", [xf.Energy_MixedSel, xf.Energy_YeohSkinFast])
@pytest.mark.parametrize("poisson", [0.5, 0.45])
@pytest.mark.parametrize("x3_gathered", [0, 1])
def test_four_lane_sweep_matches_reference_bit_for_bit(coop_emu, energy, poisson, x3_gathered):
    nodes, idx, hint = xf.GenerateTetBlock(5, 4, wonkiness=0.3)
    geo = xf.GeoLinear3dCuda(nodes, idx, device=-1, color_hint=hint)  # host-only scene: init + colouring, no stepping
    el = geo.get_elements()
    order = np.ascontiguousarray(geo.get_order(), dtype=np.uint32)
    X0, _, w = geo.get_state()
    rng = np.random.default_rng(1234 + int(energy))
    h = float(np.abs(np.diff(np.unique(np.round(X0[:, 0], 6)))).min())
    X = np.ascontiguousarray(X0 + rng.uniform(-0.12 * h, 0.12 * h, size=X0.shape))  # a deformed state: every constraint is active
    V = np.zeros_like(X)
    dt = np.float32(1.0 / 3000.0)
    compliance = 1.0

    # reference: one Substep with V = 0 and no gravity leaves X to Constrain() alone (Geo.cpp:305-331)
    if ob.have_ref("strict"):
        ref = ob.RefScene.mesh(nodes, idx, kind="strict")
    else:
        ref = ob.OracleScene(nodes, idx)
    ref.set_order(order)
    ref.set_state(X=X, V=V)
    st = ob.make_settings(energy=int(energy), simultaneous=True, poisson=poisson, compliance=compliance, gravity=(0.0, 0.0), lock_left=False)
    ref.substep(st, dt, 1)
    Xr, _, wr = ref.get_state()
    assert np.array_equal(wr, w)

    a, inv_mu, inv_lambda, dt2 = xf.substep_constants(compliance, poisson, dt)
    Xe = X.copy()
    rc = coop_emu.coop_emu_sweep(int(energy), x3_gathered, _ptr(el["idx"]), _ptr(el["Qi"]), _ptr(el["QQ"]), _ptr(el["QR"]), _ptr(el["volume"]),
                            a, inv_mu, inv_lambda, dt2, _ptr(Xe), _ptr(np.ascontiguousarray(w, dtype=np.float32)), _ptr(order), order.size)
    assert rc == 0
    assert not np.array_equal(Xe, X)                    # the sweep did something
    assert np.array_equal(Xe, Xr), "max |dX| = %g" % np.abs(Xe - Xr).max()
    geo.close()
