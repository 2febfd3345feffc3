"""GPU check of the warp-cooperative element solve (four lanes per element, xf_element_coop.cuh, measured by xf_debug_coop_element):
the device instantiation must agree bit for bit with the one-thread solve of the stepping kernels AND with the host emulation of
the same header (tests/coop_emu, itself checked against the reference in tests/test_coop_emu.py)."""
import ctypes as C

import numpy as np
import pytest

from __graft_entry__ import build, load_package

build()
xf = load_package()
pytestmark = pytest.mark.gpu


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("energy", [xf.Energy_MixedSel, xf.Energy_YeohSkinFast])
@pytest.mark.parametrize("poisson", [0.5, 0.45])
def test_four_lane_solve_on_the_device_matches_one_thread_solve_and_host_emulation(coop_emu, energy, poisson):
    nodes, idx, hint = xf.GenerateTetBlock(6, 6, wonkiness=0.3)
    geo = xf.GeoLinear3dCuda(nodes, idx, device=-1, color_hint=hint)
    el = geo.get_elements()
    X0, _, w = geo.get_state()
    rng = np.random.default_rng(99)
    h = float(np.abs(np.diff(np.unique(np.round(X0[:, 0], 6)))).min())
    X = X0 + rng.uniform(-0.12 * h, 0.12 * h, size=X0.shape)
    consts, Xg, wg = xf.gathered_elements(el, X, w)
    p4 = xf.substep_constants(1.0, poisson, 1.0 / 3000.0)
    n = consts.shape[0]
    # variant bit 0: every lane gathered vertex 3 itself (the probe gathers once: comparable for one solve only);
    # bit 1: the one-thread side runs the scalar arithmetic instead of the two-wide one
    for variant, iterations in ((0, 1), (0, 25), (2, 1), (2, 25), (1, 1), (3, 1)):
        r = xf.coop_element_probe(consts, Xg, wg, p4, energy=energy, iterations=iterations, warps_per_sm=2, variant=variant)
        assert r["compared"] == 12 * n
        assert not np.array_equal(r["x_single"], Xg)
        assert r["mismatched"] == 0 and np.array_equal(r["x_coop"], r["x_single"]), (variant, iterations, r["mismatched"])
    # host emulation of the same header on the same (independent) elements, one solve each
    r = xf.coop_element_probe(consts, Xg, wg, p4, energy=energy, iterations=1, warps_per_sm=1)
    Xe = np.ascontiguousarray(Xg.reshape(-1, 3).copy())
    own = np.arange(4 * n, dtype=np.uint32).reshape(n, 4)
    order = np.arange(n, dtype=np.uint32)
    rc = coop_emu.coop_emu_sweep(int(energy), 0, _ptr(own), _ptr(el["Qi"]), _ptr(el["QQ"]), _ptr(el["QR"]), _ptr(el["volume"]),
                                 p4[0], p4[1], p4[2], p4[3], _ptr(Xe), _ptr(np.ascontiguousarray(wg.reshape(-1))), _ptr(order), n)
    assert rc == 0
    assert np.array_equal(Xe.reshape(n, 12), r["x_coop"])
    assert r["cycles_single"] > 0 and r["cycles_coop"] > 0 and r["solves_per_s_coop"] > 0
    geo.close()
