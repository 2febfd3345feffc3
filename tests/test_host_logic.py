"""CPU-only tests of the product's host side through the C ABI: the library loads and exports every
declared symbol, element init / masses / flags are bit-identical to the reference, the colouring is
conflict-free, and errors are reported the way the header promises.  No device compute is invoked:
scenes are created host-only (device = -1), which can be inspected but never stepped."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from __graft_entry__ import ROOT, build, load_package
from oracle import bindings as ob

build()
xf = load_package()


def host_scene(nodes, idx, **kw):
    return xf.GeoLinear3dCuda(nodes, idx, device=-1, **kw)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "xpbd_fem_b200.h")).read()
    declared = set(re.findall(r"^(?:int|uint32_t|void|float|const char\*|xf_scene\*)\s+(xf_[a-z_0-9]+)\(", header, flags=re.M))
    assert len(declared) >= 20
    L = C.CDLL(xf.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(xf.EXPORTS)


def test_settings_pod_layout_matches_reference_offsets():
    # offsets probed from the reference's own struct (Settings.h:79-102): see SURVEY §8a a1
    S = xf.Settings
    assert C.sizeof(S) == 160
    expect = dict(timeScale=0, substepsPerSecond=4, volumePasses=8, gravity=16, compliance=24, damping=28, pbdDamping=32, drag=36,
                  poissonsRatio=40, wonkiness=44, leftRightSeparation=48, flags=52, areaAndTimeCorrectedPbdDamping=56,
                  volumeAndTimeCorrectedPbdDamping=60, amortizedAreaAndTimeCorrectedPbdDamping=64,
                  amortizedVolumeAndTimeCorrectedPbdDamping=68, timeCorrectedDrag=72, lockedRightTransform=80,
                  lockedRightTransform3d=96, tickId=144)
    for k, off in expect.items():
        assert getattr(S, k).offset == off, k
        assert getattr(ob.Settings, k).offset == off, k


@pytest.mark.parametrize("wonk,pattern", [(0.0, 0), (0.35, 0), (0.2, 1)])
def test_product_meshgen_matches_oracle_and_reference(wonk, pattern):
    nodes, idx, hint = xf.GenerateTetBlock(5, 3, wonkiness=wonk, pattern=pattern)
    n2, i2 = ob.generate_tet_block(5, 3, wonkiness=wonk, pattern=pattern)
    assert np.array_equal(nodes, n2) and np.array_equal(idx, i2)
    if ob.have_ref():
        r = ob.RefScene.block(5, 3, wonkiness=wonk, pattern=pattern)
        rn, ri = r.get_mesh()
        assert np.array_equal(nodes, rn) and np.array_equal(idx, ri)


@pytest.mark.parametrize("wonk,pattern,density", [(0.0, 0, 1.0), (0.3, 0, 1.0), (0.3, 1, 2.5)])
def test_host_init_bit_identical(wonk, pattern, density):
    nodes, idx, hint = xf.GenerateTetBlock(4, 3, wonkiness=wonk, pattern=pattern)
    g = host_scene(nodes, idx, density=density)
    chk = ob.RefScene.mesh(nodes, idx, density=density) if ob.have_ref() else ob.OracleScene(nodes, idx, density)
    eg, ec = g.get_elements(), chk.get_elements()
    for k in ec:
        assert np.array_equal(eg[k], ec[k]), k
    Xg, Vg, wg = g.get_state()
    Xc, Vc, wc = chk.get_state()
    assert np.array_equal(Xg, Xc) and np.array_equal(wg, wc) and not Vg.any()
    assert np.array_equal(g.get_rest()[2], chk.get_rest()[2])


def test_host_init_armadillo_autoresize():
    if not ob.have_ref():
        pytest.skip("needs the reference's embedded Armadillo tables")
    r = ob.RefScene.armadillo()
    nodes, idx = r.get_mesh()
    g = host_scene(nodes, idx, density=2.0, auto_resize=True)
    eg, er = g.get_elements(), r.get_elements()
    for k in er:
        assert np.array_equal(eg[k], er[k]), k
    assert np.array_equal(g.get_state()[0], r.get_state()[0])
    assert np.array_equal(g.get_state()[2], r.get_state()[2])


def check_coloring(idx_stream, colors, order, n_colors):
    tets = idx_stream.reshape(-1, 5)[:, 1:]
    assert colors.max() + 1 == n_colors
    for c in range(n_colors):
        verts = tets[colors == c].reshape(-1)
        assert len(np.unique(verts)) == len(verts), "colour %d has two elements sharing a vertex" % c
    # order = colour-major, stream-index-minor, a permutation
    assert np.array_equal(np.sort(order), np.arange(len(tets)))
    key = colors[order].astype(np.int64) * len(tets) + order
    assert np.all(np.diff(key) > 0)


@pytest.mark.parametrize("dims,pattern", [((8, 2), 0), ((8, 8), 0), ((7, 5), 1)])
def test_generic_coloring_is_conflict_free(dims, pattern):
    nodes, idx, hint = xf.GenerateTetBlock(*dims, pattern=pattern, wonkiness=0.1)
    g = host_scene(nodes, idx)
    check_coloring(idx, g.get_colors(), g.get_order(), g.nColors)
    assert g.nColors <= (32 if pattern == 0 else 52)


def test_lattice_hint_gives_24_balanced_colors():
    nodes, idx, hint = xf.GenerateTetBlock(8, 8)
    g = host_scene(nodes, idx, color_hint=hint)
    assert g.nColors == 24
    check_coloring(idx, g.get_colors(), g.get_order(), 24)
    info = g.info()
    assert info["minColorSize"] == info["maxColorSize"] == 8 * 8 * 8 * 6 // 24
    # ring order: colours 6k .. 6k+5 are the six tets of the cells of class k, index j of each = the same cell, and consecutive
    # tets share a face that contains the cell diagonal (what the barrier-free kernel's timing and the chained sweep rely on)
    tets = idx.reshape(-1, 5)[:, 1:]
    order = g.get_order().reshape(24, -1)
    assert np.array_equal(hint, 6 * (hint // 6) + np.arange(len(hint)) % 6)
    for c in range(24):
        assert np.array_equal(order[c] // 6, order[6 * (c // 6)] // 6) and np.all(order[c] % 6 == c % 6)
        if c % 6:
            shared = [len(set(a) & set(b)) for a, b in zip(tets[order[c - 1]], tets[order[c]])]
            assert set(shared) == {3}


@pytest.mark.parametrize("dims,pattern,wonk", [((8, 8), 0, 0.0), ((5, 3), 0, 0.2), ((6, 4), 1, 0.1)])
def test_clustered_coloring(dims, pattern, wonk):
    """Clustered colouring (one thread per cluster on the barrier-free schedule): still an ordinary conflict-free
    colouring, colours in groups of 6 on MeshGen blocks (the six tets of a cell), 8 cluster colours on the uniform lattice,
    and inside a colour group position t of cluster k sits at index k of colour C*6 + t."""
    nodes, idx, hint = xf.GenerateTetBlock(*dims, pattern=pattern, wonkiness=wonk)
    g = host_scene(nodes, idx, grouping=xf.GROUPING_CLUSTERS, color_hint=hint if pattern == 0 else None)
    tets = idx.reshape(-1, 5)[:, 1:]
    colors, order = g.get_colors(), g.get_order()
    assert g.nColors % 6 == 0 and (pattern != 0 or g.nColors == 48)
    for c in range(g.nColors):
        verts = tets[colors == c].reshape(-1)
        assert len(np.unique(verts)) == len(verts)
    assert np.array_equal(np.sort(order), np.arange(len(tets)))
    assert np.all(np.diff(colors[order].astype(np.int64)) >= 0)        # colour-major
    sizes = np.bincount(colors, minlength=g.nColors)
    start = np.concatenate([[0], np.cumsum(sizes)])
    for C in range(g.nColors // 6):
        assert len(set(sizes[6 * C:6 * C + 6])) == 1                    # every cell has six tets
        cells = [order[start[6 * C + t]:start[6 * C + t + 1]] // 6 for t in range(6)]
        for t in range(1, 6):
            assert np.array_equal(cells[0], cells[t])                   # index k of every position = the same cell
        assert np.array_equal(order[start[6 * C]:start[6 * C + 1]] % 6, np.zeros(sizes[6 * C]))  # position 0 = first tet of the cell


@pytest.mark.parametrize("kind", ["lattice_hint", "lattice_generic", "mirrored", "clusters", "armadillo"])
def test_stage_codes_replay(kind):
    """Barrier-free schedule: replay the serial order with one stage tag per vertex.  Every element must find, on each of its
    corners, exactly the tag its record expects (else the device would wait forever), and the vertex phase must find the
    tag of the vertex's last writer."""
    kw = {}
    if kind == "armadillo":
        if not ob.have_ref():
            pytest.skip("needs the reference's embedded Armadillo tables")
        nodes, idx = ob.RefScene.armadillo().get_mesh()
        kw = dict(density=2.0, auto_resize=True)
    else:
        nodes, idx, hint = xf.GenerateTetBlock(7, 4, pattern=1 if kind == "mirrored" else 0, wonkiness=0.1)
        if kind == "lattice_hint":
            kw = dict(color_hint=hint)
        if kind == "clusters":
            kw = dict(grouping=xf.GROUPING_CLUSTERS)
    g = host_scene(nodes, idx, **kw)
    tets = idx.reshape(-1, 5)[:, 1:]
    order, colors = g.get_order(), g.get_colors()
    pred, last = g.stage_codes()
    tag = np.zeros(g.nV, dtype=np.int64)          # 0 = written by the vertex phase
    for k, e in enumerate(order):
        assert np.array_equal(tag[tets[e]], pred[k]), "element %d would wait forever" % e
        tag[tets[e]] = 1 + colors[e]
    assert np.array_equal(tag, last)


@pytest.mark.parametrize("kind", ["lattice_hint", "lattice_generic", "mirrored", "armadillo"])
def test_chain_info_replay(kind):
    """XF_GROUPING_CHAINS: emulate k_substeps_chain on the CPU.  One thread per position, four private slots each, L2 records
    with a stage tag; values are symbolic (a hash of everything that was combined into them).  Every gathered record must carry
    the tag the element expects, every slot read must find the record of the right vertex, and after the last colour the L2
    records must equal those of the plain serial order, all of them tagged by their last writer.  Colouring and serial order
    must be those of XF_GROUPING_ELEMENTS."""
    kw = {}
    if kind == "armadillo":
        if not ob.have_ref():
            pytest.skip("needs the reference's embedded Armadillo tables")
        nodes, idx = ob.RefScene.armadillo().get_mesh()
        kw = dict(density=2.0, auto_resize=True)
    else:
        nodes, idx, hint = xf.GenerateTetBlock(7, 4, pattern=1 if kind == "mirrored" else 0, wonkiness=0.1)
        if kind == "lattice_hint":
            kw = dict(color_hint=hint)
    g = host_scene(nodes, idx, grouping=xf.GROUPING_CHAINS, **kw)
    plain = host_scene(nodes, idx, grouping=xf.GROUPING_ELEMENTS, **kw)
    order, colors = g.get_order(), g.get_colors()
    assert np.array_equal(order, plain.get_order()) and np.array_equal(colors, plain.get_colors())
    assert plain.chain_info()[1] == 0 and not plain.chain_info()[0].any()
    info, permille = g.chain_info()
    assert g.info()["chainedPermille"] == permille
    tets = idx.reshape(-1, 5)[:, 1:].astype(np.int64)
    pred, last = g.stage_codes()
    sizes = np.bincount(colors, minlength=g.nColors)
    start = np.concatenate([[0], np.cumsum(sizes)])
    if kind == "lattice_hint":
        assert permille == 625                                  # 30 of a cell's 48 corner uses never leave the thread
        gathers = sum(bin(int(w) & 0x42108).count("1") for w in info)
        scatters = sum(bin(int(w) & 0x84210).count("1") for w in info)
        assert gathers == scatters == 9 * len(tets) // 6
    if permille == 0:                                           # nothing lines up: the scene runs on the plain kernel
        assert not info.any()
        return

    def combine(e, vals):
        return [hash((int(e), n) + tuple(vals)) for n in range(4)]

    ref = [hash(("rest", v)) for v in range(g.nV)]
    l2 = list(ref)
    tag = np.zeros(g.nV, dtype=np.int64)
    slots = {}                                                  # (position, slot) -> (vertex, value)
    for substep in range(2):
        for k, e in enumerate(order):                           # the serial order
            new = combine(e, [ref[v] for v in tets[e]])
            for n, v in enumerate(tets[e]):
                ref[v] = new[n]
        for c in range(g.nColors):
            for j in reversed(range(sizes[c])):                 # any order inside a colour
                k = start[c] + j
                e, word = order[k], int(info[k])
                vals = []
                for n, v in enumerate(tets[e]):
                    bits = (word >> (5 * n)) & 31
                    assert bits & 7 < 4
                    if bits & 8:
                        assert tag[v] == pred[k][n], "element %d would wait forever" % e
                        vals.append(l2[v])
                    else:
                        held_v, held_val = slots.pop((j, bits & 7))
                        assert held_v == v
                        vals.append(held_val)
                new = combine(e, vals)
                for n, v in enumerate(tets[e]):
                    bits = (word >> (5 * n)) & 31
                    if bits & 16:
                        l2[v], tag[v] = new[n], 1 + c
                    else:
                        assert (j, bits & 7) not in slots
                        slots[(j, bits & 7)] = (v, new[n])
        assert not slots                                        # nothing is kept across the vertex phase
        assert l2 == ref
        assert np.array_equal(tag, last)
        tag[:] = 0                                              # the vertex phase rewrites every record


def test_bad_color_hint_is_rejected():
    nodes, idx, hint = xf.GenerateTetBlock(3, 3)
    bad = hint.copy()
    bad[:] = 0
    with pytest.raises(xf.XfError) as e:
        host_scene(nodes, idx, color_hint=bad)
    assert e.value.status == xf.XF_ERR_COLORING


def test_armadillo_coloring():
    if not ob.have_ref():
        pytest.skip("needs the reference's embedded Armadillo tables")
    r = ob.RefScene.armadillo()
    nodes, idx = r.get_mesh()
    g = host_scene(nodes, idx, density=2.0, auto_resize=True)
    check_coloring(idx, g.get_colors(), g.get_order(), g.nColors)


def test_malformed_streams_are_rejected():
    nodes, idx, _ = xf.GenerateTetBlock(2, 2)
    with pytest.raises(xf.XfError) as e:
        host_scene(nodes, idx[:-1])
    assert e.value.status == xf.XF_ERR_INVALID
    hexrec = idx.copy()
    hexrec[0] = 8  # CON_HEX record: not on the tet path
    with pytest.raises(xf.XfError) as e:
        host_scene(nodes, hexrec)
    assert e.value.status == xf.XF_ERR_UNSUPPORTED
    with pytest.raises(xf.XfError) as e:
        host_scene(nodes[:-3], idx)
    assert e.value.status == xf.XF_ERR_INVALID


def test_no_cpu_compute_path():
    """A host-only scene must refuse to step: the product has no CPU fallback."""
    nodes, idx, _ = xf.GenerateTetBlock(2, 2)
    g = host_scene(nodes, idx)
    with pytest.raises(xf.XfError) as e:
        g.Substep(xf.make_settings(), 1.0 / 3000.0, 1)
    assert e.value.status == xf.XF_ERR_CUDA
    with pytest.raises(xf.XfError):
        g.CalculateVolume()


def test_product_does_not_reference_the_oracle():
    import glob
    for path in glob.glob(os.path.join(ROOT, "xpbd-fem_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cpp", ".cu", ".cuh", ".h", "Makefile")):
            text = open(path, errors="ignore").read()
            assert "xpbd_oracle" not in text and "oracle/" not in text.replace("the oracle", ""), path


@pytest.mark.parametrize("dims,wonk,pattern,hint", [((6, 3), 0.0, 0, True), ((5, 4), 0.3, 1, False)])
def test_damping_sweep_write_counts_replay(dims, wonk, pattern, hint):
    """CPU replay of the count-versioned velocity records (k_substeps_dataflow_general): for every amortisation slice, walking the
    in-slice elements in serial order, each element must find exactly the count it expects on each corner's record in BOTH sweeps
    (expected = sweep * inSlice(v) + rank - below(v)), and after the sweeps every record must hold what the next predict waits for."""
    nodes, idx, h = xf.GenerateTetBlock(*dims, wonkiness=wonk, pattern=pattern)
    geo = host_scene(nodes, idx, color_hint=h if hint else None)
    order = geo.get_order()
    tets = idx.reshape(-1, 5)[:, 1:][order]          # corners at every serial position
    rank, below = geo.damping_codes()
    nT, nV = geo.nT, geo.nV
    valence = np.bincount(tets.reshape(-1), minlength=nV)
    assert np.array_equal(below[:, 7], valence)
    for k in list(range(8)) + [8]:                   # 8 = not amortised: the whole mesh
        lo, hi = (nT * k // 8, nT * (k + 1) // 8) if k < 8 else (0, nT)
        b = (below[:, k - 1].astype(np.int64) if 0 < k < 8 else np.zeros(nV, dtype=np.int64))
        inside = (below[:, k].astype(np.int64) if k < 8 else below[:, 7].astype(np.int64)) - b
        # the codes agree with a direct count
        direct = np.bincount(tets[lo:hi].reshape(-1), minlength=nV)
        assert np.array_equal(inside, direct)
        count = np.zeros(nV, dtype=np.int64)         # what the post phase stores
        for sweep in range(2):
            for pos in range(lo, hi):
                for j in range(4):
                    v = tets[pos, j]
                    assert count[v] == sweep * inside[v] + rank[pos, j] - b[v], (k, sweep, pos, j)
                    count[v] += 1
        assert np.array_equal(count, 2 * inside)     # the next predict's expectation
    geo.close()
