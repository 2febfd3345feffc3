"""Host-side mirror of the reference's ``Geo`` interface over the C ABI (include/xpbd_fem_b200.h).

This module is a thin ctypes binding: it owns no algorithm.  All compute happens in
``libxpbd_fem_b200.so`` (hand-written sm_100a kernels); if the library or a CUDA device is missing the
calls raise — there is no CPU fallback.

Method names follow the reference (``Geo.h:15-35``): ``Substep``, ``Transform``, ``CalculateVolume``,
``VertCount``, ``ElementCount``; the public members ``X, V, w, X0, O, flags, tOrder`` become getters.

The directory name ``xpbd-fem_b200`` is not a valid Python identifier; load it with
``importlib`` (see ``__graft_entry__.load_package``) under the module name ``xpbd_fem_b200``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XF_LIB_OVERRIDE") or os.path.join(_HERE, "libxpbd_fem_b200.so")  # override: kernel-variant experiments only

XF_ABI_VERSION = 2
XF_OK, XF_ERR_INVALID, XF_ERR_CUDA, XF_ERR_UNSUPPORTED, XF_ERR_NOMEM, XF_ERR_COLORING = 0, -1, -2, -3, -4, -5
PARTITION_SLABS, PARTITION_GRAPH = 0, 1
PRECISION_EXACT, PRECISION_FAST = 0, 1
GROUPING_AUTO, GROUPING_ELEMENTS, GROUPING_CLUSTERS, GROUPING_CHAINS = 0, 1, 2, 3
SCHEDULE_AUTO, SCHEDULE_LAUNCH_PER_COLOR, SCHEDULE_PERSISTENT, SCHEDULE_BRICKS, SCHEDULE_DATAFLOW = 0, 1, 2, 3, 4

# flag word (Settings.h:9-75)
Settings_EnergyBit = 6
Settings_XpbdSolveBit = 11
Settings_RayleighTypeBit = 20
Settings_LockLeft = 1 << 26
Settings_LockRight = 1 << 27
Element_T4 = 5
Energy_Mixed, Energy_MixedSel, Energy_YeohSkin, Energy_YeohSkinFast = 3, 4, 5, 7
Pattern_Uniform, Pattern_Mirrored = 0, 1
Rayleigh_Paper, Rayleigh_Limit, Rayleigh_Post, Rayleigh_PostAmortized = 0, 1, 2, 3


class XfError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("xf error %d: %s" % (status, message))
        self.status = status


class Settings(C.Structure):
    """xf_settings == the reference's Settings POD (Settings.h:79-102), 160 bytes."""
    _fields_ = [
        ("timeScale", C.c_float), ("substepsPerSecond", C.c_float), ("volumePasses", C.c_uint32), ("_pad0", C.c_uint32),
        ("gravity", C.c_float * 2), ("compliance", C.c_float), ("damping", C.c_float), ("pbdDamping", C.c_float),
        ("drag", C.c_float), ("poissonsRatio", C.c_float), ("wonkiness", C.c_float), ("leftRightSeparation", C.c_float),
        ("flags", C.c_uint32),
        ("areaAndTimeCorrectedPbdDamping", C.c_float), ("volumeAndTimeCorrectedPbdDamping", C.c_float),
        ("amortizedAreaAndTimeCorrectedPbdDamping", C.c_float), ("amortizedVolumeAndTimeCorrectedPbdDamping", C.c_float),
        ("timeCorrectedDrag", C.c_float), ("_pad1", C.c_uint32),
        ("lockedRightTransform", C.c_float * 4), ("lockedRightTransform3d", C.c_float * 12),
        ("tickId", C.c_uint32), ("_pad2", C.c_uint32 * 3),
    ]


class Manipulator(C.Structure):
    _fields_ = [
        ("pos", C.c_float * 3), ("manipPlaneNormal", C.c_float * 3), ("pick0", C.c_float * 3), ("pickDir", C.c_float * 3),
        ("pickDirOld", C.c_float * 3), ("pickDirTarget", C.c_float * 3), ("picked", C.c_int32), ("pickedPointIdx", C.c_uint32),
    ]


class CreateParams(C.Structure):
    _fields_ = [
        ("abiVersion", C.c_uint32), ("device", C.c_int32), ("density", C.c_float), ("autoResize", C.c_int32),
        ("precision", C.c_int32), ("schedule", C.c_int32), ("stream", C.c_void_p), ("colorHint", C.c_void_p),
        ("colorHintCount", C.c_uint32), ("grouping", C.c_uint32), ("partition", C.c_uint32), ("_reserved", C.c_uint32 * 3),
    ]


class FrameState(C.Structure):
    """Sim's persistent per-simulation scalars (Demo.h:40-44)."""
    _fields_ = [("dtResidual", C.c_float), ("tickId", C.c_uint32), ("leftRightSeparationOld", C.c_float), ("rightRotationTheta", C.c_float)]


class Info(C.Structure):
    _fields_ = [
        ("vertCount", C.c_uint32), ("elementCount", C.c_uint32), ("colorCount", C.c_uint32), ("minColorSize", C.c_uint32),
        ("maxColorSize", C.c_uint32), ("smCount", C.c_uint32), ("gridBlocks", C.c_uint32), ("blockThreads", C.c_uint32),
        ("elementRecordBytes", C.c_uint32), ("schedule", C.c_uint32), ("kernelLaunches", C.c_uint64), ("l2Bytes", C.c_uint64),
        ("chainedPermille", C.c_uint32), ("lastKernelId", C.c_uint32),
    ]


KERNEL_NAMES = {0: None, 1: "k_substeps_dataflow", 2: "k_substeps_chain", 3: "k_substeps_cluster", 4: "k_substeps_persistent",
                5: "k_substeps_bricks", 6: "k_sweep_color (x colours)", 7: "k_substeps_dataflow_general"}


def frame_constants(settings, state):
    """Sim::Update's per-frame derived constants (time-corrected drag / PBD damping, Demo.cpp:51-63) written into `settings`,
    without stepping anything (xf_frame_update with a NULL scene and dt = 0)."""
    n = C.c_uint32(0)
    _check(lib().xf_frame_update(None, C.byref(settings), None, 0.0, 1.0 / 60.0, C.byref(state), C.byref(n)))
    return settings


def new_frame_state():
    st = FrameState()
    lib().xf_frame_state_init(C.byref(st))
    return st


def make_settings(energy=Energy_MixedSel, simultaneous=True, poisson=0.5, compliance=1.0, gravity=(0.0, -0.4905), damping=0.0,
                  rayleigh=Rayleigh_Post, lock_left=True, lock_right=False, drag_tc=0.0, volume_passes=0, pbd_damping=0.0,
                  substeps_per_second=3000.0):
    """A Settings block with the web demo's defaults (wasm/ui.js:69-91) unless overridden."""
    s = Settings()
    s.timeScale = 1.0
    s.substepsPerSecond = substeps_per_second
    s.volumePasses = volume_passes
    s.gravity[0], s.gravity[1] = gravity
    s.compliance = compliance
    s.damping = damping
    s.pbdDamping = pbd_damping
    s.poissonsRatio = poisson
    s.leftRightSeparation = 1.0
    s.flags = (Element_T4 | (energy << Settings_EnergyBit) | ((1 if simultaneous else 0) << Settings_XpbdSolveBit)
               | (rayleigh << Settings_RayleighTypeBit) | (Settings_LockLeft if lock_left else 0)
               | (Settings_LockRight if lock_right else 0))
    s.timeCorrectedDrag = drag_tc
    s.lockedRightTransform[0] = s.lockedRightTransform[3] = 1.0
    s.lockedRightTransform3d[0] = s.lockedRightTransform3d[5] = s.lockedRightTransform3d[10] = 1.0
    return s


_lib = None

EXPORTS = [
    "xf_last_error", "xf_device_count", "xf_default_create_params", "xf_generate_tet_block", "xf_create", "xf_destroy",
    "xf_vert_count", "xf_element_count", "xf_color_count", "xf_get_order", "xf_get_colors", "xf_get_stage_codes", "xf_get_chain_info", "xf_get_damping_codes", "xf_get_elements", "xf_substep",
    "xf_substep_varying", "xf_sync", "xf_set_ground", "xf_set_handles", "xf_get_state", "xf_set_state", "xf_get_rest", "xf_get_origin",
    "xf_get_state_async", "xf_set_state_async", "xf_transform", "xf_volume", "xf_stats", "xf_get_info",
    "xf_frame_state_init", "xf_frame_update",
    "xf_sim_create", "xf_sim_destroy", "xf_sim_reset", "xf_sim_add_block", "xf_block_from_settings", "xf_sim_add_block_from_settings",
    "xf_sim_finish_adding_blocks", "xf_sim_set_geo_offset", "xf_sim_update", "xf_sim_geo_count", "xf_sim_geo", "xf_sim_volume0",
    "xf_sim_get_frame_state",
    "xf_batch_create", "xf_batch_destroy", "xf_batch_scene_count", "xf_batch_vert_count", "xf_batch_element_count",
    "xf_batch_color_count", "xf_batch_get_order", "xf_batch_set_ground", "xf_batch_substep", "xf_batch_sync",
    "xf_batch_get_state", "xf_batch_set_state", "xf_batch_get_info",
    "xf_part_create", "xf_part_destroy", "xf_part_local_vert_count", "xf_part_local_element_count", "xf_part_peer_count",
    "xf_part_color_count", "xf_part_global_vert_count", "xf_part_global_element_count", "xf_part_get_local_verts",
    "xf_part_get_local_elements", "xf_part_get_peers", "xf_part_get_halo", "xf_part_get_order", "xf_part_get_global_color_start",
    "xf_part_get_initial", "xf_part_get_dataflow_codes", "xf_part_ipc_export", "xf_part_ipc_connect", "xf_part_set_ground", "xf_part_substep", "xf_part_sync",
    "xf_part_get_state", "xf_part_get_info",
    "xf_debug_l2_bandwidth", "xf_debug_stage_latency", "xf_debug_torn_records", "xf_debug_scene_knob", "xf_debug_barrier_us",
    "xf_debug_coop_element",
]


def lib():
    """Load libxpbd_fem_b200.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("%s is missing: run `make -C xpbd-fem_b200` (or __graft_entry__.build())" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32, f32, i32 = C.c_void_p, C.c_uint32, C.c_float, C.c_int
    L.xf_last_error.restype = C.c_char_p
    L.xf_device_count.argtypes = [C.POINTER(C.c_int)]
    L.xf_default_create_params.argtypes = [C.POINTER(CreateParams)]
    L.xf_default_create_params.restype = None
    L.xf_generate_tet_block.argtypes = [u32, u32, u32, f32, f32, f32, u32, f32, vp, vp, vp]
    L.xf_create.argtypes = [C.POINTER(CreateParams), vp, u32, vp, u32, C.POINTER(vp)]
    L.xf_destroy.argtypes = [vp]
    for n in ("xf_vert_count", "xf_element_count", "xf_color_count"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = u32
    L.xf_get_order.argtypes = [vp, vp]
    L.xf_get_colors.argtypes = [vp, vp]
    L.xf_get_stage_codes.argtypes = [vp, vp, vp]
    L.xf_get_chain_info.argtypes = [vp, vp, vp]
    L.xf_get_elements.argtypes = [vp] * 7
    L.xf_substep.argtypes = [vp, vp, vp, f32, u32]
    L.xf_substep_varying.argtypes = [vp, vp, vp, f32, u32, vp, vp]
    L.xf_sync.argtypes = [vp]
    L.xf_set_ground.argtypes = [vp, i32, f32, f32]
    L.xf_set_handles.argtypes = [vp, u32, vp, vp]
    L.xf_get_state.argtypes = [vp, vp, vp, vp]
    L.xf_set_state.argtypes = [vp, vp, vp, vp]
    L.xf_get_rest.argtypes = [vp, vp, vp, vp]
    L.xf_get_origin.argtypes = [vp, vp]
    L.xf_get_state_async.argtypes = [vp, vp, vp]
    L.xf_set_state_async.argtypes = [vp, vp, vp]
    L.xf_transform.argtypes = [vp, vp]
    L.xf_volume.argtypes = [vp, C.POINTER(f32)]
    L.xf_stats.argtypes = [vp, vp, vp]
    L.xf_get_info.argtypes = [vp, C.POINTER(Info)]
    L.xf_frame_state_init.argtypes = [C.POINTER(FrameState)]
    L.xf_frame_state_init.restype = None
    L.xf_frame_update.argtypes = [vp, vp, vp, f32, f32, C.POINTER(FrameState), C.POINTER(u32)]
    L.xf_sim_create.argtypes = [C.POINTER(CreateParams), C.POINTER(vp)]
    L.xf_sim_destroy.argtypes = [vp]
    L.xf_sim_reset.argtypes = [vp]
    L.xf_sim_add_block.argtypes = [vp, u32, vp, u32, vp, u32, i32, vp, u32]
    L.xf_block_from_settings.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(f32), C.POINTER(f32), C.POINTER(u32)]
    L.xf_sim_add_block_from_settings.argtypes = [vp, vp]
    L.xf_sim_finish_adding_blocks.argtypes = [vp, vp]
    L.xf_sim_set_geo_offset.argtypes = [vp, f32, f32]
    L.xf_sim_update.argtypes = [vp, vp, vp, i32, f32, f32, C.POINTER(u32)]
    L.xf_sim_geo_count.argtypes = [vp]
    L.xf_sim_geo_count.restype = u32
    L.xf_sim_geo.argtypes = [vp, u32]
    L.xf_sim_geo.restype = vp
    L.xf_sim_volume0.argtypes = [vp, u32]
    L.xf_sim_volume0.restype = f32
    L.xf_sim_get_frame_state.argtypes = [vp, C.POINTER(FrameState)]
    L.xf_batch_create.argtypes = [C.POINTER(CreateParams), vp, u32, vp, u32, u32, C.POINTER(vp)]
    L.xf_batch_destroy.argtypes = [vp]
    for n in ("xf_batch_scene_count", "xf_batch_vert_count", "xf_batch_element_count", "xf_batch_color_count"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = u32
    L.xf_batch_get_order.argtypes = [vp, vp]
    L.xf_batch_set_ground.argtypes = [vp, i32, f32, f32]
    L.xf_batch_substep.argtypes = [vp, vp, u32, f32, u32]
    L.xf_batch_sync.argtypes = [vp]
    L.xf_batch_get_state.argtypes = [vp, u32, u32, vp, vp, vp]
    L.xf_batch_set_state.argtypes = [vp, u32, u32, vp, vp, vp]
    L.xf_batch_get_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_uint64)]
    L.xf_part_create.argtypes = [C.POINTER(CreateParams), vp, u32, vp, u32, u32, u32, C.POINTER(vp)]
    L.xf_part_destroy.argtypes = [vp]
    for n in ("xf_part_local_vert_count", "xf_part_local_element_count", "xf_part_peer_count", "xf_part_color_count",
              "xf_part_global_vert_count", "xf_part_global_element_count"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = u32
    L.xf_part_get_local_verts.argtypes = [vp, vp]
    L.xf_part_get_local_elements.argtypes = [vp, vp, vp]
    L.xf_part_get_peers.argtypes = [vp, vp]
    L.xf_part_get_halo.argtypes = [vp, u32, u32, i32, C.POINTER(u32), vp]
    L.xf_part_get_order.argtypes = [vp, vp]
    L.xf_part_get_global_color_start.argtypes = [vp, vp]
    L.xf_part_get_initial.argtypes = [vp, vp, vp]
    L.xf_part_get_dataflow_codes.argtypes = [vp, vp, vp, C.POINTER(C.c_int)]
    L.xf_part_ipc_export.argtypes = [vp, vp]
    L.xf_part_ipc_connect.argtypes = [vp, vp]
    L.xf_part_set_ground.argtypes = [vp, i32, f32, f32]
    L.xf_part_substep.argtypes = [vp, vp, f32, u32]
    L.xf_part_sync.argtypes = [vp]
    L.xf_part_get_state.argtypes = [vp, vp, vp, vp]
    L.xf_part_get_info.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.xf_debug_l2_bandwidth.argtypes = [i32, i32, C.c_uint64, u32, i32, i32, C.POINTER(C.c_double)]
    L.xf_debug_torn_records.argtypes = [i32, i32, u32, u32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.xf_debug_scene_knob.argtypes = [vp, i32, u32]
    L.xf_debug_barrier_us.argtypes = [i32, i32, i32, i32, u32, C.POINTER(f32)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise XfError(rc, lib().xf_last_error().decode("utf-8", "replace"))


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count():
    n = C.c_int(0)
    rc = lib().xf_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def l2_bandwidth(device=0, mode=0, mbytes=16, passes=200, reps=5, blocks_per_sm=8):
    """GB/s of an L2-resident streaming copy (mode 0, read + write bytes) or read (mode 1); see xf_debug_l2_bandwidth."""
    out = C.c_double(0.0)
    _check(lib().xf_debug_l2_bandwidth(device, mode, int(mbytes) << 20, passes, reps, blocks_per_sm, C.byref(out)))
    return out.value


def stage_latency(device=0, mode=0, iterations=2000):
    """Cycles per element solve (0), per record hand-off between two SMs (1), per dependent 256-bit L2 load (2)."""
    out = C.c_double(0.0)
    lib().xf_debug_stage_latency.argtypes = [C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_double)]
    _check(lib().xf_debug_stage_latency(device, mode, iterations, C.byref(out)))
    return out.value


def substep_constants(compliance, poisson, dt):
    """{a = 1 + mu/lambda, 1/mu, 1/lambda, dt^2} of a call, in fp32 with one rounding per operation, as FillSubstepParams
    (xf_prepare.cpp) and Fem.cpp:445-449 compute them."""
    f = np.float32
    mu = f(1.0) / f(compliance)
    with np.errstate(divide="ignore"):
        lam = (f(2.0) * mu * f(poisson)) / (f(1.0) - f(2.0) * f(poisson))
        a = f(1.0) + mu / lam
        inv_lambda = f(1.0) / lam
    return np.array([a, f(1.0) / mu, inv_lambda, f(dt) * f(dt)], dtype=np.float32)


def gathered_elements(elements, X, w):
    """Inputs of xf_debug_coop_element from xf_get_elements' dict and a state: (consts [nT,16], Xg [nT,12], wg [nT,4])."""
    idx = elements["idx"].astype(np.int64)
    consts = np.ascontiguousarray(np.concatenate([elements["Qi"], elements["volume"][:, None], elements["QQ"], elements["QR"]], axis=1),
                                  dtype=np.float32)
    Xg = np.ascontiguousarray(X[idx].reshape(idx.shape[0], 12), dtype=np.float64)
    wg = np.ascontiguousarray(w[idx], dtype=np.float32)
    return consts, Xg, wg


def coop_element_probe(consts, Xg, wg, params4, energy=Energy_YeohSkinFast, iterations=200, warps_per_sm=8, device=0, variant=0):
    """Four-lanes-per-element solve against the one-thread solve on the same gathered elements; see xf_debug_coop_element."""
    L = lib()
    L.xf_debug_coop_element.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    n = consts.shape[0]
    assert consts.shape == (n, 16) and Xg.shape == (n, 12) and wg.shape == (n, 4)
    out = np.zeros(8, dtype=np.float64)
    xs, xc = np.empty((n, 12), np.float64), np.empty((n, 12), np.float64)
    params4 = np.ascontiguousarray(params4, dtype=np.float32)
    _check(L.xf_debug_coop_element(device, int(energy), int(variant), _vp(consts), _vp(Xg), _vp(wg), n, _vp(params4), iterations, warps_per_sm, _vp(out),
                                   _vp(xs), _vp(xc)))
    return dict(cycles_single=out[0], cycles_coop=out[1], solves_per_s_single=out[2], solves_per_s_coop=out[3], mismatched=int(out[4]),
                compared=int(out[5]), sm_count=int(out[6]), clock_khz=out[7], x_single=xs, x_coop=xc)


def torn_records(device=0, remote_device=-1, n_records=1 << 16, rounds=2000):
    """(reads, torn) of the 256-bit record stress test; see xf_debug_torn_records."""
    reads, torn = C.c_uint64(0), C.c_uint64(0)
    _check(lib().xf_debug_torn_records(device, remote_device, n_records, rounds, C.byref(reads), C.byref(torn)))
    return reads.value, torn.value


def shard_scenes(n_scenes, world, rank):
    """Scenes [first, first + count) of a batch of `n_scenes` independent scenes that rank `rank` of `world` steps (BASELINE
    config 3: batches shard over GPUs with no communication).  Contiguous, ragged by at most one scene, every scene owned once;
    a scene's settings are derived from its GLOBAL index, so the result does not depend on the world size."""
    if world < 1 or not 0 <= rank < world or n_scenes < 0:
        raise ValueError("need world >= 1, 0 <= rank < world, n_scenes >= 0")
    base, extra = divmod(n_scenes, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def block_scale(dim=3.0):
    """(0.7f * kSpacing) * dim in fp32, as Demo::UpdateSettings scales its blocks (Demo.cpp:16, 318)."""
    k_spacing = np.float32(np.float32(20.0) / np.float32(100.0)) / np.float32(31.0)
    return float(np.float32(np.float32(0.7) * k_spacing) * np.float32(dim))


def GenerateTetBlock(width, height, depth=None, scale=None, pattern=Pattern_Uniform, wonkiness=0.0):
    """MeshGen.cpp:223-244 on the host.  Returns (nodes f32[3*nV], idxStream u32[30*nHex], colorHint u32[6*nHex])."""
    depth = height if depth is None else depth
    if scale is None:
        scale = (block_scale(),) * 3
    nodes = np.empty(3 * (width + 1) * (height + 1) * (depth + 1), dtype=np.float32)
    idx = np.empty(30 * width * height * depth, dtype=np.uint32)
    hint = np.empty(6 * width * height * depth, dtype=np.uint32)
    _check(lib().xf_generate_tet_block(width, height, depth, scale[0], scale[1], scale[2], pattern, wonkiness, _vp(nodes), _vp(idx),
                                       _vp(hint)))
    return nodes, idx, hint


class GeoLinear3dCuda:
    """One tet mesh resident on a B200; the drop-in for the reference's GeoLinear3d on the Substep path."""

    def __init__(self, nodes, idx_stream, density=1.0, auto_resize=False, device=0, precision=PRECISION_EXACT,
                 schedule=SCHEDULE_AUTO, color_hint=None, stream=None, grouping=GROUPING_AUTO):
        L = lib()
        nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1)
        idx_stream = np.ascontiguousarray(idx_stream, dtype=np.uint32).reshape(-1)
        p = CreateParams()
        L.xf_default_create_params(C.byref(p))
        p.device = device
        p.density = density
        p.autoResize = 1 if auto_resize else 0
        p.precision = precision
        p.schedule = schedule
        p.grouping = grouping
        p.stream = stream
        if color_hint is not None:
            color_hint = np.ascontiguousarray(color_hint, dtype=np.uint32)
            p.colorHint = color_hint.ctypes.data
            p.colorHintCount = color_hint.size
        h = C.c_void_p()
        self._h = None
        _check(L.xf_create(C.byref(p), _vp(nodes), nodes.size, _vp(idx_stream), idx_stream.size, C.byref(h)))
        self._h = h
        self.nV = L.xf_vert_count(h)
        self.nT = L.xf_element_count(h)
        self.nColors = L.xf_color_count(h)

    @classmethod
    def _view(cls, handle):
        """A scene owned by somebody else (a SimCuda): same methods, close() does not destroy it."""
        self = cls.__new__(cls)
        self._h, self._owned = handle, False
        L = lib()
        self.nV, self.nT, self.nColors = L.xf_vert_count(handle), L.xf_element_count(handle), L.xf_color_count(handle)
        return self

    def close(self):
        if self._h:
            if getattr(self, "_owned", True):
                lib().xf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Geo interface -------------------------------------------------------------------------
    def VertCount(self):
        return self.nV

    def ElementCount(self):
        return self.nT

    def Substep(self, settings, dt, n=1, manip=None):
        """Geo::Substep(settings, manip, dt), n times (asynchronous)."""
        _check(lib().xf_substep(self._h, C.byref(settings), C.byref(manip) if manip is not None else None, float(dt), int(n)))

    def SubstepVarying(self, settings, dt, lock_rows=None, dir_rows=None, manip=None):
        """n substeps in one launch with a per-substep lock transform (n x 12) and / or manipulator ray (n x 3)."""
        lock_rows = None if lock_rows is None else np.ascontiguousarray(lock_rows, dtype=np.float32).reshape(-1, 12)
        dir_rows = None if dir_rows is None else np.ascontiguousarray(dir_rows, dtype=np.float32).reshape(-1, 3)
        n = (lock_rows if lock_rows is not None else dir_rows).shape[0]
        _check(lib().xf_substep_varying(self._h, C.byref(settings), C.byref(manip) if manip is not None else None, float(dt), int(n), _vp(lock_rows),
                                        _vp(dir_rows)))

    def Sync(self):
        _check(lib().xf_sync(self._h))

    def Transform(self, m9):
        m9 = np.ascontiguousarray(m9, dtype=np.float32).reshape(9)
        _check(lib().xf_transform(self._h, _vp(m9)))

    def CalculateVolume(self):
        v = C.c_float()
        _check(lib().xf_volume(self._h, C.byref(v)))
        return float(v.value)

    def FrameUpdate(self, settings, dt, median_frame_time, state, manip=None):
        """Sim::Update for this geo (Demo.cpp:37-103); returns the number of substeps taken."""
        n = C.c_uint32()
        _check(lib().xf_frame_update(self._h, C.byref(settings), C.byref(manip) if manip is not None else None, float(dt),
                                     float(median_frame_time), C.byref(state), C.byref(n)))
        return n.value

    # ---- public members of the reference's Geo3d / GeoLinear3d ------------------------------------
    def get_order(self):
        o = np.empty(self.nT, dtype=np.uint32)
        _check(lib().xf_get_order(self._h, _vp(o)))
        return o

    def stage_codes(self):
        pred = np.empty((self.nT, 4), dtype=np.uint8)
        last = np.empty(self.nV, dtype=np.uint8)
        _check(lib().xf_get_stage_codes(self._h, _vp(pred), _vp(last)))
        return pred, last

    def damping_codes(self):
        """(rank[nT, 4] in serial order, below[nV, 8]): the write-count codes of the barrier-free damping sweeps."""
        rank = np.empty((self.nT, 4), dtype=np.uint8)
        below = np.empty((self.nV, 8), dtype=np.uint8)
        lib().xf_get_damping_codes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _check(lib().xf_get_damping_codes(self._h, _vp(rank), _vp(below)))
        return rank, below

    def chain_info(self):
        """XF_GROUPING_CHAINS: slot / first / last word of every element (xf_get_order positions) and the per-mille of
        corner uses served from a private slot."""
        info = np.empty(self.nT, dtype=np.uint32)
        permille = C.c_uint32(0)
        _check(lib().xf_get_chain_info(self._h, _vp(info), C.byref(permille)))
        return info, permille.value

    def get_colors(self):
        c = np.empty(self.nT, dtype=np.uint32)
        _check(lib().xf_get_colors(self._h, _vp(c)))
        return c

    def get_state(self):
        X = np.empty((self.nV, 3), dtype=np.float64)
        V = np.empty((self.nV, 3), dtype=np.float64)
        w = np.empty(self.nV, dtype=np.float32)
        _check(lib().xf_get_state(self._h, _vp(X), _vp(V), _vp(w)))
        return X, V, w

    def set_state(self, X=None, V=None, w=None):
        X = None if X is None else np.ascontiguousarray(X, dtype=np.float64)
        V = None if V is None else np.ascontiguousarray(V, dtype=np.float64)
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
        _check(lib().xf_set_state(self._h, _vp(X), _vp(V), _vp(w)))

    def get_rest(self):
        X0 = np.empty((self.nV, 3), dtype=np.float64)
        O = np.empty((self.nV, 3), dtype=np.float64)
        flags = np.empty(self.nV, dtype=np.uint8)
        _check(lib().xf_get_rest(self._h, _vp(X0), _vp(O), _vp(flags)))
        return X0, O, flags

    def get_origin(self):
        o = np.empty(3, dtype=np.float32)
        _check(lib().xf_get_origin(self._h, _vp(o)))
        return o

    def get_elements(self):
        n = self.nT
        out = dict(idx=np.empty((n, 4), np.uint32), Qi=np.empty((n, 9), np.float32), QQ=np.empty((n, 3), np.float32),
                   QR=np.empty((n, 3), np.float32), volume=np.empty(n, np.float32), area=np.empty(n, np.float32))
        _check(lib().xf_get_elements(self._h, _vp(out["idx"]), _vp(out["Qi"]), _vp(out["QQ"]), _vp(out["QR"]), _vp(out["volume"]),
                                     _vp(out["area"])))
        return out

    # ---- extensions / helpers ---------------------------------------------------------------------
    def set_ground(self, enabled, y0=0.0, friction=0.0):
        _check(lib().xf_set_ground(self._h, 1 if enabled else 0, y0, friction))

    def set_handles(self, vert_idx, targets):
        vert_idx = np.ascontiguousarray(vert_idx, dtype=np.uint32)
        targets = np.ascontiguousarray(targets, dtype=np.float32).reshape(-1)
        _check(lib().xf_set_handles(self._h, vert_idx.size, _vp(vert_idx), _vp(targets)))

    def stats(self, settings):
        out = np.zeros(6, dtype=np.float64)
        _check(lib().xf_stats(self._h, C.byref(settings), _vp(out)))
        return dict(volume=out[0], kinetic=out[1], gravitational=out[2], deviatoric=out[3], volumetric=out[4], nonfinite=out[5])

    def debug_knob(self, knob, value):
        _check(lib().xf_debug_scene_knob(self._h, knob, value))

    def info(self):
        i = Info()
        _check(lib().xf_get_info(self._h, C.byref(i)))
        d = {k: getattr(i, k) for k, _ in Info._fields_}
        d["lastKernel"] = KERNEL_NAMES.get(d["lastKernelId"])
        return d

    def get_state_async(self, X_ptr, V_ptr):
        """Enqueue device->host copies into caller-owned (pinned) buffers; pointers are integers or None."""
        _check(lib().xf_get_state_async(self._h, X_ptr, V_ptr))

    def set_state_async(self, X_ptr, V_ptr):
        _check(lib().xf_set_state_async(self._h, X_ptr, V_ptr))


class SimCuda:
    """The reference's Sim (Demo.h:18-52): several Geos under one Settings block and one frame clock (xf_sim_*)."""

    def __init__(self, device=0, precision=PRECISION_EXACT, schedule=SCHEDULE_AUTO, stream=None):
        p = CreateParams()
        lib().xf_default_create_params(C.byref(p))
        p.device, p.precision, p.schedule, p.stream = device, precision, schedule, stream
        h = C.c_void_p()
        self._h = None
        _check(lib().xf_sim_create(C.byref(p), C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            lib().xf_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Reset(self):
        _check(lib().xf_sim_reset(self._h))

    def AddBlock(self, nodes, idx_stream, auto_resize=False, color_hint=None):
        nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1)
        idx_stream = np.ascontiguousarray(idx_stream, dtype=np.uint32).reshape(-1)
        hint = None if color_hint is None else np.ascontiguousarray(color_hint, dtype=np.uint32)
        _check(lib().xf_sim_add_block(self._h, Element_T4, _vp(nodes), nodes.size, _vp(idx_stream), idx_stream.size, 1 if auto_resize else 0,
                                      _vp(hint), 0 if hint is None else hint.size))

    def AddBlockFromSettings(self, settings):
        _check(lib().xf_sim_add_block_from_settings(self._h, C.byref(settings)))

    def FinishAddingBlocks(self, settings):
        _check(lib().xf_sim_finish_adding_blocks(self._h, C.byref(settings)))

    def SetGeoOffset(self, x, y):
        _check(lib().xf_sim_set_geo_offset(self._h, float(x), float(y)))

    def Update(self, settings, dt, median_frame_time, manip=None, picked_geo=-1):
        n = C.c_uint32(0)
        _check(lib().xf_sim_update(self._h, C.byref(settings), C.byref(manip) if manip is not None else None, int(picked_geo), float(dt),
                                   float(median_frame_time), C.byref(n)))
        return n.value

    def geo_count(self):
        return int(lib().xf_sim_geo_count(self._h))

    def geo(self, i):
        """Geo i as a (non-owning) GeoLinear3dCuda view."""
        h = lib().xf_sim_geo(self._h, i)
        if not h:
            raise IndexError(i)
        return GeoLinear3dCuda._view(C.c_void_p(h))

    def volume0(self, i):
        return float(lib().xf_sim_volume0(self._h, i))


def block_from_settings(settings):
    """(width, height, scaleX, scaleY, pattern) of the block Demo::UpdateSettings builds for these settings (Demo.cpp:289-318)."""
    w, h, pat = C.c_uint32(), C.c_uint32(), C.c_uint32()
    sx, sy = C.c_float(), C.c_float()
    _check(lib().xf_block_from_settings(C.byref(settings), C.byref(w), C.byref(h), C.byref(sx), C.byref(sy), C.byref(pat)))
    return w.value, h.value, sx.value, sy.value, pat.value


class GeoBatchCuda:
    """nScenes independent instances of one rest mesh (Sim::Update's `for geo` loop in one launch)."""

    def __init__(self, nodes, idx_stream, n_scenes, density=1.0, auto_resize=False, device=0, precision=PRECISION_EXACT,
                 color_hint=None, stream=None):
        L = lib()
        nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1)
        idx_stream = np.ascontiguousarray(idx_stream, dtype=np.uint32).reshape(-1)
        p = CreateParams()
        L.xf_default_create_params(C.byref(p))
        p.device = device
        p.density = density
        p.autoResize = 1 if auto_resize else 0
        p.precision = precision
        p.stream = stream
        if color_hint is not None:
            color_hint = np.ascontiguousarray(color_hint, dtype=np.uint32)
            p.colorHint = color_hint.ctypes.data
            p.colorHintCount = color_hint.size
        h = C.c_void_p()
        self._h = None
        _check(L.xf_batch_create(C.byref(p), _vp(nodes), nodes.size, _vp(idx_stream), idx_stream.size, int(n_scenes), C.byref(h)))
        self._h = h
        self.nScenes = L.xf_batch_scene_count(h)
        self.nV = L.xf_batch_vert_count(h)
        self.nT = L.xf_batch_element_count(h)
        self.nColors = L.xf_batch_color_count(h)

    def close(self):
        if self._h:
            lib().xf_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_order(self):
        o = np.empty(self.nT, dtype=np.uint32)
        _check(lib().xf_batch_get_order(self._h, _vp(o)))
        return o

    def set_ground(self, enabled, y0=0.0, friction=0.0):
        _check(lib().xf_batch_set_ground(self._h, 1 if enabled else 0, y0, friction))

    def Substep(self, settings, dt, n=1):
        """settings: one Settings (shared) or a ctypes array (Settings * nScenes)."""
        if isinstance(settings, Settings):
            _check(lib().xf_batch_substep(self._h, C.byref(settings), 1, float(dt), int(n)))
        else:
            _check(lib().xf_batch_substep(self._h, C.byref(settings), len(settings), float(dt), int(n)))

    def Sync(self):
        _check(lib().xf_batch_sync(self._h))

    def get_state(self, first=0, count=None):
        count = self.nScenes - first if count is None else count
        X = np.empty((count, self.nV, 3), dtype=np.float64)
        V = np.empty((count, self.nV, 3), dtype=np.float64)
        w = np.empty((count, self.nV), dtype=np.float32)
        _check(lib().xf_batch_get_state(self._h, first, count, _vp(X), _vp(V), _vp(w)))
        return X, V, w

    def set_state(self, first, X=None, V=None, w=None):
        X = None if X is None else np.ascontiguousarray(X, dtype=np.float64)
        V = None if V is None else np.ascontiguousarray(V, dtype=np.float64)
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
        count = (X if X is not None else V if V is not None else w).shape[0]
        _check(lib().xf_batch_set_state(self._h, first, count, _vp(X), _vp(V), _vp(w)))

    def info(self):
        g, b, s, l = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        _check(lib().xf_batch_get_info(self._h, C.byref(g), C.byref(b), C.byref(s), C.byref(l)))
        return dict(groupThreads=g.value, blockThreads=b.value, smemBytes=s.value, launches=l.value)


class GeoPartitionCuda:
    """This rank's share of ONE mesh stepped on several GPUs (one process per GPU).  device=-1: host-only plan."""

    IPC_BYTES = 128

    def __init__(self, nodes, idx_stream, n_ranks, rank, density=1.0, auto_resize=False, device=0, precision=PRECISION_EXACT,
                 color_hint=None, stream=None, schedule=SCHEDULE_AUTO, partition=0):
        L = lib()
        nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1)
        idx_stream = np.ascontiguousarray(idx_stream, dtype=np.uint32).reshape(-1)
        p = CreateParams()
        L.xf_default_create_params(C.byref(p))
        p.device = device
        p.density = density
        p.autoResize = 1 if auto_resize else 0
        p.precision = precision
        p.schedule = schedule
        p.stream = stream
        p.partition = partition
        if color_hint is not None:
            color_hint = np.ascontiguousarray(color_hint, dtype=np.uint32)
            p.colorHint = color_hint.ctypes.data
            p.colorHintCount = color_hint.size
        h = C.c_void_p()
        self._h = None
        _check(L.xf_part_create(C.byref(p), _vp(nodes), nodes.size, _vp(idx_stream), idx_stream.size, int(n_ranks), int(rank), C.byref(h)))
        self._h = h
        self.nRanks, self.rank = int(n_ranks), int(rank)
        self.nV = L.xf_part_local_vert_count(h)
        self.nT = L.xf_part_local_element_count(h)
        self.nPeers = L.xf_part_peer_count(h)
        self.nColors = L.xf_part_color_count(h)
        self.nVGlobal = L.xf_part_global_vert_count(h)
        self.nTGlobal = L.xf_part_global_element_count(h)

    def close(self):
        if self._h:
            lib().xf_part_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def local_verts(self):
        a = np.empty(self.nV, dtype=np.uint32)
        _check(lib().xf_part_get_local_verts(self._h, _vp(a)))
        return a

    def local_elements(self):
        e = np.empty(self.nT, dtype=np.uint32)
        cs = np.empty(self.nColors + 1, dtype=np.uint32)
        _check(lib().xf_part_get_local_elements(self._h, _vp(e), _vp(cs)))
        return e, cs

    def peers(self):
        a = np.empty(self.nPeers, dtype=np.uint32)
        _check(lib().xf_part_get_peers(self._h, _vp(a)))
        return a

    def halo(self, color, peer_slot, send):
        n = C.c_uint32()
        _check(lib().xf_part_get_halo(self._h, color, peer_slot, 1 if send else 0, C.byref(n), None))
        a = np.empty(n.value, dtype=np.uint32)
        _check(lib().xf_part_get_halo(self._h, color, peer_slot, 1 if send else 0, C.byref(n), _vp(a)))
        return a

    def get_order(self):
        o = np.empty(self.nTGlobal, dtype=np.uint32)
        _check(lib().xf_part_get_order(self._h, _vp(o)))
        return o

    def global_color_start(self):
        cs = np.empty(self.nColors + 1, dtype=np.uint32)
        _check(lib().xf_part_get_global_color_start(self._h, _vp(cs)))
        return cs

    def initial(self):
        w = np.empty(self.nV, dtype=np.float32)
        f = np.empty(self.nV, dtype=np.uint8)
        _check(lib().xf_part_get_initial(self._h, _vp(w), _vp(f)))
        return w, f

    def dataflow_codes(self):
        pred = np.empty((self.nT, 4), dtype=np.uint8)
        last = np.empty(self.nV, dtype=np.uint8)
        ok = C.c_int()
        _check(lib().xf_part_get_dataflow_codes(self._h, _vp(pred), _vp(last), C.byref(ok)))
        return pred, last, bool(ok.value)

    def ipc_export(self):
        b = np.zeros(self.IPC_BYTES, dtype=np.uint8)
        _check(lib().xf_part_ipc_export(self._h, _vp(b)))
        return b

    def ipc_connect(self, all_blobs):
        all_blobs = np.ascontiguousarray(all_blobs, dtype=np.uint8).reshape(-1)
        assert all_blobs.size == self.IPC_BYTES * self.nRanks
        _check(lib().xf_part_ipc_connect(self._h, _vp(all_blobs)))

    def set_ground(self, enabled, y0=0.0, friction=0.0):
        _check(lib().xf_part_set_ground(self._h, 1 if enabled else 0, y0, friction))

    def Substep(self, settings, dt, n=1):
        _check(lib().xf_part_substep(self._h, C.byref(settings), float(dt), int(n)))

    def Sync(self):
        _check(lib().xf_part_sync(self._h))

    def get_state(self):
        X = np.empty((self.nV, 3), dtype=np.float64)
        V = np.empty((self.nV, 3), dtype=np.float64)
        w = np.empty(self.nV, dtype=np.float32)
        _check(lib().xf_part_get_state(self._h, _vp(X), _vp(V), _vp(w)))
        return X, V, w

    def info(self):
        l, e = C.c_uint64(), C.c_uint64()
        _check(lib().xf_part_get_info(self._h, C.byref(l), C.byref(e)))
        return dict(launches=l.value, epoch=e.value)
