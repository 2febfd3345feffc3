"""TEST INFRASTRUCTURE — ctypes bindings for the two CPU checkers.

* ``RefScene``    -> oracle/_ref/libxpbd_ref_{strict,fast}.so : the UNMODIFIED reference
  (Geo.cpp/Fem.cpp/... compiled where they lie under /root/reference) behind
  oracle/ref_harness.cpp.
* ``OracleScene`` -> oracle/libxpbd_oracle.so : the plain-C restatement (oracle/xpbd_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (xpbd-fem_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# ---- flag word, Settings.h:9-75 -------------------------------------------------------------
Settings_EnergyBit = 6
Settings_XpbdSolveBit = 11
Settings_RayleighTypeBit = 20
Settings_LockLeft = 1 << 26
Settings_LockRight = 1 << 27
Element_T4 = 5
Energy_Mixed, Energy_MixedSel, Energy_YeohSkin, Energy_YeohSkinFast = 3, 4, 5, 7
Pattern_Uniform, Pattern_Mirrored = 0, 1
Rayleigh_Paper, Rayleigh_Limit, Rayleigh_Post, Rayleigh_PostAmortized = 0, 1, 2, 3
Geo_Left, Geo_Right, Geo_Pickable = 1, 2, 16
kSpacing = np.float32(np.float32(20.0) / np.float32(100.0)) / np.float32(31.0)  # Demo.cpp:16


class Settings(C.Structure):
    """The reference's Settings POD, byte for byte (Settings.h:79-102; 160 B, align 16)."""
    _fields_ = [
        ("timeScale", C.c_float), ("substepsPerSecond", C.c_float), ("volumePasses", C.c_uint32),
        ("_pad0", C.c_uint32),
        ("gravity", C.c_float * 2), ("compliance", C.c_float), ("damping", C.c_float),
        ("pbdDamping", C.c_float), ("drag", C.c_float), ("poissonsRatio", C.c_float),
        ("wonkiness", C.c_float), ("leftRightSeparation", C.c_float), ("flags", C.c_uint32),
        ("areaAndTimeCorrectedPbdDamping", C.c_float), ("volumeAndTimeCorrectedPbdDamping", C.c_float),
        ("amortizedAreaAndTimeCorrectedPbdDamping", C.c_float),
        ("amortizedVolumeAndTimeCorrectedPbdDamping", C.c_float),
        ("timeCorrectedDrag", C.c_float), ("_pad1", C.c_uint32),
        ("lockedRightTransform", C.c_float * 4),
        ("lockedRightTransform3d", C.c_float * 12),  # 3 columns, each vec3 padded to 16 B
        ("tickId", C.c_uint32), ("_pad2", C.c_uint32 * 3),
    ]


assert C.sizeof(Settings) == 160


class Manipulator(C.Structure):
    """Flat mirror of Manipulator.h:9-13 (picked != 0 <=> pickedGeo == this geo)."""
    _fields_ = [
        ("pos", C.c_float * 3), ("manipPlaneNormal", C.c_float * 3), ("pick0", C.c_float * 3),
        ("pickDir", C.c_float * 3), ("pickDirOld", C.c_float * 3), ("pickDirTarget", C.c_float * 3),
        ("picked", C.c_int32), ("pickedPointIdx", C.c_uint32),
    ]


def make_settings(energy=Energy_MixedSel, simultaneous=True, poisson=0.5, compliance=1.0,
                  gravity=(0.0, -0.4905), damping=0.0, rayleigh=Rayleigh_Post, lock_left=True,
                  lock_right=False, drag_tc=0.0, volume_passes=0, pbd_damping=0.0,
                  substeps_per_second=3000.0):
    s = Settings()
    s.timeScale = 1.0
    s.substepsPerSecond = substeps_per_second
    s.volumePasses = volume_passes
    s.gravity[0], s.gravity[1] = gravity
    s.compliance = compliance
    s.damping = damping
    s.pbdDamping = pbd_damping
    s.drag = 0.0
    s.poissonsRatio = poisson
    s.wonkiness = 0.0
    s.leftRightSeparation = 1.0
    s.flags = (Element_T4 | (energy << Settings_EnergyBit) | ((1 if simultaneous else 0) << Settings_XpbdSolveBit)
               | (rayleigh << Settings_RayleighTypeBit) | (Settings_LockLeft if lock_left else 0)
               | (Settings_LockRight if lock_right else 0))
    s.timeCorrectedDrag = drag_tc
    # identity lock transform
    s.lockedRightTransform[0] = 1.0
    s.lockedRightTransform[3] = 1.0
    s.lockedRightTransform3d[0] = 1.0
    s.lockedRightTransform3d[5] = 1.0
    s.lockedRightTransform3d[10] = 1.0
    s.tickId = 0
    return s


def block_scale(dim=3.0):
    """(0.7f * kSpacing) * dim, evaluated in fp32 like Demo.cpp:318."""
    return float(np.float32(np.float32(0.7) * kSpacing) * np.float32(dim))


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def ref_lib_path(kind="strict"):
    return os.path.join(HERE, "_ref", "libxpbd_ref_%s.so" % kind)


def have_ref(kind="strict"):
    return os.path.exists(ref_lib_path(kind))


_ref_libs = {}


def _load_ref(kind):
    if kind in _ref_libs:
        return _ref_libs[kind]
    lib = C.CDLL(ref_lib_path(kind))
    vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
    lib.ref_create_block.restype = vp
    lib.ref_create_block.argtypes = [u32, u32, f32, f32, u32, f32, f32]
    lib.ref_create_armadillo.restype = vp
    lib.ref_create_armadillo.argtypes = [f32]
    lib.ref_create_mesh.restype = vp
    lib.ref_create_mesh.argtypes = [vp, u32, vp, u32, f32, C.c_int]
    lib.ref_destroy.argtypes = [vp]
    for n in ("ref_vert_count", "ref_tet_count", "ref_node_float_count", "ref_idx_count"):
        getattr(lib, n).restype = u32
        getattr(lib, n).argtypes = [vp]
    lib.ref_arena_used.restype = C.c_size_t
    lib.ref_arena_used.argtypes = [vp]
    lib.ref_get_mesh.argtypes = [vp, vp, vp]
    lib.ref_get_order.argtypes = [vp, vp]
    lib.ref_set_order.argtypes = [vp, vp]
    lib.ref_get_state.argtypes = [vp, vp, vp, vp]
    lib.ref_get_rest.argtypes = [vp, vp, vp, vp]
    lib.ref_set_state.argtypes = [vp, vp, vp, vp]
    lib.ref_get_origin.argtypes = [vp, vp]
    lib.ref_get_elements.argtypes = [vp] + [vp] * 6
    lib.ref_transform.argtypes = [vp, vp]
    lib.ref_volume.restype = f32
    lib.ref_volume.argtypes = [vp]
    lib.ref_pick.argtypes = [vp, vp, vp, vp, vp]
    lib.ref_substep.argtypes = [vp, vp, vp, f32, u32]
    lib.ref_substep_ext.argtypes = [vp, vp, vp, f32, u32]
    lib.ref_set_ground.argtypes = [vp, C.c_int, f32, f32]
    lib.ref_set_handles.argtypes = [vp, u32, vp, vp]
    lib.ref_time_substeps.restype = C.c_double
    lib.ref_time_substeps.argtypes = [vp, vp, f32, u32]
    _ref_libs[kind] = lib
    return lib


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pick(fn, handle, ray_origin, ray_dir):
    o = np.ascontiguousarray(ray_origin, dtype=np.float32).reshape(3)
    d = np.ascontiguousarray(ray_dir, dtype=np.float32).reshape(3)
    out5 = np.zeros(5, dtype=np.float32)
    idx = np.zeros(1, dtype=np.uint32)
    fn(handle, _vp(o), _vp(d), _vp(out5), _vp(idx))
    return bool(out5[4] != 0.0), int(idx[0]), out5[:3].copy(), float(out5[3])


class RefScene:
    """One GeoLinear3d of the unmodified reference."""

    def __init__(self, handle, kind):
        self.lib = _load_ref(kind)
        if not handle:
            raise MemoryError("reference harness could not create the scene")
        self.h = C.c_void_p(handle)
        self.nV = self.lib.ref_vert_count(self.h)
        self.nT = self.lib.ref_tet_count(self.h)

    @classmethod
    def block(cls, width, height, scale_x=None, scale_y=None, pattern=Pattern_Uniform, wonkiness=0.0,
              density=1.0, kind="strict"):
        sx = block_scale() if scale_x is None else scale_x
        sy = sx if scale_y is None else scale_y
        lib = _load_ref(kind)
        return cls(lib.ref_create_block(width, height, sx, sy, pattern, wonkiness, density), kind)

    @classmethod
    def armadillo(cls, density=2.0, kind="strict"):
        lib = _load_ref(kind)
        return cls(lib.ref_create_armadillo(density), kind)

    @classmethod
    def mesh(cls, nodes, idx_stream, density=1.0, auto_resize=False, kind="strict"):
        lib = _load_ref(kind)
        nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1)
        idx_stream = np.ascontiguousarray(idx_stream, dtype=np.uint32).reshape(-1)
        return cls(lib.ref_create_mesh(_vp(nodes), nodes.size, _vp(idx_stream), idx_stream.size, density,
                                       1 if auto_resize else 0), kind)

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_mesh(self):
        nodes = np.empty(self.lib.ref_node_float_count(self.h), dtype=np.float32)
        idx = np.empty(self.lib.ref_idx_count(self.h), dtype=np.uint32)
        self.lib.ref_get_mesh(self.h, _vp(nodes), _vp(idx))
        return nodes, idx

    def get_order(self):
        o = np.empty(self.nT, dtype=np.uint32)
        self.lib.ref_get_order(self.h, _vp(o))
        return o

    def set_order(self, order):
        order = np.ascontiguousarray(order, dtype=np.uint32)
        assert order.size == self.nT
        self.lib.ref_set_order(self.h, _vp(order))

    def get_state(self):
        X = np.empty((self.nV, 3), dtype=np.float64)
        V = np.empty((self.nV, 3), dtype=np.float64)
        w = np.empty(self.nV, dtype=np.float32)
        self.lib.ref_get_state(self.h, _vp(X), _vp(V), _vp(w))
        return X, V, w

    def get_rest(self):
        X0 = np.empty((self.nV, 3), dtype=np.float64)
        O = np.empty((self.nV, 3), dtype=np.float64)
        flags = np.empty(self.nV, dtype=np.uint8)
        self.lib.ref_get_rest(self.h, _vp(X0), _vp(O), _vp(flags))
        return X0, O, flags

    def set_state(self, X=None, V=None, w=None):
        X = None if X is None else np.ascontiguousarray(X, dtype=np.float64)
        V = None if V is None else np.ascontiguousarray(V, dtype=np.float64)
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
        self.lib.ref_set_state(self.h, _vp(X), _vp(V), _vp(w))

    def get_origin(self):
        o = np.empty(3, dtype=np.float32)
        self.lib.ref_get_origin(self.h, _vp(o))
        return o

    def get_elements(self):
        n = self.nT
        out = dict(idx=np.empty((n, 4), np.uint32), Qi=np.empty((n, 9), np.float32), QQ=np.empty((n, 3), np.float32),
                   QR=np.empty((n, 3), np.float32), volume=np.empty(n, np.float32), area=np.empty(n, np.float32))
        self.lib.ref_get_elements(self.h, _vp(out["idx"]), _vp(out["Qi"]), _vp(out["QQ"]), _vp(out["QR"]),
                                  _vp(out["volume"]), _vp(out["area"]))
        return out

    def transform(self, m9):
        m9 = np.ascontiguousarray(m9, dtype=np.float32).reshape(9)
        self.lib.ref_transform(self.h, _vp(m9))

    def volume(self):
        return float(self.lib.ref_volume(self.h))

    def pick(self, ray_origin, ray_dir):
        """Geo3d::Pick -> (found, vertex index, point xyz, distance)."""
        return _pick(self.lib.ref_pick, self.h, ray_origin, ray_dir)

    def substep(self, settings, dt, n=1, manip=None, ext=False):
        f = self.lib.ref_substep_ext if ext else self.lib.ref_substep
        f(self.h, C.byref(settings), C.byref(manip) if manip is not None else None, dt, n)

    def set_ground(self, enabled, y0=0.0, friction=0.0):
        self.lib.ref_set_ground(self.h, 1 if enabled else 0, y0, friction)

    def set_handles(self, vert_idx, targets):
        vert_idx = np.ascontiguousarray(vert_idx, dtype=np.uint32)
        targets = np.ascontiguousarray(targets, dtype=np.float32).reshape(-1)
        self.lib.ref_set_handles(self.h, vert_idx.size, _vp(vert_idx), _vp(targets))

    def time_substeps(self, settings, dt, n):
        return float(self.lib.ref_time_substeps(self.h, C.byref(settings), dt, n))


# ---------------------------------------------------------------------------------------------
# The plain-C restatement (oracle/xpbd_oracle.c)
# ---------------------------------------------------------------------------------------------
def oracle_lib_path():
    return os.path.join(HERE, "libxpbd_oracle.so")


_oracle_lib = None


def _load_oracle():
    global _oracle_lib
    if _oracle_lib is not None:
        return _oracle_lib
    lib = C.CDLL(oracle_lib_path())
    vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
    lib.xo_generate_tet_block.argtypes = [u32, u32, u32, f32, f32, f32, u32, f32, vp, vp]
    lib.xo_create.restype = vp
    lib.xo_create.argtypes = [vp, u32, vp, u32, f32, C.c_int]
    lib.xo_destroy.argtypes = [vp]
    lib.xo_vert_count.restype = u32
    lib.xo_vert_count.argtypes = [vp]
    lib.xo_tet_count.restype = u32
    lib.xo_tet_count.argtypes = [vp]
    lib.xo_get_order.argtypes = [vp, vp]
    lib.xo_set_order.argtypes = [vp, vp]
    lib.xo_get_state.argtypes = [vp, vp, vp, vp]
    lib.xo_set_state.argtypes = [vp, vp, vp, vp]
    lib.xo_get_rest.argtypes = [vp, vp, vp, vp]
    lib.xo_get_elements.argtypes = [vp] * 7
    lib.xo_get_origin.argtypes = [vp, vp]
    lib.xo_substep.argtypes = [vp, vp, vp, f32, u32]
    lib.xo_set_ground.argtypes = [vp, C.c_int, f32, f32]
    lib.xo_set_state_precision.argtypes = [vp, C.c_int]
    lib.xo_phase_predict.argtypes = [vp, vp, f32]
    lib.xo_phase_sweep.argtypes = [vp, vp, f32, u32, u32]
    lib.xo_phase_post.argtypes = [vp, vp, vp, f32]
    lib.xo_set_flags.argtypes = [vp, vp]
    lib.xo_set_handles.argtypes = [vp, u32, vp, vp]
    lib.xo_transform.argtypes = [vp, vp]
    lib.xo_volume.restype = f32
    lib.xo_volume.argtypes = [vp]
    lib.xo_energy.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.xo_time_substeps.restype = C.c_double
    lib.xo_time_substeps.argtypes = [vp, vp, f32, u32]
    _oracle_lib = lib
    return lib


def generate_tet_block(width, height, depth=None, scale=None, pattern=Pattern_Uniform, wonkiness=0.0):
    """Oracle restatement of GenerateTetBlock (MeshGen.cpp:223-244). Returns (nodes f32[3nV], idx u32[30nHex])."""
    lib = _load_oracle()
    depth = height if depth is None else depth
    if scale is None:
        scale = (block_scale(),) * 3
    nodes = np.empty(3 * (width + 1) * (height + 1) * (depth + 1), dtype=np.float32)
    idx = np.empty(30 * width * height * depth, dtype=np.uint32)
    lib.xo_generate_tet_block(width, height, depth, scale[0], scale[1], scale[2], pattern, wonkiness, _vp(nodes), _vp(idx))
    return nodes, idx


class OracleScene:
    """One scene of the C restatement; same method names as RefScene."""

    def __init__(self, nodes, idx_stream, density=1.0, auto_resize=False):
        self.lib = _load_oracle()
        nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1)
        idx_stream = np.ascontiguousarray(idx_stream, dtype=np.uint32).reshape(-1)
        h = self.lib.xo_create(_vp(nodes), nodes.size, _vp(idx_stream), idx_stream.size, density, 1 if auto_resize else 0)
        if not h:
            raise ValueError("xo_create rejected the mesh")
        self.h = C.c_void_p(h)
        self.nV = self.lib.xo_vert_count(self.h)
        self.nT = self.lib.xo_tet_count(self.h)

    @classmethod
    def block(cls, width, height, pattern=Pattern_Uniform, wonkiness=0.0, density=1.0, scale=None):
        nodes, idx = generate_tet_block(width, height, None, scale, pattern, wonkiness)
        return cls(nodes, idx, density)

    def close(self):
        if self.h:
            self.lib.xo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_order(self):
        o = np.empty(self.nT, dtype=np.uint32)
        self.lib.xo_get_order(self.h, _vp(o))
        return o

    def set_order(self, order):
        order = np.ascontiguousarray(order, dtype=np.uint32)
        assert order.size == self.nT
        self.lib.xo_set_order(self.h, _vp(order))

    def get_state(self):
        X = np.empty((self.nV, 3), dtype=np.float64)
        V = np.empty((self.nV, 3), dtype=np.float64)
        w = np.empty(self.nV, dtype=np.float32)
        self.lib.xo_get_state(self.h, _vp(X), _vp(V), _vp(w))
        return X, V, w

    def set_state(self, X=None, V=None, w=None):
        X = None if X is None else np.ascontiguousarray(X, dtype=np.float64)
        V = None if V is None else np.ascontiguousarray(V, dtype=np.float64)
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
        self.lib.xo_set_state(self.h, _vp(X), _vp(V), _vp(w))

    def get_rest(self):
        X0 = np.empty((self.nV, 3), dtype=np.float64)
        O = np.empty((self.nV, 3), dtype=np.float64)
        flags = np.empty(self.nV, dtype=np.uint8)
        self.lib.xo_get_rest(self.h, _vp(X0), _vp(O), _vp(flags))
        return X0, O, flags

    def get_origin(self):
        o = np.empty(3, dtype=np.float32)
        self.lib.xo_get_origin(self.h, _vp(o))
        return o

    def get_elements(self):
        n = self.nT
        out = dict(idx=np.empty((n, 4), np.uint32), Qi=np.empty((n, 9), np.float32), QQ=np.empty((n, 3), np.float32),
                   QR=np.empty((n, 3), np.float32), volume=np.empty(n, np.float32), area=np.empty(n, np.float32))
        self.lib.xo_get_elements(self.h, _vp(out["idx"]), _vp(out["Qi"]), _vp(out["QQ"]), _vp(out["QR"]),
                                 _vp(out["volume"]), _vp(out["area"]))
        return out

    def copy_elements_from(self, full, elems, rest=None):
        """This scene is the sub-mesh `elems` (element ids of `full`, in this scene's element order) of `full`: take its element
        constants (and, with `rest`, its rest positions) instead of the ones re-derived from this scene's own coordinates."""
        el = full.get_elements()
        vp = C.c_void_p
        self.lib.xo_set_elements.argtypes = [vp] * 6
        self.lib.xo_set_rest.argtypes = [vp, vp]
        sub = {k: np.ascontiguousarray(el[k][elems]) for k in ("Qi", "QQ", "QR", "volume", "area")}
        self.lib.xo_set_elements(self.h, _vp(sub["Qi"]), _vp(sub["QQ"]), _vp(sub["QR"]), _vp(sub["volume"]), _vp(sub["area"]))
        if rest is not None:
            rest = np.ascontiguousarray(rest, dtype=np.float64)
            self.lib.xo_set_rest(self.h, _vp(rest))

    def transform(self, m9):
        m9 = np.ascontiguousarray(m9, dtype=np.float32).reshape(9)
        self.lib.xo_transform(self.h, _vp(m9))

    def volume(self):
        return float(self.lib.xo_volume(self.h))

    def energy(self, settings):
        vals = [C.c_double() for _ in range(4)]
        self.lib.xo_energy(self.h, C.byref(settings), *[C.byref(v) for v in vals])
        return tuple(v.value for v in vals)

    def substep(self, settings, dt, n=1, manip=None, ext=False):
        self.lib.xo_substep(self.h, C.byref(settings), C.byref(manip) if manip is not None else None, dt, n)

    def phase_predict(self, settings, dt):
        self.lib.xo_phase_predict(self.h, C.byref(settings), dt)

    def phase_sweep(self, settings, dt, begin, end):
        self.lib.xo_phase_sweep(self.h, C.byref(settings), dt, begin, end)

    def phase_elems(self, settings, dt, kind, elems):
        """Elements `elems` (scene element indices, in that order) of one sweep: 0 main, 1 volume pass, 2 Rayleigh damp, 3 PBD damp."""
        elems = np.ascontiguousarray(elems, dtype=np.uint32)
        self.lib.xo_phase_elems.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_uint32]
        self.lib.xo_phase_elems(self.h, C.byref(settings), dt, int(kind), _vp(elems), elems.size)

    def phase_post(self, settings, dt, manip=None):
        self.lib.xo_phase_post(self.h, C.byref(settings), C.byref(manip) if manip is not None else None, dt)

    def set_flags(self, flags):
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        assert flags.size == self.nV
        self.lib.xo_set_flags(self.h, _vp(flags))

    def set_ground(self, enabled, y0=0.0, friction=0.0):
        self.lib.xo_set_ground(self.h, 1 if enabled else 0, y0, friction)

    def set_state_precision(self, f32):
        """Study mode: keep X, O, V as fp32 would (tools/fp32_state_study.py); not the reference's algorithm."""
        self.lib.xo_set_state_precision(self.h, int(f32))  # bit 0: X as fp32, bit 1: V and O too

    def set_handles(self, vert_idx, targets):
        vert_idx = np.ascontiguousarray(vert_idx, dtype=np.uint32)
        targets = np.ascontiguousarray(targets, dtype=np.float32).reshape(-1)
        self.lib.xo_set_handles(self.h, vert_idx.size, _vp(vert_idx), _vp(targets))

    def time_substeps(self, settings, dt, n):
        return float(self.lib.xo_time_substeps(self.h, C.byref(settings), dt, n))


class RefSim:
    """The reference's own Sim (Demo.h:18-52) holding one T4 block, stepped by Sim::Update."""

    def __init__(self, nodes, idx_stream, settings, auto_resize=False, kind="strict"):
        lib = _load_ref(kind)
        vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
        lib.ref_sim_create.restype = vp
        lib.ref_sim_create.argtypes = [vp, u32, vp, u32, vp, C.c_int]
        lib.ref_sim_destroy.argtypes = [vp]
        lib.ref_sim_set_order.argtypes = [vp, vp]
        lib.ref_sim_get_state.argtypes = [vp, vp, vp, vp]
        lib.ref_sim_update.restype = u32
        lib.ref_sim_update.argtypes = [vp, vp, vp, f32, f32]
        self.lib = lib
        nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1)
        idx_stream = np.ascontiguousarray(idx_stream, dtype=np.uint32).reshape(-1)
        self.nV = nodes.size // 3
        self.h = C.c_void_p(lib.ref_sim_create(_vp(nodes), nodes.size, _vp(idx_stream), idx_stream.size, C.byref(settings), 1 if auto_resize else 0))

    def close(self):
        if self.h:
            self.lib.ref_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_order(self, order):
        order = np.ascontiguousarray(order, dtype=np.uint32)
        self.lib.ref_sim_set_order(self.h, _vp(order))

    def update(self, settings, dt, median_frame_time, manip=None):
        return self.lib.ref_sim_update(self.h, C.byref(settings), C.byref(manip) if manip is not None else None, dt, median_frame_time)

    def get_state(self):
        X = np.empty((self.nV, 3), dtype=np.float64)
        V = np.empty((self.nV, 3), dtype=np.float64)
        w = np.empty(self.nV, dtype=np.float32)
        self.lib.ref_sim_get_state(self.h, _vp(X), _vp(V), _vp(w))
        return X, V, w


# ---------------------------------------------------------------------------------------------
# The product's reference-side adapter (GeoLinear3dCuda : Geo) hosted by the reference harness
class RefMultiSim:
    """The reference's Sim holding several T4 blocks (Sim::AddBlock / FinishAddingBlocks / Update), or the block
    Demo::UpdateSettings builds from a Settings block (from_settings: shape table of Demo.cpp:289-318)."""

    def __init__(self, settings, kind="strict", from_settings=False):
        lib = _load_ref(kind)
        vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
        lib.ref_multi_create.restype = vp
        lib.ref_multi_create.argtypes = [vp]
        lib.ref_multi_from_settings.restype = vp
        lib.ref_multi_from_settings.argtypes = [vp]
        lib.ref_multi_destroy.argtypes = [vp]
        lib.ref_multi_add_block.argtypes = [vp, vp, u32, vp, u32, C.c_int]
        lib.ref_multi_finish.argtypes = [vp]
        lib.ref_multi_set_geo_offset.argtypes = [vp, f32, f32]
        lib.ref_multi_geo_count.restype = u32
        lib.ref_multi_geo_count.argtypes = [vp]
        lib.ref_multi_geo_sizes.argtypes = [vp, u32, vp, vp]
        lib.ref_multi_volume0.restype = f32
        lib.ref_multi_volume0.argtypes = [vp, u32]
        lib.ref_multi_set_order.argtypes = [vp, u32, vp]
        lib.ref_multi_get_mesh.argtypes = [vp, u32, vp, vp]
        lib.ref_multi_get_state.argtypes = [vp, u32, vp, vp, vp]
        lib.ref_multi_update.restype = u32
        lib.ref_multi_update.argtypes = [vp, vp, vp, C.c_int, f32, f32]
        self.lib = lib
        self.h = C.c_void_p(lib.ref_multi_from_settings(C.byref(settings)) if from_settings else lib.ref_multi_create(C.byref(settings)))

    def close(self):
        if self.h:
            self.lib.ref_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_block(self, nodes, idx_stream, auto_resize=False):
        nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1)
        idx_stream = np.ascontiguousarray(idx_stream, dtype=np.uint32).reshape(-1)
        self.lib.ref_multi_add_block(self.h, _vp(nodes), nodes.size, _vp(idx_stream), idx_stream.size, 1 if auto_resize else 0)

    def finish(self):
        self.lib.ref_multi_finish(self.h)

    def set_geo_offset(self, x, y):
        self.lib.ref_multi_set_geo_offset(self.h, x, y)

    def geo_count(self):
        return int(self.lib.ref_multi_geo_count(self.h))

    def sizes(self, geo):
        nv, nt = C.c_uint32(), C.c_uint32()
        self.lib.ref_multi_geo_sizes(self.h, geo, C.byref(nv), C.byref(nt))
        return nv.value, nt.value

    def volume0(self, geo):
        return float(self.lib.ref_multi_volume0(self.h, geo))

    def set_order(self, geo, order):
        order = np.ascontiguousarray(order, dtype=np.uint32)
        self.lib.ref_multi_set_order(self.h, geo, _vp(order))

    def get_mesh(self, geo):
        nv, nt = self.sizes(geo)
        X0 = np.empty((nv, 3), dtype=np.float64)
        idx = np.empty((nt, 4), dtype=np.uint32)
        self.lib.ref_multi_get_mesh(self.h, geo, _vp(X0), _vp(idx))
        return X0, idx

    def get_state(self, geo):
        nv, _ = self.sizes(geo)
        X = np.empty((nv, 3), dtype=np.float64)
        V = np.empty((nv, 3), dtype=np.float64)
        w = np.empty(nv, dtype=np.float32)
        self.lib.ref_multi_get_state(self.h, geo, _vp(X), _vp(V), _vp(w))
        return X, V, w

    def update(self, settings, dt, median_frame_time, manip=None, picked_geo=-1):
        return self.lib.ref_multi_update(self.h, C.byref(settings), C.byref(manip) if manip is not None else None, picked_geo, dt, median_frame_time)


# ---------------------------------------------------------------------------------------------
class AdapterScene:
    """oracle/_ref/libxpbd_ref_adapter.so: the reference's `Geo` virtual interface, implemented by the CUDA library."""

    def __init__(self, nodes, idx_stream, density=1.0, auto_resize=False, color_hint=None):
        lib = C.CDLL(ref_lib_path("adapter"))
        vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
        lib.ref_adapter_create.restype = vp
        lib.ref_adapter_create.argtypes = [vp, u32, vp, u32, f32, C.c_int, vp]
        lib.ref_adapter_destroy.argtypes = [vp]
        lib.ref_adapter_vert_count.restype = u32
        lib.ref_adapter_vert_count.argtypes = [vp]
        lib.ref_adapter_element_count.restype = u32
        lib.ref_adapter_element_count.argtypes = [vp]
        lib.ref_adapter_get_order.argtypes = [vp, vp]
        lib.ref_adapter_substep.argtypes = [vp, vp, vp, f32, u32]
        lib.ref_adapter_volume.restype = f32
        lib.ref_adapter_volume.argtypes = [vp]
        lib.ref_adapter_transform.argtypes = [vp, vp]
        lib.ref_adapter_get_state.argtypes = [vp, vp, vp, vp]
        lib.ref_adapter_pick.argtypes = [vp, vp, vp, vp, vp]
        self.lib = lib
        nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1)
        idx_stream = np.ascontiguousarray(idx_stream, dtype=np.uint32).reshape(-1)
        hint = None if color_hint is None else np.ascontiguousarray(color_hint, dtype=np.uint32)
        h = lib.ref_adapter_create(_vp(nodes), nodes.size, _vp(idx_stream), idx_stream.size, density, 1 if auto_resize else 0, _vp(hint))
        if not h:
            raise RuntimeError("GeoLinear3dCuda::Init failed (no CUDA device?)")
        self.h = C.c_void_p(h)
        self.nV = lib.ref_adapter_vert_count(self.h)
        self.nT = lib.ref_adapter_element_count(self.h)

    def close(self):
        if self.h:
            self.lib.ref_adapter_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_order(self):
        o = np.empty(self.nT, dtype=np.uint32)
        self.lib.ref_adapter_get_order(self.h, _vp(o))
        return o

    def substep(self, settings, dt, n=1, manip=None):
        self.lib.ref_adapter_substep(self.h, C.byref(settings), C.byref(manip) if manip is not None else None, dt, n)

    def volume(self):
        return float(self.lib.ref_adapter_volume(self.h))

    def pick(self, ray_origin, ray_dir):
        """Geo::Pick through the virtual interface -> (found, vertex index, point xyz, distance)."""
        return _pick(self.lib.ref_adapter_pick, self.h, ray_origin, ray_dir)

    def transform(self, m9):
        m9 = np.ascontiguousarray(m9, dtype=np.float32).reshape(9)
        self.lib.ref_adapter_transform(self.h, _vp(m9))

    def get_state(self):
        X = np.empty((self.nV, 3), dtype=np.float64)
        V = np.empty((self.nV, 3), dtype=np.float64)
        w = np.empty(self.nV, dtype=np.float32)
        self.lib.ref_adapter_get_state(self.h, _vp(X), _vp(V), _vp(w))
        return X, V, w
