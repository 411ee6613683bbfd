"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from raytracinggpu_b200._abi import rt_params, rt_sphere  # noqa: E402  (struct layouts only)

ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
REF_CPU_SO = os.path.join(REF_DIR, "libref_cpu.so")
REF_CPU_BIN = os.path.join(REF_DIR, "cpu")
REF_GPU_BIN = os.path.join(REF_DIR, "ref_optimized")
CAT_REL = os.path.join("cadnav.com_model", "Models_F0202A090", "cat.obj")


def build(ref=True):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    targets = ["liboracle.so"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", _HERE] + targets, check=True, stdout=subprocess.DEVNULL)


def cat_obj_path():
    """The cat mesh: the reference tree here, the git-ignored copy under oracle/_ref on the GPU box."""
    for base in (os.environ.get("RT_REFERENCE_DIR", "/root/reference"), REF_DIR):
        p = os.path.join(base, CAT_REL)
        if os.path.exists(p):
            return p
    return None


class orc_work(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64),
        ("node_visits", C.c_uint64),
        ("tri_tests", C.c_uint64),
        ("root_miss", C.c_uint64),
        ("mesh_queries", C.c_uint64),
        ("max_stack", C.c_int32),
        ("threads", C.c_int32),
        ("seconds", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.orc_last_error.restype = C.c_char_p
        L.orc_mesh_create.argtypes = [C.POINTER(C.c_void_p)]
        L.orc_mesh_destroy.argtypes = [C.c_void_p]
        L.orc_mesh_read_obj.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_mesh_set_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.orc_mesh_rescale.argtypes = [C.c_void_p, C.c_float, C.POINTER(C.c_float)]
        L.orc_mesh_build_bvh.argtypes = [C.c_void_p]
        L.orc_mesh_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
        L.orc_mesh_bvh_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
        for f in ("orc_mesh_vertices", "orc_mesh_tri_records", "orc_mesh_arr_bvh"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_void_p
        L.orc_render.argtypes = [
            C.c_void_p, C.c_int32,
            C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
            C.POINTER(C.c_float), C.c_int32, C.c_float, C.c_float, C.c_int32,
            C.POINTER(C.c_float), C.c_float, C.POINTER(rt_params),
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.POINTER(orc_work), C.c_int32,
        ]
        L.orc_quantise.argtypes = [C.c_float, C.c_int32]
        L.orc_camera_z.argtypes = [C.c_int32, C.c_float]
        L.orc_camera_z.restype = C.c_float
        L.orc_camera_z_device.argtypes = [C.c_int32, C.c_float]
        L.orc_camera_z_device.restype = C.c_float
        L.orc_set_mesh_normals.argtypes = [C.c_void_p]
        L.orc_set_mesh_normals.restype = None
        L.orc_mesh_set_normals.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.orc_camera_basis.argtypes = [C.c_float, C.c_float, C.c_void_p]
        L.orc_camera_basis.restype = None
        L.orc_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_accumulate.restype = None
        L.orc_set_transcendentals.argtypes = [C.c_int32]
        L.orc_set_transcendentals.restype = None
        L.orc_cuda_libm.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.orc_cuda_libm.restype = None
        _lib = L
    return _lib


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _np_from(ptr, shape, dtype):
    n = int(np.prod(shape))
    if not ptr or n == 0:
        return np.zeros(shape, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


class Mesh:
    """Host mesh in the reference interchange formats: vertices (nv,3) f32, tri_records (nt,10) i32
    (post-build order), arr_bvh (nn,10) f32."""

    def __init__(self):
        self._h = C.c_void_p()
        lib().orc_mesh_create(C.byref(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_mesh_destroy(self._h)
            self._h = None

    @classmethod
    def from_obj(cls, path, rescale=None):
        m = cls()
        rc = lib().orc_mesh_read_obj(m._h, path.encode())
        if rc != 0:
            raise IOError(lib().orc_last_error().decode())
        if rescale is not None:
            m.rescale(*rescale)
        return m

    @classmethod
    def from_arrays(cls, vertices, vtx_indices):
        m = cls()
        v = np.ascontiguousarray(vertices, dtype=np.float32)
        t = np.ascontiguousarray(vtx_indices, dtype=np.int32)
        lib().orc_mesh_set_triangles(m._h, v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0])
        return m

    def rescale(self, scale, offset):
        lib().orc_mesh_rescale(self._h, float(scale), _f3(offset))
        return self

    def set_normals(self, normals, normal_indices):
        """Per-vertex normals + (nt, 3) normal indices (words 6-8 of the records); before build_bvh."""
        self.normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(normal_indices, dtype=np.int32).reshape(-1, 3)
        lib().orc_mesh_set_normals(self._h, self.normals.ctypes.data, self.normals.shape[0], idx.ctypes.data)
        return self

    def build_bvh(self):
        lib().orc_mesh_build_bvh(self._h)
        return self

    def counts(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        lib().orc_mesh_counts(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def bvh_info(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        lib().orc_mesh_bvh_info(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"leaves": a.value, "max_depth": b.value, "max_leaf": c.value}

    @property
    def vertices(self):
        nv, _, _ = self.counts()
        return _np_from(lib().orc_mesh_vertices(self._h), (nv, 3), np.float32)

    @property
    def tri_records(self):
        _, nt, _ = self.counts()
        return _np_from(lib().orc_mesh_tri_records(self._h), (nt, 10), np.int32)

    @property
    def arr_bvh(self):
        _, _, nn = self.counts()
        return _np_from(lib().orc_mesh_arr_bvh(self._h), (nn, 10), np.float32)


def camera_basis(yaw, pitch):
    """Camera::rotate (realtime_render.cu:828-849): (bx, by, bz) as three float32 triples."""
    out = np.zeros(9, np.float32)
    lib().orc_camera_basis(float(yaw), float(pitch), out.ctypes.data)
    return out[0:3].copy(), out[3:6].copy(), out[6:9].copy()


def accumulate(acc, linear, k, gamma_mode):
    """Progressive accumulation step (realtime_render.cu:1136-1140): acc (float32, in place) += linear; returns the 8-bit frame."""
    rgb = np.zeros(linear.shape, np.uint8)
    lin = np.ascontiguousarray(linear, dtype=np.float32)
    lib().orc_accumulate(acc.ctypes.data, lin.ctypes.data, lin.size, int(k), int(gamma_mode), rgb.ctypes.data)
    return rgb


def render(spheres, mesh_arrays, mesh_mat, light, params, threads=0, want=("rgb", "hit_obj", "hit_tri", "hit_t", "shadow"), normals=None):
    """Run the oracle. spheres: list of rt_sphere. mesh_arrays: None or (vertices, tri_records, arr_bvh).
    mesh_mat: dict(albedo, mirror, n_in, n_out, id). light: (L, intensity). normals: (nn, 3) per-vertex normals for
    params.smooth_normals. Returns dict of numpy arrays + 'work'."""
    L = lib()
    nrm = np.ascontiguousarray(normals, dtype=np.float32) if normals is not None else None
    L.orc_set_mesh_normals(nrm.ctypes.data if nrm is not None else None)
    p = params
    step = p.row_step if p.row_step > 0 else 1
    group = p.row_group if p.row_group > 1 else 1
    n_groups = (p.H - p.row_begin + step - 1) // step
    rows = p.row_count if p.row_count > 0 else (n_groups - 1) * group + min(group, p.H - (p.row_begin + (n_groups - 1) * step))
    n = len(spheres)
    arr = (rt_sphere * max(n, 1))(*spheres)
    if mesh_arrays is not None:
        v = np.ascontiguousarray(mesh_arrays[0], dtype=np.float32)
        t = np.ascontiguousarray(mesh_arrays[1], dtype=np.int32)
        b = np.ascontiguousarray(mesh_arrays[2], dtype=np.float32)
        vp, nv, tp, nt, bp, nn = v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], b.ctypes.data, b.shape[0]
    else:
        vp = tp = bp = None
        nv = nt = nn = 0
    mm = mesh_mat or dict(albedo=(0, 0, 0), mirror=0, n_in=1.0, n_out=1.0, id=-1)
    out = {}
    shapes = {"rgb": ((rows, p.W, 3), np.uint8), "hit_obj": ((rows, p.W), np.int32), "hit_tri": ((rows, p.W), np.int32),
              "hit_t": ((rows, p.W), np.float32), "shadow": ((rows, p.W), np.uint8), "linear": ((rows, p.W, 3), np.float32)}
    ptrs = {}
    for k, (shape, dt) in shapes.items():
        if k in want:
            out[k] = np.zeros(shape, dtype=dt)
            ptrs[k] = out[k].ctypes.data
        else:
            ptrs[k] = None
    work = orc_work()
    rc = L.orc_render(arr, n, vp, nv, tp, nt, bp, nn, _f3(mm["albedo"]), int(mm["mirror"]), float(mm["n_in"]), float(mm["n_out"]),
                      int(mm["id"]), _f3(light[0]), float(light[1]), C.byref(p),
                      ptrs["rgb"], ptrs["hit_obj"], ptrs["hit_tri"], ptrs["hit_t"], ptrs["shadow"], ptrs["linear"],
                      C.byref(work), int(threads))
    if rc != 0:
        raise RuntimeError("oracle: %s" % L.orc_last_error().decode())
    out["work"] = {f[0]: getattr(work, f[0]) for f in orc_work._fields_}
    return out


# ---- the compiled reference (oracle/_ref) ---------------------------------------------------------------
_ref = None


def camera_z(W, alpha=np.float32(np.pi / 3)):
    """z as the host evaluates it (cpu_launcher.cpp:666,694)."""
    return lib().orc_camera_z(int(W), float(alpha))


def camera_z_device(W, alpha=np.float32(np.pi / 3)):
    """z as optimized.cu's KERNEL evaluates it (CUDA's tanf, restated in the oracle): rt_camera_z_device's counterpart."""
    return lib().orc_camera_z_device(int(W), float(alpha))


def set_transcendentals(mode):
    """1: CUDA's logf / cosf / sinf restated (default: what optimized.cu calls, and the library's default); 0: double-evaluated."""
    lib().orc_set_transcendentals(int(mode))


def cuda_libm(which, x):
    """The oracle's restatement of CUDA's logf / sinf / cosf / tanf (which = 'log' | 'sin' | 'cos' | 'tan')."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().orc_cuda_libm({"log": 0, "sin": 1, "cos": 2, "tan": 3}[which], x.ctypes.data, x.size, y.ctypes.data)
    return y


def xorwow(seed, subsequence, n):
    """n curand_uniform values of XORWOW subsequence `subsequence` + the state right after curand_init (d, v0..v4)."""
    L = lib()
    L.orc_xorwow.restype = C.c_int
    L.orc_xorwow.argtypes = [C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p]
    u = np.zeros(n, np.float32)
    st = np.zeros(6, np.uint32)
    L.orc_xorwow(int(seed), int(subsequence), int(n), u.ctypes.data, st.ctypes.data)
    return u, st


def ref_cpu_available():
    return os.path.exists(REF_CPU_SO)


def ref_lib():
    global _ref
    if _ref is None:
        R = C.CDLL(REF_CPU_SO)
        R.ref_cpu_render.argtypes = [C.c_char_p] + [C.c_int] * 6 + [C.c_void_p] * 4 + [C.POINTER(C.c_double)]
        R.ref_cpu_mesh.argtypes = [C.c_char_p, C.c_int] + [C.POINTER(C.c_int)] * 6 + [C.c_void_p] * 3
        _ref = R
    return _ref


def ref_cpu_render(obj_path, scene_kind, W, H, num_rays, num_bounce, threads=0, hits=True):
    """The reference's own classes (cpu_launcher.cpp compiled from /root/reference) rendering W x H."""
    R = ref_lib()
    rgb = np.zeros((H, W, 3), np.uint8)
    obj = np.zeros((H, W), np.int32) if hits else None
    P = np.zeros((H, W, 3), np.float32) if hits else None
    N = np.zeros((H, W, 3), np.float32) if hits else None
    sec = C.c_double()
    R.ref_cpu_render((obj_path or "").encode(), scene_kind, W, H, num_rays, num_bounce, threads, rgb.ctypes.data,
                     obj.ctypes.data if hits else None, P.ctypes.data if hits else None, N.ctypes.data if hits else None, C.byref(sec))
    return {"rgb": rgb, "hit_obj": obj, "P": P, "N": N, "seconds": sec.value}


def ref_cpu_set_light(L):
    """Scene::L of the scenes the following ref_cpu_render calls build."""
    R = ref_lib()
    R.ref_cpu_set_light.argtypes = [C.c_float, C.c_float, C.c_float]
    R.ref_cpu_set_light.restype = None
    R.ref_cpu_set_light(float(L[0]), float(L[1]), float(L[2]))


def ref_cpu_mesh(obj_path, scene_kind):
    R = ref_lib()
    vals = [C.c_int() for _ in range(6)]
    R.ref_cpu_mesh(obj_path.encode(), scene_kind, *[C.byref(v) for v in vals], None, None, None)
    nv, nt, nn, leaves, depth, max_leaf = [v.value for v in vals]
    v = np.zeros((nv, 3), np.float32)
    t = np.zeros((nt, 3), np.int32)
    b = np.zeros((nn, 10), np.float32)
    R.ref_cpu_mesh(obj_path.encode(), scene_kind, *[C.byref(x) for x in vals], v.ctypes.data, t.ctypes.data, b.ctypes.data)
    return {"vertices": v, "vtx_indices": t, "arr_bvh": b, "leaves": leaves, "max_depth": depth, "max_leaf": max_leaf}
