"""ctypes binding of librtb200.so (include/rt_b200.h) — the host-side mirror used by tests and bench.py.

The library is the product: this module only marshals numpy arrays / device pointers into the C ABI. It
fails loudly when the shared library is missing (run `python -c "import __graft_entry__ as g; g.build()"`),
and the library itself fails with RT_ERR_CUDA when there is no CUDA device. There is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from ._abi import (RT_BVH_NODE_FLOATS, RT_OK, RT_RENDER_COUNT_WORK, RT_RENDER_NO_SYNC, RT_TRI_RECORD_WORDS, rt_params,
                   rt_sphere, rt_stats)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_LIB_PATH") or os.path.join(_HERE, "librtb200.so")  # RT_LIB_PATH: A/B runs of two builds in one session

# every symbol include/rt_b200.h declares: (restype, argtypes)
_vp, _i32, _f, _u64 = C.c_void_p, C.c_int32, C.c_float, C.c_uint64
_pi32, _pf = C.POINTER(C.c_int32), C.POINTER(C.c_float)
SIGNATURES = {
    "rt_last_error": (C.c_char_p, []),
    "rt_abi_version": (C.c_int, []),
    "rt_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rt_mesh_create": (C.c_int, [C.POINTER(_vp)]),
    "rt_mesh_destroy": (None, [_vp]),
    "rt_mesh_read_obj": (C.c_int, [_vp, C.c_char_p]),
    "rt_mesh_set_triangles": (C.c_int, [_vp, _vp, _i32, _vp, _i32]),
    "rt_mesh_rescale": (C.c_int, [_vp, _f, _pf]),
    "rt_mesh_keep_normals": (C.c_int, [_vp, C.c_int]),
    "rt_mesh_set_normals": (C.c_int, [_vp, _vp, _i32, _vp]),
    "rt_mesh_normal_count": (C.c_int, [_vp, _pi32]),
    "rt_mesh_normals": (_vp, [_vp]),
    "rt_camera_basis": (None, [_f, _f, _pf, _pf, _pf]),
    "rt_scene_set_mesh_normals": (C.c_int, [_vp, _vp, _i32]),
    "rt_mesh_instance": (C.c_int, [_vp, _i32, _vp, _vp]),
    "rt_mesh_build_bvh": (C.c_int, [_vp]),
    "rt_mesh_build_bvh_gpu": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "rt_mesh_counts": (C.c_int, [_vp, _pi32, _pi32, _pi32]),
    "rt_mesh_vertices": (_vp, [_vp]),
    "rt_mesh_tri_records": (_vp, [_vp]),
    "rt_mesh_arr_bvh": (_vp, [_vp]),
    "rt_mesh_bvh_info": (C.c_int, [_vp, _pi32, _pi32, _pi32]),
    "rt_camera_z": (_f, [_i32, _f]),
    "rt_camera_z_device": (C.c_int, [C.c_int, _i32, _f, _pf]),
    "rt_params_profile": (C.c_int, [C.POINTER(rt_params), C.c_char_p, _i32, _i32, _i32, _i32]),
    "rt_default_walls": (C.c_int, [C.POINTER(rt_sphere), C.c_char_p, _pi32]),
    "rt_write_png": (C.c_int, [C.c_char_p, _i32, _i32, _vp]),
    "rt_peer_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(_vp), _vp]),
    "rt_peer_open": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "rt_peer_close": (C.c_int, [C.c_int, _vp]),
    "rt_peer_free": (C.c_int, [C.c_int, _vp]),
    "rt_scene_push_rows": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32]),
    "rt_peer_signal": (C.c_int, [_vp, _vp, C.c_uint32]),
    "rt_peer_wait": (C.c_int, [_vp, _vp, C.c_uint32]),
    "rt_scene_push_row_groups": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32]),
    "rt_shard_rows": (C.c_int, [_i32, _i32, _i32, _i32, C.POINTER(rt_params)]),
    "rt_png_writer_create": (C.c_int, [C.POINTER(_vp), _i32, _i32]),
    "rt_png_writer_submit": (C.c_int, [_vp, C.c_char_p, _i32, _i32, _vp]),
    "rt_png_writer_wait": (C.c_int, [_vp]),
    "rt_png_writer_destroy": (None, [_vp]),
    "rt_move_light": (None, [_pf, _f, _f]),
    "rt_scene_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "rt_scene_destroy": (None, [_vp]),
    "rt_scene_set_stream": (C.c_int, [_vp, _vp]),
    "rt_scene_get_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "rt_scene_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "rt_scene_get_option": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int64)]),
    "rt_scene_set_spheres": (C.c_int, [_vp, C.POINTER(rt_sphere), _i32]),
    "rt_scene_set_mesh": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _i32, _pf, _i32, _f, _f, _i32]),
    "rt_scene_set_mesh_device": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _i32, _pf, _i32, _f, _f, _i32]),
    "rt_scene_set_mesh_from": (C.c_int, [_vp, _vp, _pf, _i32, _f, _f, _i32]),
    "rt_scene_set_light": (C.c_int, [_vp, _pf, _f]),
    "rt_scene_blob_size": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "rt_scene_blob_export": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "rt_scene_blob_import": (C.c_int, [_vp, _vp, C.c_size_t]),
    "rt_scene_blob_copy_out": (C.c_int, [_vp, _vp, C.c_size_t]),
    "rt_render": (C.c_int, [_vp, C.POINTER(rt_params), C.c_uint32, _vp, _vp, _vp, _vp, _vp, C.POINTER(rt_stats)]),
    "rt_scene_sync": (C.c_int, [_vp, C.POINTER(rt_stats)]),
    "rt_comm_available": (C.c_int, [C.POINTER(C.c_int)]),
    "rt_comm_unique_id": (C.c_int, [_vp]),
    "rt_comm_init": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp, C.c_int]),
    "rt_comm_init_all": (C.c_int, [C.POINTER(_vp), C.c_int, C.POINTER(C.c_int)]),
    "rt_comm_destroy": (None, [_vp]),
    "rt_comm_rank": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rt_scene_broadcast": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_size_t)]),
    "rt_gather_framebuffer": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, C.c_int]),
    "rt_gather_framebuffer_groups": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, C.c_int]),
    "rt_selftest_fma_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "rt_selftest_libm": (C.c_int, [C.c_int, C.c_int, _vp, _i32, _vp]),
    "rt_selftest_division": (C.c_int, [C.c_int, _u64, C.c_int, C.c_int, C.POINTER(_u64)]),
    "rt_selftest_division3": (C.c_int, [C.c_int, _u64, C.c_int, C.c_int, C.POINTER(_u64)]),
    "rt_selftest_xorwow": (C.c_int, [C.c_int, _u64, _vp, C.c_int32, _vp, _vp]),
}
# rt_default_walls(profile, walls, mesh_id): fix the argument order to the header's
SIGNATURES["rt_default_walls"] = (C.c_int, [C.c_char_p, C.POINTER(rt_sphere), _pi32])

_lib = None


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rt error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load librtb200.so (once). Raises if it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with __graft_entry__.build() (make -C raytracinggpu_b200/csrc). "
                              "raytracinggpu_b200 has no fallback implementation." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != RT_OK:
        raise RtError(rc, lib().rt_last_error().decode())


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if not ptr or n == 0:
        return np.zeros(shape, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


def device_count():
    n = C.c_int()
    rc = lib().rt_device_count(C.byref(n))
    return n.value if rc == RT_OK else 0


def camera_z(W, alpha=np.float32(np.pi / 3)):
    return lib().rt_camera_z(int(W), float(alpha))


def camera_basis(yaw, pitch):
    """Camera::rotate of the viewer (realtime_render.cu:828-849): (bx, by, bz) as float32 triples."""
    b = [(C.c_float * 3)() for _ in range(3)]
    lib().rt_camera_basis(float(yaw), float(pitch), b[0], b[1], b[2])
    return tuple(np.array(list(x), np.float32) for x in b)


def camera_z_device(W, alpha=np.float32(np.pi / 3), device=0):
    """z as optimized.cu's kernel evaluates it (CUDA tanf on the device), see rt_camera_z_device."""
    z = C.c_float()
    _check(lib().rt_camera_z_device(int(device), int(W), float(alpha), C.byref(z)))
    return z.value


def shard_rows(params, rank, nranks, row_group=1):
    """rt_shard_rows: set row_begin / row_step / row_group / row_count of `params` for `rank` of `nranks`; returns the number of rows."""
    n = lib().rt_shard_rows(int(params.H), int(rank), int(nranks), int(row_group), C.byref(params))
    if n < 0:
        _check(n)
    return n


def shard_row_count(p):
    """Rows a call with these rt_params renders (row_count, or every row the shard holds when it is 0)."""
    if p.row_count > 0:
        return p.row_count
    step = p.row_step if p.row_step > 0 else 1
    group = p.row_group if p.row_group > 1 else 1
    n_groups = (p.H - p.row_begin + step - 1) // step
    return (n_groups - 1) * group + min(group, p.H - (p.row_begin + (n_groups - 1) * step))


def params_profile(profile, W, H, num_rays=1, num_bounce=1):
    p = rt_params()
    _check(lib().rt_params_profile(C.byref(p), profile.encode(), W, H, num_rays, num_bounce))
    return p


def default_walls(profile):
    walls = (rt_sphere * 6)()
    mesh_id = C.c_int32()
    _check(lib().rt_default_walls(profile.encode(), walls, C.byref(mesh_id)))
    return list(walls), mesh_id.value


def write_png(path, rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    H, W, _ = rgb.shape
    _check(lib().rt_write_png(path.encode(), W, H, rgb.ctypes.data))


def peer_alloc(device, nbytes):
    """(device pointer, 64-byte CUDA IPC handle) of a buffer other processes can open with peer_open."""
    p, h = C.c_void_p(), (C.c_uint8 * 64)()
    _check(lib().rt_peer_alloc(int(device), int(nbytes), C.byref(p), h))
    return p.value, bytes(h)


def peer_open(device, handle):
    p = C.c_void_p()
    buf = (C.c_uint8 * 64).from_buffer_copy(handle)
    _check(lib().rt_peer_open(int(device), buf, C.byref(p)))
    return p.value


def peer_close(device, ptr):
    _check(lib().rt_peer_close(int(device), C.c_void_p(int(ptr))))


def peer_free(device, ptr):
    _check(lib().rt_peer_free(int(device), C.c_void_p(int(ptr))))


class PngWriter:
    """Asynchronous PNG output (rt_png_writer_*): submit() copies the frame and returns, a background thread encodes and writes."""

    def __init__(self, threads=0, max_pending=0):
        self._h = C.c_void_p()
        _check(lib().rt_png_writer_create(C.byref(self._h), int(threads), int(max_pending)))

    def submit(self, path, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        H, W, _ = rgb.shape
        _check(lib().rt_png_writer_submit(self._h, path.encode(), W, H, rgb.ctypes.data))

    def wait(self):
        _check(lib().rt_png_writer_wait(self._h))

    def close(self):
        if self._h:
            lib().rt_png_writer_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def move_light(L, angular_speed, dt=0.02):
    v = _f3(L)
    lib().rt_move_light(v, float(angular_speed), float(dt))
    return (v[0], v[1], v[2])


class Mesh:
    """Host mesh: the TriangleMeshHost surface (readOBJ / rescale / buildBVH / bvhTreeToArray)."""

    def __init__(self):
        self._h = C.c_void_p()
        _check(lib().rt_mesh_create(C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rt_mesh_destroy(self._h)
            self._h = None

    @classmethod
    def read_obj(cls, path, keep_normals=False):
        """readOBJ. keep_normals: the viewer's loader (realtime_render.cu:489-493, 538-545) — `vn` lines + normal indices kept."""
        m = cls()
        if keep_normals:
            _check(lib().rt_mesh_keep_normals(m._h, 1))
        _check(lib().rt_mesh_read_obj(m._h, path.encode()))
        return m

    def set_normals(self, normals, normal_indices):
        n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(normal_indices, dtype=np.int32).reshape(-1, 3)
        _check(lib().rt_mesh_set_normals(self._h, n.ctypes.data, n.shape[0], i.ctypes.data))
        return self

    @property
    def normals(self):
        n = C.c_int32()
        _check(lib().rt_mesh_normal_count(self._h, C.byref(n)))
        return _view(lib().rt_mesh_normals(self._h), (n.value, 3), np.float32)

    @classmethod
    def from_arrays(cls, vertices, vtx_indices):
        m = cls()
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(vtx_indices, dtype=np.int32).reshape(-1, 3)
        _check(lib().rt_mesh_set_triangles(m._h, v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0]))
        return m

    def rescale(self, scale, offset):
        _check(lib().rt_mesh_rescale(self._h, float(scale), _f3(offset)))
        return self

    def instance(self, scales, offsets):
        s = np.ascontiguousarray(scales, dtype=np.float32)
        o = np.ascontiguousarray(offsets, dtype=np.float32).reshape(-1, 3)
        assert s.shape[0] == o.shape[0]
        _check(lib().rt_mesh_instance(self._h, s.shape[0], s.ctypes.data, o.ctypes.data))
        return self

    def build_bvh(self):
        _check(lib().rt_mesh_build_bvh(self._h))
        return self

    def build_bvh_gpu(self, device=0):
        """The reference's builder on the device (identical tree); self.build_ms = device time of the build."""
        ms = C.c_double()
        _check(lib().rt_mesh_build_bvh_gpu(self._h, int(device), C.byref(ms)))
        self.build_ms = ms.value
        return self

    def counts(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        _check(lib().rt_mesh_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def bvh_info(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        _check(lib().rt_mesh_bvh_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"leaves": a.value, "max_depth": b.value, "max_leaf": c.value}

    @property
    def vertices(self):
        return _view(lib().rt_mesh_vertices(self._h), (self.counts()[0], 3), np.float32)

    @property
    def tri_records(self):
        return _view(lib().rt_mesh_tri_records(self._h), (self.counts()[1], RT_TRI_RECORD_WORDS), np.int32)

    @property
    def arr_bvh(self):
        return _view(lib().rt_mesh_arr_bvh(self._h), (self.counts()[2], RT_BVH_NODE_FLOATS), np.float32)


def _ptr(x):
    """None, an int device pointer, a numpy array or anything with data_ptr() (a torch tensor)."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


class Scene:
    """Device-resident scene (the Scene object + uploaded mesh of optimized.cu:679-726, 814-826)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().rt_scene_create(C.byref(self._h), int(device)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rt_scene_destroy(self._h)
            self._h = None

    __del__ = close

    def set_stream(self, cuda_stream):
        _check(lib().rt_scene_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def get_stream(self):
        """The cudaStream_t (as an int) every call on this scene enqueues on."""
        p = C.c_void_p()
        _check(lib().rt_scene_get_stream(self._h, C.byref(p)))
        return p.value or 0

    def set_option(self, key, value):
        """rt_scene_set_option: tuning / cross-check options are scene state (never read from the environment per call)."""
        _check(lib().rt_scene_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key):
        v = C.c_int64()
        _check(lib().rt_scene_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    def set_spheres(self, spheres):
        arr = (rt_sphere * max(1, len(spheres)))(*spheres)
        _check(lib().rt_scene_set_spheres(self._h, arr, len(spheres)))

    def set_light(self, L, intensity):
        _check(lib().rt_scene_set_light(self._h, _f3(L), float(intensity)))

    def set_mesh(self, vertices, tri_records, arr_bvh, albedo=(0.25, 0.25, 0.25), mirror=0, n_in=1.0, n_out=1.0, id=1):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(tri_records, dtype=np.int32).reshape(-1, RT_TRI_RECORD_WORDS)
        b = np.ascontiguousarray(arr_bvh, dtype=np.float32).reshape(-1, RT_BVH_NODE_FLOATS)
        _check(lib().rt_scene_set_mesh(self._h, v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], b.ctypes.data, b.shape[0],
                                       _f3(albedo), int(mirror), float(n_in), float(n_out), int(id)))
        self.mesh_h2d_bytes = v.nbytes + t.nbytes + b.nbytes

    def set_mesh_from(self, mesh, albedo=(0.25, 0.25, 0.25), mirror=0, n_in=1.0, n_out=1.0, id=1):
        """The mesh of a Mesh handle: from the device when build_bvh_gpu left its arrays there (no host round trip), else from the host."""
        _check(lib().rt_scene_set_mesh_from(self._h, mesh._h, _f3(albedo), int(mirror), float(n_in), float(n_out), int(id)))

    def set_mesh_device(self, d_vertices, nv, d_tri_records, nt, d_arr_bvh, n_nodes, albedo=(0.25, 0.25, 0.25), mirror=0, n_in=1.0, n_out=1.0, id=1):
        """rt_scene_set_mesh_device: the three interchange arrays as DEVICE pointers (ints)."""
        _check(lib().rt_scene_set_mesh_device(self._h, C.c_void_p(int(d_vertices)), int(nv), C.c_void_p(int(d_tri_records)), int(nt), C.c_void_p(int(d_arr_bvh)), int(n_nodes),
                                              _f3(albedo), int(mirror), float(n_in), float(n_out), int(id)))

    def set_mesh_normals(self, normals):
        """Per-vertex normals for rt_params.smooth_normals (after set_mesh; indices are words 6-8 of its triangle records)."""
        n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3) if normals is not None else np.zeros((0, 3), np.float32)
        _check(lib().rt_scene_set_mesh_normals(self._h, n.ctypes.data if n.shape[0] else None, n.shape[0]))

    def clear_mesh(self):
        _check(lib().rt_scene_set_mesh(self._h, None, 0, None, 0, None, 0, _f3((0, 0, 0)), 0, 1.0, 1.0, 0))

    def blob_export(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib().rt_scene_blob_export(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def blob_size(self):
        n = C.c_size_t()
        _check(lib().rt_scene_blob_size(self._h, C.byref(n)))
        return n.value

    def blob_copy_out(self, device_ptr, nbytes):
        _check(lib().rt_scene_blob_copy_out(self._h, C.c_void_p(int(device_ptr)), int(nbytes)))

    def blob_import(self, device_ptr, nbytes):
        _check(lib().rt_scene_blob_import(self._h, C.c_void_p(int(device_ptr)), int(nbytes)))

    def push_rows(self, band_ptr, frame_ptr, W, bytes_per_pixel, row_begin, row_step, rows, row_group=1):
        """Strided device-to-device copy of this rank's row band into a (peer) frame buffer, on the scene's stream."""
        _check(lib().rt_scene_push_row_groups(self._h, C.c_void_p(int(band_ptr)), C.c_void_p(int(frame_ptr)), int(W), int(bytes_per_pixel), int(row_begin),
                                              int(row_step), int(row_group), int(rows)))

    def peer_signal(self, flag_ptr, value):
        """rt_peer_signal: stream-ordered write of `value` to the 32-bit word at flag_ptr (behind everything enqueued on the scene's stream)."""
        _check(lib().rt_peer_signal(self._h, C.c_void_p(int(flag_ptr)), int(value) & 0xffffffff))

    def peer_wait(self, flag_ptr, value):
        """rt_peer_wait: the scene's stream waits until the 32-bit word at flag_ptr is >= value."""
        _check(lib().rt_peer_wait(self._h, C.c_void_p(int(flag_ptr)), int(value) & 0xffffffff))

    def render_into(self, params, rgb=None, hit_obj=None, hit_tri=None, hit_t=None, shadow=None, flags=0):
        """rt_render with caller-provided buffers (numpy = host, torch cuda tensor / int = device)."""
        st = rt_stats()
        _check(lib().rt_render(self._h, C.byref(params), int(flags), _ptr(rgb), _ptr(hit_obj), _ptr(hit_tri), _ptr(hit_t), _ptr(shadow),
                               C.byref(st)))
        return st

    def sync(self):
        st = rt_stats()
        _check(lib().rt_scene_sync(self._h, C.byref(st)))
        return st

    def render(self, params, want=("rgb", "hit_obj", "hit_tri", "hit_t", "shadow"), count_work=False):
        """Render into fresh host arrays; returns dict of numpy arrays + 'stats'."""
        p = params
        rows = shard_row_count(p)
        shapes = {"rgb": ((rows, p.W, 3), np.uint8), "hit_obj": ((rows, p.W), np.int32), "hit_tri": ((rows, p.W), np.int32),
                  "hit_t": ((rows, p.W), np.float32), "shadow": ((rows, p.W), np.uint8)}
        out = {k: np.zeros(*shapes[k]) for k in want}
        st = self.render_into(p, out.get("rgb"), out.get("hit_obj"), out.get("hit_tri"), out.get("hit_t"), out.get("shadow"),
                              RT_RENDER_COUNT_WORK if count_work else 0)
        out["stats"] = {f[0]: getattr(st, f[0]) for f in rt_stats._fields_}
        return out


class Comm:
    """rt_comm: NCCL communicator behind the C ABI (one process or thread per GPU, or one process driving N devices)."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        _check(lib().rt_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def init(cls, nranks, rank, uid, device):
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        _check(lib().rt_comm_init(C.byref(h), int(nranks), int(rank), buf, int(device)))
        return cls(h)

    @classmethod
    def init_all(cls, ndev, devices=None):
        hs = (C.c_void_p * ndev)()
        devs = (C.c_int * ndev)(*devices) if devices is not None else None
        _check(lib().rt_comm_init_all(hs, int(ndev), devs))
        return [cls(C.c_void_p(h)) for h in hs]

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rt_comm_destroy(self._h)
            self._h = None

    def rank(self):
        r, n = C.c_int(), C.c_int()
        _check(lib().rt_comm_rank(self._h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def broadcast_scene(self, scene, root=0):
        n = C.c_size_t()
        _check(lib().rt_scene_broadcast(scene._h, self._h, int(root), C.byref(n)))
        return n.value

    def gather_framebuffer(self, scene, band_ptr, W, H, bytes_per_pixel, frame_ptr, root=0, row_group=1):
        _check(lib().rt_gather_framebuffer_groups(scene._h, self._h, C.c_void_p(int(band_ptr)), int(W), int(H), int(bytes_per_pixel), int(row_group),
                                                  C.c_void_p(int(frame_ptr)) if frame_ptr else None, int(root)))


def comm_available():
    v = C.c_int()
    rc = lib().rt_comm_available(C.byref(v))
    return v.value if rc == RT_OK else 0


def fma_peak_tflops(device=0, reps=5):
    """Measured FP32 FMA throughput of the device (TFLOP/s)."""
    v = C.c_double()
    _check(lib().rt_selftest_fma_peak(int(device), int(reps), C.byref(v)))
    return v.value


def selftest_libm(which, x, device=0):
    """CUDA's logf / sinf / cosf / tanf ('log' | 'sin' | 'cos' | 'tan') evaluated on the device for the float32 array x."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    _check(lib().rt_selftest_libm(int(device), {"log": 0, "sin": 1, "cos": 2, "tan": 3}[which], x.ctypes.data, x.size, y.ctypes.data))
    return y


def selftest_division(device=0, seed=1, blocks=148 * 8, per_thread=4096):
    out = (C.c_uint64 * 3)()
    _check(lib().rt_selftest_division(device, seed, blocks, per_thread, out))
    return {"mismatch_1step": out[0], "mismatch_2step": out[1], "pairs": out[2]}


def selftest_division3(device=0, seed=1, blocks=148 * 8, per_thread=2048):
    out = (C.c_uint64 * 2)()
    _check(lib().rt_selftest_division3(device, seed, blocks, per_thread, out))
    return {"mismatch": out[0], "components": out[1]}


def selftest_xorwow(subsequences, seed=123456, device=0):
    """(states [n, 6] uint32, uniforms [n, 4] float32) from the cuRAND device library for the given subsequences."""
    sub = np.ascontiguousarray(subsequences, dtype=np.uint32)
    st = np.zeros((len(sub), 6), np.uint32)
    u = np.zeros((len(sub), 4), np.float32)
    _check(lib().rt_selftest_xorwow(device, seed, sub.ctypes.data, len(sub), st.ctypes.data, u.ctypes.data))
    return st, u
