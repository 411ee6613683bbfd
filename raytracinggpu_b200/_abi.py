"""ctypes mirror of include/rt_b200.h (struct layouts and constants only; no library is loaded here)."""
import ctypes as C

RT_OK = 0
RT_ERR_INVALID = -1
RT_ERR_CUDA = -2
RT_ERR_NOMEM = -3
RT_ERR_IO = -4
RT_ERR_STATE = -5
RT_ERR_UNSUPPORTED = -6
RT_ERR_AGAIN = -7

RT_TRI_RECORD_WORDS = 10
RT_BVH_NODE_FLOATS = 10

RT_RENDER_COUNT_WORK = 1
RT_RENDER_NO_SYNC = 2


class rt_sphere(C.Structure):
    _fields_ = [
        ("C", C.c_float * 3),
        ("R", C.c_float),
        ("albedo", C.c_float * 3),
        ("mirror", C.c_int32),
        ("n_in", C.c_float),
        ("n_out", C.c_float),
        ("id", C.c_int32),
    ]


class rt_params(C.Structure):
    _fields_ = [
        ("W", C.c_int32),
        ("H", C.c_int32),
        ("num_rays", C.c_int32),
        ("num_bounce", C.c_int32),
        ("cam", C.c_float * 3),
        ("z", C.c_float),
        ("eps_surface", C.c_float),
        ("eps_tri", C.c_float),
        ("push_order", C.c_int32),
        ("extra_segment", C.c_int32),
        ("aa_sigma", C.c_float),
        ("indirect", C.c_int32),
        ("gamma_mode", C.c_int32),
        ("row_begin", C.c_int32),
        ("row_step", C.c_int32),
        ("row_count", C.c_int32),
        ("reserved", C.c_int32),
        ("camera_mode", C.c_int32),
        ("cam_bx", C.c_float * 3),
        ("cam_by", C.c_float * 3),
        ("cam_bz", C.c_float * 3),
        ("smooth_normals", C.c_int32),
        ("accumulate", C.c_int32),
        ("row_group", C.c_int32),
    ]


class rt_stats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double),
        ("rays", C.c_uint64),
        ("node_visits", C.c_uint64),
        ("tri_tests", C.c_uint64),
        ("launches", C.c_int32),
        ("max_stack", C.c_int32),
        ("slab_fallbacks", C.c_uint64),
        ("tri_exact", C.c_uint64),
    ]
