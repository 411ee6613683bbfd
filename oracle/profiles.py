"""The per-program knobs of the reference, stated independently of the product's
rt_params_profile / rt_default_walls so tests can cross-check them. TEST INFRASTRUCTURE ONLY."""
import math

import numpy as np

from raytracinggpu_b200._abi import rt_params, rt_sphere
from . import pyoracle

# walls in the order cpu_launcher.cpp:673-678 adds them (C, R, albedo)
WALLS = [
    ((0, 0, -1000), 940, (0., 1., 0.)),   # green fore wall
    ((0, -1000, 0), 990, (0., 0., 1.)),   # blue floor
    ((0, 1000, 0), 940, (1., 0., 0.)),    # red ceiling
    ((-1000, 0, 0), 940, (0., 1., 1.)),   # cyan left wall
    ((1000, 0, 0), 940, (1., 1., 0.)),    # yellow right wall
    ((0, 0, 1000), 940, (1., 0., 1.)),    # magenta back wall
]
LIGHT = ((-10., 20., 40.), 3e10)          # optimized.cu:681-683 / cpu_launcher.cpp:650-651
CAT_ALBEDO = (0.25, 0.25, 0.25)           # optimized.cu:692
ALPHA = np.float32(math.pi / 3)           # optimized.cu:748 (float alpha = PI/3)

PROFILES = {
    # name: (eps_surface, eps_tri, push_order, extra_segment, gamma_mode, mesh_id, mesh rescale)
    "cpu": dict(eps_surface=1e-3, eps_tri=1e-4, push_order=0, extra_segment=1, gamma_mode=0, mesh_id=6, rescale=None),
    "optimized": dict(eps_surface=1e-4, eps_tri=0.0, push_order=1, extra_segment=0, gamma_mode=1, mesh_id=1, rescale=(0.6, (0., -4., 0.))),
    "array_bvh": dict(eps_surface=1e-4, eps_tri=1e-4, push_order=0, extra_segment=0, gamma_mode=0, mesh_id=6, rescale=(0.6, (0., -10., 0.))),
    # the GLUT viewer (realtime_render.cu:908,298,288-289,1021,809,311): viewer camera (pitch 0.3, 90 degree field of view), smooth normals
    "realtime": dict(eps_surface=1e-3, eps_tri=1e-3, push_order=0, extra_segment=0, gamma_mode=1, mesh_id=6, rescale=(0.6, (0., -10., 0.))),
}


def sphere(C, R, albedo, id, mirror=0, n_in=1.0, n_out=1.0):
    s = rt_sphere()
    s.C[:] = [float(x) for x in C]
    s.R = float(R)
    s.albedo[:] = [float(x) for x in albedo]
    s.mirror = int(mirror)
    s.n_in = float(n_in)
    s.n_out = float(n_out)
    s.id = int(id)
    return s


def walls(profile, with_mesh=True):
    """Wall spheres with the object ids of the profile; the mesh takes PROFILES[profile]['mesh_id']."""
    mesh_id = PROFILES[profile]["mesh_id"] if with_mesh else None
    out, nxt = [], 0
    for (C, R, a) in WALLS:
        if mesh_id is not None and nxt == mesh_id:
            nxt += 1
        out.append(sphere(C, R, a, nxt))
        nxt += 1
    return out


def params(profile, W, H, num_rays=1, num_bounce=1):
    k = PROFILES[profile]
    p = rt_params()
    p.W, p.H, p.num_rays, p.num_bounce = W, H, num_rays, num_bounce
    p.cam[:] = [0., 0., 55.]
    p.z = pyoracle.lib().orc_camera_z(W, float(ALPHA))
    p.eps_surface = k["eps_surface"]
    p.eps_tri = k["eps_tri"]
    p.push_order = k["push_order"]
    p.extra_segment = k["extra_segment"]
    p.aa_sigma = 0.0
    p.indirect = 0
    p.gamma_mode = k["gamma_mode"]
    p.row_begin, p.row_step, p.row_count, p.row_group = 0, 1, 0, 1
    if profile == "realtime":
        p.z = pyoracle.lib().orc_camera_z(W, float(np.float32(math.pi / 2)))
        p.camera_mode = 1
        bx, by, bz = pyoracle.camera_basis(0.0, 0.3)
        p.cam_bx[:], p.cam_by[:], p.cam_bz[:] = [float(x) for x in bx], [float(x) for x in by], [float(x) for x in bz]
        p.smooth_normals = 1
    return p


def cat_mesh(profile, obj_path=None):
    """Oracle-loaded cat with the profile's transform and BVH (None when the asset is unavailable)."""
    obj_path = obj_path or pyoracle.cat_obj_path()
    if obj_path is None:
        return None
    m = pyoracle.Mesh.from_obj(obj_path, rescale=PROFILES[profile]["rescale"])
    return m.build_bvh()


def mesh_material(profile, mirror=0):
    return dict(albedo=CAT_ALBEDO, mirror=mirror, n_in=1.0, n_out=1.0, id=PROFILES[profile]["mesh_id"])
