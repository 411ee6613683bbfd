"""Viewer-derived features (realtime_render.cu), CPU side: the host helpers of the product against the oracle's
restatements, and internal consistency of the restatements. The GLUT program cannot be built here (no GL headers), so these
features are RESTATED, NOT PINNED by any output of the reference; the GPU tests (tests/test_gpu_viewer.py) compare kernel and
oracle."""
import numpy as np
import pytest

import raytracinggpu_b200 as rt
from oracle import profiles, pyoracle, scenes


def test_camera_basis_matches_the_restatement(built):
    for yaw, pitch in ((0.0, 0.3), (0.0, 0.0), (0.7, -0.2), (-2.5, 1.1), (3.0, 0.05)):
        a = rt.camera_basis(yaw, pitch)
        b = pyoracle.camera_basis(yaw, pitch)
        for u, v in zip(a, b):
            assert np.array_equal(u.view(np.uint32), v.view(np.uint32)), (yaw, pitch)
        bx, by, bz = [x.astype(np.float64) for x in a]
        for x in (bx, by, bz):
            assert abs(np.linalg.norm(x) - 1) < 1e-6
        assert abs(bx @ by) < 1e-6 and abs(bx @ bz) < 1e-6 and abs(by @ bz) < 1e-6
    bx, by, bz = rt.camera_basis(0.0, 0.0)
    assert bx.tolist() == [1, 0, 0] and by.tolist() == [0, 1, 0] and bz.tolist() == [0, 0, 1]  # Camera::rotate ends with bz = bx x by


def test_realtime_profile_knobs(built):
    p = rt.params_profile("realtime", 640, 360, 1, 3)
    q = profiles.params("realtime", 640, 360, 1, 3)
    for f, _ in rt.rt_params._fields_:
        a, b = getattr(p, f), getattr(q, f)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), f
        else:
            assert a == b, f
    assert p.camera_mode == 1 and p.smooth_normals == 1 and p.eps_surface == np.float32(1e-3)
    walls, mesh_id = rt.default_walls("realtime")
    assert mesh_id == 6 and walls[1].R == 940.0  # the viewer's floor, realtime_render.cu:1027


def test_loader_keeps_normals_like_the_viewers(cat_path, built):
    m = rt.Mesh.read_obj(cat_path, keep_normals=True)
    n, idx = scenes.obj_normals(cat_path)
    assert n.shape == (2152, 3) and np.array_equal(m.normals.view(np.uint32), n.view(np.uint32))
    assert np.array_equal(m.tri_records[:, 6:9], idx)
    plain = rt.Mesh.read_obj(cat_path)
    assert plain.normals.shape[0] == 0 and (plain.tri_records[:, 3:] == -1).all()  # optimized.cu's loader drops them
    assert np.array_equal(plain.tri_records[:, :3], m.tri_records[:, :3])
    # the build reorders whole records: normal indices travel with their triangle
    m.build_bvh()
    plain.build_bvh()
    assert np.array_equal(plain.tri_records[:, :3], m.tri_records[:, :3])
    assert sorted(map(tuple, m.tri_records[:, 6:9].tolist())) == sorted(map(tuple, idx.tolist()))


def test_viewer_camera_reduces_to_the_launchers(built):
    """With the identity basis and the camera at the origin, realtime_render.cu:1113 is optimized.cu:751."""
    d = scenes.torus_scene("optimized")
    p = profiles.params("optimized", 96, 54, 1, 1)
    p.cam[:] = [0.0, 0.0, 0.0]
    a = scenes.run_oracle(d, p)
    p.camera_mode = 1
    p.cam_bx[:], p.cam_by[:], p.cam_bz[:] = [1, 0, 0], [0, 1, 0], [0, 0, 1]
    b = scenes.run_oracle(d, p)
    for k in ("rgb", "hit_obj", "hit_tri", "shadow"):
        assert np.array_equal(a[k], b[k]), k
    # away from the origin the camera position enters the direction (the reference adds cam.C to u_center): a different image
    p.cam[:] = [0.0, 0.0, 55.0]
    c = scenes.run_oracle(d, p)
    p.camera_mode = 0
    assert not np.array_equal(c["hit_obj"], scenes.run_oracle(d, p)["hit_obj"])


def test_smooth_normals_and_accumulation_restatements(cat_path, built):
    d = scenes.viewer_cat_scene(cat_path)
    p = profiles.params("realtime", 320, 180, 1, 1)
    smooth = scenes.run_oracle(d, p, want=("rgb", "hit_obj", "hit_tri", "linear"))
    p.smooth_normals = 0
    flat = scenes.run_oracle(d, p, want=("rgb", "hit_obj", "hit_tri"))
    assert np.array_equal(smooth["hit_tri"], flat["hit_tri"])  # the normal changes the shading, not the hit
    cat = smooth["hit_obj"] == 6
    assert cat.sum() > 150
    assert (smooth["rgb"][cat] != flat["rgb"][cat]).any() and np.array_equal(smooth["rgb"][~cat], flat["rgb"][~cat])
    # accumulation of identical frames: frame 1 is the frame itself, later frames stay within one level
    acc = np.zeros_like(smooth["linear"])
    f1 = pyoracle.accumulate(acc, smooth["linear"], 1, p.gamma_mode)
    assert np.array_equal(f1, smooth["rgb"])
    f3 = None
    for k in (2, 3):
        f3 = pyoracle.accumulate(acc, smooth["linear"], k, p.gamma_mode)
    assert np.abs(f3.astype(int) - smooth["rgb"].astype(int)).max() <= 1
