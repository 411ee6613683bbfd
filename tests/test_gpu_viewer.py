"""Viewer-derived features (realtime_render.cu) on the GPU: kernel == oracle restatement. RESTATED, NOT PINNED by the
reference (its GLUT program cannot be built here). Camera basis (yaw / pitch, realtime_render.cu:805-861, 1113), interpolated
vertex normals (:221-245, 311), progressive accumulation (:1136-1140), with the light orbit of :1072-1090 between frames."""
import numpy as np
import pytest

import raytracinggpu_b200 as rt
from oracle import profiles, pyoracle, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built):
    if rt.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box (there is no CPU fallback)")
    return 0


@pytest.fixture()
def scene(gpu):
    sc = rt.Scene(gpu)
    yield sc
    sc.close()


def viewer_scene(mirror=0):
    d = scenes.viewer_cat_scene()
    if d is None:  # without the asset: a torus with analytic vertex normals
        v, t = scenes.torus(48, 24)
        m = pyoracle.Mesh.from_arrays(v, t)
        c = v.mean(axis=0)
        n = (v - c) / np.linalg.norm(v - c, axis=1, keepdims=True)
        m.set_normals(n.astype(np.float32), t)
        m.build_bvh()
        sp = profiles.walls("cpu")
        sp[1].R = 940.0
        d = scenes._scene(sp, m, profiles.mesh_material("cpu", 0), ((0., 15., 40.), 3e10))
        d["normals"] = m.normals
    d["mesh_mat"] = dict(d["mesh_mat"], mirror=mirror)
    return d


def with_basis(p, yaw, pitch):
    bx, by, bz = rt.camera_basis(yaw, pitch)
    p.cam_bx[:], p.cam_by[:], p.cam_bz[:] = [float(x) for x in bx], [float(x) for x in by], [float(x) for x in bz]
    return p


@pytest.mark.parametrize("yaw,pitch,bounce,mirror", [(0.0, 0.3, 1, 0), (0.6, -0.1, 1, 0), (-1.2, 0.45, 3, 1), (3.0, 0.0, 2, 0)])
def test_viewer_camera_and_smooth_normals_match_the_oracle(scene, yaw, pitch, bounce, mirror):
    d = viewer_scene(mirror)
    scenes.upload(scene, d)
    p = with_basis(profiles.params("realtime", 480, 270, 1, bounce), yaw, pitch)
    got = scene.render(p)
    ora = scenes.run_oracle(d, p)
    res = scenes.compare(got, ora)
    assert res["rgb_exact_mismatch"] == 0, res
    assert (ora["hit_obj"] == 6).sum() > 0 or abs(yaw) > 1  # the mesh is in view for the forward-looking cameras
    # geometric normals on the same view: same hits, different shading on the mesh
    p.smooth_normals = 0
    flat = scene.render(p)
    assert np.array_equal(flat["hit_tri"], got["hit_tri"])
    scenes.compare(flat, scenes.run_oracle(d, p))


def test_smooth_normals_need_the_normals(scene):
    d = viewer_scene()
    scene.set_spheres(d["spheres"])
    mm = d["mesh_mat"]
    scene.set_mesh(*d["mesh"], albedo=mm["albedo"], id=mm["id"])
    p = profiles.params("realtime", 64, 36, 1, 1)
    with pytest.raises(rt.RtError) as e:
        scene.render(p)
    assert e.value.code == -5  # RT_ERR_STATE
    scene.set_mesh_normals(d["normals"])
    scene.render(p)
    scene.set_option("variant", 1)  # the one-kernel variants do not carry the viewer features
    with pytest.raises(rt.RtError) as e:
        scene.render(p)
    assert e.value.code == -6  # RT_ERR_UNSUPPORTED


def test_progressive_accumulation_with_a_moving_light(scene):
    """Frames k = 1..4 of a stochastic render (new seed per frame, as the viewer re-seeds with the frame number) accumulate into
    buffer / k; the light moves between accumulations and frame 1 restarts the buffer."""
    d = viewer_scene()
    scenes.upload(scene, d)
    W, H = 320, 180
    L = d["light"][0]
    for restart in range(2):
        acc = np.zeros((H, W, 3), np.float32)
        for k in range(1, 5):
            p = with_basis(profiles.params("realtime", W, H, 2, 2), 0.1, 0.3)
            p.aa_sigma, p.indirect, p.reserved, p.accumulate = 0.2, 1, 1000 + k, k
            got = scene.render(p, want=("rgb",))["rgb"]
            q = with_basis(profiles.params("realtime", W, H, 2, 2), 0.1, 0.3)
            q.aa_sigma, q.indirect, q.reserved = 0.2, 1, 1000 + k
            lin = scenes.run_oracle(dict(d, light=(L, 3e10)), q, want=("linear",))["linear"]
            want = pyoracle.accumulate(acc, lin, k, q.gamma_mode)
            diff = np.abs(got.astype(int) - want.astype(int)).max(axis=2)
            assert (diff == 0).mean() >= 0.999 and diff.max() <= 1, (restart, k, (diff == 0).mean(), diff.max())
        L = rt.move_light(L, 1.309, 0.5)
        scene.set_light(L, 3e10)
    # deterministic frames accumulate too, and through host outputs with several row bands
    p = with_basis(profiles.params("realtime", W, H, 1, 1), 0.1, 0.3)
    base = scene.render(p, want=("rgb",))["rgb"]
    scene.set_option("strips", 3)
    for k in (1, 2, 3):
        p.accumulate = k
        f = scene.render(p, want=("rgb", "hit_obj"))
        assert np.abs(f["rgb"].astype(int) - base.astype(int)).max() <= (0 if k == 1 else 1)
