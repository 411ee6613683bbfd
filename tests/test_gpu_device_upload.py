"""rt_scene_set_mesh_device / rt_scene_set_mesh_from: the interchange arrays taken from the DEVICE (node relayout, leaf table and
triangle repack as kernels, csrc/rt_relayout.cuh) give the same frames as the host upload — ids, t bits, shadow flags, colours."""
import numpy as np
import pytest
import torch

import raytracinggpu_b200 as rt
from oracle import profiles, pyoracle, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built):
    if rt.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box (there is no CPU fallback)")
    return 0


def same(a, b):
    for k in ("rgb", "hit_obj", "hit_tri", "shadow"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["hit_t"].view(np.uint32), b["hit_t"].view(np.uint32))
    assert a["stats"]["rays"] == b["stats"]["rays"]


def meshes():
    cat = pyoracle.cat_obj_path()
    if cat:
        yield "cat", lambda: rt.Mesh.read_obj(cat).rescale(0.6, (0.0, -4.0, 0.0))
    v, t = scenes.torus(40, 20)
    yield "torus", lambda: rt.Mesh.from_arrays(v, t)
    # leaves far larger than RT_LEAF_MAX: 600 triangles with one exact centroid (a chunk tree of 150 chunks) among 3000 others
    rng = np.random.RandomState(5)
    q = rng.randint(-256, 257, size=(600, 4)).astype(np.float32) / 64.0
    verts, tris = [], []
    for a_, b_, c_, d_ in q:
        k = len(verts)
        verts += [np.float32([a_, b_, 1.0]), np.float32([c_, d_, -0.5]), np.float32([-a_ - c_, -b_ - d_, -0.5])]
        tris.append((k, k + 1, k + 2))
    for b_ in rng.rand(3000, 3).astype(np.float32) * 20 - 10:
        k = len(verts)
        verts += [b_, b_ + np.float32([0.6, 0, 0]), b_ + np.float32([0, 0.6, 0.2])]
        tris.append((k, k + 1, k + 2))
    V, T = np.asarray(verts, np.float32), np.asarray(tris, np.int32)
    yield "big_leaf", lambda: rt.Mesh.from_arrays(V, T)


@pytest.mark.parametrize("profile,bounce,mirror", [("optimized", 1, 0), ("cpu", 0, 0), ("optimized", 3, 1)])
def test_device_upload_equals_host_upload(gpu, profile, bounce, mirror):
    walls = profiles.walls(profile)
    mesh_id = profiles.PROFILES[profile]["mesh_id"]
    p = profiles.params(profile, 480, 270, 1, bounce)
    for name, make in meshes():
        host_mesh = make().build_bvh()
        dev_mesh = make().build_bvh_gpu(gpu)          # post-build arrays stay on the device
        a, b = rt.Scene(gpu), rt.Scene(gpu)
        try:
            a.set_spheres(walls)
            b.set_spheres(walls)
            a.set_mesh(host_mesh.vertices, host_mesh.tri_records, host_mesh.arr_bvh, mirror=mirror, id=mesh_id)
            b.set_mesh_from(dev_mesh, mirror=mirror, id=mesh_id)   # rt_scene_set_mesh_device underneath
            ra, rb = a.render(p), b.render(p)
            same(ra, rb)
            for opt in ((("anchored", 0),), (("anchored", 1),)):   # both searches on the device-built layout
                for k, v in opt:
                    b.set_option(k, v)
                same(ra, b.render(p))
            # the lazy host mirror of the device build equals the host build
            assert np.array_equal(dev_mesh.tri_records, host_mesh.tri_records), name
            assert np.array_equal(dev_mesh.arr_bvh.view(np.uint32), host_mesh.arr_bvh.view(np.uint32)), name
        finally:
            a.close()
            b.close()


def test_callers_device_arrays_and_errors(gpu):
    v, t = scenes.torus(32, 16)
    m = rt.Mesh.from_arrays(v, t).build_bvh()
    dv = torch.from_numpy(m.vertices).cuda()
    dr = torch.from_numpy(m.tri_records).cuda()
    db = torch.from_numpy(m.arr_bvh).cuda()
    walls = profiles.walls("optimized")
    a, b = rt.Scene(gpu), rt.Scene(gpu)
    try:
        a.set_spheres(walls)
        b.set_spheres(walls)
        a.set_mesh(m.vertices, m.tri_records, m.arr_bvh, id=1)
        nv, nt, nn = m.counts()
        b.set_mesh_device(dv.data_ptr(), nv, dr.data_ptr(), nt, db.data_ptr(), nn, id=1)
        p = profiles.params("optimized", 320, 180, 1, 1)
        same(a.render(p), b.render(p))
        with pytest.raises(rt.RtError) as e:   # host pointers are refused, not dereferenced on the device
            b.set_mesh_device(m.vertices.ctypes.data, nv, dr.data_ptr(), nt, db.data_ptr(), nn, id=1)
        assert e.value.code == -1
        bad = db.clone()
        bad[0, 0] = 5000.0   # child index out of range
        with pytest.raises(rt.RtError) as e:
            b.set_mesh_device(dv.data_ptr(), nv, dr.data_ptr(), nt, bad.data_ptr(), nn, id=1)
        assert e.value.code == -1
        same(a.render(p), b.render(p))  # the failed call left the scene as it was
    finally:
        a.close()
        b.close()
