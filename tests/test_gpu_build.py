"""the reference's BVH builder (compute_bbox / buildBVH / bvhTreeToArray, optimized.cu:466-534) run on the
device, a tree level at a time, must produce the host builder's arrays element for element: the 10-float array BVH and the
triangle order (the reference's swap-to-pivot partition is not a stable partition; the device resolves its permutation
in closed form). The host builder itself is pinned to the reference's compiled classes in tests/test_host_mesh.py."""
import numpy as np
import pytest

import raytracinggpu_b200 as rt
from oracle import pyoracle, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built):
    if rt.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box (there is no CPU fallback)")
    return 0


def both(make):
    a, b = make().build_bvh(), make().build_bvh_gpu(0)
    assert a.counts() == b.counts()
    assert a.bvh_info() == b.bvh_info()
    assert np.array_equal(a.tri_records, b.tri_records)
    assert np.array_equal(a.arr_bvh, b.arr_bvh)  # value equality: a zero bound may differ in sign (rt_bvh_build.cuh)
    return a, b


def test_cat_tree_is_the_host_builders(gpu):
    cat = pyoracle.cat_obj_path()
    if cat is None:
        pytest.skip("cat asset unavailable")
    a, b = both(lambda: rt.Mesh.read_obj(cat).rescale(0.6, (0.0, -4.0, 0.0)))
    info = a.bvh_info()
    assert a.counts()[2] == 2019 and info["leaves"] == 1010 and info["max_leaf"] == 73  # SURVEY.md §8c pins (its depth 24 counts edges: 25 levels)
    assert info["max_depth"] == 25
    both(lambda: rt.Mesh.read_obj(cat))  # the cpu_launcher.cpp placement


@pytest.mark.parametrize("nu,nv", [(48, 24), (7, 5), (200, 150)])
def test_torus_trees(gpu, nu, nv):
    v, t = scenes.torus(nu, nv)
    both(lambda: rt.Mesh.from_arrays(v, t))


def test_degenerate_inputs(gpu):
    v, t = scenes.grid_quads(9)  # flat boxes, many equal centroids: long runs of one side, leaves by the pivot rules
    both(lambda: rt.Mesh.from_arrays(v, t))
    rng = np.random.RandomState(3)
    v = rng.rand(400, 3).astype(np.float32) * 20 - 10
    t = rng.randint(0, 400, size=(3000, 3)).astype(np.int32)  # random soup: deep, unbalanced
    both(lambda: rt.Mesh.from_arrays(v, t))
    v = np.array([[-8, -8, 0], [8, -8, 0], [0, 8, 2], [0, 0, 9]], np.float32)
    t = np.array([[0, 1, 2], [0, 1, 3], [1, 2, 3]], np.int32)  # the root is a leaf
    both(lambda: rt.Mesh.from_arrays(v, t))
    t5 = np.array([[0, 1, 2]] * 7, np.int32)  # identical triangles: nothing goes left of the split
    both(lambda: rt.Mesh.from_arrays(v, t5))


def test_ten_million_triangles(gpu):
    """BASELINE.json configs[4] mesh: 9,999,666 triangles, 5.05 M nodes, depth 40. Host builder: seconds; device: the build
    time is printed by tools/build_bench.py. Equality of every node and of the whole triangle order."""
    from raytracinggpu_b200 import synthetic
    cat = pyoracle.cat_obj_path()
    if cat is None:
        pytest.skip("cat asset unavailable")
    scales, offs = synthetic.instance_lattice()
    a, b = both(lambda: rt.Mesh.read_obj(cat).instance(scales, offs))
    assert a.counts()[1] == 9999666
    assert b.build_ms > 0
