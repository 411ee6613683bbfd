"""The anchored-ray path (rt_bins.cuh: camera rays and shadow rays find their leaves through per-anchor direction bins,
wf_leaves tests the (ray, candidate leaf) tasks) and the optional wide index of wf_traverse, through the C ABI: both must
give exactly what the tree search over the reference's two-child records gives, which the other GPU tests pin to the
oracle. Covers the bins' edge cases: a moving light (lists rebuilt without a read-back), a light inside the mesh's box
(lists not usable: exact search), a task buffer that is too small (the frame is repeated), rays with a zero direction
component (outside the bins' contract)."""
import numpy as np
import pytest

import raytracinggpu_b200 as rt
from oracle import profiles, scenes

pytestmark = pytest.mark.gpu
KEYS = ("rgb", "hit_obj", "hit_tri", "hit_t", "shadow")


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


def same(a, b):
    for k in KEYS:
        x, y = a[k], b[k]
        if x.dtype == np.float32:
            x, y = x.view(np.uint32), y.view(np.uint32)
        assert np.array_equal(x, y), k


def cat_or_torus(profile, mirror=0):
    return scenes.cat_scene(profile, mirror=mirror) or scenes.torus_scene(profile, mirror=mirror)


@pytest.mark.parametrize("profile,W,H,bounce,mirror", [
    ("optimized", 1920, 1080, 1, 0),   # BASELINE.json configs[1] at full size
    ("cpu", 800, 450, 0, 0),           # push order 0 (leaf-order tie-break rank), eps_tri 1e-4, extra segment
    ("optimized", 960, 540, 4, 1),     # mirror mesh: bounce rays through wf_traverse, shadow rays through the bins
    ("array_bvh", 333, 187, 2, 0),     # odd sizes: a column and a row of rays with a zero direction component
])
def test_bins_equal_tree_search(scene, profile, W, H, bounce, mirror):
    desc = cat_or_torus(profile, mirror)
    scenes.upload(scene, desc)
    p = profiles.params(profile, W, H, 1, bounce)
    scene.set_option("anchored", 1)
    a = scene.render(p)
    scene.set_option("anchored", 0)
    b = scene.render(p)
    same(a, b)
    assert a["stats"]["rays"] == b["stats"]["rays"]
    if W * H <= 960 * 540:  # and both equal the oracle
        scenes.compare(a, scenes.run_oracle(desc, p))


def test_bins_follow_a_moving_light(scene):
    """Every frame of a light orbit rebuilds the light's bins (no read-back after the first build); frames equal the oracle's."""
    desc = cat_or_torus("optimized")
    scenes.upload(scene, desc)
    scene.set_option("anchored", 1)
    p = profiles.params("optimized", 480, 270, 1, 1)
    L = (-10.0, 20.0, 40.0)
    for k in range(6):
        L = rt.move_light(L, 1.309, 0.4)  # realtime_render.cu:1072-1090, big steps
        scene.set_light(L, 3e10)
        d = dict(desc, light=(L, 3e10))
        scenes.compare(scene.render(p), scenes.run_oracle(d, p))


def test_light_inside_the_mesh_box(scene):
    """A leaf box around the anchor has no bounded set of cells: the status word sends the shadow rays to the exact search."""
    desc = cat_or_torus("optimized")
    bvh = desc["mesh"][2]
    leaf = bvh[bvh[:, 0] < 0][len(bvh) // 5]
    L = tuple(float(x) for x in 0.5 * (leaf[2:5] + leaf[5:8]))
    d = dict(desc, light=(L, 3e10))
    scenes.upload(scene, d)
    scene.set_option("anchored", 1)
    p = profiles.params("optimized", 200, 112, 1, 1)
    scenes.compare(scene.render(p), scenes.run_oracle(d, p))


def test_task_buffer_overflow_repeats_the_frame(gpu):
    sc = rt.Scene(gpu)
    try:
        sc.set_option("anchored", 1)
        sc.set_option("task_factor", 1)  # one task per pixel: too few where the mesh fills the view
        desc = cat_or_torus("cpu")  # the un-rescaled cat covers a quarter of the frame
        scenes.upload(sc, desc)
        p = profiles.params("cpu", 640, 360, 1, 0)
        p.cam[2] = 30.0  # closer: longer candidate lists per pixel
        a = sc.render(p)
        assert sc.get_option("task_factor") > 1  # the synchronous call repeated the frame with a doubled buffer
        sc.set_option("anchored", 0)
        same(a, sc.render(p))
    finally:
        sc.close()


def test_wide_index_equals_two_child_records(scene):
    """Option wide = 1: wf_traverse over the four-child collapse of the reference tree (leaf decisions exact, inner ones
    conservative) gives the same frames, tree search only and behind the bins."""
    desc = cat_or_torus("optimized", mirror=1)
    scenes.upload(scene, desc)
    p = profiles.params("optimized", 640, 360, 1, 3)
    scene.set_option("anchored", 0)
    ref = scene.render(p)
    scene.set_option("wide", 1)
    same(scene.render(p), ref)
    scene.set_option("anchored", 1)
    same(scene.render(p), ref)
    q = profiles.params("array_bvh", 111, 77, 1, 2)  # zero direction components: answered by the exact search at admission
    scene.set_option("anchored", 0)
    d2 = cat_or_torus("array_bvh")
    scenes.upload(scene, d2)
    scenes.compare(scene.render(q), scenes.run_oracle(d2, q))


@pytest.mark.parametrize("profile,W,H,bounce,mirror,stoch,anchored", [
    ("optimized", 960, 540, 1, 0, False, 0),   # every ray through the tree search
    ("cpu", 640, 360, 0, 0, False, 0),         # push order 0, extra segment
    ("optimized", 960, 540, 4, 1, False, -1),  # mirror mesh: bounce rays through wf_traverse beside the bins
    ("optimized", 480, 270, 3, 0, True, -1),   # stochastic passes (the wide index is switched off below: TOPS is a two-child variant)
])
def test_top_levels_from_shared_memory(scene, profile, W, H, bounce, mirror, stoch, anchored):
    """Option top_smem: wf_traverse takes the records of the top four levels from a breadth-first table staged in shared memory
    (build_top_table: child references inside the table rewritten to n_inner + slot). Same boxes, same references, same order:
    every output equals the global-memory traversal's, and the table follows a mesh change."""
    desc = cat_or_torus(profile, mirror)
    scenes.upload(scene, desc)
    scene.set_option("anchored", anchored)
    scene.set_option("wide", 0)
    p = profiles.params(profile, W, H, 2 if stoch else 1, bounce)
    if stoch:
        p.aa_sigma, p.indirect = 0.2, 1
    a = scene.render(p)
    scene.set_option("top_smem", 1)
    b = scene.render(p)
    same(a, b)
    assert a["stats"]["rays"] == b["stats"]["rays"]
    # another tree: the table is rebuilt
    other = scenes.torus_scene(profile, mirror=mirror, nu=40, nv=20)
    scenes.upload(scene, other)
    c = scene.render(p)
    scene.set_option("top_smem", 0)
    d = scene.render(p)
    same(c, d)
    if not stoch:
        scenes.compare(c, scenes.run_oracle(other, p))
