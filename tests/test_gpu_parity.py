"""Parity tests proper: the CUDA path, called through the C ABI (librtb200.so), against the CPU oracle on the
same inputs and against the committed golden fixtures. Bar (BASELINE.json): object ids, triangle ids and t bits
exact; shadow flags exact; 8-bit colour within 1 LSB per channel on >= 99.9 % of pixels (in practice the
gamma table makes the colours exact too, which is asserted as well where noted)."""
import os

import numpy as np
import pytest

import raytracinggpu_b200 as rt
from oracle import profiles, scenes

import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


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


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_case_matches_oracle_and_golden(scene, name):
    case = cases.CASES[name]
    desc = case["scene"]()
    gold_path = os.path.join(GOLD, "oracle_%s.npz" % name)
    if desc is None:
        pytest.skip("cat asset unavailable")
    p = case["params"]()
    scenes.upload(scene, desc)
    got = scene.render(p)
    ora = scenes.run_oracle(desc, p)
    res = scenes.compare(got, ora)
    assert res["rgb_exact_mismatch"] == 0, res  # the table-based transfer function is exact, not just <= 1 LSB
    # identical deterministic samples are traced once on the device; the oracle retraces every sample
    assert got["stats"]["rays"] * p.num_rays == ora["work"]["rays"]
    g = np.load(gold_path)
    gold = dict(rgb=g["rgb"], hit_obj=g["hit_obj"].astype(np.int32), hit_tri=g["hit_tri"], hit_t=g["hit_t"], shadow=g["shadow"])
    scenes.compare(got, gold)


def test_full_size_config2_cat_1080p(scene):
    """BASELINE.json config 2 at full size: cat, 1920x1080, primary + shadow (4,147,200 rays)."""
    desc = scenes.cat_scene("optimized")
    if desc is None:
        pytest.skip("cat asset unavailable")
    p = profiles.params("optimized", 1920, 1080, 1, 1)
    scenes.upload(scene, desc)
    got = scene.render(p, count_work=True)
    ora = scenes.run_oracle(desc, p)
    res = scenes.compare(got, ora)
    assert res["rgb_exact_mismatch"] == 0
    assert got["stats"]["rays"] == 2 * 1920 * 1080 == ora["work"]["rays"]
    # SURVEY.md §8c pin (3)
    assert list(np.bincount(got["hit_obj"].ravel(), minlength=7)) == [1099781, 196633, 662766, 0, 57210, 57210, 0]
    assert int((got["shadow"] == 1).sum()) == 80113


def test_full_size_cpu_profile_1080p(scene):
    """The cpu_launcher scene (cat not rescaled: 550 k cat pixels) at 1080p: the one ray whose winner depends on
    the traversal order must come out as the reference's order decides it."""
    desc = scenes.cat_scene("cpu")
    if desc is None:
        pytest.skip("cat asset unavailable")
    p = profiles.params("cpu", 1920, 1080, 1, 0)
    scenes.upload(scene, desc)
    got = scene.render(p)
    ora = scenes.run_oracle(desc, p)
    scenes.compare(got, ora)
    p.push_order = 1
    got1 = scene.render(p)
    ora1 = scenes.run_oracle(desc, p)
    scenes.compare(got1, ora1)


def test_4k_mirror_depth4_properties(scene):
    """Config 3 at full size (3840x2160, depth 4): size-independent properties instead of a full oracle run —
    idempotence, agreement of a row-sharded render with the whole frame, and oracle parity on a row sample."""
    desc = scenes.cat_scene("optimized", mirror=1) or scenes.torus_scene("optimized", mirror=1)
    W, H = 3840, 2160
    p = profiles.params("optimized", W, H, 1, 4)
    scenes.upload(scene, desc)
    a = scene.render(p)
    b = scene.render(p)
    for k in ("rgb", "hit_obj", "hit_tri", "shadow"):
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(a["hit_t"].view(np.uint32), b["hit_t"].view(np.uint32))
    # 8-way row interleave reassembles to the same frame and the same ray count
    parts, rays = [], 0
    for r in range(8):
        q = profiles.params("optimized", W, H, 1, 4)
        q.row_begin, q.row_step, q.row_count = rt.sharding.rows_for_rank(H, r, 8)
        o = scene.render(q, want=("rgb",))
        parts.append(o["rgb"])
        rays += o["stats"]["rays"]
    frame = rt.sharding.assemble(np.stack(parts), H, 8)
    assert np.array_equal(frame, a["rgb"])
    assert rays == a["stats"]["rays"]
    # the same in groups of 8 consecutive rows (rt_params.row_group): all five outputs
    parts, rays = {k: [] for k in ("rgb", "hit_obj", "hit_tri", "hit_t", "shadow")}, 0
    pad = rt.sharding.padded_rows(H, 8, group=8)
    for r in range(8):
        q = profiles.params("optimized", W, H, 1, 4)
        assert rt.shard_rows(q, r, 8, 8) == rt.sharding.rows_for_rank(H, r, 8, group=8)[2]
        o = scene.render(q)
        for k in parts:
            band = np.zeros((pad,) + o[k].shape[1:], o[k].dtype)
            band[:o[k].shape[0]] = o[k]
            parts[k].append(band)
        rays += o["stats"]["rays"]
    for k in parts:
        whole = rt.sharding.assemble(np.stack(parts[k]), H, 8, group=8)
        assert np.array_equal(whole.view(np.uint32) if k == "hit_t" else whole, a[k].view(np.uint32) if k == "hit_t" else a[k]), k
    assert rays == a["stats"]["rays"]
    # oracle on every 40th row
    q = profiles.params("optimized", W, H, 1, 4)
    q.row_begin, q.row_step, q.row_count = 7, 40, 0
    scenes.compare(scene.render(q), scenes.run_oracle(desc, q))


def test_device_pointer_outputs_and_external_stream(scene):
    """Outputs may be device pointers (written in place) and the scene may run on the caller's stream."""
    import torch
    desc = scenes.torus_scene("optimized")
    p = profiles.params("optimized", 320, 200, 1, 1)
    scenes.upload(scene, desc)
    host = scene.render(p)
    st = torch.cuda.Stream()
    scene.set_stream(st.cuda_stream)
    rgb = torch.zeros((200, 320, 3), dtype=torch.uint8, device="cuda")
    obj = torch.full((200, 320), -7, dtype=torch.int32, device="cuda")
    with torch.cuda.stream(st):
        stats = scene.render_into(p, rgb=rgb, hit_obj=obj)
    st.synchronize()
    assert np.array_equal(rgb.cpu().numpy(), host["rgb"])
    assert np.array_equal(obj.cpu().numpy(), host["hit_obj"])
    assert stats.launches >= 1 and stats.rays == host["stats"]["rays"]


def test_scene_blob_roundtrip(gpu):
    """The packed scene travels as one blob (what a rank-0 broadcast sends): import it into a second scene."""
    desc = scenes.torus_scene("cpu", mirror=1)
    p = profiles.params("cpu", 160, 100, 1, 2)
    a = scenes.upload(rt.Scene(gpu), desc)
    ptr, n = a.blob_export()
    b = rt.Scene(gpu)
    b.blob_import(ptr, n)
    ra, rb = a.render(p), b.render(p)
    for k in ("rgb", "hit_obj", "hit_tri", "shadow"):
        assert np.array_equal(ra[k], rb[k])
    a.close()
    b.close()


def test_spheres_only_and_mesh_removal(scene):
    desc = scenes.torus_scene("cpu")
    scenes.upload(scene, desc)
    sp = scenes.spheres_scene()
    scenes.upload(scene, sp)  # clears the mesh
    p = profiles.params("cpu", 96, 64, 1, 3)
    scenes.compare(scene.render(p), scenes.run_oracle(sp, p))


def test_error_paths(scene):
    p = profiles.params("optimized", 64, 64, 1, 1)
    with pytest.raises(rt.RtError) as e:  # empty scene
        scene.render(p)
    assert e.value.code == -5
    scenes.upload(scene, scenes.spheres_scene())
    p.aa_sigma, p.indirect, p.num_bounce = 0.2, 1, 40
    with pytest.raises(rt.RtError) as e:  # more path segments than the stochastic kernel's fold arrays hold: say so
        scene.render(p)
    assert e.value.code == -6
    bad = scenes.spheres_scene()
    bad["spheres"][0].id = 42
    scene.set_spheres(bad["spheres"])
    with pytest.raises(rt.RtError):
        scene.render(profiles.params("optimized", 64, 64, 1, 1))
    v, t = scenes.torus(8, 4)
    from oracle import pyoracle
    m = pyoracle.Mesh.from_arrays(v, t).build_bvh()
    bvh = m.arr_bvh.copy()
    bvh[0, 1] = 0  # right child pointing back at the root
    with pytest.raises(rt.RtError):
        scene.set_mesh(m.vertices, m.tri_records, bvh)


def test_node_pool_overflow_spills_to_global_memory(scene):
    """wf_traverse parks the older half of a full node pool in global memory. Option npool_cap = 64 (test hook of
    rt_render) makes that happen all the time; results must not change."""
    desc = scenes.cat_scene("cpu") or scenes.torus_scene("cpu")
    p = profiles.params("cpu", 640, 360, 1, 0)
    scenes.upload(scene, desc)
    ora = scenes.run_oracle(desc, p)
    scene.set_option("anchored", 0)  # the tree search is the user of the pool
    scene.set_option("npool_cap", 64)
    scenes.compare(scene.render(p), ora)
    scene.set_option("npool_cap", 0)
    scenes.compare(scene.render(p), ora)


def test_config5_ten_million_triangles_4k_shadows(scene):
    """BASELINE.json configs[4]: the cat instanced to 9,999,666 triangles (5.05 M BVH nodes, depth 40, largest leaf
    3,075 triangles), 3840x2160 primary + shadow. The scene is built by the product's host code; parity through
    size-independent properties: idempotence, 8-way row-interleaved shards reassemble to the whole frame with the same
    ray count, and the oracle (given the same interchange arrays) on a sample of rows."""
    from raytracinggpu_b200 import synthetic
    from oracle import pyoracle
    cat = pyoracle.cat_obj_path()
    if cat is None:
        pytest.skip("cat asset unavailable")
    scales, offs = synthetic.instance_lattice()
    mesh = rt.Mesh.read_obj(cat).instance(scales, offs).build_bvh_gpu(0)   # the reference's tree, built on the device (tests/test_gpu_build.py)
    assert mesh.counts()[1] == 9999666
    mesh_id = profiles.mesh_material("optimized", 0)["id"]
    scene.set_spheres(profiles.walls("optimized"))
    scene.set_light(*profiles.LIGHT)
    scene.set_mesh_from(mesh, id=mesh_id)                                   # arrays never leave the device (rt_scene_set_mesh_device)
    desc = dict(spheres=profiles.walls("optimized"), mesh=(mesh.vertices, mesh.tri_records, mesh.arr_bvh),  # host mirror, for the oracle
                mesh_mat=profiles.mesh_material("optimized", 0), light=profiles.LIGHT)
    W, H = 3840, 2160
    p = profiles.params("optimized", W, H, 1, 1)
    a = scene.render(p)
    b = scene.render(p, want=("rgb", "hit_tri"))
    assert np.array_equal(a["rgb"], b["rgb"]) and np.array_equal(a["hit_tri"], b["hit_tri"])
    assert a["stats"]["rays"] >= W * H
    assert (a["hit_obj"] == mesh_id).sum() > 100000  # the instances are actually visible
    parts, rays = [], 0
    for r in range(8):
        q = profiles.params("optimized", W, H, 1, 1)
        q.row_begin, q.row_step, q.row_count = rt.sharding.rows_for_rank(H, r, 8)
        o = scene.render(q, want=("rgb",))
        parts.append(o["rgb"])
        rays += o["stats"]["rays"]
    assert np.array_equal(rt.sharding.assemble(np.stack(parts), H, 8), a["rgb"])
    assert rays == a["stats"]["rays"]
    # oracle on every 240th row (9 rows), same arrays
    q = profiles.params("optimized", W, H, 1, 1)
    q.row_begin, q.row_step, q.row_count = 100, 240, 0
    scenes.compare(scene.render(q), scenes.run_oracle(desc, q))


def test_overflow_of_an_earlier_no_sync_frame_is_not_lost(gpu):
    """the overflow flags are sticky on the device — a frame enqueued with RT_RENDER_NO_SYNC that overflowed the (ray,
    leaf) task buffer is reported by rt_scene_sync (RT_ERR_AGAIN) even when later frames did not overflow."""
    import torch
    sc = rt.Scene(gpu)
    try:
        sc.set_option("anchored", 1)
        sc.set_option("task_factor", 1)
        desc = scenes.cat_scene("cpu") or scenes.torus_scene("cpu")
        scenes.upload(sc, desc)
        big = profiles.params("cpu", 640, 360, 1, 0)
        big.cam[2] = 30.0                      # the mesh fills the view: more than one task per pixel
        small = profiles.params("cpu", 640, 360, 1, 0)
        small.cam[2] = 400.0                   # far away: a handful of tasks
        buf = torch.zeros((360, 640, 3), dtype=torch.uint8, device="cuda")
        sc.render_into(big, rgb=buf, flags=rt.RT_RENDER_NO_SYNC)     # overflows
        for _ in range(3):
            sc.render_into(small, rgb=buf, flags=rt.RT_RENDER_NO_SYNC)  # do not
        with pytest.raises(rt.RtError) as e:
            sc.sync()
        assert e.value.code == -7  # RT_ERR_AGAIN
        assert sc.get_option("task_factor") == 2
        sc.sync()  # the flags were consumed
        # rendered again (a few doublings later) the frame is complete and equals the tree search's
        sc.set_option("task_factor", 16)
        a = sc.render(big)
        sc.set_option("anchored", 0)
        b = sc.render(big)
        for k in ("rgb", "hit_tri", "shadow"):
            assert np.array_equal(a[k], b[k]), k
    finally:
        sc.close()


def test_big_leaf_guard_travels_with_the_blob(gpu):
    """max_leaf is part of the scene header, so a scene adopted with rt_scene_blob_import takes the same kernel variant as
    its source. 70,000 triangles of which 40,000 share one centroid (a leaf the builder cannot split): with push order 0 the in-leaf
    offset does not fit the tie-break rank and both scenes must fall back to the one-kernel variant."""
    import torch
    rng = np.random.RandomState(3)
    n_fan, n_rest = 40000, 30000
    verts, tris = [], []
    q = rng.randint(-256, 257, size=(n_fan, 4)).astype(np.float32) / 64.0  # multiples of 1/64 in [-4, 4]: every sum below is exact
    for a_, b_, c_, d_ in q:  # triangle (p, r, -(p + r)): its centroid is exactly (0, 0, 0), so the builder cannot split the group
        k = len(verts)
        verts += [np.float32([a_, b_, 1.0]), np.float32([c_, d_, -0.5]), np.float32([-a_ - c_, -b_ - d_, -0.5])]
        tris.append((k, k + 1, k + 2))
    base = rng.rand(n_rest, 3).astype(np.float32) * 20 - 10
    for b_ in base:
        k = len(verts)
        verts += [b_, b_ + np.float32([0.3, 0, 0]), b_ + np.float32([0, 0.3, 0.1])]
        tris.append((k, k + 1, k + 2))
    m = rt.Mesh.from_arrays(np.asarray(verts, np.float32), np.asarray(tris, np.int32)).build_bvh()
    assert m.bvh_info()["max_leaf"] > 32768
    src, dst = rt.Scene(gpu), rt.Scene(gpu)
    try:
        src.set_spheres(profiles.walls("cpu"))
        src.set_mesh(m.vertices, m.tri_records, m.arr_bvh, id=6)
        ptr, n = src.blob_export()
        dst.blob_import(ptr, n)
        p = profiles.params("cpu", 160, 90, 1, 0)
        a, b = src.render(p), dst.render(p)
        assert a["stats"]["launches"] == 1 and b["stats"]["launches"] == 1  # the one-kernel variant on BOTH
        for k in ("rgb", "hit_obj", "hit_tri", "shadow"):
            assert np.array_equal(a[k], b[k]), k
    finally:
        src.close()
        dst.close()


def test_full_size_config0_spheres_800x600(scene):
    """BASELINE.json configs[0] at its named size: the spheres scene of cpu_launcher.cpp:668-678 (mirror sphere, refractive shell), point
    light, 800x600, 1 spp, cpu_launcher.cpp knobs with num_bounce 5 (6 path segments): every output against the oracle."""
    desc = scenes.spheres_scene()
    scenes.upload(scene, desc)
    p = profiles.params("cpu", 800, 600, 1, 5)
    got = scene.render(p)
    ora = scenes.run_oracle(desc, p)
    res = scenes.compare(got, ora)
    assert res["rgb_exact_mismatch"] == 0, res
    assert got["stats"]["rays"] == ora["work"]["rays"]


def test_full_size_config3_animation_frames_1080p(scene):
    """BASELINE.json configs[3] at its named size: frames of the 240-frame light orbit (realtime_render.cu:1072-1090, one revolution) of the
    spheres scene at 1920x1080 — frames 0, 80 and 160 against the oracle, every rank deriving the same light positions in float."""
    desc = scenes.spheres_scene()
    scenes.upload(scene, desc)
    omega = 2 * np.pi / (240 * 0.02)
    orbit = rt.sharding.light_positions((-10.0, 20.0, 40.0), 240, omega, 0.02, rt.move_light)
    p = profiles.params("cpu", 1920, 1080, 1, 5)
    for f in (0, 80, 160):
        scene.set_light(orbit[f], 3e10)
        d = scenes.spheres_scene(light=(orbit[f], 3e10))
        got = scene.render(p, want=("rgb", "hit_obj", "shadow"))
        ora = scenes.run_oracle(d, p, want=("rgb", "hit_obj", "shadow"))
        res = scenes.compare(got, ora)
        assert res["rgb_exact_mismatch"] == 0, (f, res)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("group,world,H", [(4, 3, 187), (8, 8, 200), (64, 2, 150)])
def test_row_group_shards_match_the_oracle(scene, variant, group, world, H):
    """rt_params.row_group: a rank's share of groups of consecutive rows (compact row k = image row begin + (k / G) step + k % G), rendered by
    the wavefront pipeline and by the thread-per-pixel kernel, equals the oracle's rendering of the same rt_params in every output; ragged
    frame heights (a last group cut short, a rank without rows) included."""
    desc = scenes.cat_scene("optimized", mirror=1) or scenes.torus_scene("optimized", mirror=1)
    scene.set_option("variant", variant)
    scenes.upload(scene, desc)
    W = 256
    total = 0
    for r in range(world):
        p = profiles.params("optimized", W, H, 1, 3)
        n = rt.shard_rows(p, r, world, group)
        total += n
        if n == 0:
            continue
        got = scene.render(p)
        assert got["rgb"].shape[0] == n
        ora = scenes.run_oracle(desc, p)
        res = scenes.compare(got, ora)
        assert res["rgb_exact_mismatch"] == 0, (r, res)
    assert total == H
    # a group that is not a power of two, or larger than the step, is refused
    p = profiles.params("optimized", W, H, 1, 1)
    p.row_group = 3
    with pytest.raises(rt.RtError):
        scene.render(p)
    p.row_group, p.row_step = 8, 4
    with pytest.raises(rt.RtError):
        scene.render(p)
