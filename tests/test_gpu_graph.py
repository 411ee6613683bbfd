"""A frame whose plan repeats is recorded into a CUDA graph and replayed (rt_render): the replayed frames must equal
directly launched ones — device outputs, pinned host outputs copied back band by band, side-stream tree searches, several sample
passes — and any change of the scene or the call must be picked up."""
import numpy as np
import pytest
import torch

import raytracinggpu_b200 as rt
from oracle import profiles, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built):
    if rt.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box (there is no CPU fallback)")
    return 0


def stoch(p, sigma=0.2):
    p.aa_sigma, p.indirect = sigma, 1
    return p


@pytest.mark.parametrize("name,mirror,make", [
    ("primary_shadow", 0, lambda: profiles.params("optimized", 640, 360, 1, 1)),
    ("mirror_depth3", 1, lambda: profiles.params("optimized", 480, 270, 1, 3)),
    ("stochastic_2_3", 0, lambda: stoch(profiles.params("optimized", 320, 180, 2, 3))),
    ("stochastic_1_1", 0, lambda: stoch(profiles.params("optimized", 640, 360, 1, 1))),
])
def test_replayed_frames_equal_direct_launches(gpu, name, mirror, make):
    desc = scenes.cat_scene("optimized", mirror=mirror) or scenes.torus_scene("optimized", mirror=mirror)
    ref_scene, sc = rt.Scene(gpu), rt.Scene(gpu)
    try:
        ref_scene.set_option("graph", 0)
        scenes.upload(ref_scene, desc)
        scenes.upload(sc, desc)
        p = make()
        want = ref_scene.render(p, want=("rgb",))["rgb"]
        dev = torch.zeros((p.H, p.W, 3), dtype=torch.uint8, device="cuda")
        host = torch.zeros((p.H, p.W, 3), dtype=torch.uint8).pin_memory()
        for out in (dev, host.numpy()):
            for k in range(5):  # direct, recorded, replayed x3
                if isinstance(out, torch.Tensor):
                    out.zero_()
                else:
                    out[:] = 0
                st = sc.render_into(p, rgb=out)
                got = out.cpu().numpy() if isinstance(out, torch.Tensor) else out
                assert np.array_equal(got, want), (name, k)
        # a moved light is a different frame: picked up (direct launch), then recorded and replayed again
        L = rt.move_light((-10.0, 20.0, 40.0), 1.309, 0.4)
        ref_scene.set_light(L, 3e10)
        sc.set_light(L, 3e10)
        want2 = ref_scene.render(p, want=("rgb",))["rgb"]
        assert not np.array_equal(want2, want)
        for k in range(4):
            sc.render_into(p, rgb=dev)
            assert np.array_equal(dev.cpu().numpy(), want2), (name, "moved light", k)
        # so is a different band count, and frames enqueued without a sync in between
        sc.set_option("strips", 3)
        for k in range(3):
            sc.render_into(p, rgb=host.numpy(), flags=rt.RT_RENDER_NO_SYNC)
        sc.sync()
        assert np.array_equal(host.numpy(), want2), (name, "no-sync replays")
    finally:
        ref_scene.close()
        sc.close()


def test_pipelined_host_outputs(gpu):
    """Frames enqueued with RT_RENDER_NO_SYNC into HOST buffers leave on the library's copy stream from two alternating scratch sets,
    and rt_scene_set_mesh's fast path alternates between two pinned halves (rt_device.cu): whatever the interleaving — a changing
    light, a mesh re-uploaded with moved vertices, pageable and pinned destinations, all five outputs, a synchronous frame in
    between — every host buffer must hold exactly the frame a synchronous render of the same state returns."""
    desc = scenes.cat_scene("optimized") or scenes.torus_scene("optimized")
    ref_scene, sc = rt.Scene(gpu), rt.Scene(gpu)
    try:
        ref_scene.set_option("graph", 0)
        scenes.upload(ref_scene, desc)
        scenes.upload(sc, desc)
        p = profiles.params("optimized", 640, 360, 1, 1)
        v0, recs, bvh = desc["mesh"]
        mm = desc["mesh_mat"]
        kw = dict(albedo=mm["albedo"], mirror=mm["mirror"], n_in=mm["n_in"], n_out=mm["n_out"], id=mm["id"])
        verts = [np.ascontiguousarray(v0, dtype=np.float32)]
        moved = verts[0].copy()
        moved[::7] += np.float32(0.01)  # same tree (arr_bvh unchanged: boxes are not re-fitted), different triangles
        verts.append(moved)
        n = 9
        lights = [rt.move_light((-10.0, 20.0, 40.0), 1.309, 0.05 * (k // 3)) for k in range(n)]
        names = ("rgb", "hit_obj", "hit_tri", "hit_t", "shadow")
        want = []
        for k in range(n):
            ref_scene.set_light(lights[k], 3e10)
            ref_scene.set_mesh(verts[k & 1], recs, bvh, **kw)
            want.append(ref_scene.render(p, want=names))
        # distinct destinations for every frame: nothing may be torn by a later frame
        outs = []
        for k in range(n):
            o = {}
            for nm in names:
                w = want[k][nm]
                t = torch.zeros(w.shape, dtype=torch.from_numpy(w).dtype)
                o[nm] = (t.pin_memory() if k % 2 == 0 else t).numpy()
            outs.append(o)
        for k in range(n):
            sc.set_light(lights[k], 3e10)
            sc.set_mesh(verts[k & 1], recs, bvh, **kw)
            if k == 5:  # a synchronous frame in the middle of the pipeline (uses scratch set 0 as well)
                sc.render_into(p, **outs[k])
            else:
                sc.render_into(p, flags=rt.RT_RENDER_NO_SYNC, **outs[k])
        sc.sync()
        for k in range(n):
            for nm in names:
                assert np.array_equal(outs[k][nm], want[k][nm]), (k, nm)
        # a static scene: the two scratch sets each record a graph and replay it
        a = torch.zeros((p.H, p.W, 3), dtype=torch.uint8).pin_memory()
        b = torch.zeros((p.H, p.W, 3), dtype=torch.uint8).pin_memory()
        for k in range(8):
            sc.set_mesh(verts[0], recs, bvh, **kw)
            sc.render_into(p, rgb=(a if k % 2 == 0 else b).numpy(), flags=rt.RT_RENDER_NO_SYNC)
        st = sc.sync()
        ref_scene.set_mesh(verts[0], recs, bvh, **kw)
        w = ref_scene.render(p, want=("rgb",))["rgb"]
        assert np.array_equal(a.numpy(), w) and np.array_equal(b.numpy(), w)
        assert st.rays > 0
    finally:
        ref_scene.close()
        sc.close()
