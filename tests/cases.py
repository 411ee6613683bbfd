"""The fixed parity cases: scene + params, shared by the golden generator, the CPU tests and the GPU tests."""
import numpy as np

from oracle import profiles, scenes


def _p(profile, W, H, rays=1, bounce=1, **kw):
    def make():
        p = profiles.params(profile, W, H, rays, bounce)
        for k, v in kw.items():
            setattr(p, k, v)
        return p
    return make


def _lens_mesh():
    v, t = scenes.torus(nu=24, nv=12, R=7.0, r=4.0, center=(2.0, 1.0, 5.0), tilt=1.1)
    return v, t


CASES = {
    # BASELINE.json config 2 at reduced size: cat, primary + shadow, optimized.cu knobs
    "cat_optimized_240x135": dict(scene=lambda: scenes.cat_scene("optimized"), params=_p("optimized", 240, 135, 1, 1), needs_cat=True),
    # the cpu_launcher program (`./cpu 1 0` content) at reduced size
    "cat_cpu_240x135": dict(scene=lambda: scenes.cat_scene("cpu"), params=_p("cpu", 240, 135, 1, 0), needs_cat=True),
    "cat_arraybvh_160x120": dict(scene=lambda: scenes.cat_scene("array_bvh"), params=_p("array_bvh", 160, 120, 1, 1), needs_cat=True),
    # config 3 at reduced size: mirror cat, reflection depth 4
    "cat_mirror_depth4_192x108": dict(scene=lambda: scenes.cat_scene("optimized", mirror=1), params=_p("optimized", 192, 108, 1, 4), needs_cat=True),
    # config 1: spheres scene (mirror + refractive shells), cpu knobs, 6 segments; 2 samples exercise the average
    "spheres_cpu_200x150_b5": dict(scene=scenes.spheres_scene, params=_p("cpu", 200, 150, 2, 5)),
    "spheres_opt_160x120_b1": dict(scene=scenes.spheres_scene, params=_p("optimized", 160, 120, 1, 1)),
    # synthetic meshes (no asset needed)
    "torus_optimized_256x144": dict(scene=lambda: scenes.torus_scene("optimized"), params=_p("optimized", 256, 144, 1, 1)),
    "torus_cpu_mirror_160x90_b3": dict(scene=lambda: scenes.torus_scene("cpu", mirror=1), params=_p("cpu", 160, 90, 1, 3)),
    "torus_glass_160x90_b6": dict(scene=lambda: scenes.mesh_scene("optimized", *_lens_mesh(), n_in=1.5, n_out=1.0), params=_p("optimized", 160, 90, 1, 6)),
    # flat axis-aligned quads: every leaf box has zero thickness and the reference's strict slab test rejects it
    "flatgrid_optimized_128x96": dict(scene=lambda: scenes.mesh_scene("optimized", *scenes.grid_quads()), params=_p("optimized", 128, 96, 1, 1)),
    # ragged size (not a multiple of the 16x8 tile) and a 3-triangle mesh whose root is a leaf
    "tiny_mesh_37x23": dict(scene=lambda: scenes.mesh_scene("array_bvh", np.array([[-8, -8, 0], [8, -8, 0], [0, 8, 2], [0, 0, 9]], np.float32),
                                                             np.array([[0, 1, 2], [0, 1, 3], [1, 2, 3]], np.int32)),
                            params=_p("array_bvh", 37, 23, 1, 2)),
}
