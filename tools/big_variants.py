"""configs[4] (10 M triangles, 4K, primary + shadow) through the render variants: wavefront task pools (2) against the thread-per-pixel
kernel with a per-lane stack (1: certified fast paths, 0: exact arithmetic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raytracinggpu_b200 as rt
from raytracinggpu_b200 import synthetic
from oracle import profiles, pyoracle
scales, offs = synthetic.instance_lattice()
mesh = rt.Mesh.read_obj(pyoracle.cat_obj_path()).instance(scales, offs).build_bvh_gpu(0)
sc = rt.Scene(0)
sc.set_spheres(profiles.walls("optimized"))
sc.set_light(*profiles.LIGHT)
sc.set_mesh_from(mesh, id=1)
p = profiles.params("optimized", 3840, 2160, 1, 1)
rgb = torch.empty((2160, 3840, 3), dtype=torch.uint8, device="cuda")
ref = None
for variant, extra in ((2, {}), (2, {"top_smem": 1}), (2, {"top_smem": 0}), (1, {})):
    sc.set_option("variant", variant)
    for k, v in extra.items():
        sc.set_option(k, v)
    ms = []
    for i in range(5):
        st = sc.render_into(p, rgb=rgb)
        ms.append(st.kernel_ms)
    img = rgb.cpu().numpy().copy()
    if ref is None:
        ref = img
    print("variant", variant, extra, "kernel_ms", [round(m, 3) for m in ms], "rays", st.rays, "identical", bool(np.array_equal(img, ref)), flush=True)
