"""Short program for ncu: BASELINE.json configs[4] (9,999,666-triangle instanced cat, 3840x2160 primary + shadow)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import raytracinggpu_b200 as rt
from raytracinggpu_b200 import synthetic
from oracle import profiles, scenes, pyoracle
scales, offs = synthetic.instance_lattice()
mesh = rt.Mesh.read_obj(pyoracle.cat_obj_path()).instance(scales, offs).build_bvh_gpu(0)
sc = rt.Scene(0)
sc.set_option("graph", 0)  # ncu: plain launches
sc.set_spheres(profiles.walls("optimized"))
sc.set_light(*profiles.LIGHT)
sc.set_mesh_from(mesh, id=1)
p = profiles.params("optimized", 3840, 2160, 1, 1)
rgb = torch.empty((2160, 3840, 3), dtype=torch.uint8, device="cuda")
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    st = sc.render_into(p, rgb=rgb)
print("kernel_ms", st.kernel_ms, "rays", st.rays, "launches", st.launches)
