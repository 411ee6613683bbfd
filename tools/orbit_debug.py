import os, sys
os.environ["RT_DEBUG_BINS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import raytracinggpu_b200 as rt
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0); sc.set_spheres(walls); sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
p = rt.params_profile("optimized", 1920, 1080, 1, 1)
dev = torch.empty((1080, 1920, 3), dtype=torch.uint8, device="cuda")
omega = 2 * np.pi / (240 * 0.02)
orbit = rt.sharding.light_positions((-10.0, 20.0, 40.0), 23, omega, 0.02, rt.move_light)
print(orbit[:3], orbit[-1])
for i in range(23):
    sc.set_light(orbit[i], 3e10)
    st = sc.render_into(p, rgb=dev)
    print(i, round(st.kernel_ms, 4))
