"""Cost of (re)building the anchored-ray bins: frames with a fixed light vs frames of a light orbit (bins rebuilt every frame)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import raytracinggpu_b200 as rt
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0); sc.set_spheres(walls); sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
p = rt.params_profile("optimized", 1920, 1080, 1, 1)
dev = torch.empty((1080, 1920, 3), dtype=torch.uint8, device="cuda")
def run(move, n=40):
    L = (-10.0, 20.0, 40.0); ms = []
    for i in range(n + 5):
        if move: L = rt.move_light(L, 1.309, 0.02)
        sc.set_light(L, 3e10)
        st = sc.render_into(p, rgb=dev)
        if i >= 5: ms.append(st.kernel_ms)
    return float(np.median(ms))
print("fixed light  %.4f ms/frame" % run(False))
print("moving light %.4f ms/frame (light bins rebuilt every frame)" % run(True))
