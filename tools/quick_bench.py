"""Short timing loop for A/B experiments: cat 1080p primary+shadow, N frames back to back, kernel_ms per frame (events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import raytracinggpu_b200 as rt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0)
sc.set_spheres(walls)
sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
p = rt.params_profile("optimized", 1920, 1080, 1, 1)
rgb = torch.empty((1080, 1920, 3), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for i in range(n + 5):
    flush.zero_()
    torch.cuda.synchronize()
    st = sc.render_into(p, rgb=rgb)
    if i >= 5:
        ms.append(st.kernel_ms)
print("strips=%s variant=%s  kernel_ms median %.4f  min %.4f  max %.4f  (L2 flushed, %d frames)" % (os.environ.get("RT_STRIPS", "default"), os.environ.get("RT_VARIANT", "2"), np.median(ms), min(ms), max(ms), n))
