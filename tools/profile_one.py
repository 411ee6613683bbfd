"""Short program for ncu: the bench workload (cat 1080p primary+shadow), 3 warm-up + N frames, nothing else."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import raytracinggpu_b200 as rt  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Wd = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
Hd = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0)
sc.set_option("graph", 0)  # ncu: plain launches, one per kernel
for kv in os.environ.get("RT_OPTS", "").split(","):  # e.g. RT_OPTS=six=0,strips=1
    if kv:
        sc.set_option(kv.split("=")[0], int(kv.split("=")[1]))
sc.set_spheres(walls)
sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
p = rt.params_profile("optimized", Wd, Hd, 1, 1)
import torch  # noqa: E402
rgb = torch.empty((Hd, Wd, 3), dtype=torch.uint8, device="cuda")  # a device buffer, as in bench.py's timed loop (two row bands)
st = None
for i in range(3 + n):
    st = sc.render_into(p, rgb=rgb)
print(name, st.kernel_ms, st.rays, st.launches)
