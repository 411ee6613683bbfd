"""Per-launch times of one frame (RT_DEBUG_TIMES: one strip, an event after every launch).
    python tools/times_debug.py [det|stoch|mirror4k]"""
import os, sys
os.environ["RT_DEBUG_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import raytracinggpu_b200 as rt
mode = sys.argv[1] if len(sys.argv) > 1 else "det"
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0)
sc.set_spheres(walls)
sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, mirror=1 if mode == "mirror4k" else 0, id=mesh_id)
p = rt.params_profile("optimized", 1920, 1080, 1, 1)
if mode == "stoch":
    p = rt.params_profile("optimized", 1920, 1080, 2, 3)
    p.aa_sigma, p.indirect = 0.2, 1
if mode == "mirror4k":
    p = rt.params_profile("optimized", 3840, 2160, 1, 4)
for i in range(4):
    o = sc.render(p, want=("rgb",))
print(o["stats"])
