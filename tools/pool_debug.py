import os, sys
os.environ["RT_DEBUG_POOL"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import raytracinggpu_b200 as rt
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0)
sc.set_spheres(walls)
sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
p = rt.params_profile("optimized", 1920, 1080, 1, 1)
for i in range(2):
    o = sc.render(p, want=("rgb",), count_work=True)
print(o["stats"])
