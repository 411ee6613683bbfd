"""Where the e2e step (mesh re-upload + render + frame D2H) spends its wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import raytracinggpu_b200 as rt
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
verts, recs, bvh = mesh.vertices, mesh.tri_records, mesh.arr_bvh
sc = rt.Scene(0); sc.set_spheres(walls); sc.set_mesh(verts, recs, bvh, id=mesh_id)
p = rt.params_profile("optimized", 1920, 1080, 1, 1)
host = torch.empty((1080, 1920, 3), dtype=torch.uint8).pin_memory()
dev = torch.empty((1080, 1920, 3), dtype=torch.uint8, device="cuda")
def t(f, n=30):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("set_mesh                 %.3f ms" % t(lambda: sc.set_mesh(verts, recs, bvh, id=mesh_id)))
print("render -> device buffer  %.3f ms (sync)" % t(lambda: sc.render_into(p, rgb=dev)))
print("render -> pinned host    %.3f ms (sync, incl. D2H 6.2 MB)" % t(lambda: sc.render_into(p, rgb=host.numpy())))
print("D2H 6.2 MB alone         %.3f ms" % t(lambda: host.copy_(dev)))
print("full e2e step            %.3f ms" % t(lambda: (sc.set_mesh(verts, recs, bvh, id=mesh_id), sc.render_into(p, rgb=host.numpy()))))
