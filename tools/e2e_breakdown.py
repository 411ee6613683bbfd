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

# pipelined steps (RT_RENDER_NO_SYNC into two alternating pinned frames): host enqueue time against completion time
hosts = [torch.empty((1080, 1920, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
def pipe(n, upload=True, out="host"):
    t0 = time.perf_counter()
    for i in range(n):
        if upload:
            sc.set_mesh(verts, recs, bvh, id=mesh_id)
        sc.render_into(p, rgb=hosts[i & 1].numpy() if out == "host" else dev, flags=rt.RT_RENDER_NO_SYNC)
    t1 = time.perf_counter()
    sc.sync()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3
for upload in (True, False):
    for out in ("host", "device"):
        pipe(6, upload, out)
        e, c = pipe(40, upload, out)
        print("pipelined upload=%d out=%-6s  enqueue %.3f ms/step  complete %.3f ms/step" % (upload, out, e, c))
def host_only(n=200):
    t0 = time.perf_counter()
    for i in range(n):
        sc.set_mesh(verts, recs, bvh, id=mesh_id)
    t1 = time.perf_counter(); sc.sync()
    return (t1 - t0) / n * 1e3
host_only(5)
print("set_mesh enqueue only    %.3f ms" % host_only())
