"""Scene set-up time of BASELINE.json configs[4] (10 M triangles): host upload vs device-resident upload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raytracinggpu_b200 as rt
from raytracinggpu_b200 import synthetic
from oracle import pyoracle
cat = pyoracle.cat_obj_path()
scales, offs = synthetic.instance_lattice()
walls, mesh_id = rt.default_walls("optimized")
p = rt.params_profile("optimized", 3840, 2160, 1, 1)
rgb = torch.empty((2160, 3840, 3), dtype=torch.uint8, device="cuda")
out = {}
for mode in ("device", "host", "device"):
    t0 = time.perf_counter(); mesh = rt.Mesh.read_obj(cat).instance(scales, offs); t1 = time.perf_counter()
    mesh.build_bvh_gpu(0); t2 = time.perf_counter()
    sc = rt.Scene(0); sc.set_spheres(walls)
    t3 = time.perf_counter()
    if mode == "device":
        sc.set_mesh_from(mesh, id=mesh_id)
    else:
        sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
    sc.sync(); torch.cuda.synchronize(); t4 = time.perf_counter()
    ms = []
    for i in range(6):
        st = sc.render_into(p, rgb=rgb)
        ms.append(st.kernel_ms)
    out[mode] = rgb.cpu().numpy().copy()
    print("%s: instancing %.0f ms, upload + build %.0f ms (device %.1f), scene upload %.0f ms, frame %.3f ms, rays %d, blob %.0f MB" % (
        mode, (t1 - t0) * 1e3, (t2 - t1) * 1e3, mesh.build_ms, (t4 - t3) * 1e3, float(np.median(ms[2:])), st.rays, sc.blob_size() / 1e6), flush=True)
    sc.close(); del mesh
print("frames identical:", bool(np.array_equal(out["device"], out["host"])))
