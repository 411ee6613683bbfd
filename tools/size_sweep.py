"""Where do the anchored-ray bins stop paying? Instanced cats of growing count (and a single big cat), 1080p primary + shadow:
frame time with the bins (RT_ANCHOR=1) and with the tree search (RT_ANCHOR=0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raytracinggpu_b200 as rt
from raytracinggpu_b200 import synthetic
from oracle import profiles, pyoracle
cat = pyoracle.cat_obj_path()
def timeit(sc, p, n=12):
    rgb = torch.empty((p.H, p.W, 3), dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(n + 3):
        st = sc.render_into(p, rgb=rgb)
        if i >= 3: ms.append(st.kernel_ms)
    return float(np.median(ms))
p = profiles.params("optimized", 1920, 1080, 1, 1)
for copies, scale in ((1, 1.0), (8, 0.5), (64, 0.25), (256, 0.125), (512, 0.125), (1024, 0.0625)):
    if copies == 1:
        mesh = rt.Mesh.read_obj(cat).rescale(0.6, (0.0, -4.0, 0.0)).build_bvh()
    else:
        scales, offs = synthetic.instance_lattice(copies, scale=scale)
        mesh = rt.Mesh.read_obj(cat).instance(scales, offs).build_bvh_gpu(0)
    sc = rt.Scene(0)
    sc.set_spheres(profiles.walls("optimized"))
    sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=1)
    out = []
    for mode in ("1", "0"):
        os.environ["RT_ANCHOR"] = mode
        out.append(timeit(sc, p))
    print("%5d copies %8d triangles %7d leaves: bins %.3f ms  tree %.3f ms" % (copies, mesh.counts()[1], mesh.bvh_info()["leaves"], out[0], out[1]))
    sc.close()
