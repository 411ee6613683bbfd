"""A/B timing of the three representative workloads: configs[1] (latency/tail-bound), stochastic 4 3 at 1080p and
configs[4] (issue-bound). Median kernel ms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raytracinggpu_b200 as rt
from raytracinggpu_b200 import synthetic
from oracle import profiles, scenes, pyoracle
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(sc, p, n=20):
    rgb = torch.empty((p.H, p.W, 3), dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(n + 4):
        flush.zero_(); torch.cuda.synchronize()
        st = sc.render_into(p, rgb=rgb)
        if i >= 4: ms.append(st.kernel_ms)
    return float(np.median(ms)), float(min(ms))
sc = scenes.upload(rt.Scene(0), scenes.cat_scene("optimized"))
print("configs[1] deterministic 1080p      median %.4f  min %.4f" % timeit(sc, profiles.params("optimized", 1920, 1080, 1, 1), 40))
p = profiles.params("optimized", 1920, 1080, 4, 3); p.aa_sigma, p.indirect = 0.2, 1
if len(sys.argv) > 1 and sys.argv[1] == "det": sys.exit(0)
print("stochastic 4 3 1080p                median %.4f  min %.4f" % timeit(sc, p, 8))
scenes.upload(sc, scenes.cat_scene("optimized", mirror=1))
print("configs[2] mirror 4K depth 4        median %.4f  min %.4f" % timeit(sc, profiles.params("optimized", 3840, 2160, 1, 4), 10))
if len(sys.argv) > 1:
    scales, offs = synthetic.instance_lattice()
    mesh = rt.Mesh.read_obj(pyoracle.cat_obj_path()).instance(scales, offs).build_bvh()
    desc = dict(spheres=profiles.walls("optimized"), mesh=(mesh.vertices, mesh.tri_records, mesh.arr_bvh), mesh_mat=profiles.mesh_material("optimized", 0), light=profiles.LIGHT)
    scenes.upload(sc, desc)
    print("configs[4] 10M tris 4K              median %.4f  min %.4f" % timeit(sc, profiles.params("optimized", 3840, 2160, 1, 1), 6))
