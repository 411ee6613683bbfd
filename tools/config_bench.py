"""Kernel time (CUDA events inside rt_render, L2 flushed between frames) of every BASELINE.json config on ONE GPU.
Not the bench contract (bench.py is); a table for profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import raytracinggpu_b200 as rt
from raytracinggpu_b200 import synthetic
from oracle import profiles, scenes, pyoracle

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(sc, p, n=15, count=True):
    rows = p.H
    rgb = torch.empty((rows, p.W, 3), dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(n + 3):
        flush.zero_(); torch.cuda.synchronize()
        st = sc.render_into(p, rgb=rgb)
        if i >= 3:
            ms.append(st.kernel_ms)
    work = sc.render(p, want=("rgb",), count_work=True)["stats"] if count else {}
    return float(np.median(ms)), int(st.rays), int(st.launches), work


out = []
sc = rt.Scene(0)
# config 1: spheres scene 800x600 (cpu profile, 6 segments)
desc = scenes.spheres_scene(); scenes.upload(sc, desc)
for b in (0, 5):
    p = profiles.params("cpu", 800, 600, 1, b)
    ms, rays, l, w = timeit(sc, p)
    t0 = time.time(); o = scenes.run_oracle(desc, p, want=("rgb",)); cpu_s = o["work"]["seconds"]
    out.append(("configs[0] spheres 800x600 cpu knobs num_bounce=%d" % b, ms, rays, l, w, "oracle port %.1f ms on %d threads" % (cpu_s * 1e3, o["work"]["threads"])))
# config 2
desc = scenes.cat_scene("optimized"); scenes.upload(sc, desc)
p = profiles.params("optimized", 1920, 1080, 1, 1)
out.append(("configs[1] cat 1920x1080 primary+shadow", *timeit(sc, p), ""))
# config 3
desc = scenes.cat_scene("optimized", mirror=1); scenes.upload(sc, desc)
p = profiles.params("optimized", 3840, 2160, 1, 4)
out.append(("configs[2] mirror cat 3840x2160 depth 4 (1 GPU)", *timeit(sc, p), ""))
# config 4: one frame of the spheres animation at 1080p
desc = scenes.spheres_scene(); scenes.upload(sc, desc)
p = profiles.params("cpu", 1920, 1080, 1, 5)
out.append(("configs[3] spheres 1920x1080, one animation frame (6 segments)", *timeit(sc, p), ""))
# config 5
cat = pyoracle.cat_obj_path()
scales, offs = synthetic.instance_lattice()
t0 = time.time(); mesh = rt.Mesh.read_obj(cat).instance(scales, offs).build_bvh(); build_s = time.time() - t0
desc = dict(spheres=profiles.walls("optimized"), mesh=(mesh.vertices, mesh.tri_records, mesh.arr_bvh), mesh_mat=profiles.mesh_material("optimized", 0), light=profiles.LIGHT)
t0 = time.time(); scenes.upload(sc, desc); torch.cuda.synchronize(); up_s = time.time() - t0
p = profiles.params("optimized", 3840, 2160, 1, 1)
out.append(("configs[4] 9,999,666-triangle instanced cat 3840x2160 primary+shadow (1 GPU)", *timeit(sc, p, n=8), "host BVH build %.1f s, upload+repack %.2f s, blob %.0f MB" % (build_s, up_s, sc.blob_size() / 1e6)))
print("| config | ms/frame | rays/frame | Mrays/s | launches | node visits | triangle tests | note |")
print("|---|---|---|---|---|---|---|---|")
for name, ms, rays, l, w, note in out:
    print("| %s | %.4f | %d | %.0f | %d | %s | %s | %s |" % (name, ms, rays, rays / ms / 1e3, l, w.get("node_visits", ""), w.get("tri_tests", ""), note))
