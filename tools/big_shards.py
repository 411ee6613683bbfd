"""configs[4] (10 M triangles, 4K): one rank's share of the frame for 8 ranks, rows interleaved (row % 8 == rank) against contiguous bands
(rank k renders rows k*270 .. k*270+269): how much of the 4.7x at 8 GPUs is lost coherence, and whether contiguous bands are balanced."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raytracinggpu_b200 as rt
from raytracinggpu_b200 import synthetic
from oracle import profiles, pyoracle
scales, offs = synthetic.instance_lattice()
mesh = rt.Mesh.read_obj(pyoracle.cat_obj_path()).instance(scales, offs).build_bvh_gpu(0)
sc = rt.Scene(0)
sc.set_spheres(profiles.walls("optimized"))
sc.set_light(*profiles.LIGHT)
sc.set_mesh_from(mesh, id=1)
W, H, N = 3840, 2160, 8
def run(p, rows):
    rgb = torch.empty((rows, W, 3), dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(5):
        st = sc.render_into(p, rgb=rgb)
        ms.append(st.kernel_ms)
    return float(np.median(ms[1:])), int(st.rays)
p = profiles.params("optimized", W, H, 1, 1)
print("whole frame", run(p, H), flush=True)
for mode in ("interleaved", "contiguous", "groups of 4 rows", "groups of 8 rows", "groups of 16 rows", "groups of 32 rows"):
    res = []
    for k in range(N):
        p = profiles.params("optimized", W, H, 1, 1)
        if mode == "interleaved":
            p.row_begin, p.row_step, p.row_count = k, N, H // N
        elif mode == "contiguous":
            p.row_begin, p.row_step, p.row_count = k * (H // N), 1, H // N
        else:
            rt.shard_rows(p, k, N, int(mode.split()[2]))
        res.append(run(p, p.row_count if p.row_count > 0 else H // N))
    if res:
        print(mode, "ms per rank", [round(r[0], 3) for r in res], "max %.3f" % max(r[0] for r in res), "rays", [r[1] for r in res], flush=True)
