"""configs[4] (10 M triangles, 4K, primary + shadow) across scene options: row bands, admission run length, guided self-scheduling, pool size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raytracinggpu_b200 as rt
from raytracinggpu_b200 import synthetic
from oracle import profiles, pyoracle
scales, offs = synthetic.instance_lattice()
mesh = rt.Mesh.read_obj(pyoracle.cat_obj_path()).instance(scales, offs).build_bvh_gpu(0)
p = profiles.params("optimized", 3840, 2160, 1, 1)
rgb = torch.empty((2160, 3840, 3), dtype=torch.uint8, device="cuda")
base = None
for opts in ({}, {"strips": 1}, {"strips": 3}, {"strips": 4}, {"run_shift": 2}, {"run_shift": 4}, {"run_shift": 5}, {"gss": 0}, {"gss": 8}, {"fair_share": 0}, {"side_stream": 0},
             {"npool_cap": 128}, {"npool_cap": 512}):
    sc = rt.Scene(0)
    for k, v in opts.items():
        sc.set_option(k, v)
    sc.set_spheres(profiles.walls("optimized"))
    sc.set_light(*profiles.LIGHT)
    sc.set_mesh_from(mesh, id=1)
    ms = []
    for i in range(5):
        st = sc.render_into(p, rgb=rgb)
        ms.append(st.kernel_ms)
    img = rgb.cpu().numpy().copy()
    if base is None:
        base = img
    print(opts, "kernel_ms median %.3f" % float(np.median(ms[1:])), "identical", bool(np.array_equal(img, base)), flush=True)
    sc.close()
