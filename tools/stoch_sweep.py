"""./optimized 4 3 at 1080p (stochastic mode) across scene options."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import raytracinggpu_b200 as rt
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
p = rt.params_profile("optimized", 1920, 1080, 4, 3)
p.aa_sigma, p.indirect = 0.2, 1
p.z = rt.camera_z_device(1920)
rgb = torch.empty((1080, 1920, 3), dtype=torch.uint8, device="cuda")
for opts in ({}, {"wide": 0}, {"strips": 1}, {"strips": 2}, {"strips": 3}, {"side_stream": 0}, {"fair_share": 0}, {"run_shift": 2}, {"run_shift": 3}, {"gss": 0}, {"gss": 8}, {"bins_r": 2048}):
    sc = rt.Scene(0)
    for k, v in opts.items():
        sc.set_option(k, v)
    sc.set_spheres(walls)
    sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
    ms = []
    for i in range(6):
        st = sc.render_into(p, rgb=rgb)
        if i >= 2:
            ms.append(st.kernel_ms)
    print(opts, "kernel_ms median %.4f" % np.median(ms), "crc %08x" % zlib.crc32(rgb.cpu().numpy().tobytes()), "launches", st.launches, flush=True)
    sc.close()
