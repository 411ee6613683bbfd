"""configs[2] (mirror cat 3840x2160, depth 4): every rank's share of 8 rendered on one GPU, by rows per group: the slowest rank is what an
8-GPU frame waits for."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import raytracinggpu_b200 as rt
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0)
sc.set_spheres(walls)
sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, mirror=1, id=mesh_id)
W, H, N = 3840, 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 8
for G in (1, 2, 4, 8, 16, 32):
    res = []
    for r in range(N):
        p = rt.params_profile("optimized", W, H, 1, 4)
        rows = rt.shard_rows(p, r, N, G)
        rgb = torch.empty((rows, W, 3), dtype=torch.uint8, device="cuda")
        ms = []
        for i in range(12):
            st = sc.render_into(p, rgb=rgb)
            if i >= 4:
                ms.append(st.kernel_ms)
        res.append((float(np.median(ms)), int(st.rays)))
    print("G=%2d  ms per rank %s  max %.4f  mean %.4f  rays min/max %d/%d" % (G, [round(x[0], 4) for x in res], max(x[0] for x in res), np.mean([x[0] for x in res]),
                                                                          min(x[1] for x in res), max(x[1] for x in res)), flush=True)
