"""A/B of scene options on configs[1] (device-timed, L2 flushed between frames):
    python tools/opt_sweep.py key=v1,v2,... [key2=...] [mode=det|stoch11|mirror4k|shard8]"""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import raytracinggpu_b200 as rt
args = dict(a.split("=") for a in sys.argv[1:])
mode = args.pop("mode", "det")
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
keys = list(args)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for combo in itertools.product(*[args[k].split(",") for k in keys]):
    sc = rt.Scene(0)
    for k, v in zip(keys, combo):
        sc.set_option(k, int(v))
    sc.set_spheres(walls)
    sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, mirror=1 if mode in ("mirror4k", "shard8") else 0, id=mesh_id)
    W, H = (3840, 2160) if mode in ("mirror4k", "shard8") else (1920, 1080)
    p = rt.params_profile("optimized", W, H, 1, 4 if mode in ("mirror4k", "shard8") else 1)
    if mode == "stoch11":
        p.aa_sigma, p.indirect = 0.2, 1
    if mode == "shard8":
        rt.shard_rows(p, 0, 8, int(os.environ.get("RT_ROW_GROUP", "1")))
    rows = p.row_count if p.row_count > 0 else H
    rgb = torch.empty((rows, W, 3), dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(24):
        flush.zero_()
        torch.cuda.synchronize()
        st = sc.render_into(p, rgb=rgb)
        if i >= 4:
            ms.append(st.kernel_ms)
    print(mode, dict(zip(keys, combo)), "kernel_ms median %.4f min %.4f  launches %d rays %d" % (np.median(ms), min(ms), st.launches, st.rays), flush=True)
    sc.close()
