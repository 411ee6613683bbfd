"""Concurrent timeline of one frame's launches (globaltimer stamps written by the kernels of a -DRT_TIMELINE build):
    make -C raytracinggpu_b200/csrc OBJ=build_tl OUT=../../ab/tl EXTRA=-DRT_TIMELINE ../../ab/tl/librtb200.so
    RT_LIB_PATH=$PWD/ab/tl/librtb200.so RT_TIMELINE_PRINT=1 python tools/timeline.py [mode=det|stoch11|shard8] [key=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import raytracinggpu_b200 as rt
args = dict(a.split("=") for a in sys.argv[1:])
mode = args.pop("mode", "det")
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0)
for k, v in args.items():
    sc.set_option(k, int(v))
sc.set_spheres(walls)
sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, mirror=1 if mode in ("shard8", "mirror4k") else 0, id=mesh_id)
W, H = (3840, 2160) if mode in ("shard8", "mirror4k") else (1920, 1080)
p = rt.params_profile("optimized", W, H, 1, 4 if mode in ("shard8", "mirror4k") else 1)
if mode == "stoch11":
    p.aa_sigma, p.indirect = 0.2, 1
if mode == "stoch13":
    p = rt.params_profile("optimized", W, H, 1, 3)
    p.aa_sigma, p.indirect = 0.2, 1
if mode == "shard8":
    rt.shard_rows(p, 0, 8, int(os.environ.get("RT_ROW_GROUP", "16")))
rows = p.row_count if p.row_count > 0 else H
rgb = torch.empty((rows, W, 3), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(8):
    flush.zero_()
    torch.cuda.synchronize()
    st = sc.render_into(p, rgb=rgb)
    print("frame", i, "kernel_ms %.4f" % st.kernel_ms, file=sys.stderr, flush=True)
