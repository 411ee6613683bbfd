"""Cost of (re)building the anchored-ray bins: frames with a fixed light vs frames of a light orbit (bins rebuilt every frame)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import raytracinggpu_b200 as rt
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0); sc.set_spheres(walls); sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
p = rt.params_profile("optimized", 1920, 1080, 1, 1)
dev = torch.empty((1080, 1920, 3), dtype=torch.uint8, device="cuda")
def run(move, n=40):
    L = (-10.0, 20.0, 40.0); ms = []
    for i in range(n + 5):
        if move: L = rt.move_light(L, 1.309, 0.02)
        sc.set_light(L, 3e10)
        st = sc.render_into(p, rgb=dev)
        if i >= 5: ms.append(st.kernel_ms)
    return float(np.median(ms))
print("fixed light  %.4f ms/frame" % run(False))
print("moving light %.4f ms/frame (light bins rebuilt every frame)" % run(True))

# the bench's way: frames enqueued without synchronising, L2 flushed between frames, one event pair per frame
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream(); sc.set_stream(stream.cuda_stream)
def run_async(move, do_flush, n=30):
    L = (-10.0, 20.0, 40.0)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n + 5):
        if move: L = rt.move_light(L, 1.309, 0.02)
        sc.set_light(L, 3e10)
        with torch.cuda.stream(stream):
            if do_flush: flush.zero_()
            if i >= 5: evs[i - 5][0].record(stream)
            sc.render_into(p, rgb=dev, flags=rt.RT_RENDER_NO_SYNC)
            if i >= 5: evs[i - 5][1].record(stream)
    sc.sync(); torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return ms[len(ms) // 2], ms[-1], sum(ms) / len(ms)
for move in (False, True):
    for fl in (False, True):
        print("enqueued, move %d flush %d: median %.4f max %.4f mean %.4f ms" % ((move, fl) + run_async(move, fl)))
