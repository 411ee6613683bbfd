"""Investigation aid: per-warp timeline of wf_traverse (RT_DEBUG_WARPS): when each persistent warp started and
finished, how many N/T steps and tasks it ran, how many batches it admitted, on which SM."""
import os, sys
os.environ["RT_DEBUG_WARPS"] = "/tmp/warps.bin"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytracinggpu_b200 as rt
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0)
sc.set_spheres(walls)
sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
p = rt.params_profile("optimized", 1920, 1080, 1, 1)
NR = 2
if os.environ.get("RT_TL_SHARD"):  # one rank's share of configs[2]: 4K mirror cat, depth 4, rows r, r+world, ...
    world = int(os.environ["RT_TL_SHARD"])
    sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, mirror=1, id=mesh_id)
    p = rt.params_profile("optimized", 3840, 2160, 1, 4)
    p.row_begin, p.row_step, p.row_count = rt.sharding.rows_for_rank(2160, 0, world)
    NR = 5
for _ in range(4):
    o = sc.render(p, want=("rgb",), count_work=True)
print(o["stats"])
d = np.fromfile("/tmp/warps.bin", dtype=np.int32).reshape(NR, -1, 16)
for r in range(NR):
    w = d[r]
    t0 = w[:, 0].astype(np.uint32).astype(np.int64); t1 = w[:, 1].astype(np.uint32).astype(np.int64)
    base = t0.min()
    s, e = (t0 - base) / 1e3, (t1 - base) / 1e3
    steps = w[:, 2] + w[:, 4]
    print("round %d: warps %d  start us p0/p50/p100 %.1f %.1f %.1f   end us p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (
        r, len(w), s.min(), np.median(s), s.max(), *np.percentile(e, [10, 50, 90, 99]), e.max()))
    print("   steps/warp mean %.1f p50 %.0f p90 %.0f p99 %.0f max %d   N tasks/step %.1f  T tasks/step %.1f  batches/warp mean %.1f max %d" % (
        steps.mean(), np.median(steps), np.percentile(steps, 90), np.percentile(steps, 99), steps.max(),
        w[:, 3].sum() / max(w[:, 2].sum(), 1), w[:, 5].sum() / max(w[:, 4].sum(), 1), w[:, 6].mean(), w[:, 6].max()))
    dur = e - s
    print("   us per step: mean %.2f ; by step-count decile:" % (dur.sum() / steps.sum()), np.round([dur[steps >= q].sum() / steps[steps >= q].sum() for q in np.percentile(steps, [0, 50, 90, 99])], 2))
    # how many warps are still running at time t
    for frac in (0.5, 0.6, 0.7, 0.8, 0.9, 0.95):
        t = e.max() * frac
        print("   t=%.0f us (%.0f%%): %d warps running on %d SMs" % (t, frac * 100, int(((s <= t) & (e > t)).sum()), len(set(w[(s <= t) & (e > t), 7]))))
    cyc = w[:, 8:12].astype(np.int64) * 16
    print("   cycles per N step %.0f  per T step %.0f  per admission %.0f  retire per loop %.0f ; share of warp time N %.2f T %.2f A %.2f R %.2f" % (
        cyc[:, 0].sum() / max(w[:, 2].sum(), 1), cyc[:, 1].sum() / max(w[:, 4].sum(), 1), cyc[:, 2].sum() / max(w[:, 6].sum(), 1), cyc[:, 3].sum() / max(steps.sum(), 1),
        *(cyc.sum(axis=0) / cyc.sum())))
    tail = e > np.percentile(e, 99)
    print("   tail warps (last 1%%): cycles per N step %.0f  per T step %.0f  per admission %.0f" % (
        cyc[tail, 0].sum() / max(w[tail, 2].sum(), 1), cyc[tail, 1].sum() / max(w[tail, 4].sum(), 1), cyc[tail, 2].sum() / max(w[tail, 6].sum(), 1)))
    late = np.argsort(e)[-8:]
    print("   last finishers: end us", np.round(e[late], 1), "steps", steps[late], "batches", w[late, 6])
np.save("gpurun_out/warps.npy", d)
