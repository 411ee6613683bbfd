"""Investigation aid: per-pixel cost maps of the render kernel (RT_DEBUG_COST=1)."""
import os, sys
os.environ["RT_DEBUG_COST"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import raytracinggpu_b200 as rt
mesh, walls, mesh_id, name = bench.build_scene_host(rt)
sc = rt.Scene(0)
sc.set_spheres(walls)
sc.set_mesh(mesh.vertices, mesh.tri_records, mesh.arr_bvh, id=mesh_id)
p = rt.params_profile("optimized", 1920, 1080, 1, 1)
for _ in range(3):
    o = sc.render(p, count_work=True)
clk, work, smid = o["hit_t"], o["hit_tri"], o["hit_obj"]
print("stats", o["stats"])
print("clk: mean %.0f median %.0f p99 %.0f max %.0f" % (clk.mean(), np.median(clk), np.percentile(clk, 99), clk.max()))
print("work: mean %.1f p99 %.0f max %d" % (work.mean(), np.percentile(work, 99), work.max()))
i = np.unravel_index(np.argmax(clk), clk.shape); print("argmax clk at", i, "work there", work[i], "smid", smid[i])
# per-block (16x8 tiles) max clock
H, W = clk.shape
bl = clk[:H // 8 * 8, :W // 16 * 16].reshape(H // 8, 8, W // 16, 16).max(axis=(1, 3))
print("block clk: mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f" % (bl.mean(), np.median(bl), np.percentile(bl, 90), np.percentile(bl, 99), bl.max()))
rows = bl.mean(axis=1)
print("tile-row mean block clk:", np.array2string(rows[::5], precision=0))
# per-SM total of block clocks
sm = smid[:H // 8 * 8:8, :W // 16 * 16:16]
tot = np.bincount(sm.ravel(), weights=bl.ravel(), minlength=148)
cnt = np.bincount(sm.ravel(), minlength=148)
print("per-SM blocks min/mean/max", cnt.min(), cnt.mean(), cnt.max(), " per-SM sum(block clk) min/mean/max %.0f %.0f %.0f" % (tot.min(), tot.mean(), tot.max()))
np.savez_compressed("gpurun_out/cost_map.npz", clk=clk.astype(np.float32), work=work.astype(np.int32), smid=smid.astype(np.int16))
