"""Small renders of every kernel family for compute-sanitizer (memcheck / racecheck): wavefront (2 strips, multi-bounce,
forced pool spill), render_mega variants, stochastic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinggpu_b200 as rt
from oracle import profiles, scenes
sc = rt.Scene(0)
d = scenes.cat_scene("optimized", mirror=1) or scenes.torus_scene("optimized", mirror=1)
scenes.upload(sc, d)
p = profiles.params("optimized", 320, 192, 1, 3)
a = sc.render(p)
os.environ["RT_NPOOL_CAP"] = "64"
b = sc.render(p)
del os.environ["RT_NPOOL_CAP"]
assert (a["rgb"] == b["rgb"]).all()
q = profiles.params("optimized", 160, 96, 2, 3)
q.aa_sigma, q.indirect = 0.2, 1
sc.render(q)
sc.render(profiles.params("optimized", 160, 96, 1, 2), count_work=True)
# the tree search for everything (RT_ANCHOR=0), the wide index, a moving light (bins rebuilt), odd sizes (zero direction components)
os.environ["RT_ANCHOR"] = "0"
c = sc.render(p)
assert (a["rgb"] == c["rgb"]).all()
os.environ["RT_WIDE"] = "1"
c = sc.render(p)
assert (a["rgb"] == c["rgb"]).all()
del os.environ["RT_ANCHOR"], os.environ["RT_WIDE"]
L = (-10.0, 20.0, 40.0)
for k in range(3):
    L = rt.move_light(L, 1.309, 0.3)
    sc.set_light(L, 3e10)
    sc.render(profiles.params("optimized", 161, 97, 1, 2))
sc.close()
# the device BVH builder
v, t = scenes.torus(24, 12)
m1 = rt.Mesh.from_arrays(v, t).build_bvh()
m2 = rt.Mesh.from_arrays(v, t).build_bvh_gpu(0)
assert (m1.arr_bvh == m2.arr_bvh).all() and (m1.tri_records == m2.tri_records).all()
os.environ["RT_VARIANT"] = "1"
sc = rt.Scene(0)
scenes.upload(sc, d)
sc.render(profiles.params("optimized", 160, 96, 1, 2))
print("sanitize_small ok")
