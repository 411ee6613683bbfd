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
sc.close()
os.environ["RT_VARIANT"] = "1"
sc = rt.Scene(0)
scenes.upload(sc, d)
sc.render(profiles.params("optimized", 160, 96, 1, 2))
print("sanitize_small ok")
