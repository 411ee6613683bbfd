"""Render the same frame N times and report any pixel that differs from the first render (races show up here)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracinggpu_b200 as rt
from oracle import profiles, scenes

prof = sys.argv[1] if len(sys.argv) > 1 else "cpu"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bounce = int(sys.argv[3]) if len(sys.argv) > 3 else (0 if prof == "cpu" else 1)
desc = scenes.cat_scene(prof)
p = profiles.params(prof, 1920, 1080, 1, bounce)
sc = scenes.upload(rt.Scene(0), desc)
ref = sc.render(p)
bad = 0
for it in range(n):
    if it % 2 and prof == "cpu":
        p.push_order = 1 - p.push_order
        ref2 = None
    o = sc.render(p)
    if prof == "cpu" and it % 2:
        p.push_order = 1 - p.push_order
        continue
    for k in ("rgb", "hit_obj", "hit_tri", "shadow"):
        d = np.argwhere(o[k] != ref[k])
        if len(d):
            bad += 1
            print("iter", it, k, "differs at", len(d), "places; first", d[0], "got", o[k][tuple(d[0])], "ref", ref[k][tuple(d[0])])
    d = np.argwhere(o["hit_t"].view(np.uint32) != ref["hit_t"].view(np.uint32))
    if len(d):
        bad += 1
        print("iter", it, "hit_t differs at", len(d), "first", d[0], o["hit_t"][tuple(d[0])], ref["hit_t"][tuple(d[0])])
print("done", n, "iterations,", bad, "differences; stats", o["stats"])
