import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import raytracinggpu_b200 as rt
from oracle import scenes
import cases
case = cases.CASES["tiny_mesh_37x23"]
desc = case["scene"](); p = case["params"]()
ora = scenes.run_oracle(desc, p)
for v in (1, 2):
    os.environ["RT_VARIANT"] = str(v)
    sc = scenes.upload(rt.Scene(0), desc)
    got = sc.render(p)
    bad = np.argwhere(got["hit_obj"] != ora["hit_obj"])
    print("variant", v, "mismatch", len(bad), "stats", got["stats"]["rays"], ora["work"]["rays"])
    for (y, x) in bad[:6]:
        print("  px", y, x, "gpu obj/tri/t", got["hit_obj"][y, x], got["hit_tri"][y, x], got["hit_t"][y, x], " oracle", ora["hit_obj"][y, x], ora["hit_tri"][y, x], ora["hit_t"][y, x])
    sc.close()
print("---- work counters")
for v in (1, 2):
    os.environ["RT_VARIANT"] = str(v)
    sc = scenes.upload(rt.Scene(0), desc)
    got = sc.render(p, count_work=True)
    print("variant", v, got["stats"])
    sc.close()
print("oracle", ora["work"])
