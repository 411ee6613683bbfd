import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import raytracinggpu_b200.api as api
api.LIB_PATH = api.LIB_PATH.replace("librtb200.so", "librtb200_trace.so")
import raytracinggpu_b200 as rt
from oracle import scenes
import cases
case = cases.CASES["tiny_mesh_37x23"]
desc = case["scene"](); p = case["params"]()
sc = scenes.upload(rt.Scene(0), desc)
got = sc.render(p)
print(got["stats"])
import numpy as np
ora = scenes.run_oracle(desc, p)
bad = np.argwhere(got["hit_obj"] != ora["hit_obj"])
print("TRACE BUILD mismatches:", len(bad), bad[:5].tolist())
