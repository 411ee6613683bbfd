import os, sys
sys.path.insert(0, ".")
os.environ["RT_DEBUG_POOL"] = "1"; os.environ["RT_TASK_FACTOR"] = "1"; os.environ["RT_ANCHOR"] = "1"
import raytracinggpu_b200 as rt
from oracle import profiles, scenes
sc = rt.Scene(0)
scenes.upload(sc, scenes.cat_scene("cpu"))
p = profiles.params("cpu", 640, 360, 1, 0); p.cam[2] = 30.0
o = sc.render(p); print(o["stats"])
