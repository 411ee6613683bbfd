"""Stochastic mode vs the unmodified reference kernel on the same box: `./optimized R B` frames (sigma 0.2, indirect)."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import raytracinggpu_b200 as rt
from oracle import profiles, scenes, pyoracle
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "oracle", "_ref", "ref_optimized")
cat = pyoracle.cat_obj_path()
sc = scenes.upload(rt.Scene(0), scenes.cat_scene("optimized"))
print("| W x H | rays | bounce | reference optimized.cu ms | this library ms | speed-up | rays/frame | Mrays/s |")
print("|---|---|---|---|---|---|---|---|")
for W, H, rays, bounce in ((512, 512, 1, 1), (512, 512, 4, 3), (512, 512, 16, 5), (1920, 1080, 1, 1), (1920, 1080, 4, 3)):
    p = profiles.params("optimized", W, H, rays, bounce)
    p.aa_sigma, p.indirect = 0.2, 1
    rgb = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(8):
        st = sc.render_into(p, rgb=rgb)
        if i >= 3:
            ms.append(st.kernel_ms)
    ours = float(np.median(ms))
    r = subprocess.run([exe, cat, str(W), str(H), str(rays), str(bounce), "5"], capture_output=True, text=True)
    ref = json.loads(r.stdout.strip().splitlines()[-1])["kernel_ms_median"]
    print("| %dx%d | %d | %d | %.3f | %.3f | %.1fx | %d | %.0f |" % (W, H, rays, bounce, ref, ours, ref / ours, st.rays, st.rays / ours / 1e3))
