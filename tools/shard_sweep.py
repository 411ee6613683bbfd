"""One rank's share of configs[2] (4K mirror cat, depth 4, rows r, r+N, ...): frame time for N = 1, 2, 4, 8 with the two-child
records and with the wide index (RT_WIDE=1) behind the bins, and with 1 or 2 row bands."""
import os, subprocess, sys
code = r'''
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import raytracinggpu_b200 as rt
from oracle import profiles, scenes
sc = scenes.upload(rt.Scene(0), scenes.cat_scene("optimized", mirror=1))
out = []
for world in (1, 2, 4, 8):
    p = profiles.params("optimized", 3840, 2160, 1, 4)
    p.row_begin, p.row_step, p.row_count = rt.sharding.rows_for_rank(2160, 0, world)
    rgb = torch.empty((p.row_count, 3840, 3), dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(14):
        st = sc.render_into(p, rgb=rgb)
        if i >= 4: ms.append(st.kernel_ms)
    out.append("N=%d %.3f" % (world, float(np.median(ms))))
print("  ".join(out))
'''
for env in (dict(), dict(RT_WIDE="1"), dict(RT_STRIPS="1"), dict(RT_WIDE="1", RT_STRIPS="1"), dict(RT_ANCHOR="0")):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    print(env, r.stdout.strip() or r.stderr[-300:])
