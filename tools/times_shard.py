"""Per-kernel times (RT_DEBUG_TIMES) of one rank's share of configs[2]: 4K mirror cat depth 4, rows r, r+8, ..."""
import os, sys
os.environ["RT_DEBUG_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raytracinggpu_b200 as rt
from oracle import profiles, scenes
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
desc = scenes.cat_scene("optimized", mirror=1)
sc = scenes.upload(rt.Scene(0), desc)
p = profiles.params("optimized", 3840, 2160, 1, 4)
p.row_begin, p.row_step, p.row_count = rt.sharding.rows_for_rank(2160, 0, world)
rgb = torch.empty((p.row_count, 3840, 3), dtype=torch.uint8, device="cuda")
for i in range(5):
    st = sc.render_into(p, rgb=rgb)
print("world", world, "kernel_ms", st.kernel_ms, "rays", st.rays, "launches", st.launches)
