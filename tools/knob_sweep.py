"""Sweep of the launch-shape knobs (blocks per SM of generate / leaves / shade, number of row bands) on configs[1]."""
import os, subprocess, sys
combos = [dict(), dict(RT_GEN_BLOCKS="4", RT_LEAVES_BLOCKS="5", RT_SHADE_BLOCKS="4"), dict(RT_GEN_BLOCKS="5", RT_LEAVES_BLOCKS="4", RT_SHADE_BLOCKS="4"),
          dict(RT_GEN_BLOCKS="4", RT_LEAVES_BLOCKS="5", RT_SHADE_BLOCKS="4", RT_STRIPS="4"), dict(RT_GEN_BLOCKS="6", RT_LEAVES_BLOCKS="3", RT_SHADE_BLOCKS="3", RT_STRIPS="4"),
          dict(RT_GEN_BLOCKS="3", RT_LEAVES_BLOCKS="6", RT_SHADE_BLOCKS="3", RT_STRIPS="4"), dict(RT_STRIPS="4"), dict(RT_GEN_BLOCKS="9"), dict(RT_GEN_BLOCKS="5", RT_LEAVES_BLOCKS="4", RT_SHADE_BLOCKS="4", RT_STRIPS="3")]
for c in combos:
    env = dict(os.environ); env.update(c)
    out = subprocess.run([sys.executable, "tools/perf_matrix.py", "det"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
    print(c, out[0] if out else "?")
