"""Host builder vs device builder of the reference's BVH (rt_mesh_build_bvh / rt_mesh_build_bvh_gpu), wall time and the
device-side build time, for the cat and for the 10 M-triangle instanced cat of BASELINE.json configs[4]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinggpu_b200 as rt
from raytracinggpu_b200 import synthetic
from oracle import pyoracle
cat = pyoracle.cat_obj_path()
def bench(name, make):
    m = make(); t0 = time.perf_counter(); m.build_bvh(); th = time.perf_counter() - t0
    g = make(); g.build_bvh_gpu(0)  # warm-up (context, allocations)
    g = make(); t0 = time.perf_counter(); g.build_bvh_gpu(0); tg = time.perf_counter() - t0
    print("%-28s %9d triangles %8d nodes: host %.1f ms, device builder %.1f ms wall (upload + build + download + record reorder), %.2f ms on the device" % (
        name, m.counts()[1], m.counts()[2], th * 1e3, tg * 1e3, g.build_ms))
bench("cat", lambda: rt.Mesh.read_obj(cat).rescale(0.6, (0.0, -4.0, 0.0)))
scales, offs = synthetic.instance_lattice()
bench("cat x 2529 (configs[4])", lambda: rt.Mesh.read_obj(cat).instance(scales, offs))
