"""Output path: PNG encode of a rendered-like 4K frame with 1 thread (what stbi_write_png amounts to) vs
parallel bands, and the asynchronous writer's submit() cost. No GPU needed."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracinggpu_b200 as rt
H, W = 2160, 3840
rng = np.random.RandomState(1)
yy, xx = np.mgrid[0:H, 0:W]
img = np.stack([(xx // 7 + yy // 5) % 256, (xx * yy // 4096) % 256, rng.randint(0, 40, size=(H, W))], -1).astype(np.uint8)  # smooth + a noisy channel
d = tempfile.mkdtemp()
for th in (1, 4, 16, os.cpu_count()):
    os.environ["RT_PNG_THREADS"] = str(th)
    t0 = time.perf_counter(); rt.write_png(os.path.join(d, "a.png"), img); t = time.perf_counter() - t0
    print("rt_write_png 3840x2160, %2d threads: %.1f ms, %d KB" % (th, t * 1e3, os.path.getsize(os.path.join(d, "a.png")) // 1024))
w = rt.PngWriter()
t0 = time.perf_counter()
for k in range(8):
    w.submit(os.path.join(d, "f%d.png" % k), img)
t1 = time.perf_counter(); w.wait(); t2 = time.perf_counter()
print("async writer: 8 frames submitted in %.1f ms (%.1f ms each: the copy), all on disk after %.1f ms" % ((t1 - t0) * 1e3, (t1 - t0) * 1e3 / 8, (t2 - t0) * 1e3))
