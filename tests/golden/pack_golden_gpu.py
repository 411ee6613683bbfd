"""Pack what tests/golden/make_golden_gpu.py wrote on the GPU box (gpurun_out/golden/) into the committed fixtures:
PNG for frames, one compressed npz for the first-segment dump of the reference kernel (t kept for mesh pixels only, the
full t plane by its SHA-256)."""
import hashlib, os, sys
import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = os.path.join(ROOT, "gpurun_out", "golden")
dst = os.path.dirname(os.path.abspath(__file__))
for name in ("ref_gpu_ieee_512_1_1", "ref_gpu_ieee_512_4_3"):
    a = np.fromfile(os.path.join(src, name + ".raw"), np.uint8).reshape(512, 512, 3)
    Image.fromarray(a).save(os.path.join(dst, name + ".png"), optimize=True)
W, H = 960, 540
pre = os.path.join(src, "ref_gpu_ids_960x540")
rgb = np.fromfile(pre + ".raw", np.uint8).reshape(H, W, 3)
obj = np.fromfile(pre + ".obj.i32", np.int32).reshape(H, W)
tri = np.fromfile(pre + ".tri.i32", np.int32).reshape(H, W)
t = np.fromfile(pre + ".t.f32", np.float32).reshape(H, W)
shadow = np.fromfile(pre + ".shadow.u8", np.uint8).reshape(H, W)
mesh = obj == 1
Image.fromarray(rgb).save(os.path.join(dst, "ref_gpu_sigma0_960x540.png"), optimize=True)
np.savez_compressed(os.path.join(dst, "ref_gpu_ids_960x540.npz"), obj=obj.astype(np.int8), tri_mesh=tri[mesh], t_mesh_bits=t[mesh].view(np.uint32), shadow=shadow,
                    t_sha256=np.frombuffer(hashlib.sha256(t.tobytes()).digest(), np.uint8),
                    source="optimized.cu sigma-0 copy + id dump (oracle/make_ref_variants.py), IEEE flags, NVIDIA B200, 960x540 `1 1`")
print("mesh pixels", int(mesh.sum()), "files:", [f for f in os.listdir(dst) if f.startswith("ref_gpu_")])
