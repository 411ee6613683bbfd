"""Generate the fixtures that only a GPU box can produce (run once under gpurun; outputs land in gpurun_out/golden/
and are then committed under tests/golden/):

  xorwow_vectors.json            cuRAND device library: XORWOW start states + first four curand_uniform values of a
                                 list of subsequences, seed 123456 (the stream of optimized.cu:745, 32-37)
  ref_gpu_optimized_512_R_B.raw  the UNMODIFIED reference kernel (oracle/_ref/ref_optimized = optimized.cu included
                                 from where it lies, retargeted to sm_100a, reference flags incl. --use_fast_math):
                                 512x512 frames for `./optimized R B`
  ref_gpu_ieee_512_R_B.raw       the same unmodified source without --use_fast_math and with -fmad=false (ref_optimized_ieee)
  ref_gpu_ids_960x540.*          sigma-0 copy with the first-segment id dump (ref_optimized_ids): rgb, object id, triangle
                                 index, t, shadow flag of every pixel; packed into tests/golden/ref_gpu_ids_960x540.npz by
                                 tests/golden/pack_golden_gpu.py
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import raytracinggpu_b200 as rt

out = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out, exist_ok=True)
subs = [0, 1, 2, 3, 5, 255, 256, 511, 512, 12345, 262143, 262144, 2073599, 8294399, 16777215]
st, u = rt.selftest_xorwow(subs, seed=123456)
json.dump({"seed": 123456, "subsequences": subs, "states_d_v0_v4": st.tolist(), "uniforms_bits": u.view(np.uint32).tolist(),
           "source": "curand_init(seed, subsequence, 0) + 4 x curand_uniform, CUDA 12.9 curand_kernel.h on an NVIDIA B200"},
          open(os.path.join(out, "xorwow_vectors.json"), "w"), indent=1)
cat = os.path.join(ROOT, "oracle", "_ref", "cadnav.com_model", "Models_F0202A090", "cat.obj")
exe = os.path.join(ROOT, "oracle", "_ref", "ref_optimized")
for rays, bounce in ((1, 1), (4, 3)):
    raw = os.path.join(out, "ref_gpu_optimized_512_%d_%d.raw" % (rays, bounce))
    r = subprocess.run([exe, cat, "512", "512", str(rays), str(bounce), "1", raw], capture_output=True, text=True)
    print(r.stdout.strip(), r.stderr.strip()[-200:])
# round 2: the IEEE build of the same unmodified source (the parity pin of the `optimized` profile), and the sigma-0 copy
# with the first-segment id dump (oracle/make_ref_variants.py)
for rays, bounce in ((1, 1), (4, 3)):
    raw = os.path.join(out, "ref_gpu_ieee_512_%d_%d.raw" % (rays, bounce))
    r = subprocess.run([exe + "_ieee", cat, "512", "512", str(rays), str(bounce), "1", raw], capture_output=True, text=True)
    print(r.stdout.strip(), r.stderr.strip()[-200:])
pre = os.path.join(out, "ref_gpu_ids_960x540")
r = subprocess.run([exe + "_ids", cat, "960", "540", "1", "1", "1", pre + ".raw", pre], capture_output=True, text=True)
print(r.stdout.strip(), r.stderr.strip()[-200:])
rng = np.random.default_rng(20261018)
u = (rng.integers(1, 2**32, 4096, dtype=np.uint64).astype(np.float64) * 2.0**-32 + 2.0**-33).astype(np.float32)  # curand_uniform values
ang = (2 * np.pi * u.astype(np.float64)).astype(np.float32)                                                      # (float)(2 pi u)
libm = {"x_log_bits": u.view(np.uint32).tolist(), "x_trig_bits": ang.view(np.uint32).tolist(),
        "log_bits": rt.selftest_libm("log", u).view(np.uint32).tolist(), "sin_bits": rt.selftest_libm("sin", ang).view(np.uint32).tolist(),
        "cos_bits": rt.selftest_libm("cos", ang).view(np.uint32).tolist(),
        "tan_pi_6_bits": int(rt.selftest_libm("tan", np.float32([np.float32(np.float32(np.pi / 3) / np.float32(2))])).view(np.uint32)[0]),
        "source": "logf / sinf / cosf / tanf of CUDA 12.9 (nvcc -fmad=false, no fast math) evaluated on an NVIDIA B200 through rt_selftest_libm"}
json.dump(libm, open(os.path.join(out, "cuda_libm_vectors.json"), "w"))
json.dump({"z_device_960": rt.camera_z_device(960), "z_device_512": rt.camera_z_device(512), "z_device_1920": rt.camera_z_device(1920),
           "z_host_960": rt.camera_z(960), "z_host_512": rt.camera_z(512), "z_host_1920": rt.camera_z(1920),
           "source": "-W / (2 * tanf(alpha / 2)), alpha = (float)(pi/3), evaluated on an NVIDIA B200 (rt_camera_z_device) and by the host libm (rt_camera_z)"},
          open(os.path.join(out, "camera_z.json"), "w"), indent=1)
print("written", os.listdir(out))
