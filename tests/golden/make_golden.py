"""Regenerates tests/golden/*. Run HERE (the container that has /root/reference):  python tests/golden/make_golden.py

  ref_cpu_1_0_512.png      output of the UNMODIFIED reference binary (oracle/_ref/cpu, built from
                           /root/reference/cpu_launcher.cpp exactly as its Makefile:38) run as `./cpu 1 0`:
                           the one bit-deterministic image an unmodified reference program produces.
  ref_shim_k{0,1}_*.npz    the reference's compiled classes (oracle/_ref/libref_cpu.so) at 192x108 for the
                           cpu_launcher scene (k0) and the optimized.cu object order / mesh transform (k1):
                           8-bit image, primary object id, primary hit point.
  oracle_*.npz             oracle outputs (rgb, hit ids, t bits, shadow flags) for the fixed parity cases of
                           tests/cases.py — what the CUDA path is compared with on the GPU box, where neither
                           /root/reference nor (necessarily) the cat asset exist.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle, scenes  # noqa: E402
import cases  # noqa: E402


def main():
    pyoracle.build(ref=True)
    ref_root = os.environ.get("RT_REFERENCE_DIR", "/root/reference")
    # 1. the unmodified binary
    with tempfile.TemporaryDirectory() as d:
        os.symlink(os.path.join(ref_root, "cadnav.com_model"), os.path.join(d, "cadnav.com_model"))
        subprocess.run([pyoracle.REF_CPU_BIN, "1", "0"], cwd=d, check=True, stdout=subprocess.DEVNULL)
        shutil.copy(os.path.join(d, "image.png"), os.path.join(HERE, "ref_cpu_1_0_512.png"))
    # 2. the reference's classes at another resolution, both scene layouts
    cat = pyoracle.cat_obj_path()
    for kind in (0, 1):
        r = pyoracle.ref_cpu_render(cat, kind, 192, 108, 1, 0)
        np.savez_compressed(os.path.join(HERE, "ref_shim_k%d_192x108.npz" % kind), rgb=r["rgb"], hit_obj=r["hit_obj"].astype(np.int8), P=r["P"])
    # 3. oracle outputs for the fixed cases
    for name, case in cases.CASES.items():
        desc = case["scene"]()
        if desc is None:
            print("skip", name)
            continue
        o = scenes.run_oracle(desc, case["params"]())
        np.savez_compressed(os.path.join(HERE, "oracle_%s.npz" % name), rgb=o["rgb"], hit_obj=o["hit_obj"].astype(np.int8),
                            hit_tri=o["hit_tri"], hit_t=o["hit_t"], shadow=o["shadow"], rays=np.uint64(o["work"]["rays"]))
        print(name, o["work"])


if __name__ == "__main__":
    main()
