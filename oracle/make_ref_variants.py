#!/usr/bin/env python
"""make_ref_variants.py — TEST INFRASTRUCTURE ONLY. Writes patched COPIES of the reference's optimized.cu into the
git-ignored oracle/_ref/ (never committed; SURVEY.md 8c allows a patched copy as a cross-check):

  oracle/_ref/optimized_sigma0.cu   `float sigma = 0.2;` (optimized.cu:753) -> 0.0: the deterministic image of the
                                    reference's own GPU kernel (the jitter terms become exactly 0)
  oracle/_ref/optimized_ids.cu      sigma 0 + a dump of what the kernel decides for the FIRST segment of every pixel:
                                    object id, winning triangle index (post-build order), t, shadow flag — the
                                    quantities the reference computes but never emits

Every edit is an exact-string substitution that must match exactly once, so a changed reference fails loudly instead of
silently producing an unpatched copy. The arithmetic is untouched: the dump only stores values the kernel already holds.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def sub_once(src, old, new, what):
    n = src.count(old)
    if n != 1:
        raise SystemExit("make_ref_variants: anchor for %s matches %d times (expected 1)" % (what, n))
    return src.replace(old, new)


GLOBALS = """
/* ---- injected by oracle/make_ref_variants.py: first-segment dump (test infrastructure) ---- */
__device__ int* rtb_dump_obj;            /* per pixel, initialised to -2 by the shim */
__device__ int* rtb_dump_tri;
__device__ float* rtb_dump_t;
__device__ unsigned char* rtb_dump_shadow; /* initialised to 2 */
__device__ int* rtb_last_tri;            /* per thread: winner of the last TriangleMesh::intersect */
__device__ float* rtb_last_t;            /* per thread: t_min of the last intersect_all */
#define RTB_GID ((size_t)blockIdx.x * blockDim.x + threadIdx.x)
"""


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, "_ref")
    src = open(os.path.join(ref, "optimized.cu")).read()
    os.makedirs(out_dir, exist_ok=True)

    sigma0 = sub_once(src, "float sigma = 0.2;", "float sigma = 0.0;", "sigma (optimized.cu:753)")
    open(os.path.join(out_dir, "optimized_sigma0.cu"), "w").write(sigma0)

    ids = sub_once(sigma0, "#define MAX_RAY_DEPTH 10\n", "#define MAX_RAY_DEPTH 10\n" + GLOBALS, "globals")
    # TriangleMesh::intersect (optimized.cu:220-285): remember the triangle that set t_min
    ids = sub_once(ids, "\t\tfloat t_min = INF;\n\t\twhile (s_size) {", "\t\tfloat t_min = INF;\n\t\trtb_last_tri[RTB_GID] = -1;\n\t\twhile (s_size) {", "mesh query start")
    ids = sub_once(ids, "\t\t\t\t\t\tt_min = t_cur;\n\t\t\t\t\t\tN = N_triangle;\n", "\t\t\t\t\t\tt_min = t_cur;\n\t\t\t\t\t\tN = N_triangle;\n\t\t\t\t\t\trtb_last_tri[RTB_GID] = i;\n",
                   "triangle accept (optimized.cu:275-278)")
    # Scene::intersect_all (optimized.cu:539-559): remember t_min
    ids = sub_once(ids, "\t\tP = r.O + t_min * r.u;\n\t\tobjectId = id_min;", "\t\trtb_last_t[RTB_GID] = t_min;\n\t\tP = r.O + t_min * r.u;\n\t\tobjectId = id_min;", "intersect_all result")
    # getColorIterative (optimized.cu:561-661): first segment of the first sample
    ids = sub_once(ids, "\t\t\tbool inter = intersect_all(ray, P, N, sphere_id);\n",
                   "\t\t\tbool inter = intersect_all(ray, P, N, sphere_id);\n"
                   "\t\t\tconst bool rtb_first = ray_depth == 0 && rtb_dump_obj[RTB_GID] == -2;\n"
                   "\t\t\tif (rtb_first) { rtb_dump_obj[RTB_GID] = sphere_id; rtb_dump_t[RTB_GID] = rtb_last_t[RTB_GID];\n"
                   "\t\t\t\trtb_dump_tri[RTB_GID] = (sphere_id == 1 /* the mesh, optimized.cu:690-699 */) ? rtb_last_tri[RTB_GID] : -1; }\n",
                   "first intersect_all of the path")
    ids = sub_once(ids, "\t\t\t\t\t\t// Is shadow\n", "\t\t\t\t\t\t// Is shadow\n\t\t\t\t\t\tif (rtb_first) rtb_dump_shadow[RTB_GID] = 1;\n", "shadow branch")
    ids = sub_once(ids, "\t\t\t\t\t\t// Get direct color\n", "\t\t\t\t\t\t// Get direct color\n\t\t\t\t\t\tif (rtb_first) rtb_dump_shadow[RTB_GID] = 0;\n", "lit branch")
    open(os.path.join(out_dir, "optimized_ids.cu"), "w").write(ids)
    print("wrote", os.path.join(out_dir, "optimized_sigma0.cu"), "and optimized_ids.cu")


if __name__ == "__main__":
    main()
