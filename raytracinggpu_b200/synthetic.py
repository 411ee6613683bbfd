"""Synthetic scene data for BASELINE.json configs[4]: the cat mesh instanced to ~10 M triangles.

The reference has no instancing: instances are baked into one TriangleMeshHost with
rt_mesh_instance (v * scale + offset per copy, unfused float), then the reference BVH algorithm runs on the merged
mesh. 2,529 copies x 3,954 triangles = 9,999,666 triangles.
"""
import numpy as np

COPIES_10M = 2529


def lcg(seed):
    """The classic 31-bit LCG (a = 1103515245, c = 12345), uniform in [0, 1)."""
    x = seed & 0x7fffffff
    while True:
        x = (1103515245 * x + 12345) & 0x7fffffff
        yield x / 2147483648.0


def instance_lattice(copies=COPIES_10M, seed=12345, scale=1.0 / 16.0):
    """(scales[copies], offsets[copies, 3]) float32: a 19 x 11 x 13 lattice filling the room in front of the camera
    (x in [-50, 50], y in [-6, 46], z in [-52, 28]; walls at +-60, floor y = -10, camera z = 55), jittered by a
    fixed-seed LCG so that every rank and the oracle derive identical instances."""
    nx, ny, nz = 19, 11, 13
    assert copies <= nx * ny * nz
    g = lcg(seed)
    scales = np.full(copies, scale, np.float32)
    offs = np.empty((copies, 3), np.float32)
    for c in range(copies):
        ix, iy, iz = c % nx, (c // nx) % ny, c // (nx * ny)
        jx, jy, jz = next(g) - 0.5, next(g) - 0.5, next(g) - 0.5
        offs[c, 0] = -50.0 + (ix + 0.5) * (100.0 / nx) + jx
        offs[c, 1] = -6.0 + (iy + 0.5) * (52.0 / ny) + jy
        offs[c, 2] = -52.0 + (iz + 0.5) * (80.0 / nz) + jz
    return scales, offs


def instanced_arrays(vertices, vtx_indices, scales, offsets):
    """numpy restatement of rt_mesh_instance (used by the tests to cross-check the C++ one): copy c is
    v * scales[c] + offsets[c] in unfused float32; triangle indices are shifted by c * nv."""
    v = np.asarray(vertices, np.float32)
    t = np.asarray(vtx_indices, np.int32)
    nv = v.shape[0]
    V = (v[None, :, :] * scales[:, None, None].astype(np.float32) + offsets[:, None, :].astype(np.float32)).astype(np.float32)
    T = t[None, :, :] + (np.arange(len(scales), dtype=np.int32) * nv)[:, None, None]
    return V.reshape(-1, 3), T.reshape(-1, 3).astype(np.int32)
