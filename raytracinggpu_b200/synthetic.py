"""Synthetic scene data for BASELINE.json configs[4]: the cat mesh instanced to ~10 M triangles.

The reference has no instancing: instances are baked into one TriangleMeshHost with
rt_mesh_instance (v * scale + offset per copy, unfused float), then the reference BVH algorithm runs on the merged
mesh. 2,529 copies x 3,954 triangles = 9,999,666 triangles.
"""
import math

import numpy as np

COPIES_10M = 2529


def torus(nu=48, nv=24, R=9.0, r=3.5, center=(0.0, 0.0, 0.0), tilt=0.6, wobble=0.35, seed=7):
    """A lumpy tilted torus (nu*nv*2 triangles): a closed mesh with silhouettes, concavity and grazing hits — the stand-in
    mesh of bench.py when the (non-redistributable) cat asset is absent. Returns (vertices [nu*nv, 3] f32, triangles [.., 3] i32)."""
    rng = np.random.RandomState(seed)
    us = np.linspace(0, 2 * math.pi, nu, endpoint=False)
    vs = np.linspace(0, 2 * math.pi, nv, endpoint=False)
    bump = 1.0 + wobble * 0.5 * (np.sin(3 * us)[:, None] * np.cos(2 * vs)[None, :]) + 0.05 * rng.rand(nu, nv)
    x = (R + r * bump * np.cos(vs)[None, :]) * np.cos(us)[:, None]
    y = r * bump * np.sin(vs)[None, :] * np.ones((nu, 1))
    z = (R + r * bump * np.cos(vs)[None, :]) * np.sin(us)[:, None]
    c, s = math.cos(tilt), math.sin(tilt)
    y2, z2 = c * y - s * z, s * y + c * z
    verts = np.stack([x + center[0], y2 + center[1], z2 + center[2]], axis=-1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a, b = i * nv + j, ((i + 1) % nu) * nv + j
    c2, d = ((i + 1) % nu) * nv + (j + 1) % nv, i * nv + (j + 1) % nv
    idx = np.stack([np.stack([a, b, c2], -1), np.stack([a, c2, d], -1)], axis=2).reshape(-1, 3)
    return verts, idx.astype(np.int32)


def spheres_scene_spheres(rt):
    """BASELINE.json configs[0] / configs[3]: the six walls of cpu_launcher.cpp:673-678 (ids 0-5) + the demo spheres of its commented
    lines :668-672 — white diffuse (0,0,0) R 10, mirror (-20,0,0) R 10, refractive shell: inner (20,0,0) R 9 (n 1 -> 1.5) inside
    outer R 10 (n 1.5 -> 1) — as a list of rt_sphere (ids 6-9). No mesh."""
    walls, _ = rt.default_walls("cpu")
    out = list(walls)

    def add(C, R, albedo, mirror=0, n_in=1.0, n_out=1.0):
        s = rt.rt_sphere()
        s.C[:] = [float(x) for x in C]
        s.R = float(R)
        s.albedo[:] = [float(x) for x in albedo]
        s.mirror, s.n_in, s.n_out, s.id = int(mirror), float(n_in), float(n_out), len(out)
        out.append(s)
    add((0, 0, 0), 10, (1., 1., 1.))
    add((-20, 0, 0), 10, (0., 0., 0.), mirror=1)
    add((20, 0, 0), 9, (0., 0., 0.), n_in=1.0, n_out=1.5)
    add((20, 0, 0), 10, (0., 0., 0.), n_in=1.5, n_out=1.0)
    return out


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
