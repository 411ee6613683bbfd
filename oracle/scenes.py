"""Scene builders shared by tests/, smoke() and bench.py. TEST/BENCH INFRASTRUCTURE ONLY.

Every builder returns a plain dict description (spheres, mesh arrays in the reference interchange formats,
mesh material, light) that can be handed both to the oracle (pyoracle.render) and to the product
(raytracinggpu_b200.Scene), so both sides always see byte-identical inputs. Mesh arrays come from the
ORACLE's loader/builder here; tests/test_host_mesh.py separately proves the product's loader/builder
produce the same arrays.
"""
import math

import numpy as np

from . import profiles, pyoracle


def torus(nu=48, nv=24, R=9.0, r=3.5, center=(0.0, 0.0, 0.0), tilt=0.6, wobble=0.35, seed=7):
    """A lumpy tilted torus (nu*nv*2 triangles): a closed mesh with silhouettes, concavity and grazing hits."""
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
    idx = []
    for i in range(nu):
        for j in range(nv):
            a = i * nv + j
            b = ((i + 1) % nu) * nv + j
            c2 = ((i + 1) % nu) * nv + (j + 1) % nv
            d = i * nv + (j + 1) % nv
            idx.append((a, b, c2))
            idx.append((a, c2, d))
    return verts, np.asarray(idx, dtype=np.int32)


def grid_quads(n=6, size=20.0, y=-5.0):
    """Axis-aligned flat grid: every leaf box is degenerate (zero thickness), the case the reference's strict
    slab test always rejects."""
    xs = np.linspace(-size / 2, size / 2, n + 1)
    verts = np.array([(x, y, z) for x in xs for z in xs], dtype=np.float32)
    idx = []
    for i in range(n):
        for j in range(n):
            a = i * (n + 1) + j
            idx.append((a, a + 1, a + n + 2))
            idx.append((a, a + n + 2, a + n + 1))
    return verts, np.asarray(idx, dtype=np.int32)


def _scene(spheres, mesh, mesh_mat, light=profiles.LIGHT):
    arrays = None
    if mesh is not None:
        arrays = (mesh.vertices, mesh.tri_records, mesh.arr_bvh)
    return dict(spheres=spheres, mesh=arrays, mesh_mat=mesh_mat if mesh is not None else None, light=light)


def cat_scene(profile, mirror=0, obj_path=None):
    """Walls + cat of the given reference program (None when the cat asset is unavailable)."""
    m = profiles.cat_mesh(profile, obj_path)
    if m is None:
        return None
    return _scene(profiles.walls(profile), m, profiles.mesh_material(profile, mirror))


def torus_scene(profile, mirror=0, nu=48, nv=24, **kw):
    """Walls + synthetic torus with the object ids of the profile (used when the cat is unavailable and for
    extra geometry coverage)."""
    v, t = torus(nu, nv, **kw)
    m = pyoracle.Mesh.from_arrays(v, t).build_bvh()
    return _scene(profiles.walls(profile), m, profiles.mesh_material(profile, mirror))


def mesh_scene(profile, verts, idx, mirror=0, n_in=1.0, n_out=1.0):
    m = pyoracle.Mesh.from_arrays(verts, idx).build_bvh()
    mm = profiles.mesh_material(profile, mirror)
    mm["n_in"], mm["n_out"] = n_in, n_out
    return _scene(profiles.walls(profile), m, mm)


def instanced_cat_scene(profile, copies, seed=12345, obj_path=None):
    """BASELINE.json configs[4]: `copies` baked instances of the loader-transformed cat on
    the lattice of raytracinggpu_b200.synthetic.instance_lattice, merged into one mesh, reference BVH on top. Built
    here with numpy + the oracle's own host builder (independent of the product's rt_mesh_instance / build)."""
    from raytracinggpu_b200 import synthetic
    path = obj_path or pyoracle.cat_obj_path()
    if path is None:
        return None
    base = pyoracle.Mesh.from_obj(path)
    scales, offs = synthetic.instance_lattice(copies, seed)
    v, t = synthetic.instanced_arrays(base.vertices, base.tri_records[:, :3], scales, offs)
    m = pyoracle.Mesh.from_arrays(v, t).build_bvh()
    return _scene(profiles.walls(profile), m, profiles.mesh_material(profile, 0))


def spheres_scene(light=profiles.LIGHT):
    """BASELINE.json config 1/4: the six walls + the demo spheres of the commented lines cpu_launcher.cpp:668-672
    (white diffuse, mirror, refractive shell = inner R 9 (n 1 -> 1.5) inside outer R 10 (n 1.5 -> 1)). No mesh."""
    sp = profiles.walls("cpu", with_mesh=False)
    nid = len(sp)
    sp.append(profiles.sphere((0, 0, 0), 10, (1., 1., 1.), nid)); nid += 1
    sp.append(profiles.sphere((-20, 0, 0), 10, (0., 0., 0.), nid, mirror=1)); nid += 1
    sp.append(profiles.sphere((20, 0, 0), 9, (0., 0., 0.), nid, n_in=1.0, n_out=1.5)); nid += 1
    sp.append(profiles.sphere((20, 0, 0), 10, (0., 0., 0.), nid, n_in=1.5, n_out=1.0)); nid += 1
    return _scene(sp, None, None, light)


def run_oracle(scene, params, threads=0, want=("rgb", "hit_obj", "hit_tri", "hit_t", "shadow")):
    return pyoracle.render(scene["spheres"], scene["mesh"], scene["mesh_mat"], scene["light"], params, threads=threads, want=want,
                           normals=scene.get("normals"))


def obj_normals(path):
    """The viewer loader's extra (realtime_render.cu:489-493, 538-545), restated independently in Python: the `vn` lines of an OBJ
    file as (nn, 3) float32 and the normal indices of its faces, fan-triangulated in file order, as (nt, 3) int32 (-1: none)."""
    normals, idx = [], []
    for line in open(path, "r", errors="replace"):
        if line.startswith("vn "):
            normals.append([np.float32(x) for x in line.split()[1:4]])
        elif line.startswith("f"):
            toks = line[1:].split()
            ni = []
            for t in toks:
                parts = t.split("/")
                if not parts[0].lstrip("-").isdigit():
                    break
                k = int(parts[2]) if len(parts) >= 3 and parts[2] else 0
                ni.append(k - 1 if k > 0 else (len(normals) + k if k < 0 else -1))
            for k in range(2, len(ni)):
                tri = [ni[0], ni[k - 1], ni[k]]
                idx.append(tri if min(tri) >= 0 else [-1, -1, -1])
    return np.asarray(normals, np.float32).reshape(-1, 3), np.asarray(idx, np.int32).reshape(-1, 3)


def viewer_cat_scene(obj_path=None):
    """The viewer's scene (realtime_render.cu:1018-1050): walls 0-5 with the R = 940 floor, cat = object 6 with vertex normals, mesh
    transform 0.6 / (0, -10, 0) after the loader's (:1309), light (0, 15, 40). None without the cat asset."""
    path = obj_path or pyoracle.cat_obj_path()
    if path is None:
        return None
    m = pyoracle.Mesh.from_obj(path, rescale=(0.6, (0., -10., 0.)))
    m.set_normals(*obj_normals(path))
    m.build_bvh()
    sp = profiles.walls("cpu")
    sp[1].R = 940.0
    d = _scene(sp, m, profiles.mesh_material("cpu", 0), ((0., 15., 40.), 3e10))
    d["normals"] = m.normals
    return d


def upload(rt_scene, scene):
    """Put a scene description on the device through the product's C ABI."""
    rt_scene.set_spheres(scene["spheres"])
    rt_scene.set_light(*scene["light"])
    if scene["mesh"] is not None:
        mm = scene["mesh_mat"]
        rt_scene.set_mesh(*scene["mesh"], albedo=mm["albedo"], mirror=mm["mirror"], n_in=mm["n_in"], n_out=mm["n_out"], id=mm["id"])
        if scene.get("normals") is not None:
            rt_scene.set_mesh_normals(scene["normals"])
    else:
        rt_scene.clear_mesh()
    return rt_scene


def compare(gpu, ora, lsb_frac=0.999):
    """The parity bar of BASELINE.json: hit ids / t bit-exact, 8-bit colour within 1 LSB on >= 99.9 % of pixels.
    Returns a dict of mismatch counts; raises AssertionError with a summary when the bar is missed."""
    res = {}
    for k in ("hit_obj", "hit_tri", "shadow"):
        if k in gpu and k in ora:
            res[k] = int((gpu[k] != ora[k]).sum())
    if "hit_t" in gpu and "hit_t" in ora:
        res["hit_t"] = int((gpu["hit_t"].view(np.uint32) != ora["hit_t"].view(np.uint32)).sum())
    if "rgb" in gpu and "rgb" in ora:
        d = np.abs(gpu["rgb"].astype(np.int32) - ora["rgb"].astype(np.int32)).max(axis=-1)
        res["rgb_exact_mismatch"] = int((d > 0).sum())
        res["rgb_gt1_lsb"] = int((d > 1).sum())
        res["rgb_max_diff"] = int(d.max()) if d.size else 0
        res["pixels"] = int(d.size)
    bad = [k for k in ("hit_obj", "hit_tri", "hit_t", "shadow") if res.get(k, 0) != 0]
    if "pixels" in res and res["pixels"]:
        if res["rgb_gt1_lsb"] > (1 - lsb_frac) * res["pixels"]:
            bad.append("rgb")
    assert not bad, "parity failed on %s: %r" % (bad, res)
    return res
