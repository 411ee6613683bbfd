"""Host scene surface of the product (librtb200.so host code, no GPU needed): loader, BVH builder, flattener,
profiles, PNG writer, error behaviour — against the oracle and the reference's own loader/builder."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from PIL import Image

import raytracinggpu_b200 as rt
from oracle import profiles, pyoracle, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(built):
    """The C-ABI library loads and exports every function include/rt_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    L = C.CDLL(rt.LIB_PATH)
    for n in sorted(names):
        assert hasattr(L, n), "missing export %s" % n
    assert set(rt.api.SIGNATURES) == names
    assert L.rt_abi_version() == 3


def test_struct_layouts_match_header(built, tmp_path):
    """The header is valid plain C, and the ctypes mirrors have the sizes a C compiler gives the structs."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "rt_b200.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(rt_sphere), sizeof(rt_params), sizeof(rt_stats));return 0;}\n')
    exe = str(tmp_path / "sz")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe], check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(rt.rt_sphere), C.sizeof(rt.rt_params), C.sizeof(rt.rt_stats)] == [44, 128, 56]


@pytest.mark.parametrize("profile", ["cpu", "optimized", "array_bvh"])
def test_loader_and_builder_match_oracle(cat_path, profile):
    m = rt.Mesh.read_obj(cat_path)
    k = profiles.PROFILES[profile]["rescale"]
    if k:
        m.rescale(*k)
    m.build_bvh()
    o = profiles.cat_mesh(profile, cat_path)
    assert np.array_equal(m.vertices.view(np.uint32), o.vertices.view(np.uint32))
    assert np.array_equal(m.tri_records, o.tri_records)
    assert np.array_equal(m.arr_bvh.view(np.uint32), o.arr_bvh.view(np.uint32))
    assert m.bvh_info() == o.bvh_info()
    assert m.counts() == (2247, 3954, 2019)


@pytest.mark.skipif(not pyoracle.ref_cpu_available() or not os.path.exists("/root/reference"), reason="needs the live compiled reference")
def test_loader_and_builder_match_live_reference(cat_path):
    rm = pyoracle.ref_cpu_mesh(cat_path, 1)
    m = rt.Mesh.read_obj(cat_path).rescale(0.6, (0, -4, 0)).build_bvh()
    assert np.array_equal(rm["vertices"].view(np.uint32), m.vertices.view(np.uint32))
    assert np.array_equal(rm["vtx_indices"], m.tri_records[:, :3])
    assert np.array_equal(rm["arr_bvh"].view(np.uint32), m.arr_bvh.view(np.uint32))


def test_builder_on_synthetic_meshes_matches_oracle(built):
    for (v, t) in (scenes.torus(), scenes.torus(96, 40), scenes.grid_quads(9)):
        m = rt.Mesh.from_arrays(v, t).build_bvh()
        o = pyoracle.Mesh.from_arrays(v, t).build_bvh()
        assert np.array_equal(m.tri_records, o.tri_records)
        assert np.array_equal(m.arr_bvh.view(np.uint32), o.arr_bvh.view(np.uint32))
        assert m.bvh_info() == o.bvh_info()


OBJ_FORMS = """# forms of optimized.cu:366-447
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0.5 0.5 1 0.2 0.3 0.4
vt 0 0
vn 0 0 1
f 1 2 3
f 1/1 2/1 3/1
f 1//1 2//1 3//1
f 1/1/1 2/1/1 3/1/1 4/1/1
f -5 -4 -3 -2
f 1 2 3 4 5
"""


def test_obj_forms_match_oracle_loader(built, tmp_path):
    """All index forms, a polygon fan, negative (relative) indices and a coloured 6-field vertex (not transformed)."""
    p = tmp_path / "forms.obj"
    p.write_text(OBJ_FORMS.replace("\n", "\r\n"))
    m = rt.Mesh.read_obj(str(p))
    o = pyoracle.Mesh.from_obj(str(p))
    assert m.counts()[:2] == o.counts()[:2] == (5, 1 + 1 + 1 + 2 + 2 + 3)
    assert np.array_equal(m.vertices.view(np.uint32), o.vertices.view(np.uint32))
    assert np.array_equal(m.tri_records, o.tri_records)
    assert np.allclose(m.vertices[1], (0.8, -10, 0)) and np.allclose(m.vertices[4], (0.5, 0.5, 1))


def test_missing_obj_is_an_error_not_an_exit(built):
    with pytest.raises(rt.RtError) as e:
        rt.Mesh.read_obj("/nonexistent/cat.obj")
    assert e.value.code == -4


def test_empty_and_tiny_meshes(built):
    m = rt.Mesh.from_arrays(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)).build_bvh()
    assert m.counts() == (0, 0, 1)
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    m = rt.Mesh.from_arrays(v, np.array([[0, 1, 2]], np.int32)).build_bvh()
    assert m.counts() == (3, 1, 1) and m.bvh_info() == {"leaves": 1, "max_depth": 1, "max_leaf": 1}
    with pytest.raises(rt.RtError):
        rt.Mesh.from_arrays(v, np.array([[0, 1, 3]], np.int32))


def test_instancing_bakes_copies(built):
    v, t = scenes.torus(12, 6)
    m = rt.Mesh.from_arrays(v, t).instance([1.0, 0.5, 0.25], [(0, 0, 0), (10, 0, 0), (0, 5, 0)])
    nv, nt, _ = m.counts()
    assert (nv, nt) == (3 * len(v), 3 * len(t))
    assert np.array_equal(m.vertices[len(v):2 * len(v)], (v * np.float32(0.5) + np.array([10, 0, 0], np.float32)).astype(np.float32))
    assert np.array_equal(m.tri_records[2 * len(t):, :3], t + 2 * len(v))


def test_profiles_and_walls_match_independent_statement(built):
    for prof in ("cpu", "optimized", "array_bvh"):
        assert bytes(rt.params_profile(prof, 1920, 1080, 3, 4)) == bytes(profiles.params(prof, 1920, 1080, 3, 4))
        w, mid = rt.default_walls(prof)
        assert mid == profiles.PROFILES[prof]["mesh_id"]
        assert all(bytes(a) == bytes(b) for a, b in zip(w, profiles.walls(prof)))
    with pytest.raises(rt.RtError):
        rt.params_profile("nope", 8, 8)


def test_camera_constant(built):
    # the values the reference's compilers fold: SURVEY.md §7 "Camera constant"
    assert abs(rt.camera_z(512) - (-443.405029)) < 1e-4
    assert abs(rt.camera_z(1920) - (-1662.7688)) < 1e-3
    assert rt.camera_z(512) == pyoracle.lib().orc_camera_z(512, float(profiles.ALPHA))


def test_png_writer_roundtrip(built, tmp_path):
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(37, 53, 3)).astype(np.uint8)
    path = str(tmp_path / "x.png")
    rt.write_png(path, img)
    assert np.array_equal(np.array(Image.open(path)), img)


def test_move_light_orbits(built):
    L = (-10.0, 20.0, 40.0)
    w = 2 * np.pi / (240 * 0.02)
    for _ in range(240):
        L = rt.move_light(L, w, 0.02)
    assert abs(L[0] + 10) < 0.05 and L[1] == 20.0 and abs(L[2] - 40) < 0.05


@pytest.mark.skipif(rt.device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback(built):
    with pytest.raises(rt.RtError) as e:
        rt.Scene(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_instancing_and_bvh_build_match_the_oracle_at_scale(cat_path):
    """rt_mesh_instance + rt_mesh_build_bvh (product host code) against numpy instancing + the oracle's builder:
    64 baked cat instances (253,056 triangles), the reduced form of BASELINE.json configs[4]."""
    import raytracinggpu_b200 as rt
    from raytracinggpu_b200 import synthetic
    from oracle import scenes
    copies = 64
    scales, offs = synthetic.instance_lattice(copies)
    m = rt.Mesh.read_obj(cat_path).instance(scales, offs).build_bvh()
    desc = scenes.instanced_cat_scene("optimized", copies, obj_path=cat_path)
    v, recs, bvh = desc["mesh"]
    assert m.counts()[1] == copies * 3954 == recs.shape[0]
    assert np.array_equal(m.vertices.view(np.uint32), np.asarray(v).view(np.uint32))
    assert np.array_equal(m.tri_records[:, :3], np.asarray(recs)[:, :3])
    assert np.array_equal(m.arr_bvh.view(np.uint32), np.asarray(bvh).view(np.uint32))


def test_instance_lattice_is_deterministic_and_inside_the_room():
    from raytracinggpu_b200 import synthetic
    s1, o1 = synthetic.instance_lattice()
    s2, o2 = synthetic.instance_lattice()
    assert np.array_equal(o1, o2) and len(s1) == 2529 and 2529 * 3954 == 9999666
    assert o1[:, 0].min() > -52 and o1[:, 0].max() < 52 and o1[:, 1].min() > -8 and o1[:, 1].max() < 48 and o1[:, 2].min() > -54 and o1[:, 2].max() < 30
