"""CPU-side checks of the product library: it loads, exports every symbol include/trb.h declares, validates
scene descriptions with the reference's error conditions, and its JSON/OBJ loader flattens the shipped scenes
into descriptions the oracle renders. No compute entry point is called without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tray_rust_b200 import _ffi as F, api, scenebuild as SB
from oracle import pyoracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(REPO, "tests", "golden", "scenes")


def test_library_exports_every_declared_symbol(trb):
    hdr = open(os.path.join(REPO, "include", "trb.h")).read()
    declared = set(re.findall(r"\b(trb_[a-z0-9_]+)\s*\(", hdr)) - {"trb_status"}
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(trb, name), name
    assert set(F.TRB_SYMBOLS) <= declared | {"trb_last_error"}
    assert trb.trb_abi_version() == F.TRB_ABI_VERSION


def test_struct_layouts_match_the_header():
    assert C.sizeof(F.Ray) == 32 and C.sizeof(F.Hit) == 16 and C.sizeof(F.Sample) == 20 and C.sizeof(F.BvhNode) == 32
    assert C.sizeof(F.Keyframe) == 40 and C.sizeof(F.Instance) == 40 and C.sizeof(F.Material) == 56
    assert F.RAY_DTYPE.itemsize == 32 and F.HIT_DTYPE.itemsize == 16 and F.SAMPLE_DTYPE.itemsize == 20 and F.NODE_DTYPE.itemsize == 32
    assert C.sizeof(F.Stats) == 8 * 8 + 8


def _create(trb, desc):
    h = C.c_void_p()
    rc = trb.trb_scene_create(C.byref(desc), 0, C.byref(h))
    if rc == F.TRB_OK:
        trb.trb_scene_destroy(h)
    return rc, (trb.trb_last_error() or b"").decode()


def test_reference_panics_become_status_codes(trb):
    """The conditions the reference panics on are reported before any device work (so this runs without a GPU)."""
    def base():
        return SB.scene_smallpt_like(16, 16, 4)
    b = base(); b.film["width"] = 20
    rc, msg = _create(trb, b.finish())
    assert rc == F.TRB_INVALID_ARG and "evenly divided" in msg                      # block_queue.rs:29-31
    b = SB.SceneBuilder(16, 16, 4)
    m = b.add_material(F.MAT_MATTE, (1, 1, 1), roughness=1.0)
    b.receiver(F.SHAPE_SPHERE, m, [SB.trs()], p0=1.0); b.add_camera([SB.trs(t=(0, 0, -5))])
    rc, msg = _create(trb, b.finish())
    assert rc == F.TRB_INVALID_ARG and "At least one light is required" in msg      # multithreaded.rs:39
    b = SB.SceneBuilder(16, 16, 4); b.add_camera([SB.trs()])
    rc, msg = _create(trb, b.finish())
    assert rc == F.TRB_INVALID_ARG and "does not have any objects" in msg           # scene.rs:134
    b = base(); b.cameras = []
    assert _create(trb, b.finish())[0] == F.TRB_INVALID_ARG
    b = base(); b.integrator = (7, 1, 2)
    rc, msg = _create(trb, b.finish())
    assert rc == F.TRB_INVALID_ARG and "Unrecognized integrator type" in msg          # scene.rs:313
    b = base(); b.integrator = (F.INTEGRATOR_WHITTED, 0, 40)
    assert _create(trb, b.finish())[0] == F.TRB_UNSUPPORTED                            # deeper than the device recursion stack provided for
    d = base().finish(); d.abi_version = 99
    assert _create(trb, d)[0] == F.TRB_INVALID_ARG
    b = base(); b.instances[5] = b.instances[5][:5] + (99,) + b.instances[5][6:]      # material out of range
    assert _create(trb, b.finish())[0] == F.TRB_INVALID_ARG
    b = base(); mm = b.add_mesh(*SB.icosphere_mesh(0)); b.area_light(F.SHAPE_MESH, 0, [SB.trs()], (1, 1, 1), mesh=mm)
    rc, msg = _create(trb, b.finish())
    assert rc == F.TRB_INVALID_ARG and "not sampleable" in msg                       # scene.rs:577-579
    assert trb.trb_scene_create(None, 0, C.byref(C.c_void_p())) == F.TRB_INVALID_ARG


def test_no_cpu_fallback(trb):
    """On a machine without a CUDA device a valid scene must be refused, not rendered on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc, msg = _create(trb, SB.scene_smallpt_like(16, 16, 4).finish())
    assert rc in (F.TRB_NO_DEVICE, F.TRB_CUDA) and rc != F.TRB_OK


def test_null_and_unbuilt_handles(trb):
    assert trb.trb_render(None, None, None, None) == F.TRB_INVALID_ARG
    assert trb.trb_scene_update_frame(None, 0, 0.0, 0.0) == F.TRB_INVALID_ARG
    trb.trb_scene_destroy(None)


def _load(trb, name, w=0, h=0, spp=0):
    d = C.POINTER(F.SceneDesc)()
    rc = trb.trb_desc_load_json(os.path.join(SCENES, name).encode(), w, h, spp, C.byref(d))
    return rc, d


def test_loader_flattens_cornell_box(trb):
    rc, d = _load(trb, "c1_cornell_box.json", 64, 48, 8)
    assert rc == F.TRB_OK, trb.trb_last_error()
    desc = d.contents
    assert (desc.film.width, desc.film.height, desc.film.samples) == (64, 48, 8)
    assert desc.n_instances == 8 and desc.n_meshes == 1 and desc.n_materials == 4 and desc.n_cameras == 1
    kinds = [desc.instances[i].kind for i in range(8)]
    assert kinds == [0, 0, 0, 0, 0, 1, 0, 0]                                   # JSON object order == light order (Q20)
    assert [desc.instances[i].n_splines for i in range(8)] == [2, 2, 2, 2, 2, 1, 1, 1]  # group levels stack (Q18)
    # own spline first, then the group's (AnimatedTransform::mul, animated_transform.rs:78-87)
    own = desc.keyframes[desc.splines[desc.instances[0].spline_first].ctrl_first]
    grp = desc.keyframes[desc.splines[desc.instances[0].spline_first + 1].ctrl_first]
    assert list(own.translation) == [0, 0, 20] and list(own.scaling) == [15, 12, 1] and list(grp.translation) == [0, 12, 0]
    rot = desc.keyframes[desc.splines[desc.instances[1].spline_first].ctrl_first]   # left wall: rotate_y(90)
    assert np.allclose(list(rot.rotation), [0, np.sqrt(0.5), 0, np.sqrt(0.5)], atol=1e-6) and np.allclose(list(rot.scaling), [20, 12, 1], atol=1e-5)
    me = desc.meshes[0]
    assert (me.n_verts, me.n_tris) == (24, 12)                                  # cube.obj: 6 quads, unique (v,vt,vn) per face
    assert [me.indices[i] for i in range(6)] == [0, 1, 2, 0, 2, 3]              # fan triangulation
    em = desc.color_keys[desc.instances[5].emission_first]
    assert np.allclose(list(em.rgba), [40.0, 0.772549 * 40, 0.560784 * 40, 40.0], rtol=1e-6)   # load_color scales by [3] (Q15)
    o = O.OracleScene(desc)
    film, st = o.render(seed=1)
    img = film[..., :3] / np.maximum(film[..., 3:], 1e-9)
    assert np.isfinite(film).all() and 0.1 < img.mean() < 0.5
    left, right = img[:, :8].mean((0, 1)), img[:, -8:].mean((0, 1))
    assert left[0] > 2 * left[1] and right[1] > 1.5 * right[0]                 # red wall left, green wall right
    trb.trb_desc_free(d)


def test_loader_smallpt_and_errors(trb):
    rc, d = _load(trb, "c2_smallpt.json", 32, 32, 4)
    assert rc == F.TRB_OK
    desc = d.contents
    assert desc.n_instances == 8 and desc.n_meshes == 0 and desc.n_materials == 6
    assert [desc.materials[i].type for i in range(6)] == [F.MAT_MATTE] * 3 + [F.MAT_METAL, F.MAT_PLASTIC, F.MAT_GLASS]
    assert abs(desc.materials[5].eta - 1.52) < 1e-6
    trb.trb_desc_free(d)
    rc, d = _load(trb, "does_not_exist.json")
    assert rc == F.TRB_IO and b"Failed to open scene file" in trb.trb_last_error()


def test_loader_rejects_bad_json(tmp_path, trb):
    p = tmp_path / "bad.json"
    p.write_text('{"film": {"width": 8}')
    d = C.POINTER(F.SceneDesc)()
    assert trb.trb_desc_load_json(str(p).encode(), 0, 0, 0, C.byref(d)) == F.TRB_INVALID_ARG
    p.write_text('{"film": {"width": 8, "height": 8, "samples": 1, "frames": 1, "start_frame": 0, "end_frame": 0, "scene_time": 0,'
                 ' "filter": {"type": "box", "width": 1, "height": 1}}}')
    assert trb.trb_desc_load_json(str(p).encode(), 0, 0, 0, C.byref(d)) == F.TRB_INVALID_ARG
    assert b"Unrecognized filter type" in trb.trb_last_error()


def test_cpp_host_mirror_compiles_and_reports_no_device(tmp_path):
    """include/tray_exec.hpp (the C++ mirror of Exec / Scene / RenderTarget / Config) builds against libtrb.so; without a
    GPU the example must fail loudly with a status, never render on the host."""
    import subprocess, torch
    exe = str(tmp_path / "render_cornell")
    lib = os.path.join(REPO, "tray_rust_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(REPO, "include"), os.path.join(REPO, "examples", "render_cornell.cpp"),
                    "-L" + lib, "-ltrb", "-Wl,-rpath," + lib, "-o", exe], check=True)
    r = subprocess.run([exe, os.path.join(SCENES, "c1_cornell_box.json"), "32", "32", "2"], capture_output=True, text=True, cwd=str(tmp_path))
    if torch.cuda.is_available():
        assert r.returncode == 0 and "rendering took" in r.stdout
    else:
        assert r.returncode == 1 and "status" in r.stderr


def _build_abi_check(tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_check")
    lib = os.path.join(REPO, "tray_rust_b200", "lib")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(REPO, "include"), os.path.join(REPO, "tests", "c", "abi_check.c"),
                    "-L" + lib, "-ltrb", "-Wl,-rpath," + lib, "-o", exe], check=True)
    return exe


CTYPES_OF = {"trb_keyframe": F.Keyframe, "trb_spline": F.Spline, "trb_color_key": F.ColorKey, "trb_instance": F.Instance, "trb_mesh": F.Mesh,
             "trb_material": F.Material, "trb_image": F.Image, "trb_texture": F.Texture, "trb_camera": F.Camera, "trb_film": F.Film, "trb_integrator": F.Integrator, "trb_scene_desc": F.SceneDesc,
             "trb_render_cfg": F.RenderCfg, "trb_stats": F.Stats, "trb_ray": F.Ray, "trb_hit": F.Hit, "trb_sample": F.Sample, "trb_bvh_node": F.BvhNode}


def test_plain_c_caller_layout_matches_ctypes_and_the_documented_rust_binding(tmp_path):
    """A C11 translation unit including only include/trb.h: its _Static_asserts pin ABI v4; every sizeof / offsetof it prints
    must equal the ctypes mirror, and the #[repr(C)] structs INTEGRATION.md documents for the Rust side must list the same
    fields in the same order (round 1 shipped a Rust TrbRenderCfg three fields short)."""
    import subprocess
    out = subprocess.run([_build_abi_check(tmp_path), "layout"], capture_output=True, text=True, check=True).stdout
    sizes, offsets = {}, {}
    for line in out.splitlines():
        a, b, *c = line.split()
        if b == "sizeof":
            sizes[a] = int(c[0])
        elif "." in a:
            offsets.setdefault(a.split(".")[0], []).append((a.split(".")[1], int(b)))
    assert "abi_version %d" % F.TRB_ABI_VERSION in out
    assert set(sizes) == set(CTYPES_OF)
    for name, ct in CTYPES_OF.items():
        assert C.sizeof(ct) == sizes[name], name
        assert [(f, getattr(ct, f).offset) for f, _ in ct._fields_] == offsets[name], name
    # the Rust declarations in INTEGRATION.md: same field names, same order, matching scalar widths
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    width = {"u32": 4, "f32": 4, "u64": 8}
    for rust, cname in (("TrbRenderCfg", "trb_render_cfg"), ("TrbStats", "trb_stats")):
        m = re.search(r"pub struct %s \{(.*?)\}" % rust, doc, re.S)
        assert m, rust
        fields = [(n.strip(), t.strip()) for n, t in re.findall(r"(\w+)\s*:\s*(\w+)", m.group(1))]
        assert [n for n, _ in fields] == [f for f, _ in offsets[cname]], rust
        off = 0
        for (n, t), (_, o) in zip(fields, offsets[cname]):
            off = (off + width[t] - 1) // width[t] * width[t]
            assert off == o, (rust, n)
            off += width[t]
        assert (off + 7) // 8 * 8 == sizes[cname] or off == sizes[cname], rust


def test_plain_c_caller_reports_no_device_without_a_gpu(tmp_path):
    import subprocess, torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_plain_c_caller_renders_through_the_documented_sequence")
    r = subprocess.run([_build_abi_check(tmp_path), "render", os.path.join(SCENES, "c1_cornell_box.json"), "32", "32", "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_plain_c_caller_renders_through_the_documented_sequence(tmp_path):
    """INTEGRATION.md's call sequence from C: trb_scene_load_json -> trb_render (whole frame, one call) -> trb_film_to_srgb8."""
    import subprocess
    r = subprocess.run([_build_abi_check(tmp_path), "render", os.path.join(SCENES, "c1_cornell_box.json"), "400", "400", "64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "camera_samples %d" % (400 * 400 * 64) in r.stdout
