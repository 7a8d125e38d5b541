"""image textures — texture::Image bilinear sampling (/root/reference/src/texture/image.rs:36-48, mod.rs:21-41),
AnimatedImage (animated_image.rs:18-60), the JSON "textures" section with "image" / "animated_image" / "movie" entries
(scene.rs:317-394), named-texture material parameters (scene.rs:45-87) and the PNG reader that stands in for image::open."""
import ctypes as C
import json
import math
import os
import struct
import zlib

import numpy as np
import pytest

from tray_rust_b200 import _ffi as F, api, scenebuild as SB
from oracle import pyoracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def write_png(path, img, filters=(0, 1, 2, 3, 4), palette=None):
    """A PNG written with zlib's dynamic-Huffman deflate and a mix of scanline filters (so the loader's inflate and un-filtering are
    exercised). img: H x W (grey), H x W x 2 / 3 / 4 uint8, or H x W palette indices with `palette` (N x 3)."""
    img = np.asarray(img, np.uint8)
    if img.ndim == 2:
        img = img[..., None]
    h, w, ch = img.shape
    ctype = 3 if palette is not None else {1: 0, 2: 4, 3: 2, 4: 6}[ch]
    rows = img.reshape(h, w * ch).astype(np.int32)
    raw = bytearray()
    for y in range(h):
        ft = filters[y % len(filters)]
        cur = rows[y]
        left = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
        up = rows[y - 1] if y else np.zeros_like(cur)
        ul = np.concatenate([np.zeros(ch, np.int32), up[:-ch]])
        if ft == 0:
            f = cur
        elif ft == 1:
            f = cur - left
        elif ft == 2:
            f = cur - up
        elif ft == 3:
            f = cur - (left + up) // 2
        else:
            p = left + up - ul
            pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
            f = cur - pred
        raw.append(ft)
        raw += (f & 0xff).astype(np.uint8).tobytes()

    def chunk(ty, data):
        return struct.pack(">I", len(data)) + ty + data + struct.pack(">I", zlib.crc32(ty + data) & 0xffffffff)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0))
    if palette is not None:
        out += chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    z = zlib.compress(bytes(raw), 6)
    out += chunk(b"IDAT", z[: len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b"")
    open(path, "wb").write(out)


def checker(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    img[..., 3] = 255
    return img


def bilinear(img, u, v, channel=None):
    """texture/mod.rs:21-41 in float64 (tolerance checks only)."""
    h, w = img.shape[:2]
    x, y = u * w, v * h
    x0, y0 = int(x), int(y)
    g = lambda xx, yy: img[min(yy, h - 1), min(xx, w - 1), :3].astype(np.float64) / 255.0
    sx, sy = x - x0, y - y0
    return g(x0, y0) * (1 - sx) * (1 - sy) + g(x0 + 1, y0) * sx * (1 - sy) + g(x0, y0 + 1) * (1 - sx) * sy + g(x0 + 1, y0 + 1) * sx * sy


def textured_floor(img, depth=0, spp=16):
    b = SB.SceneBuilder(8, 8, spp)
    b.integrator = (F.INTEGRATOR_WHITTED, 0, depth)
    t = b.add_texture(img)
    m = b.add_material(F.MAT_MATTE, (0, 0, 0), roughness=0.0, tex_c0=t)
    b.receiver(F.SHAPE_RECT, m, [SB.trs(q=SB.quat_axis_angle((1, 0, 0), -90))], p0=8.0, p1=6.0)       # floor y = 0, 8 x 6
    b.point_light([SB.trs(t=(0, 4, 0))], (1, 1, 1, 8))
    b.add_camera([SB.trs(t=(0.7, 3, -6), q=SB.quat_axis_angle((1, 0, 0), 26.565))], fov=12.0)
    return b


def test_oracle_samples_the_image_bilinearly_at_the_hit_uv(oracle):
    """Whitted on a Lambertian floor whose diffuse colour is an image: L = tex(u, v) / pi * I / d^2 * cos, with (u, v) the rectangle's
    parameterisation (rectangle.rs:54-55) and tex the bilinear lookup of image.rs:36-48."""
    img = checker(5, 7, 1)
    o = O.OracleScene(textured_floor(img).finish())
    o.update_frame(0, 0.0, 0.0)
    s, _ = o.render_samples(seed=4)
    rays, _ = o.camera_rays(seed=4)
    hits, _ = o.intersect(rays)
    p = rays["o"].astype(np.float64) + rays["d"].astype(np.float64) * hits["t"][:, None]
    # object space of the floor: rotate_x(-90) maps object (x, y, 0) to world (x, 0, -y)
    u = (p[:, 0] + 4.0) / 8.0
    v = (-p[:, 2] + 3.0) / 6.0
    d2 = p[:, 0] ** 2 + (4.0 - p[:, 1]) ** 2 + p[:, 2] ** 2
    geo = 8.0 / d2 * (4.0 / np.sqrt(d2)) / math.pi
    want = np.stack([bilinear(img, uu, vv) for uu, vv in zip(u, v)]) * geo[:, None]
    got = np.stack([s["r"], s["g"], s["b"]], axis=1)
    assert np.allclose(got, want, rtol=3e-3, atol=2e-4)
    assert got.std() > 0.01          # the texture is visible


def test_loader_reads_png_textures_and_named_parameters(tmp_path, trb):
    """The "textures" section and material parameters that NAME a texture (scene.rs:45-87, 317-394); the PNG reader against images
    written by zlib (dynamic Huffman, all five scanline filters, split IDAT) in every 8-bit colour type."""
    rgba = checker(9, 13, 2); rgba[..., 3] = np.arange(9 * 13, dtype=np.uint8).reshape(9, 13)
    rgb = checker(6, 5, 3)[..., :3]
    grey = checker(7, 4, 4)[..., 0]
    ga = checker(3, 8, 5)[..., :2]
    pal_idx = (np.arange(4 * 6).reshape(4, 6) % 5).astype(np.uint8)
    pal = checker(1, 5, 6)[0, :, :3]
    write_png(tmp_path / "rgba.png", rgba); write_png(tmp_path / "rgb.png", rgb); write_png(tmp_path / "grey.png", grey)
    write_png(tmp_path / "ga.png", ga); write_png(tmp_path / "pal.png", pal_idx, palette=pal)
    for k in range(3):
        write_png(tmp_path / ("mov%05d.png" % k), checker(4, 4, 10 + k)[..., :3], filters=(4,))
    d = json.load(open(os.path.join(HERE, "golden", "scenes", "c2_smallpt.json")))
    d["textures"] = [{"name": "t_rgba", "type": "image", "file": "rgba.png"}, {"name": "t_rgb", "type": "image", "file": "rgb.png"},
                     {"name": "t_grey", "type": "image", "file": "grey.png"}, {"name": "t_ga", "type": "image", "file": "ga.png"},
                     {"name": "t_pal", "type": "image", "file": "pal.png"},
                     {"name": "t_anim", "type": "animated_image", "keyframes": [{"file": "rgb.png", "time": 0.0}, {"file": "grey.png", "time": 0.25}]},
                     {"name": "t_movie", "type": "movie", "file_prefix": "mov", "file_suffix": ".png", "frames": 3, "framerate": 24}]
    mats = {m["name"]: m for m in d["materials"]}
    first = d["materials"][0]
    assert first["type"] == "matte"
    first["diffuse"] = "t_rgba"; first["roughness"] = "t_grey"
    d["materials"].append({"name": "tex_glass", "type": "glass", "reflect": "t_anim", "transmit": [1, 1, 1], "eta": "t_movie"})
    p = tmp_path / "tex.json"
    p.write_text(json.dumps(d))
    dp = C.POINTER(F.SceneDesc)()
    assert trb.trb_desc_load_json(str(p).encode(), 0, 0, 0, C.byref(dp)) == F.TRB_OK, trb.trb_last_error()
    try:
        desc = dp.contents
        assert desc.n_textures == 7 and desc.n_images == 5 + 2 + 3
        px = lambda i: np.ctypeslib.as_array(desc.images[i].rgba8, shape=(desc.images[i].height, desc.images[i].width, 4))
        assert np.array_equal(px(0), rgba)
        assert np.array_equal(px(1)[..., :3], rgb) and (px(1)[..., 3] == 255).all()
        assert np.array_equal(px(2)[..., 0], grey) and np.array_equal(px(2)[..., 1], grey) and (px(2)[..., 3] == 255).all()
        assert np.array_equal(px(3)[..., 2], ga[..., 0]) and np.array_equal(px(3)[..., 3], ga[..., 1])
        assert np.array_equal(px(4)[..., :3], pal[pal_idx])
        assert [(desc.textures[k].first_image, desc.textures[k].n_images) for k in range(7)] == [(0, 1), (1, 1), (2, 1), (3, 1), (4, 1), (5, 2), (7, 3)]
        assert [desc.images[k].time for k in (5, 6, 7, 8, 9)] == [0.0, 0.25, 0.0, np.float32(1) / np.float32(24), np.float32(2) / np.float32(24)]
        m0 = desc.materials[0]
        assert list(m0.tex) == [1, 0, 3, 0]
        mg = desc.materials[desc.n_materials - 1]
        assert mg.type == F.MAT_GLASS and list(mg.tex) == [6, 0, 0, 7] and list(mg.c1) == [1.0, 1.0, 1.0]
    finally:
        trb.trb_desc_free(dp)
    d["materials"][0]["diffuse"] = "no_such_texture"
    p.write_text(json.dumps(d))
    assert trb.trb_desc_load_json(str(p).encode(), 0, 0, 0, C.byref(dp)) == F.TRB_INVALID_ARG and b"texture" in trb.trb_last_error()


def textured_zoo(spp=8, size=64):
    """Every shape's (u, v) and every textured parameter kind: colour / roughness / eta textures, an animated image, a mesh."""
    b = SB.scene_materials_zoo(size, size, spp, SB.synthetic_merl_table())
    b.film.update(frames=4, end_frame=3, scene_time=2.0)
    t_col = b.add_texture(checker(16, 12, 21))
    rough = checker(8, 8, 22); rough[..., 0] = rough[..., 0] // 4          # roughness in "degrees" (matte) / Beckmann width
    t_rough = b.add_texture(rough)
    eta = checker(4, 4, 23); eta[..., 0] = 255                             # eta = 1.0 from channel 0... stays physical with the scale below
    t_anim = b.add_texture([(checker(6, 9, 24), 0.0), (checker(6, 9, 25), 0.2), (checker(5, 5, 26), 0.6)])
    m_a = b.add_material(F.MAT_MATTE, (0, 0, 0), roughness=0.0, tex_c0=t_col, tex_roughness=t_rough)
    m_b = b.add_material(F.MAT_PLASTIC, (0, 0, 0), (0.6, 0.6, 0.6), roughness=0.2, tex_c0=t_anim)
    m_c = b.add_material(F.MAT_METAL, (0.155265, 0.116723, 0.138381), (4.82835, 3.12225, 2.14696), roughness=0.0, tex_roughness=t_rough)
    m_d = b.add_material(F.MAT_GLASS, (1, 1, 1), (0, 0, 0), eta=1.5, tex_c1=t_col)
    b.receiver(F.SHAPE_SPHERE, m_a, [SB.trs(t=(-9, 9, 6), s=2.5)], p0=1.0)
    b.receiver(F.SHAPE_DISK, m_b, [SB.trs(t=(9, 9, 10), q=SB.quat_axis_angle((0, 1, 0), 180))], p0=3.0, p1=0.5)
    b.receiver(F.SHAPE_RECT, m_a, [SB.trs(t=(0, 0.05, 4), q=SB.quat_axis_angle((1, 0, 0), -90))], p0=10.0, p1=6.0)
    b.receiver(F.SHAPE_SPHERE, m_c, [SB.trs(t=(3, 9, 4), s=2.0)], p0=1.0)
    b.receiver(F.SHAPE_SPHERE, m_d, [SB.trs(t=(-3, 9, 0), s=2.0)], p0=1.0)
    m = b.add_mesh(*SB.icosphere_mesh(2, 1.0, 0.1, 9))
    b.receiver(F.SHAPE_MESH, m_b, [SB.trs(t=(0, 14, 6), s=2.5)], mesh=m)
    return b


def test_oracle_textured_zoo_renders_and_differs_from_constants(oracle):
    b = textured_zoo(4, 32)
    o = O.OracleScene(b.finish())
    o.update_frame(1, 0.5, 1.0)
    s, st = o.render_samples(seed=3)
    assert np.isfinite(s["r"]).all() and s["r"].std() > 0.01 and st.rays_total() > 5 * len(s)


@pytest.mark.gpu
@pytest.mark.parametrize("integ", [(F.INTEGRATOR_PATH, 2, 6), (F.INTEGRATOR_WHITTED, 0, 3)])
def test_textured_scene_gpu_vs_oracle(integ):
    """Image textures on the device, bit for bit: (u, v) of spheres / disks / rectangles / triangles, bilinear Image sampling,
    AnimatedImage keyframe blending at the ray's time, textured colour / roughness parameters (derived Beckmann width and Oren-Nayar
    coefficients recomputed per hit), fused and split shade kernels."""
    b = textured_zoo()
    b.integrator = integ
    desc = b.finish()
    g, o = api.Scene(desc), O.OracleScene(desc)
    for fr in (0, 2):
        g.update_frame(fr, fr * 0.5, (fr + 1) * 0.5); o.update_frame(fr, fr * 0.5, (fr + 1) * 0.5)
        os_, ost = o.render_samples(seed=7)
        gs, gst = g.render_samples(seed=7, flags=F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW)
        assert gs.tobytes() == os_.tobytes()
        assert gst.rays_total() == ost.rays_total()
        if integ[0] == F.INTEGRATOR_PATH:
            g.set_option("shade.split", 1)
            assert g.render_samples(seed=7)[0].tobytes() == os_.tobytes()
            g.set_option("shade.split", 0)
            assert g.render_samples(seed=7, flags=F.RENDER_MEGAKERNEL)[0].tobytes() == os_.tobytes()
    gf, _ = g.render(seed=7, flags=F.RENDER_NO_UPDATE); of, _ = o.render(seed=7, flags=F.RENDER_NO_UPDATE)
    ig = gf[..., :3] / np.maximum(gf[..., 3:], 1e-6); io = of[..., :3] / np.maximum(of[..., 3:], 1e-6)
    assert np.sqrt(np.mean((ig - io) ** 2)) < 1e-5
