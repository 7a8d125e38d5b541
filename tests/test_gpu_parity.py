"""Parity tests proper (run on the B200 box): the CUDA path, called through the C ABI, against the oracle and the
committed golden vectors. Bars: ray-scene hit records, traversal counters, camera rays, LD samples and
per-camera-sample radiance BIT-EXACT; the film within 1e-5 RMSE (its accumulation order is atomic-dependent)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from tray_rust_b200 import _ffi as F, api, scenebuild as SB
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

GOLD = make_golden.golden_scenes()
KEYS = ["camera_samples", "rays_primary", "rays_shadow", "rays_mis", "rays_continuation", "node_tests", "tri_tests", "inst_tests"]


def both(desc):
    g, o = api.Scene(desc), O.OracleScene(desc)
    g.update_frame(0, 0.0, 0.0); o.update_frame(0, 0.0, 0.0)
    return g, o


@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_matches_golden_vectors(name):
    mk, kw = GOLD[name]
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    g = api.Scene(mk())
    g.update_frame(*make_golden.frame_of(name))
    rays, xy = g.camera_rays(**kw)
    assert rays.tobytes() == gold["rays"].tobytes() and xy.tobytes() == gold["xy"].tobytes()
    hits, _ = g.intersect(gold["rays"])
    assert hits.tobytes() == gold["hits"].tobytes()
    s, st = g.render_samples(flags=F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW, **kw)
    assert s.tobytes() == gold["samples"].tobytes()
    assert [st.rays_primary, st.rays_shadow, st.rays_mis, st.rays_continuation, st.node_tests, st.tri_tests, st.inst_tests] == gold["counters"].tolist()
    film, _ = g.render(flags=F.RENDER_NO_UPDATE, **kw)
    img_g = film[..., :3] / np.maximum(film[..., 3:], 1e-6); img_o = gold["film"][..., :3] / np.maximum(gold["film"][..., 3:], 1e-6)
    assert np.sqrt(np.mean((img_g - img_o) ** 2)) < 1e-5 and np.allclose(film, gold["film"], rtol=1e-4, atol=1e-5)


SCENES = {
    "zoo": lambda: SB.scene_materials_zoo(64, 64, 16, SB.synthetic_merl_table()),
    "smallpt": lambda: SB.scene_smallpt_like(64, 64, 16),
    "c4_20k": lambda: SB.scene_c4(20000, 128, 72, 8),
    "c3": lambda: SB.scene_c3(96, 72, 8, subdiv=4),
}


@pytest.mark.gpu
def test_animated_scene_gpu_vs_oracle():
    """keyframed instances (one- and two-level stacks), keyframed camera, moving lights with keyframed
    emission — transforms recomposed per ray on the device (receiver.rs:30, emitter.rs:122,170,176,197, camera.rs:156)."""
    desc = SB.scene_animated(64, 64, 8, frames=4, scene_time=1.0, animated_fov=True).finish()  # + CameraFov::Animated (camera.rs:134-141)
    g, o = api.Scene(desc), O.OracleScene(desc)
    kw = dict(sample_first=0, sample_count=4, seed=21)
    rng = np.random.default_rng(5)
    for fr in range(4):
        t0, t1 = fr * 0.25, (fr + 1) * 0.25
        g.update_frame(fr, t0, t1); o.update_frame(fr, t0, t1)
        gn, go = g.bvh(-1); on, oo = o.bvh(-1)
        assert gn.tobytes() == on.tobytes() and np.array_equal(go, oo)       # animation_bounds (128 time samples / Q22)
        for i in range(desc.n_instances):
            for x, y in zip(g.transform(i), o.transform(i)):
                assert np.array_equal(api.bits(x), api.bits(y))
        gr, gxy = g.camera_rays(**kw); orr, oxy = o.camera_rays(**kw)
        assert gr.tobytes() == orr.tobytes() and gxy.tobytes() == oxy.tobytes()
        assert len(np.unique(gr["o"], axis=0)) > 1                            # the camera moves inside the shutter interval
        gh, gs = g.intersect(orr); oh, os_ = o.intersect(orr)
        assert gh.tobytes() == oh.tobytes()
        assert (gs.node_tests, gs.tri_tests, gs.inst_tests) == (os_.node_tests, os_.tri_tests, os_.inst_tests)
        rays = np.zeros(20000, F.RAY_DTYPE)
        rays["o"] = rng.uniform((-12, 1, -15), (12, 22, 15), size=(20000, 3)).astype(np.float32)
        d = rng.normal(size=(20000, 3)).astype(np.float32)
        rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["min_t"] = 0.001; rays["max_t"] = np.inf
        assert g.intersect(rays)[0].tobytes() == o.intersect(rays)[0].tobytes()
        gsamp, gst = g.render_samples(flags=F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW, **kw)
        osamp, ost = o.render_samples(**kw)
        assert gsamp.tobytes() == osamp.tobytes()
        assert [getattr(gst, k) for k in KEYS] == [getattr(ost, k) for k in KEYS]
        gsamp2, _ = g.render_samples(**kw)
        assert gsamp2.tobytes() == gsamp.tobytes()
        gsamp3, gst3 = g.render_samples(flags=F.RENDER_MEGAKERNEL | F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW, **kw)
        assert gsamp3.tobytes() == gsamp.tobytes() and [getattr(gst3, k) for k in KEYS] == [getattr(gst, k) for k in KEYS]
        # Exec::render for this frame: update_frame from (current_frame, scene_time / frames) then the film
        gf, _ = g.render(current_frame=fr, **kw); of, _ = o.render(current_frame=fr, **kw)
        ig = gf[..., :3] / np.maximum(gf[..., 3:], 1e-6); io = of[..., :3] / np.maximum(of[..., 3:], 1e-6)
        assert np.isfinite(gf).all() and np.sqrt(np.mean((ig - io) ** 2)) < 1e-5


@pytest.mark.parametrize("name", sorted(SCENES))
def test_gpu_vs_oracle(name):
    desc = SCENES[name]().finish()
    g, o = both(desc)
    gn, go = g.bvh(-1); on, oo = o.bvh(-1)
    assert gn.tobytes() == on.tobytes() and np.array_equal(go, oo)                       # BVH<Instance>
    for m in range(desc.n_meshes):
        a, b = g.bvh(m), o.bvh(m)
        assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1])           # BVH<Triangle>
    for i in range(desc.n_instances):
        for x, y in zip(g.transform(i), o.transform(i)):
            assert np.array_equal(api.bits(x), api.bits(y))
    assert np.array_equal(api.bits(g.filter_table()), api.bits(o.filter_table()))
    assert np.array_equal(g.block_list(), o.block_list()) and np.array_equal(g.block_list(3, 5), o.block_list(3, 5))
    kw = dict(sample_first=0, sample_count=4, seed=7)
    gr, gxy = g.camera_rays(**kw); orr, oxy = o.camera_rays(**kw)
    assert gr.tobytes() == orr.tobytes() and gxy.tobytes() == oxy.tobytes()
    gh, gs = g.intersect(orr); oh, os_ = o.intersect(orr)
    assert gh.tobytes() == oh.tobytes()
    assert (gs.node_tests, gs.tri_tests, gs.inst_tests) == (os_.node_tests, os_.tri_tests, os_.inst_tests)
    # incoherent secondary-like rays from the hit points
    rng = np.random.default_rng(3)
    hit = np.nonzero(oh["inst"] != F.MISS)[0]
    sel = rng.choice(hit, size=min(20000, len(hit)), replace=False)
    rays = np.zeros(len(sel), F.RAY_DTYPE)
    rays["o"] = orr["o"][sel] + orr["d"][sel] * oh["t"][sel, None]
    d = rng.normal(size=(len(sel), 3)).astype(np.float32)
    rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["min_t"] = 0.001; rays["max_t"] = np.inf
    assert g.intersect(rays)[0].tobytes() == o.intersect(rays)[0].tobytes()
    # whole paths, bit-exact, with the reference's closest-hit shadow rays so the counters are comparable
    gsamp, gst = g.render_samples(flags=F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW, **kw)
    osamp, ost = o.render_samples(**kw)
    assert gsamp.tobytes() == osamp.tobytes()
    assert [getattr(gst, k) for k in KEYS] == [getattr(ost, k) for k in KEYS]
    # default mode (shadow rays stop at the first accepted hit): identical radiance, fewer tests
    gsamp2, gst2 = g.render_samples(flags=F.RENDER_STATS, **kw)
    assert gsamp2.tobytes() == gsamp.tobytes() and gst2.node_tests <= gst.node_tests
    # the megakernel execution shape agrees bit for bit with the wavefront pipeline, counters included
    gsamp3, gst3 = g.render_samples(flags=F.RENDER_MEGAKERNEL | F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW, **kw)
    assert gsamp3.tobytes() == gsamp.tobytes() and [getattr(gst3, k) for k in KEYS] == [getattr(gst, k) for k in KEYS]
    # film (splat order is atomic-dependent): tolerance
    gf, _ = g.render(**kw); of, _ = o.render(**kw)
    ig = gf[..., :3] / np.maximum(gf[..., 3:], 1e-6); io = of[..., :3] / np.maximum(of[..., 3:], 1e-6)
    assert np.isfinite(gf).all() and np.sqrt(np.mean((ig - io) ** 2)) < 1e-5
    assert np.array_equal(g.to_srgb8(of), o.to_srgb8(of))


def test_json_scenes_through_the_loader():
    """C1 / C2 inputs through trb_scene_load_json-equivalent path (reference-shaped host API)."""
    from tray_rust_b200 import exec as X
    for fn, dims in (("c1_cornell_box.json", (40, 32, 8)), ("c2_smallpt.json", (32, 32, 8))):
        path = os.path.join(HERE, "golden", "scenes", fn)
        scene, rt, spp, fi = X.Scene.load_file(path, 0, *dims)
        assert rt.dimensions() == dims[:2] and spp == dims[2] and fi.frames == 1
        cfg = X.Config(spp=spp, frame_info=fi, seed=9)
        st = X.B200().render(scene, rt, cfg)
        lib = F.load_trb()
        d = C.POINTER(F.SceneDesc)()
        assert lib.trb_desc_load_json(path.encode(), *dims, C.byref(d)) == 0
        o = O.OracleScene(d.contents)
        of, ost = o.render(seed=9)
        ig = rt.pixels[..., :3] / np.maximum(rt.pixels[..., 3:], 1e-6); io = of[..., :3] / np.maximum(of[..., 3:], 1e-6)
        assert np.sqrt(np.mean((ig - io) ** 2)) < 1e-5
        assert st.rays_primary == ost.rays_primary and st.rays_continuation == ost.rays_continuation
        # two passes accumulate to the same film (Exec mirror with samples_per_pass)
        rt2 = X.RenderTarget(*dims[:2])
        X.B200(samples_per_pass=spp // 2).render(scene, rt2, cfg)
        assert np.allclose(rt2.pixels, rt.pixels, rtol=1e-4, atol=1e-5)
        assert np.array_equal(X.get_render(scene, rt), o.to_srgb8(rt.pixels))
        lib.trb_desc_free(d); scene.close()


def test_tr15_like_json_scene_gpu_vs_oracle():
    """C5-shaped input through the JSON loader: keyframed camera / groups / objects, keyed emission, OBJ mesh, MERL file."""
    import make_scenes
    merl = os.path.join(HERE, "golden", "scenes", "merl", "synthetic.binary")
    if not os.path.exists(merl):
        make_scenes.write_synthetic_merl(merl)
    path = os.path.join(HERE, "golden", "scenes", "c5_tr15_like.json")
    lib = F.load_trb()
    d = C.POINTER(F.SceneDesc)()
    assert lib.trb_desc_load_json(path.encode(), 96, 56, 4, C.byref(d)) == 0
    desc = d.contents
    assert desc.n_merl == 1 and desc.film.frames == 50 and desc.n_instances == 12
    g, o = api.Scene(desc), O.OracleScene(desc)
    step = desc.film.scene_time / desc.film.frames
    for fr in (0, 9, 14, 49):
        g.update_frame(fr, fr * step, (fr + 1) * step); o.update_frame(fr, fr * step, (fr + 1) * step)
        gn, go = g.bvh(-1); on, oo = o.bvh(-1)
        assert gn.tobytes() == on.tobytes() and np.array_equal(go, oo)
        gr, gxy = g.camera_rays(seed=31); orr, oxy = o.camera_rays(seed=31)
        assert gr.tobytes() == orr.tobytes()
        assert g.intersect(orr)[0].tobytes() == o.intersect(orr)[0].tobytes()
        gs, gst = g.render_samples(seed=31, flags=F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW); os_, ost = o.render_samples(seed=31)
        assert gs.tobytes() == os_.tobytes()
        assert [getattr(gst, k) for k in KEYS] == [getattr(ost, k) for k in KEYS]
        assert g.render_samples(seed=31)[0].tobytes() == gs.tobytes()
    # Exec::render on a frame in the middle of the emission ramp, through the reference-shaped host mirror
    from tray_rust_b200 import exec as X
    scene, rt, spp, fi = X.Scene.load_file(path, 0, 96, 56, 4)
    assert fi.frames == 50
    cfg = X.Config(spp=spp, frame_info=fi, seed=5, current_frame=9)
    X.B200().render(scene, rt, cfg)
    of, _ = o.render(seed=5, current_frame=9)
    ig = rt.pixels[..., :3] / np.maximum(rt.pixels[..., 3:], 1e-6); io = of[..., :3] / np.maximum(of[..., 3:], 1e-6)
    assert np.sqrt(np.mean((ig - io) ** 2)) < 1e-5 and ig.mean() > 0.005
    lib.trb_desc_free(d); scene.close()


def test_film_kernels_agree():
    """The default film kernel (per-warp private tiles, no shared atomics) against the shared-atomics one (TRB_FILM_V2=0):
    same weights and products, different float summation order (measured on B200: profiles/r01_film_v2_check.json)."""
    desc = SB.scene_materials_zoo(64, 64, 8, SB.synthetic_merl_table()).finish()
    g = api.Scene(desc)
    try:
        os.environ["TRB_FILM_V2"] = "0"; f1, _ = g.render(seed=3)
        os.environ["TRB_FILM_V2"] = "1"; f2, _ = g.render(seed=3)
    finally:
        os.environ.pop("TRB_FILM_V2", None)
    assert np.allclose(f1, f2, rtol=1e-4, atol=1e-5) and np.allclose(f1[..., 3], f2[..., 3], rtol=1e-5, atol=1e-6)


def test_wide_gaussian_filter_film_vs_oracle():
    """A filter whose reach (width / inv_width = 9 px) exceeds filter_pixel_width (6): RenderTarget::write's per-2x2-block sample
    filter rejects contributions the per-pixel distance test would accept. No scene of the reference uses such a filter."""
    b = SB.scene_c4(20000, 128, 72, 8)
    b.film.update(filter_type=F.FILTER_GAUSSIAN, filter_w=3.0, filter_h=2.5, filter_b=0.5, filter_c=0.0)
    desc = b.finish()
    g, o = api.Scene(desc), O.OracleScene(desc)
    gf, _ = g.render(seed=3); of, _ = o.render(seed=3)
    ig = gf[..., :3] / np.maximum(gf[..., 3:], 1e-6); io = of[..., :3] / np.maximum(of[..., 3:], 1e-6)
    assert np.sqrt(np.mean((ig - io) ** 2)) < 1e-5 and np.allclose(gf[..., 3], of[..., 3], rtol=1e-4, atol=1e-5)


def test_edge_cases():
    desc = SB.scene_smallpt_like(16, 16, 4).finish()
    g, o = both(desc)
    assert len(g.intersect(np.zeros(0, F.RAY_DTYPE))[0]) == 0                              # empty batch
    rays = np.zeros(6, F.RAY_DTYPE)
    rays["o"] = [0, 12, -60]; rays["d"] = [[0, 0, 1], [0, 0, -1], [0, 0, 0], [np.nan, 0, 1], [0, 1e-30, 1], [1e30, 0, 1]]
    rays["max_t"] = [np.inf, np.inf, np.inf, np.inf, np.inf, 1e-3]
    assert g.intersect(rays)[0].tobytes() == o.intersect(rays)[0].tobytes()               # zero / NaN / denormal directions
    with pytest.raises(api.TrbError):
        g.render(sample_first=3, sample_count=4)                                           # beyond spp
    film, st = g.render(block_start=10 ** 6, block_count=5)                                # "This block queue is empty!"
    assert st.camera_samples == 0 and not film.any()
    f1, s1 = g.render(block_start=1, block_count=2, seed=5); f2, _ = o.render(block_start=1, block_count=2, seed=5)
    assert s1.camera_samples == 2 * 64 * 4 and np.allclose(f1, f2, rtol=1e-4, atol=1e-5)
    fresh = api.Scene(desc)
    with pytest.raises(api.TrbError):
        fresh.render_samples()                                                             # update_frame not called (scene.rs:179)


def test_box_hit_finite_selftest_and_forced_literal_path():
    """The trace kernel's 23-instruction box test (all-finite rays) against the literal transcription of BBox::fast_intersect:
    generated awkward cases on the device, then a whole render with every ray forced onto the literal path."""
    trb = F.load_trb()
    out = (C.c_uint64 * 4)()
    assert trb.trb_selftest_box(1 << 22, 7, out) == F.TRB_OK, trb.trb_last_error()
    finite, hits, bad, literal = [int(x) for x in out]
    assert bad == 0 and finite > 1 << 20 and hits > 1 << 14 and literal > 1 << 16, (finite, hits, bad, literal)
    desc = SB.scene_c4(20000, 128, 72, 8).finish()
    g, o = both(desc)
    kw = dict(block_start=0, block_count=40, sample_first=0, sample_count=4, seed=9)
    ref, st_o = o.render_samples(**kw)
    for exact in (0, 1):
        g.set_option("trace.exact_box", exact)
        s, _ = g.render_samples(**kw)                                   # the 9-CTA variant
        s2, st = g.render_samples(flags=F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW, **kw)   # the variant with counters, closest-hit shadow rays
        assert s.tobytes() == ref.tobytes() and s2.tobytes() == ref.tobytes()
        assert all(getattr(st, k) == getattr(st_o, k) for k in KEYS)
    g.set_option("trace.pipe", 0)                                       # round-1 kernel: same results
    assert g.render_samples(**kw)[0].tobytes() == ref.tobytes()


def test_split_shade_buckets_and_kind_instantiations_agree():
    """Split shading with / without the material buckets and the matte instantiations, and the fused kernel: one result."""
    desc = SB.scene_materials_zoo(64, 64, 8, SB.synthetic_merl_table()).finish()
    g, o = both(desc)
    kw = dict(sample_first=0, sample_count=4, seed=11)
    ref = o.render_samples(**kw)[0].tobytes()
    for split, sort, kind in ((-1, 1, 1), (1, 1, 0), (1, 0, 0), (0, 1, 1)):
        g.set_option("shade.split", split); g.set_option("shade.sort", sort); g.set_option("shade.kind", kind)
        assert g.render_samples(**kw)[0].tobytes() == ref, (split, sort, kind)
        film, st = g.render(**kw)
        assert st.camera_samples == 64 * 64 * 4 and np.isfinite(film).all()


def test_full_size_properties_c4():
    """BASELINE-size workload (1M triangles, 1920x1080): size-independent properties instead of an oracle run."""
    g = api.Scene(SB.scene_c4(1_000_000, 1920, 1080, 4096).finish())
    g.update_frame(0, 0.0, 0.0)
    nb = g.n_blocks()
    assert nb == 32400
    # determinism: a camera sample is a pure function of (seed, pixel, sample index), whatever the launch shape
    a, _ = g.render_samples(block_start=500, block_count=40, sample_first=8, sample_count=2, seed=3)
    b, _ = g.render_samples(block_start=500, block_count=40, sample_first=8, sample_count=2, seed=3)
    c, _ = g.render_samples(block_start=520, block_count=10, sample_first=9, sample_count=1, seed=3)
    assert a.tobytes() == b.tobytes()
    assert c.tobytes() == a.reshape(40, 64, 2)[20:30, :, 1].reshape(-1).tobytes()
    assert (a["r"] >= 0).all() and (a["r"] <= 1).all() and np.isfinite(a["r"]).all()       # per-sample clamp
    # tile sharding is additive: blocks [0,n/2) + [n/2,n) == all blocks
    full, sf = g.render(sample_first=0, sample_count=1, seed=3)
    half, s0 = g.render(block_start=0, block_count=nb // 2, sample_first=0, sample_count=1, seed=3)
    _, s1 = g.render(half, block_start=nb // 2, block_count=nb - nb // 2, sample_first=0, sample_count=1, seed=3)
    assert np.allclose(full, half, rtol=1e-4, atol=1e-5)
    assert s0.rays_total() + s1.rays_total() == sf.rays_total() and sf.camera_samples == 1920 * 1080
    # interleaved sharding (dist.shard_interleaved): the union of the 3 shards is the whole block list, films add up
    from tray_rust_b200.dist import shard_interleaved
    acc = np.zeros_like(full); rays_sh = 0
    for r in range(3):
        _, ss = g.render(acc, sample_first=0, sample_count=1, seed=3, **shard_interleaved(r, 3, chunk=32))
        rays_sh += ss.rays_total()
    assert rays_sh == sf.rays_total() and np.allclose(full, acc, rtol=1e-4, atol=1e-5)
    # every pixel got weight, weights are positive in the interior (Mitchell lobes sum > 0)
    assert (full[8:-8, 8:-8, 3] > 0).all()
    # hits: t > 0, triangle ids in range, and closest-hit == the minimum over a second, any-order query of the same ray
    rays, _ = g.camera_rays(block_start=1000, block_count=50, sample_first=0, sample_count=1, seed=3)
    h, st = g.intersect(rays)
    hit = h["inst"] != F.MISS
    assert (h["t"][hit] > 0).all() and (h["prim"][hit] < 1_000_000).all() and st.node_tests > 0
    short = rays.copy(); short["max_t"] = np.where(hit, h["t"] * 0.999, 1.0)
    assert (g.intersect(short)[0]["inst"] == F.MISS).all()                                 # nothing closer than the closest hit


@pytest.mark.parametrize("integ", [(F.INTEGRATOR_WHITTED, 0, 4), (F.INTEGRATOR_NORMALS_DEBUG, 0, 0)])
def test_whitted_and_normals_debug_gpu_vs_oracle(integ):
    """integrator/whitted.rs:41-70 (+ specular_reflection / specular_transmission, integrator/mod.rs:41-103) and
    integrator/normals_debug.rs:28-36 on the device, bit for bit against the oracle: every material / light kind, a mesh, MERL."""
    b = SB.scene_materials_zoo(64, 64, 8, SB.synthetic_merl_table())
    b.integrator = integ
    g, o = both(b.finish())
    gs, gst = g.render_samples(seed=9, flags=F.RENDER_REFERENCE_SHADOW)
    os_, ost = o.render_samples(seed=9)
    assert gs.tobytes() == os_.tobytes()
    assert [getattr(gst, k) for k in KEYS] == [getattr(ost, k) for k in KEYS]
    assert g.render_samples(seed=9)[0].tobytes() == os_.tobytes()          # any-hit shadow rays: same booleans
    gf, _ = g.render(seed=9); of, _ = o.render(seed=9)
    ig = gf[..., :3] / np.maximum(gf[..., 3:], 1e-6); io = of[..., :3] / np.maximum(of[..., 3:], 1e-6)
    assert np.sqrt(np.mean((ig - io) ** 2)) < 1e-5 and ig.mean() > 0.01
    if integ[0] == F.INTEGRATOR_WHITTED:
        assert ost.rays_continuation > 0 and ost.rays_shadow > 0 and ost.rays_mis == 0
        # the keyframed kernel variant: moving lights / objects / camera, per-ray transforms
        ba = SB.scene_animated(48, 48, 4, frames=4, scene_time=1.0)
        ba.integrator = (F.INTEGRATOR_WHITTED, 0, 3)
        desc = ba.finish()
        ga, oa = api.Scene(desc), O.OracleScene(desc)
        ga.update_frame(2, 0.5, 0.75); oa.update_frame(2, 0.5, 0.75)
        assert ga.render_samples(seed=4)[0].tobytes() == oa.render_samples(seed=4)[0].tobytes()


def test_whitted_json_scene_through_the_loader(tmp_path):
    """scene.rs:305-309: {"type": "whitted", "min_depth": N} — the recursion limit is read from "min_depth"."""
    import json
    d = json.load(open(os.path.join(HERE, "golden", "scenes", "c2_smallpt.json")))
    d["integrator"] = {"type": "whitted", "min_depth": 3}
    p = tmp_path / "whitted.json"
    p.write_text(json.dumps(d))
    lib = F.load_trb()
    dp = C.POINTER(F.SceneDesc)()
    assert lib.trb_desc_load_json(str(p).encode(), 96, 96, 4, C.byref(dp)) == 0, lib.trb_last_error()
    try:
        desc = dp.contents
        assert (desc.integrator.type, desc.integrator.max_depth) == (F.INTEGRATOR_WHITTED, 3)
        g, o = both(desc)
        gs, gst = g.render_samples(seed=2, flags=F.RENDER_REFERENCE_SHADOW); os_, ost = o.render_samples(seed=2)
        assert gs.tobytes() == os_.tobytes() and [getattr(gst, k) for k in KEYS] == [getattr(ost, k) for k in KEYS]
        assert ost.rays_continuation > 0       # the metal and glass spheres recurse
    finally:
        lib.trb_desc_free(dp)


def test_update_frame_on_device_matches_the_host_path_and_the_oracle():
    """Scene::update_frame on the device (k_frame_instances + k_tlas_build: instance transforms at shutter-open,
    animation bounds over the shutter, the reference's SAH build with max_geom 4 and the traversal records) against the same
    work done by the host code (option "frame.device" = 0) and by the oracle, frame by frame of a keyframed scene."""
    desc = SB.scene_animated(48, 48, 4, frames=6, scene_time=1.5, animated_fov=True).finish()
    g, o = api.Scene(desc), O.OracleScene(desc)
    step = 1.5 / 6
    for fr in range(6):
        o.update_frame(fr, fr * step, (fr + 1) * step)
        on, oo = o.bvh(-1)
        per_path = {}
        for dev in (1, 0):
            g.set_option("frame.device", dev)
            g.update_frame(fr, fr * step, (fr + 1) * step)
            gn, go = g.bvh(-1)
            assert gn.tobytes() == on.tobytes() and np.array_equal(go, oo), (fr, dev)
            for i in range(desc.n_instances):
                gm, gi = g.transform(i); om, oi = o.transform(i)
                assert gm.tobytes() == om.tobytes() and gi.tobytes() == oi.tobytes(), (fr, dev, i)
            per_path[dev] = g.render_samples(seed=6)[0]
        assert per_path[0].tobytes() == per_path[1].tobytes() == o.render_samples(seed=6)[0].tobytes()
    g.set_option("frame.device", 1)
    # the per-path transform table against per-ray evaluation
    g.update_frame(2, 2 * step, 3 * step)
    a = g.render_samples(seed=8)[0]                   # default: each distinct keyframed spline once per path, then the stacks
    for mode in (1, 0):                               # one thread per (path, instance) evaluates the whole stack; per ray per instance
        g.set_option("anim.table", mode)
        assert g.render_samples(seed=8)[0].tobytes() == a.tobytes(), mode
