"""Analytic pins for the oracle's shading code (the reference holds no vectors for it):
BSDF reciprocity/energy/pdf consistency, a closed-form direct-lighting integral, film filter properties,
detmath-vs-glibc neutrality."""
import ctypes as C
import math

import numpy as np

from tray_rust_b200 import _ffi as F, api, scenebuild as SB
from oracle import pyoracle as O

ALL = 31
NON_SPECULAR = 4 | 8 | 1 | 2


def mat(t, c0=(0, 0, 0), c1=(0, 0, 0), roughness=0.0, eta=1.0):
    m = F.Material()
    m.type = t; m.c0[:] = c0; m.c1[:] = c1; m.roughness = roughness; m.eta = eta; m.merl = 0
    return m


def probe(oracle, m, wo, wi, flags=ALL, u=(0.3, 0.6, 0.1), merl=None):
    out = np.zeros(12, np.float32)
    wo = np.asarray(wo, np.float32); wi = np.asarray(wi, np.float32); u = np.asarray(u, np.float32)
    oracle.orc_bsdf_probe(C.byref(m), F.ptr(merl) if merl is not None else None, F.ptr(wo), F.ptr(wi), flags, F.ptr(u), F.ptr(out))
    return dict(f=out[0:3], pdf=out[3], fs=out[4:7], wi=out[7:10], pdf_s=out[10], type=int(out[11]))


def unit(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v)).astype(np.float32)


def test_lambertian_is_albedo_over_pi(oracle):  # lambertian.rs:32-34
    m = mat(F.MAT_MATTE, (0.2, 0.5, 0.8), roughness=0.0)
    r = probe(oracle, m, unit((0.3, 0.1, 1)), unit((-0.5, 0.2, 0.7)))
    assert np.allclose(r["f"], np.array([0.2, 0.5, 0.8]) / math.pi, rtol=1e-6)
    assert abs(r["pdf"] - unit((-0.5, 0.2, 0.7))[2] / math.pi) < 1e-6
    assert np.all(probe(oracle, m, unit((0.3, 0.1, 1)), unit((0.1, 0.1, -1)))["f"] == 0)  # below the surface: no BRDF lobe


def test_reciprocity(oracle):
    rng = np.random.default_rng(0)
    mats = [mat(F.MAT_MATTE, (0.5, 0.5, 0.5), roughness=20.0), mat(F.MAT_PLASTIC, (0.5, 0.2, 0.2), (0.6, 0.6, 0.6), roughness=0.3),
            mat(F.MAT_METAL, (0.15, 0.11, 0.13), (4.8, 3.1, 2.1), roughness=0.2)]
    for m in mats:
        for _ in range(50):
            a = unit(np.abs(rng.normal(size=3)) + [0, 0, 0.1]); b = unit(rng.normal(size=3) * [1, 1, 0] + [0, 0, abs(rng.normal()) + 0.1])
            f1, f2 = probe(oracle, m, a, b)["f"], probe(oracle, m, b, a)["f"]
            assert np.allclose(f1, f2, rtol=2e-4, atol=1e-7)


def test_sampling_consistency_and_energy(oracle):
    """sample() must return exactly eval()/pdf() of its own direction, and E[f cos / pdf] <= 1 (+MC noise)."""
    rng = np.random.default_rng(1)
    mats = [mat(F.MAT_MATTE, (0.9, 0.9, 0.9), roughness=0.0), mat(F.MAT_MATTE, (0.9, 0.9, 0.9), roughness=30.0),
            mat(F.MAT_PLASTIC, (0.5, 0.5, 0.5), (0.4, 0.4, 0.4), roughness=0.2), mat(F.MAT_METAL, (0.2, 0.9, 1.1), (3.9, 2.4, 2.2), roughness=0.3),
            mat(F.MAT_ROUGH_GLASS, (1, 1, 1), (1, 1, 1), roughness=0.3, eta=1.5)]
    wo = unit((0.4, -0.2, 0.8))
    for m in mats:
        acc, n = np.zeros(3), 4000
        for _ in range(n):
            u = rng.uniform(0, 0.999999, 3)
            r = probe(oracle, m, wo, wo, ALL, u)
            if r["pdf_s"] <= 0:
                continue
            again = probe(oracle, m, wo, r["wi"], ALL, u)
            # 1-ulp slack: sample() evaluates a single lobe in shading space, eval()/pdf() re-project from world space
            assert np.allclose(again["f"], r["fs"], rtol=1e-4, atol=1e-7) and np.isclose(again["pdf"], r["pdf_s"], rtol=1e-4)
            acc += r["fs"].astype(np.float64) * abs(r["wi"][2]) / r["pdf_s"]
        # reflective models cannot create energy; the Walter BTDF carries the eta^2 radiance scaling, so it may
        bound = 2.5 if m.type == F.MAT_ROUGH_GLASS else 1.08
        assert (acc / n < bound).all(), (m.type, acc / n)
        assert (acc / n > 0.05).all()


def test_specular_glass(oracle):  # specular_reflection.rs / specular_transmission.rs / fresnel.rs
    m = mat(F.MAT_GLASS, (1, 1, 1), (1, 1, 1), eta=1.5)
    wo = unit((0, 0, 1))
    r = probe(oracle, m, wo, wo, ALL, (0.5, 0.5, 0.1))  # first lobe = reflection
    assert r["type"] == 16 | 1 and np.allclose(r["wi"], [0, 0, 1]) and r["pdf_s"] == 1.0
    assert np.allclose(r["fs"], 0.04, atol=1e-6)  # ((1.5-1)/(1.5+1))^2 at normal incidence
    t = probe(oracle, m, wo, wo, ALL, (0.5, 0.5, 0.9))  # second lobe = transmission
    assert t["type"] == 16 | 2 and np.allclose(t["wi"], [0, 0, -1]) and np.allclose(t["fs"], 0.96, atol=1e-6)
    assert np.all(probe(oracle, m, wo, unit((0.1, 0, 1)), NON_SPECULAR)["f"] == 0)  # eval of specular lobes is 0
    # total internal reflection from inside
    inside = unit((0.9, 0, -0.3))
    assert probe(oracle, m, inside, inside, 16 | 2, (0.5, 0.5, 0.5))["pdf_s"] == 0.0


def test_merl_lookup_indexing(oracle):  # bxdf/merl.rs:47-82
    table = np.zeros(F.MERL_TABLE_FLOATS, np.float32)
    table.reshape(90, 90, 180, 3)[0, 0, 0] = (1, 2, 3)       # theta_h = 0, theta_d = 0 -> wo == wi == n
    table.reshape(90, 90, 180, 3)[0, 45, :] = (4, 5, 6)      # wo, wi mirrored at 45 degrees about n
    m = mat(F.MAT_MERL)
    r = probe(oracle, m, (0, 0, 1), (0, 0, 1), ALL, merl=table)
    assert r["f"].tolist() == [1, 2, 3]
    a = unit((math.sin(math.pi / 4) * 1.0001, 0, math.cos(math.pi / 4)))
    b = a * np.array([-1, 1, 1], np.float32)
    assert probe(oracle, m, a, b, ALL, merl=table)["f"].tolist() == [4, 5, 6]


def direct_lighting_scene(spp):
    """Lambertian floor lit by a rectangle light, max_depth 0: the estimator is sample_one_light only."""
    b = SB.SceneBuilder(8, 8, spp, min_depth=0, max_depth=0)
    white = b.add_material(F.MAT_MATTE, (0.8, 0.8, 0.8), roughness=0.0)
    b.receiver(F.SHAPE_RECT, white, [SB.trs(q=SB.quat_axis_angle((1, 0, 0), -90), s=50)], p0=2, p1=2)  # floor y=0, normal +y
    b.area_light(F.SHAPE_RECT, white, [SB.trs(t=(0, 4, 0), q=SB.quat_axis_angle((1, 0, 0), 90))], (0.5, 0.5, 0.5), p0=2, p1=2)  # facing down
    b.add_camera([SB.trs(t=(0, 3, -6), q=SB.quat_axis_angle((1, 0, 0), 26.565))], fov=4.0)  # looks at the origin
    return b


def test_direct_lighting_matches_closed_form(oracle):
    """E = rho/pi * L * integral over the light of cos cos' / r^2 dA, evaluated numerically in float64."""
    spp = 256
    o = O.OracleScene(direct_lighting_scene(spp).finish())
    o.update_frame(0, 0.0, 0.0)
    s, st = o.render_samples(seed=5)
    rays, _ = o.camera_rays(seed=5)
    hits, _ = o.intersect(rays)
    assert (hits["inst"] == 0).all()
    p = rays["o"] + rays["d"] * hits["t"][:, None]
    assert np.abs(p[:, 1]).max() < 1e-3 and np.abs(p[:, [0, 2]]).max() < 0.7  # all samples near the origin on the floor
    g = (np.arange(400) + 0.5) / 400 * 2 - 1
    gx, gz = np.meshgrid(g, g)
    px, pz = float(p[:, 0].mean()), float(p[:, 2].mean())
    dx, dz, dy = gx - px, gz - pz, 4.0
    r2 = dx * dx + dz * dz + dy * dy
    integral = np.sum((dy / np.sqrt(r2)) ** 2 / r2) * (2.0 / 400) ** 2
    expect = 0.8 / math.pi * 0.5 * integral
    got = float(s["r"].mean())
    assert abs(got - expect) < 0.03 * expect, (got, expect)
    assert st.rays_continuation == 0 and st.rays_shadow == len(s)


def test_libm_choice_is_statistically_neutral(oracle, oracle_sys):
    """detmath and glibc builds of the oracle agree at the level of Monte-Carlo noise."""
    d = SB.scene_smallpt_like(16, 16, 64).finish()
    a, b = O.OracleScene(d, "det"), O.OracleScene(d, "sys")
    for o in (a, b):
        o.update_frame(0, 0.0, 0.0)
    sa, _ = a.render_samples(seed=2)
    sb, _ = b.render_samples(seed=2)
    assert np.array_equal(sa["x"], sb["x"])  # no transcendental before the first hit
    same = np.mean((sa["r"] == sb["r"]) & (sa["g"] == sb["g"]))
    close = np.mean(np.abs(sa["r"] - sb["r"]) < 1e-4)
    assert close > 0.9, (same, close)          # ulp-level differences, a few paths branch differently
    assert abs(sa["r"].mean() - sb["r"].mean()) < 0.01


def test_film_filter_properties(oracle):
    o = O.OracleScene(SB.scene_smallpt_like(16, 16, 1).finish())
    t = o.filter_table()
    assert t.shape == (16, 16) and np.allclose(t, t.T) and t[0, 0] > 0.75  # Mitchell b=c=1/3 at the centre ~ (8/9)^2
    assert (np.diff(t[0]) <= 1e-7).all() or t[0].min() < 0  # decreasing into the negative lobe
    assert t.min() < 0  # negative lobes exist: weights must be carried as RGBW, not normalised per sample
    # splatting a constant colour: rgb/weight recovers it exactly wherever weight != 0 (render_target.rs:198-203)
    b = SB.SceneBuilder(16, 16, 4, 0, 0)
    m = b.add_material(F.MAT_MATTE, (0, 0, 0), roughness=0.0)
    b.area_light(F.SHAPE_RECT, m, [SB.trs(t=(0, 0, 5), q=SB.quat_axis_angle((1, 0, 0), 180), s=100)], (0.25, 0.5, 0.75), p0=2, p1=2)
    b.add_camera([SB.trs()], fov=40)
    o = O.OracleScene(b.finish())
    film, st = o.render(seed=1)
    img = film[..., :3] / film[..., 3:]
    assert np.allclose(img, [0.25, 0.5, 0.75], atol=1e-5)
    assert abs(film[..., 3].sum() - 0) > 1  # weights accumulated
    srgb = o.to_srgb8(film)
    exp = [int((1.055 * c ** (1 / 2.4) - 0.055) * 255.0) for c in (0.25, 0.5, 0.75)]
    assert np.abs(srgb.reshape(-1, 3).astype(int) - exp).max() <= 1


# ---- the reference's other integrators -------------------------------------------------

def test_normals_debug_is_half_the_shading_normal_plus_half(oracle):
    """NormalsDebug::illumination = (bsdf.n + 1) / 2 (integrator/normals_debug.rs:28-36): a sphere seen head-on."""
    b = SB.SceneBuilder(8, 8, 4)
    b.integrator = (F.INTEGRATOR_NORMALS_DEBUG, 0, 0)
    white = b.add_material(F.MAT_MATTE, (0.8, 0.8, 0.8), roughness=0.0)
    b.receiver(F.SHAPE_SPHERE, white, [SB.trs(s=2.0)], p0=1.0)
    b.point_light([SB.trs(t=(0, 10, 0))], (1, 1, 1, 10))
    b.add_camera([SB.trs(t=(0, 0, -10))], fov=1.0)      # a needle of rays at the sphere's near pole: n = (0, 0, -1)
    o = O.OracleScene(b.finish())
    o.update_frame(0, 0.0, 0.0)
    s, st = o.render_samples(seed=2)
    assert np.allclose(s["r"], 0.5, atol=0.02) and np.allclose(s["g"], 0.5, atol=0.02) and np.allclose(s["b"], 0.0, atol=1e-3) and abs(float(s["r"].mean()) - 0.5) < 2e-3
    assert st.rays_shadow == 0 and st.rays_continuation == 0 and st.rays_primary == len(s)


def whitted_floor_scene(mat, depth, spp=16, light_height=4.0):
    b = SB.SceneBuilder(8, 8, spp)
    b.integrator = (F.INTEGRATOR_WHITTED, 0, depth)
    m = b.add_material(*mat[0], **mat[1])
    b.receiver(F.SHAPE_RECT, m, [SB.trs(q=SB.quat_axis_angle((1, 0, 0), -90), s=50)], p0=2, p1=2)        # floor y = 0, normal +y
    b.point_light([SB.trs(t=(0, light_height, 0))], (1, 1, 1, 8))
    b.add_camera([SB.trs(t=(0, 3, -6), q=SB.quat_axis_angle((1, 0, 0), 26.565))], fov=2.0)             # looks at the origin
    return b


def test_whitted_direct_term_matches_closed_form(oracle):
    """Whitted on a Lambertian floor under a point light: L = rho/pi * I / d^2 * cos(theta) exactly (whitted.rs:56-62, emitter.rs:169-174),
    one shadow ray per light and camera sample, no specular recursion."""
    o = O.OracleScene(whitted_floor_scene(((F.MAT_MATTE, (0.6, 0.6, 0.6)), dict(roughness=0.0)), 5).finish())
    o.update_frame(0, 0.0, 0.0)
    s, st = o.render_samples(seed=4)
    rays, _ = o.camera_rays(seed=4)
    hits, _ = o.intersect(rays)
    p = rays["o"] + rays["d"] * hits["t"][:, None]
    d2 = p[:, 0] ** 2 + (4.0 - p[:, 1]) ** 2 + p[:, 2] ** 2
    expect = 0.6 / math.pi * 8.0 / d2 * (4.0 / np.sqrt(d2))
    assert np.allclose(s["r"], expect, rtol=2e-4)
    assert st.rays_shadow == len(s) and st.rays_continuation == 0


def test_whitted_mirror_recursion_and_depth_limit(oracle):
    """A specular-metal floor under a point light and a Lambertian ceiling: the reflection ray (integrator/mod.rs:41-71) carries the
    ceiling's direct light back, weighted by the Fresnel reflectance; max_depth 0 cuts the recursion (whitted.rs:63-66)."""
    def build(depth):
        b = whitted_floor_scene(((F.MAT_SPECULAR_METAL, (0.2, 0.9, 1.1), (3.9, 2.4, 2.2)), {}), depth)
        grey = b.add_material(F.MAT_MATTE, (0.5, 0.5, 0.5), roughness=0.0)
        b.receiver(F.SHAPE_RECT, grey, [SB.trs(t=(0, 9, 0), q=SB.quat_axis_angle((1, 0, 0), 90), s=50)], p0=2, p1=2)   # ceiling y = 9, normal -y
        return b
    o0, o3 = O.OracleScene(build(0).finish()), O.OracleScene(build(3).finish())
    o0.update_frame(0, 0.0, 0.0); o3.update_frame(0, 0.0, 0.0)
    s0, st0 = o0.render_samples(seed=4)
    s3, st3 = o3.render_samples(seed=4)
    assert st0.rays_continuation == 0 and np.all(s0["r"] == 0.0)                  # a mirror has no non-specular response to the light sample
    assert st3.rays_continuation >= len(s3)                                        # at least the first reflection ray per sample
    assert (s3["r"] > 0).all() and (s3["r"] < 0.5 / math.pi * 8.0 / 25.0 * 1.01).all()   # <= the ceiling's own direct light (rho/pi * I / d^2, d >= 5)


def test_whitted_glass_branches_into_reflection_and_transmission(oracle):
    """Glass has both specular lobes: every visit spawns a reflection AND a transmission ray (a binary recursion tree)."""
    b = SB.SceneBuilder(8, 8, 4)
    b.integrator = (F.INTEGRATOR_WHITTED, 0, 2)
    glass = b.add_material(F.MAT_GLASS, (1, 1, 1), (1, 1, 1), eta=1.5)
    grey = b.add_material(F.MAT_MATTE, (0.5, 0.5, 0.5), roughness=0.0)
    b.receiver(F.SHAPE_SPHERE, glass, [SB.trs(s=2.0)], p0=1.0)
    b.receiver(F.SHAPE_SPHERE, grey, [SB.trs(s=60.0)], p0=1.0)                    # an enclosing diffuse shell: every ray lands somewhere
    b.point_light([SB.trs(t=(0, 10, -10))], (1, 1, 1, 200))
    b.add_camera([SB.trs(t=(0, 0, -10))], fov=8.0)
    o = O.OracleScene(b.finish())
    o.update_frame(0, 0.0, 0.0)
    s, st = o.render_samples(seed=3)
    n = len(s)
    assert st.rays_continuation > 3 * n                                             # depth 0: 2 rays, depth 1: up to 4 more
    assert st.rays_continuation <= (2 + 4) * n and np.isfinite(s["r"]).all() and (s["r"] > 0).mean() > 0.9
