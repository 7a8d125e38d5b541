"""Keyframed transforms / emission / camera on the CPU oracle.

The B-spline evaluation is the third-party `bspline 0.2.2` crate (Cargo.lock) — de Boor over Keyframe::interpolate
(keyframe.rs:66-72: lerp translation/scaling, Quaternion::slerp for rotation, quaternion.rs:101-113). The reference holds
no test vectors for it, so these are property / closed-form checks: parity for this row is pinned only between our two
implementations (oracle vs GPU, tests/test_gpu_parity.py::test_animated_scene_gpu_vs_oracle)."""
import math

import numpy as np
import pytest

from tray_rust_b200 import _ffi as F, api, scenebuild as SB
from oracle import pyoracle as O
from tray_rust_b200.scenebuild import Anim, trs, quat_axis_angle


def one_instance_scene(levels, emission=None, frames=8, scene_time=1.0, cam=None):
    b = SB.SceneBuilder(16, 16, 2, 2, 4)
    b.film.update(frames=frames, start_frame=0, end_frame=frames - 1, scene_time=scene_time)
    white = b.add_material(F.MAT_MATTE, (0.7, 0.7, 0.7), roughness=0.0)
    b.receiver(F.SHAPE_SPHERE, white, levels, p0=1.0)
    b.area_light(F.SHAPE_SPHERE, white, [trs(t=(0, 10, 0))], emission if emission is not None else (1, 1, 1, 10), p0=1.0)
    b.add_camera(cam if cam is not None else [trs(t=(0, 0, -20))], fov=40.0, shutter_size=1.0)
    return b


def xf_at(o, inst, time):
    o.update_frame(0, time, time)
    return o.transform(inst)


def test_degree1_translation_is_a_lerp_and_time_is_clamped_to_the_knot_domain():
    a, c = (1.0, 2.0, 3.0), (5.0, -2.0, 11.0)
    o = O.OracleScene(one_instance_scene([Anim([trs(t=a), trs(t=c)], degree=1)]).finish())
    for t in (0.0, 0.125, 0.5, 0.75, 1.0):
        m, inv = xf_at(o, 0, t)
        want = np.float32(1.0 - np.float32(t)) * np.asarray(a, np.float32) + np.float32(t) * np.asarray(c, np.float32)
        assert np.array_equal(m[:3, 3], want), t
        assert np.allclose(m @ inv, np.eye(4), atol=1e-6)
    assert np.array_equal(xf_at(o, 0, -3.0)[0], xf_at(o, 0, 0.0)[0])   # linalg::clamp(time, domain) (animated_transform.rs:50)
    assert np.array_equal(xf_at(o, 0, 7.0)[0], xf_at(o, 0, 1.0)[0])


def test_slerp_midpoint_and_unit_rotation():
    o = O.OracleScene(one_instance_scene([Anim([trs(q=quat_axis_angle((0, 1, 0), 0)), trs(q=quat_axis_angle((0, 1, 0), 90))], degree=1)]).finish())
    m, _ = xf_at(o, 0, 0.5)
    c = math.cos(math.radians(45)); s = math.sin(math.radians(45))
    assert np.allclose(m[:3, :3], [[c, 0, s], [0, 1, 0], [-s, 0, c]], atol=2e-6)
    for t in np.linspace(0, 1, 9):
        r = xf_at(o, 0, float(t))[0][:3, :3]
        assert np.allclose(r @ r.T, np.eye(3), atol=3e-6)
    # nearly parallel quaternions take the normalised-lerp branch (cos_theta > 0.9995, quaternion.rs:104-105)
    o2 = O.OracleScene(one_instance_scene([Anim([trs(q=quat_axis_angle((0, 0, 1), 10)), trs(q=quat_axis_angle((0, 0, 1), 10.5))], degree=1)]).finish())
    r = xf_at(o2, 0, 0.5)[0][:3, :3]
    a = math.radians(10.25)
    assert np.allclose(r, [[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]], atol=2e-6)


def test_bspline_partition_of_unity_endpoints_and_symmetry():
    k = trs(t=(3, -1, 2), q=quat_axis_angle((1, 2, 3), 40), s=(2, 1, 0.5))
    o = O.OracleScene(one_instance_scene([Anim([k] * 6, degree=3)]).finish())
    ref = xf_at(o, 0, 0.0)[0]
    for t in (0.1, 0.33, 0.5, 0.9, 1.0):
        assert np.allclose(xf_at(o, 0, t)[0], ref, atol=2e-6)
    pts = [(0, 0, 0), (1, 4, 0), (3, 5, 1), (6, 4, 2), (8, 0, 3)]
    o = O.OracleScene(one_instance_scene([Anim([trs(t=p) for p in pts], degree=3)]).finish())
    assert np.allclose(xf_at(o, 0, 0.0)[0][:3, 3], pts[0], atol=1e-6)   # clamped knot vector interpolates the end control points
    assert np.allclose(xf_at(o, 0, 1.0)[0][:3, 3], pts[-1], atol=1e-5)
    # closed form at the interior knot 0.5 of knots [0,0,0,0,.5,1,1,1,1]: basis (1/4, 1/2, 1/4) on control points 1..3
    want = 0.25 * np.asarray(pts[1]) + 0.5 * np.asarray(pts[2]) + 0.25 * np.asarray(pts[3])
    assert np.allclose(xf_at(o, 0, 0.5)[0][:3, 3], want, atol=1e-5)
    # the curve stays in the convex hull of its control points
    for t in np.linspace(0, 1, 33):
        p = xf_at(o, 0, float(t))[0][:3, 3]
        assert (p >= np.min(pts, axis=0) - 1e-5).all() and (p <= np.max(pts, axis=0) + 1e-5).all()


def test_stack_order_animated_level_below_a_static_level():
    # transform = t_last * ... * t_first (animated_transform.rs:42-54): spin first, then place
    spin = Anim([trs(q=quat_axis_angle((0, 1, 0), 0)), trs(q=quat_axis_angle((0, 1, 0), 90))], degree=1)
    o = O.OracleScene(one_instance_scene([spin, trs(t=(5, 0, 0), s=2.0)]).finish())
    m, _ = xf_at(o, 0, 1.0)
    assert np.allclose(m[:3, 3], (5, 0, 0), atol=1e-6)
    assert np.allclose(m[:3, :3], 2.0 * np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0]]), atol=3e-6)


def test_tlas_bounds_follow_animation_bounds_q22():
    fly = Anim([trs(t=(-8, 0, 0)), trs(t=(8, 0, 0))], degree=1)
    b = one_instance_scene([fly], frames=1)
    o = O.OracleScene(b.finish())
    o.update_frame(0, 0.0, 1.0)   # shutter_size 1: the box must cover the whole sweep (128 time samples, animated_transform.rs:62-68)
    nodes, _ = o.bvh(-1)
    assert nodes["bmin"][:, 0].min() <= -9.0 and nodes["bmax"][:, 0].max() >= 9.0
    # Q22: a stack holding one static level is "not animated" for bounds -> only the box of transform(start)
    o2 = O.OracleScene(one_instance_scene([fly, trs(s=1.0)], frames=1).finish())
    o2.update_frame(0, 0.0, 1.0)
    n2, _ = o2.bvh(-1)
    assert n2["bmin"][:, 0].min() == -9.0 and n2["bmax"][:, 0].max() <= 1.0 + 1e-5   # sphere at x=-8 (r=1) and the light at x in [-1, 1]
    assert np.allclose(xf_at(o2, 0, 1.0)[0][:3, 3], (8, 0, 0))                        # ... although the instance does move


def test_keyframed_emission_lerps_and_clamps():
    keys = [((1.0, 0.0, 0.0, 10), 0.25), ((0.0, 1.0, 0.0, 20), 0.75)]
    b = one_instance_scene([trs(t=(0, 0, 0), s=0.01)], emission=keys, cam=[trs(t=(0, 10, -20))])
    o = O.OracleScene(b.finish())

    def seen(time):
        o.update_frame(0, time, time)
        s, _ = o.render_samples(seed=3)
        lit = s[(s["r"] + s["g"]) > 0]
        return lit

    # looking straight at the light (first hit emission, path.rs:73): radiance is emission.color(time), clamped to 1 per channel
    early, late = seen(0.0), seen(1.0)
    assert len(early) and np.all(early["r"] == 1.0) and np.all(late["g"] == 1.0)
    assert np.all(early["g"] < 0.3) and np.all(late["r"] < 0.3)


def test_animated_scene_renders_and_frames_differ():
    d = SB.scene_animated(32, 32, 4).finish()
    o = O.OracleScene(d)
    means, cams = [], []
    for fr in range(4):
        o.update_frame(fr, fr * 0.25, (fr + 1) * 0.25)
        rays, _ = o.camera_rays(seed=5)
        s, st = o.render_samples(seed=5)
        assert np.isfinite(s["r"]).all() and s["r"].mean() > 0.01
        assert st.rays_primary == len(s) and st.rays_shadow > 0
        means.append(s["r"].mean()); cams.append(rays["o"].copy())
        # keyframed camera: origins vary inside one frame (cam_world.transform(frame_time) per ray, camera.rs:156)
        assert len(np.unique(rays["o"], axis=0)) > 1
    assert len({round(float(m), 6) for m in means}) == 4
    # determinism across thread counts
    a, _ = o.render_samples(seed=5, threads=1); b, _ = o.render_samples(seed=5, threads=4)
    assert a.tobytes() == b.tobytes()


def test_product_host_animated_transform_matches_oracle_bit_for_bit(trb):
    """trb_host_animated_transform is the code the device runs per ray (csrc/trb_anim.h) compiled for the host."""
    b = SB.scene_animated(32, 32, 4)
    d = b.finish()
    o = O.OracleScene(d)
    import ctypes as C
    for time in (0.0, 0.03, 0.25, 0.3333, 0.5, 0.77, 1.0, 1.5, -0.5):
        o.update_frame(0, time, time)
        for i in range(d.n_instances):
            inst = d.instances[i]
            m, inv = np.zeros(16, np.float32), np.zeros(16, np.float32)
            assert trb.trb_host_animated_transform(C.byref(d), inst.spline_first, inst.n_splines, time, F.ptr(m), F.ptr(inv)) == 0
            om, oi = o.transform(i)
            assert np.array_equal(api.bits(m), api.bits(om.reshape(-1))) and np.array_equal(api.bits(inv), api.bits(oi.reshape(-1))), (time, i)


def test_product_host_animated_color(trb):
    import ctypes as C
    keys = [((1.0, 0.0, 0.0, 10), 0.25), ((0.0, 1.0, 0.0, 20), 0.75), ((0.0, 0.0, 1.0, 5), 0.9)]
    d = one_instance_scene([trs()], emission=keys).finish()
    first = d.instances[1].emission_first
    assert d.instances[1].n_emission == 3

    def col(t):
        c = np.zeros(3, np.float32)
        assert trb.trb_host_animated_color(C.byref(d), first, 3, t, F.ptr(c)) == 0
        return c
    assert np.array_equal(col(0.0), (10, 0, 0)) and np.array_equal(col(0.25), (10, 0, 0))   # before the first key: first colour
    assert np.array_equal(col(2.0), (0, 0, 5))                                               # after the last key: last colour
    assert np.allclose(col(0.5), (5, 10, 0), atol=1e-5)                                      # lerp between the bracketing keys
    assert np.allclose(col(0.825), (0, 10, 2.5), atol=1e-4)


def _tr15_like_desc(trb, w=48, h=32, spp=2):
    import ctypes as C
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_scenes
    merl = os.path.join(here, "golden", "scenes", "merl", "synthetic.binary")
    if not os.path.exists(merl):
        make_scenes.write_synthetic_merl(merl)
    d = C.POINTER(F.SceneDesc)()
    assert trb.trb_desc_load_json(os.path.join(here, "golden", "scenes", "c5_tr15_like.json").encode(), w, h, spp, C.byref(d)) == 0, trb.trb_last_error()
    return d


def test_tr15_like_json_through_the_loader(trb):
    """scene.rs:832-850 load_keyframes: control_points / knots / default degree 3; emission keyframes (scene.rs:722-747);
    a group's keyframes stack above its children's (scene.rs:600-640)."""
    d = _tr15_like_desc(trb)
    desc = d.contents
    assert (desc.film.frames, desc.film.scene_time, desc.integrator.min_depth, desc.integrator.max_depth) == (50, 25.0, 5, 10)
    assert desc.n_cameras == 1 and desc.n_instances == 12 and desc.n_merl == 1 and desc.n_meshes == 1
    cam = desc.cameras[0]
    assert cam.n_splines == 1 and desc.splines[cam.spline_first].n_ctrl == 7 and desc.splines[cam.spline_first].degree == 3
    # walls: object transform (static) below the group's 4-point cubic with repeated knots
    w0 = desc.instances[0]
    assert w0.n_splines == 2
    s0, s1 = desc.splines[w0.spline_first], desc.splines[w0.spline_first + 1]
    assert (s0.n_ctrl, s1.n_ctrl, s1.degree, s1.n_knots) == (1, 4, 3, 8)
    assert [desc.knots[s1.knot_first + k] for k in range(8)] == [2.5] * 4 + [6.0] * 4
    lights = [desc.instances[i] for i in range(desc.n_instances) if desc.instances[i].kind != F.INST_RECEIVER]
    assert sorted(l.n_emission for l in lights) == [1, 2, 4]
    big = [l for l in lights if l.n_emission == 4][0]
    ck = [desc.color_keys[big.emission_first + k] for k in range(4)]
    assert [k.time for k in ck] == [0.0, 3.5, 5.0, 7.0] and ck[0].rgba[0] == 0.0 and abs(ck[2].rgba[0] - 110.0) < 1e-4  # rgb * 4th component
    o = O.OracleScene(desc)
    step = desc.film.scene_time / desc.film.frames
    means = []
    for fr in (0, 9, 14, 30):
        o.update_frame(fr, fr * step, (fr + 1) * step)
        s, st = o.render_samples(seed=3)
        assert np.isfinite(s["r"]).all()
        means.append(float(s["r"].mean()))
    assert means[0] > 0.003            # the static panel light
    assert means[2] > 1.5 * means[0]   # the keyed disk light has ramped up by t = 7
    # the walls sit still until t = 2.5, then sink (group keyframes): instance 0's transform
    o.update_frame(0, 0.0, 0.5); y0 = o.transform(0)[0][1, 3]
    o.update_frame(4, 2.0, 2.5); y1 = o.transform(0)[0][1, 3]
    o.update_frame(20, 10.0, 10.5); y2 = o.transform(0)[0][1, 3]
    assert y0 == y1 == 12.0 and y2 == 10.0
    trb.trb_desc_free(d)


def test_animated_fov_is_sampled_once_per_frame_at_the_clamped_midpoint():
    """camera.rs:134-141: fov = spline.point(clamp((start + end) / 2, knot domain)); scaling = tan(fov / 2)."""
    b = one_instance_scene([trs()], frames=10, scene_time=10.0)
    b.cameras = []
    b.add_camera([trs(t=(0, 0, -20))], fov=[20.0, 60.0], fov_knots=[2.0, 2.0, 6.0, 6.0], fov_degree=1, shutter_size=1.0)
    o = O.OracleScene(b.finish())

    def half_width(start, end):
        o.update_frame(0, start, end)
        rays, xy = o.camera_rays(seed=1)
        # the image-space x of a sample maps linearly to d.x / d.z = tan(fov/2) * aspect * ndc: recover tan(fov/2) from one sample
        k = int(np.argmax(xy[:, 0]))
        d = rays["d"][k]
        ndc = (xy[k, 0] / 16.0) * 2.0 - 1.0
        return abs(d[0] / d[2]) / abs(ndc)

    t20, t40, t60 = (math.tan(math.radians(f) / 2) for f in (20.0, 40.0, 60.0))
    assert abs(half_width(0.0, 1.0) - t20) < 1e-4      # mid 0.5 clamps to the domain start (2.0) -> 20 degrees
    assert abs(half_width(3.0, 5.0) - t40) < 1e-4      # mid 4.0 -> halfway -> 40 degrees
    assert abs(half_width(3.9, 4.1) - t40) < 1e-4      # only the midpoint matters, not the shutter interval
    assert abs(half_width(8.0, 9.0) - t60) < 1e-4      # clamps to the domain end


def test_invalid_splines_are_rejected_like_bspline_new(trb):
    """BSpline::new asserts knots.len() == control_points.len() + degree + 1 (bspline 0.2.2); the C ABI returns
    TRB_INVALID_ARG with that message instead of aborting; degrees above the device evaluator's cap are TRB_UNSUPPORTED."""
    import ctypes as C
    m, inv = np.zeros(16, np.float32), np.zeros(16, np.float32)

    def status(b):
        d = b.finish()
        rc = trb.trb_host_animated_transform(C.byref(d), 0, 1, 0.5, F.ptr(m), F.ptr(inv))
        return rc, (trb.trb_last_error() or b"").decode()

    ok = one_instance_scene([Anim([trs(), trs(t=(1, 0, 0))], degree=1)])
    assert status(ok)[0] == F.TRB_OK
    bad = one_instance_scene([Anim([trs(), trs(t=(1, 0, 0))], degree=1)])
    bad.splines[0] = (1, 2, bad.splines[0][2], 3, bad.splines[0][4])          # 3 knots for 2 control points of degree 1
    rc, msg = status(bad)
    assert rc == F.TRB_INVALID_ARG and "knots.len() != control_points.len() + degree + 1" in msg
    # BSpline::new's first check: control_points.len() > degree ("Too few control points for curve"). Two control points
    # of degree 3 with 6 knots satisfy the knot-count rule but would index before the control points in de Boor.
    few = one_instance_scene([Anim([trs(), trs(t=(1, 0, 0))], degree=1)])
    few.knots = few.knots[:few.splines[0][4]] + [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
    few.splines[0] = (3, 2, few.splines[0][2], 6, few.splines[0][4])
    rc, msg = status(few)
    assert rc == F.TRB_INVALID_ARG and "Too few control points" in msg
    # BSpline::new sorts its knots: an unsorted knot vector evaluates like the sorted one
    pts = [trs(t=(float(k), 0.5 * k, 0)) for k in range(5)]
    srt = one_instance_scene([Anim(pts, knots=[0, 0, 0, 0.3, 0.6, 1, 1, 1], degree=2)])
    uns = one_instance_scene([Anim(pts, knots=[1, 0, 0.6, 0, 1, 0.3, 0, 1], degree=2)])
    assert status(srt)[0] == F.TRB_OK
    m_sorted = m.copy()
    assert status(uns)[0] == F.TRB_OK and m.tobytes() == m_sorted.tobytes()
    nan = one_instance_scene([Anim(pts, knots=[0, 0, 0, float("nan"), 0.6, 1, 1, 1], degree=2)])
    assert status(nan)[0] == F.TRB_INVALID_ARG
    deep = one_instance_scene([Anim([trs(t=(k, 0, 0)) for k in range(8)], degree=6)])
    rc, msg = status(deep)
    assert rc == F.TRB_UNSUPPORTED and "degree" in msg
    # the same rules for the camera's field-of-view spline
    cam = one_instance_scene([trs()])
    cam.cameras = []
    cam.add_camera([trs(t=(0, 0, -20))], fov=[20.0, 60.0], fov_knots=[0.0, 0.0, 1.0, 1.0], fov_degree=1)
    assert status(cam)[0] == F.TRB_OK
    c = list(cam.cameras[0]); c[8] = 3; cam.cameras[0] = tuple(c)                # n_fov_knots 3 instead of 4
    rc, msg = status(cam)
    assert rc == F.TRB_INVALID_ARG and "knots.len()" in msg
    c[8] = 4; c[5] = 2; cam.cameras[0] = tuple(c)                                # fov degree 2 with 2 control points
    rc, msg = status(cam)
    assert rc == F.TRB_INVALID_ARG and "Too few control points" in msg
