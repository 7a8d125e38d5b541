// trb_anim.h — AnimatedTransform::transform(time) (src/linalg/animated_transform.rs:40-56) for keyframed instances and
// cameras, shared by the host (TLAS animation bounds, src/linalg/animated_transform.rs:58-71) and the device (the
// reference evaluates it per ray per instance, receiver.rs:30 / emitter.rs:122,176,197 / camera.rs:156).
//   slerp            src/linalg/quaternion.rs:101-113 (acos / sin / cos through the detmath contract)
//   kf_interpolate   src/linalg/keyframe.rs:66-72 (bspline::Interpolate)
//   spline_point     bspline 0.2.2 BSpline::point -> de Boor (third-party crate, restated; "parity unpinned")
//   animated_xf      stack of splines: transform = t_i * transform
#pragma once
#include "trb_host.h"
#include "trb_detmath.cuh"

namespace trbh {

constexpr int kMaxSplineDegree = 5;

TRB_HD inline float clampf_hd(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

TRB_HD inline void quat_slerp(float t, const float a[4], const float b[4], float out[4]) {
    const float cos_theta = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; // quaternion::dot: dot(v) + w*w
    if (cos_theta > 0.9995f) {
        float q[4];
        for (int i = 0; i < 4; ++i) q[i] = (1.0f - t) * a[i] + t * b[i];
        const float l = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (int i = 0; i < 4; ++i) out[i] = q[i] / l;
        return;
    }
    const float theta = trb::dacos(clampf_hd(cos_theta, -1.0f, 1.0f));
    const float theta_t = theta * t;
    float perp[4];
    for (int i = 0; i < 4; ++i) perp[i] = b[i] - a[i] * cos_theta;
    const float l = sqrtf(perp[0] * perp[0] + perp[1] * perp[1] + perp[2] * perp[2] + perp[3] * perp[3]);
    float sn, cs;
    trb::dsincos(theta_t, sn, cs);
    for (int i = 0; i < 4; ++i) out[i] = a[i] * cs + (perp[i] / l) * sn;
}

TRB_HD inline trb_keyframe kf_interpolate(const trb_keyframe& a, const trb_keyframe& b, float t) {
    trb_keyframe k;
    for (int i = 0; i < 3; ++i) {
        k.translation[i] = (1.0f - t) * a.translation[i] + t * b.translation[i];
        k.scaling[i] = (1.0f - t) * a.scaling[i] + t * b.scaling[i];
    }
    quat_slerp(t, a.rotation, b.rotation, k.rotation);
    return k;
}

// BSpline::point(t) for t inside the knot domain [knots[degree], knots[n - 1 - degree]]
TRB_HD inline trb_keyframe spline_point(const trb_spline& sp, const trb_keyframe* kfs, const float* knots, float t) {
    const uint32_t n = sp.n_knots, deg = sp.degree;
    const float* kn = knots + sp.knot_first;
    uint32_t ub = n; // first index with knot > t
    for (uint32_t i = 0; i < n; ++i) if (kn[i] > t) { ub = i; break; }
    uint32_t i0;
    if (ub == n) i0 = n - deg - 1;
    else if (ub == 0) i0 = deg;
    else if (ub >= n - deg - 1) i0 = n - deg - 1;
    else i0 = ub;
    trb_keyframe tmp[kMaxSplineDegree + 1];
    for (uint32_t j = 0; j <= deg; ++j) tmp[j] = kfs[sp.ctrl_first + j + i0 - deg - 1];
    for (uint32_t lvl = 0; lvl < deg; ++lvl) {
        const uint32_t k = lvl + 1;
        for (uint32_t j = 0; j < deg - lvl; ++j) {
            const uint32_t i = j + k + i0 - deg;
            const float alpha = (t - kn[i - 1]) / (kn[i + deg - k] - kn[i - 1]);
            tmp[j] = kf_interpolate(tmp[j], tmp[j + 1], alpha);
        }
    }
    return tmp[0];
}

// BSpline<f32>::point (the camera's animated field of view, camera.rs:136-139): Interpolate for f32 is a * (1 - t) + b * t
inline float spline_point_f32(uint32_t deg, const float* ctrl, const float* kn, uint32_t n, float t) {
    uint32_t ub = n;
    for (uint32_t i = 0; i < n; ++i) if (kn[i] > t) { ub = i; break; }
    uint32_t i0;
    if (ub == n) i0 = n - deg - 1;
    else if (ub == 0) i0 = deg;
    else if (ub >= n - deg - 1) i0 = n - deg - 1;
    else i0 = ub;
    float tmp[kMaxSplineDegree + 1];
    for (uint32_t j = 0; j <= deg; ++j) tmp[j] = ctrl[j + i0 - deg - 1];
    for (uint32_t lvl = 0; lvl < deg; ++lvl) {
        const uint32_t k = lvl + 1;
        for (uint32_t j = 0; j < deg - lvl; ++j) {
            const uint32_t i = j + k + i0 - deg;
            const float alpha = (t - kn[i - 1]) / (kn[i + deg - k] - kn[i - 1]);
            tmp[j] = tmp[j] * (1.0f - alpha) + tmp[j + 1] * alpha;
        }
    }
    return tmp[0];
}

// `level_xf` (optional): Keyframe::transform of every one-control-point level, computed once per scene with the same
// operations (time-independent, so exact re-use): a keyframed stack usually mixes static and animated levels.
TRB_HD inline Xf animated_xf(const trb_spline* splines, uint32_t first, uint32_t count, const trb_keyframe* kfs, const float* knots, float time,
                             const Xf* level_xf = nullptr) {
    Xf acc = xf_identity();
    for (uint32_t s = first; s < first + count; ++s) {
        const trb_spline& sp = splines[s];
        Xf t;
        if (sp.n_ctrl == 1) t = level_xf ? level_xf[s] : keyframe_xf(kfs[sp.ctrl_first]);
        else {
            const float lo = knots[sp.knot_first + sp.degree], hi = knots[sp.knot_first + sp.n_knots - 1 - sp.degree];
            t = keyframe_xf(spline_point(sp, kfs, knots, clampf_hd(time, lo, hi)));
        }
        acc = xf_compose(t, acc);
    }
    return acc;
}

// AnimatedTransform::is_animated (animated_transform.rs:73-75): true only if EVERY stacked spline has > 1 control point (Q22)
TRB_HD inline bool xf_is_animated(const trb_spline* splines, uint32_t first, uint32_t count) {
    if (count == 0) return true;
    bool b = true;
    for (uint32_t s = first; s < first + count; ++s) b = b && splines[s].n_ctrl > 1;
    return b;
}
TRB_HD inline bool xf_is_static(const trb_spline* splines, uint32_t first, uint32_t count) {
    for (uint32_t s = first; s < first + count; ++s) if (splines[s].n_ctrl != 1) return false;
    return true;
}

// film::AnimatedColor::color (src/film/animated_color.rs:52-78), rgb only
TRB_HD inline void animated_color(const trb_color_key* keys, uint32_t first, uint32_t n, float time, float out[3]) {
    if (n == 0) { out[0] = out[1] = out[2] = 0.0f; return; }
    if (n == 1) { for (int i = 0; i < 3; ++i) out[i] = keys[first].rgba[i]; return; }
    // first = last key of the leading run with key.time < time (take_while().last()), second = first key with !(key.time < time)
    // (find()): the run ends exactly where the second begins, so one scan gives both
    int fi = -1;
    for (uint32_t k = 0; k < n; ++k) { if (keys[first + k].time < time) fi = (int)k; else break; }
    const int si = fi + 1 < (int)n ? fi + 1 : -1;
    if (fi < 0) { for (int i = 0; i < 3; ++i) out[i] = keys[first].rgba[i]; return; }
    if (si < 0) { for (int i = 0; i < 3; ++i) out[i] = keys[first + n - 1].rgba[i]; return; }
    const trb_color_key& a = keys[first + fi];
    const trb_color_key& b = keys[first + si];
    const float t = (time - a.time) / (b.time - a.time);
    for (int i = 0; i < 3; ++i) out[i] = a.rgba[i] * (1.0f - t) + b.rgba[i] * t; // linalg::lerp
}

} // namespace trbh
