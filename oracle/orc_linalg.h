/* oracle/orc_linalg.h — TEST INFRASTRUCTURE (parity oracle), not product code.
 * CPU restatement of /root/reference/src/linalg/* and src/geometry/bbox.rs.
 * Expression order follows the Rust source (left-associative, no FMA contraction). */
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include <algorithm>
#include "detmath.h"

#ifdef ORC_SYSTEM_LIBM
#define M_SIN(x) sinf(x)
#define M_COS(x) cosf(x)
#define M_ACOS(x) acosf(x)
#define M_ATAN2(y, x) atan2f(y, x)
#define M_EXP(x) expf(x)
#define M_LOG(x) logf(x)
#define M_POW(x, y) powf(x, y)
#else
#define M_SIN(x) dm_sinf(x)
#define M_COS(x) dm_cosf(x)
#define M_ACOS(x) dm_acosf(x)
#define M_ATAN2(y, x) dm_atan2f(y, x)
#define M_EXP(x) dm_expf(x)
#define M_LOG(x) dm_logf(x)
#define M_POW(x, y) dm_powf(x, y)
#endif

namespace orc {

static const float PI = 3.14159265358979323846f;      /* f32::consts::PI */
static const float FRAC_1_PI = 0.318309886183790671f; /* f32::consts::FRAC_1_PI */
static const float FRAC_PI_4 = 0.785398163397448309f; /* f32::consts::FRAC_PI_4 */
static const float F32_EPSILON = 1.1920929e-7f;       /* f32::EPSILON */
static const float F32_INF = INFINITY;

/* Rust `f as usize` (saturating, NaN -> 0) */
static inline uint32_t f2u(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}
/* linalg::clamp (src/linalg/mod.rs:51-53) */
static inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
/* linalg::lerp (mod.rs:47-49): a*(1-t) + b*t */
static inline float lerpf(float t, float a, float b) { return a * (1.0f - t) + b * t; }
/* linalg::to_radians (mod.rs:34-36) */
static inline float to_radians(float d) { return PI / 180.0f * d; }

/* Vector / Point / Normal share arithmetic (vector.rs, point.rs, normal.rs) */
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit V3(float a) : x(a), y(a), z(a) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
/* linalg::dot / cross (mod.rs:38-45) */
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length_sqr(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline float length(V3 a) { return sqrtf(length_sqr(a)); }
/* Vector::normalized (vector.rs:32-35): three divides */
static inline V3 normalized(V3 a) { float l = length(a); return V3(a.x / l, a.y / l, a.z / l); }
static inline float distance_sqr(V3 a, V3 b) { return length_sqr(a - b); }

/* linalg::coordinate_system (mod.rs:96-108) */
static inline void coordinate_system(V3 e1, V3& e2, V3& e3) {
    if (fabsf(e1.x) > fabsf(e1.y)) {
        float inv_len = 1.0f / sqrtf(e1.x * e1.x + e1.z * e1.z);
        e2 = V3(-e1.z * inv_len, 0.0f, e1.x * inv_len);
    } else {
        float inv_len = 1.0f / sqrtf(e1.y * e1.y + e1.z * e1.z);
        e2 = V3(0.0f, e1.z * inv_len, -e1.y * inv_len);
    }
    e3 = cross(e1, e2);
}
/* linalg::reflect (mod.rs:110-112) */
static inline V3 reflect(V3 w, V3 v) { return 2.0f * dot(w, v) * v - w; }
/* linalg::refract (mod.rs:117-127); powf(x, 2.0) == x*x exactly */
static inline bool refract(V3 w, V3 n, float eta, V3& out) {
    float cos_t1 = dot(n, w);
    float sin_t1_sqr = fmaxf(0.0f, 1.0f - cos_t1 * cos_t1);
    float sin_t2_sqr = eta * eta * sin_t1_sqr;
    if (sin_t2_sqr >= 1.0f) return false;
    float cos_t2 = sqrtf(1.0f - sin_t2_sqr);
    out = eta * -w + (eta * cos_t1 - cos_t2) * n;
    return true;
}
/* linalg::solve_quadratic (mod.rs:78-94) */
static inline bool solve_quadratic(float a, float b, float c, float& t0, float& t1) {
    float discrim_sqr = b * b - 4.0f * a * c;
    if (discrim_sqr < 0.0f) return false;
    float discrim = sqrtf(discrim_sqr);
    float q = b < 0.0f ? -0.5f * (b - discrim) : -0.5f * (b + discrim);
    float x = q / a, y = c / q;
    if (x > y) { t0 = y; t1 = x; } else { t0 = x; t1 = y; }
    return true;
}
static inline V3 spherical_dir(float sin_theta, float cos_theta, float phi) {
    return V3(sin_theta * M_COS(phi), sin_theta * M_SIN(phi), cos_theta);
}
static inline float spherical_theta(V3 v) { return M_ACOS(clampf(v.z, -1.0f, 1.0f)); }
static inline float spherical_phi(V3 v) {
    float x = M_ATAN2(v.y, v.x);
    return x < 0.0f ? x + PI * 2.0f : x;
}

/* linalg::Ray (ray.rs) */
struct Ray {
    V3 o, d;
    float min_t, max_t;
    float time;
    Ray() : min_t(0), max_t(F32_INF), time(0) {}
    Ray(V3 o_, V3 d_, float t) : o(o_), d(d_), min_t(0.0f), max_t(F32_INF), time(t) {}
    static Ray segment(V3 o, V3 d, float mn, float mx, float t) { Ray r(o, d, t); r.min_t = mn; r.max_t = mx; return r; }
    V3 at(float t) const { return o + d * t; }
};

/* linalg::Matrix4 (matrix4.rs), row-major */
struct M4 {
    float m[16];
    static M4 zero() { M4 r; for (int i = 0; i < 16; ++i) r.m[i] = 0.0f; return r; }
    static M4 identity() { M4 r = zero(); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
    float at(int i, int j) const { return m[4 * i + j]; }
    float& at(int i, int j) { return m[4 * i + j]; }
    M4 transpose() const { M4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.at(i, j) = at(j, i); return r; }
    /* matrix4.rs:232-247 */
    M4 operator*(const M4& r) const {
        M4 o;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                o.at(i, j) = at(i, 0) * r.at(0, j) + at(i, 1) * r.at(1, j) + at(i, 2) * r.at(2, j) + at(i, 3) * r.at(3, j);
        return o;
    }
    /* matrix4.rs:48-172 (MESA gluInvertMatrix), term order as written there */
    M4 inverse() const {
        const float* a = m;
        M4 inv;
        float* v = inv.m;
        v[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
        v[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
        v[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
        v[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
        v[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
        v[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
        v[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
        v[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
        v[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
        v[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
        v[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
        v[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
        v[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
        v[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
        v[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
        v[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
        float det = a[0] * v[0] + a[1] * v[4] + a[2] * v[8] + a[3] * v[12];
        det = 1.0f / det;
        for (int i = 0; i < 16; ++i) v[i] *= det;
        return inv;
    }
};

struct BBox;

/* linalg::Transform (transform.rs) */
struct Transform {
    M4 mat, inv;
    static Transform identity() { Transform t; t.mat = M4::identity(); t.inv = M4::identity(); return t; }
    static Transform from_mat(const M4& m) { Transform t; t.mat = m; t.inv = m.inverse(); return t; }
    static Transform translate(V3 v) {
        Transform t = identity();
        t.mat.at(0, 3) = v.x; t.mat.at(1, 3) = v.y; t.mat.at(2, 3) = v.z;
        t.inv.at(0, 3) = -v.x; t.inv.at(1, 3) = -v.y; t.inv.at(2, 3) = -v.z;
        return t;
    }
    static Transform scale(V3 v) {
        Transform t = identity();
        t.mat.at(0, 0) = v.x; t.mat.at(1, 1) = v.y; t.mat.at(2, 2) = v.z;
        t.inv.at(0, 0) = 1.0f / v.x; t.inv.at(1, 1) = 1.0f / v.y; t.inv.at(2, 2) = 1.0f / v.z;
        return t;
    }
    Transform inverse() const { Transform t; t.mat = inv; t.inv = mat; return t; }
    /* transform.rs:191-197 */
    Transform operator*(const Transform& r) const { Transform t; t.mat = mat * r.mat; t.inv = r.inv * inv; return t; }
    /* transform.rs:199-216 (note: divides by w only when |w-1| < EPSILON, sic) */
    static V3 mul_point(const M4& a, V3 p) {
        V3 res;
        for (int i = 0; i < 3; ++i) res[i] = a.at(i, 0) * p.x + a.at(i, 1) * p.y + a.at(i, 2) * p.z + a.at(i, 3);
        float w = a.at(3, 0) * p.x + a.at(3, 1) * p.y + a.at(3, 2) * p.z + a.at(3, 3);
        if (fabsf(w - 1.0f) < F32_EPSILON) return res / w;
        return res;
    }
    static V3 mul_vector(const M4& a, V3 v) {
        V3 res;
        for (int i = 0; i < 3; ++i) res[i] = a.at(i, 0) * v.x + a.at(i, 1) * v.y + a.at(i, 2) * v.z;
        return res;
    }
    /* normals: transpose of the other matrix (transform.rs:231-242, 171-180) */
    static V3 mul_normal_T(const M4& a, V3 n) {
        V3 res;
        for (int i = 0; i < 3; ++i) res[i] = a.at(0, i) * n.x + a.at(1, i) * n.y + a.at(2, i) * n.z;
        return res;
    }
    V3 point(V3 p) const { return mul_point(mat, p); }
    V3 vector(V3 v) const { return mul_vector(mat, v); }
    V3 normal(V3 n) const { return mul_normal_T(inv, n); }
    V3 inv_point(V3 p) const { return mul_point(inv, p); }   /* transform.rs:150-162 */
    V3 inv_vector(V3 v) const { return mul_vector(inv, v); } /* transform.rs:164-171 */
    Ray ray(const Ray& r) const { Ray o = r; o.o = point(r.o); o.d = vector(r.d); return o; }
    Ray inv_ray(const Ray& r) const { Ray o = r; o.o = inv_point(r.o); o.d = inv_vector(r.d); return o; } /* :183-188 */
};

/* linalg::Quaternion (quaternion.rs) */
struct Quat {
    V3 v; float w;
    /* quaternion.rs:65-84: the literal array then .transpose() */
    M4 to_matrix() const {
        M4 a = M4::zero();
        float* m = a.m;
        m[0] = 1.0f - 2.0f * (v.y * v.y + v.z * v.z);
        m[1] = 2.0f * (v.x * v.y + v.z * w);
        m[2] = 2.0f * (v.x * v.z - v.y * w);
        m[4] = 2.0f * (v.x * v.y - v.z * w);
        m[5] = 1.0f - 2.0f * (v.x * v.x + v.z * v.z);
        m[6] = 2.0f * (v.y * v.z + v.x * w);
        m[8] = 2.0f * (v.x * v.z + v.y * w);
        m[9] = 2.0f * (v.y * v.z - v.x * w);
        m[10] = 1.0f - 2.0f * (v.x * v.x + v.y * v.y);
        m[15] = 1.0f;
        return a.transpose();
    }
};
static inline float qdot(const Quat& a, const Quat& b) { return dot(a.v, b.v) + a.w * b.w; }
static inline Quat qadd(const Quat& a, const Quat& b) { return Quat{a.v + b.v, a.w + b.w}; }
static inline Quat qsub(const Quat& a, const Quat& b) { return Quat{a.v - b.v, a.w - b.w}; }
static inline Quat qmul(const Quat& a, float s) { return Quat{a.v * s, a.w * s}; }
static inline Quat qnorm(const Quat& a) { float l = sqrtf(qdot(a, a)); return Quat{a.v / l, a.w / l}; }
/* quaternion.rs:101-113 */
static inline Quat slerp(float t, const Quat& a, const Quat& b) {
    float cos_theta = qdot(a, b);
    if (cos_theta > 0.9995f) return qnorm(qadd(qmul(a, 1.0f - t), qmul(b, t)));
    float theta = M_ACOS(clampf(cos_theta, -1.0f, 1.0f));
    float theta_t = theta * t;
    Quat q_perp = qnorm(qsub(b, qmul(a, cos_theta)));
    return qadd(qmul(a, M_COS(theta_t)), qmul(q_perp, M_SIN(theta_t)));
}

/* linalg::Keyframe (keyframe.rs) */
struct Keyframe {
    V3 translation; Quat rotation; V3 scaling;
    /* keyframe.rs:60-63 */
    Transform transform() const {
        M4 m = rotation.to_matrix();
        return Transform::translate(translation) * Transform::from_mat(m) * Transform::scale(scaling);
    }
    /* bspline::Interpolate (keyframe.rs:66-72) */
    Keyframe interpolate(const Keyframe& o, float t) const {
        Keyframe k;
        k.translation = (1.0f - t) * translation + t * o.translation;
        k.rotation = slerp(t, rotation, o.rotation);
        k.scaling = (1.0f - t) * scaling + t * o.scaling;
        return k;
    }
};

/* bspline 0.2.2 BSpline<Keyframe> (third-party crate, pinned in Cargo.lock; restated
 * from its published de Boor algorithm — "parity unpinned") */
struct Spline {
    uint32_t degree;
    std::vector<Keyframe> ctrl;
    std::vector<float> knots;
    void knot_domain(float& lo, float& hi) const { lo = knots[degree]; hi = knots[knots.size() - 1 - degree]; }
    Keyframe point(float t) const {
        size_t n = knots.size();
        size_t ub = n; /* first index with knot > t */
        for (size_t i = 0; i < n; ++i) if (knots[i] > t) { ub = i; break; }
        size_t i0;
        if (ub == n) i0 = n - degree - 1;
        else if (ub == 0) i0 = degree;
        else if (ub >= n - degree - 1) i0 = n - degree - 1;
        else i0 = ub;
        std::vector<Keyframe> tmp(degree + 1);
        for (size_t j = 0; j <= degree; ++j) tmp[j] = ctrl[j + i0 - degree - 1];
        for (size_t lvl = 0; lvl < degree; ++lvl) {
            size_t k = lvl + 1;
            for (size_t j = 0; j < degree - lvl; ++j) {
                size_t i = j + k + i0 - degree;
                float alpha = (t - knots[i - 1]) / (knots[i + degree - k] - knots[i - 1]);
                tmp[j] = tmp[j].interpolate(tmp[j + 1], alpha);
            }
        }
        return tmp[0];
    }
};

/* geometry::BBox (bbox.rs) */
struct BBox {
    V3 min, max;
    BBox() : min(V3(F32_INF)), max(V3(-F32_INF)) {}
    BBox(V3 a, V3 b) : min(a), max(b) {}
    BBox box_union(const BBox& b) const {
        return BBox(V3(fminf(min.x, b.min.x), fminf(min.y, b.min.y), fminf(min.z, b.min.z)),
                    V3(fmaxf(max.x, b.max.x), fmaxf(max.y, b.max.y), fmaxf(max.z, b.max.z)));
    }
    BBox point_union(V3 p) const {
        return BBox(V3(fminf(min.x, p.x), fminf(min.y, p.y), fminf(min.z, p.z)),
                    V3(fmaxf(max.x, p.x), fmaxf(max.y, p.y), fmaxf(max.z, p.z)));
    }
    /* bbox.rs:47-56 */
    int max_extent() const {
        V3 d = max - min;
        if (d.x > d.y && d.x > d.z) return 0;
        if (d.y > d.z) return 1;
        return 2;
    }
    V3 lerp(float tx, float ty, float tz) const { return V3(lerpf(tx, min.x, max.x), lerpf(ty, min.y, max.y), lerpf(tz, min.z, max.z)); }
    float surface_area() const { V3 d = max - min; return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z); }
    const V3& operator[](int i) const { return i == 0 ? min : max; }
    /* bbox.rs:75-104, compares and conditional assignments transcribed literally (NaN behaviour) */
    bool fast_intersect(const Ray& r, V3 inv_dir, const int neg_dir[3]) const {
        float tmin = ((*this)[neg_dir[0]].x - r.o.x) * inv_dir.x;
        float tmax = ((*this)[1 - neg_dir[0]].x - r.o.x) * inv_dir.x;
        float tymin = ((*this)[neg_dir[1]].y - r.o.y) * inv_dir.y;
        float tymax = ((*this)[1 - neg_dir[1]].y - r.o.y) * inv_dir.y;
        if (tmin > tymax || tymin > tmax) return false;
        if (tymin > tmin) tmin = tymin;
        if (tymax < tmax) tmax = tymax;
        float tzmin = ((*this)[neg_dir[2]].z - r.o.z) * inv_dir.z;
        float tzmax = ((*this)[1 - neg_dir[2]].z - r.o.z) * inv_dir.z;
        if (tmin > tzmax || tzmin > tmax) return false;
        if (tzmin > tmin) tmin = tzmin;
        if (tzmax < tmax) tmax = tzmax;
        return tmin < r.max_t && tmax > r.min_t;
    }
};
/* Transform * BBox, Arvo 1990 (transform.rs:256-281) */
static inline BBox transform_bbox(const Transform& t, const BBox& b) {
    BBox out;
    for (int i = 0; i < 3; ++i) { out.min[i] = t.mat.at(i, 3); out.max[i] = t.mat.at(i, 3); }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float x = t.mat.at(i, j) * b.min[j];
            float y = t.mat.at(i, j) * b.max[j];
            if (x < y) { out.min[i] += x; out.max[i] += y; }
            else { out.min[i] += y; out.max[i] += x; }
        }
    return out;
}

/* linalg::AnimatedTransform (animated_transform.rs) */
struct AnimatedTransform {
    std::vector<Spline> keyframes;
    /* animated_transform.rs:40-56 */
    Transform transform(float time) const {
        Transform tr = Transform::identity();
        for (const Spline& s : keyframes) {
            Transform t;
            if (s.ctrl.size() == 1) t = s.ctrl[0].transform();
            else {
                float lo, hi;
                s.knot_domain(lo, hi);
                t = s.point(clampf(time, lo, hi)).transform();
            }
            tr = t * tr;
        }
        return tr;
    }
    /* animated_transform.rs:73-75 */
    bool is_animated() const {
        if (keyframes.empty()) return true;
        bool b = true;
        for (const Spline& s : keyframes) b = b && s.ctrl.size() > 1;
        return b;
    }
    /* animated_transform.rs:58-71 */
    BBox animation_bounds(const BBox& b, float start, float end) const {
        if (!is_animated()) return transform_bbox(transform(start), b);
        BBox ret;
        for (int i = 0; i < 128; ++i) {
            float time = lerpf((float)i / 127.0f, start, end);
            ret = ret.box_union(transform_bbox(transform(time), b));
        }
        return ret;
    }
};

} // namespace orc
