/* oracle/orc_geom.h — TEST INFRASTRUCTURE (parity oracle), not product code.
 * CPU restatement of /root/reference/src/geometry/{bvh,mesh,sphere,disk,rectangle,
 * receiver,emitter,instance,differential_geometry}.rs and src/partition.rs. */
#pragma once
#include "orc_linalg.h"
#include <functional>

namespace orc {

struct Counters {
    uint64_t node_tests = 0, tri_tests = 0, inst_tests = 0;
    uint64_t rays[4] = {0, 0, 0, 0}; /* primary, shadow, mis, continuation */
    uint64_t camera_samples = 0;
    void add(const Counters& o) {
        node_tests += o.node_tests; tri_tests += o.tri_tests; inst_tests += o.inst_tests;
        for (int i = 0; i < 4; ++i) rays[i] += o.rays[i];
        camera_samples += o.camera_samples;
    }
};

/* ---- geometry::BVH (bvh.rs) ------------------------------------------------------ */
struct FlatNode {
    BBox bounds;
    bool leaf;
    uint32_t a; /* interior: second_child; leaf: geom_offset */
    uint32_t b; /* interior: axis;         leaf: ngeom       */
};

struct BVH {
    std::vector<uint32_t> ordered_geom;
    std::vector<FlatNode> tree;
    uint32_t max_geom = 0;

    struct GeomInfo { uint32_t geom_idx; V3 center; BBox bounds; };
    struct BuildNode {
        BBox bounds; bool leaf; int axis; uint32_t ngeom, geom_offset;
        BuildNode* c[2] = {nullptr, nullptr};
        ~BuildNode() { delete c[0]; delete c[1]; }
    };

    /* partition.rs:9-38 (two-ended, unstable) */
    template <class Pred>
    static size_t partition(GeomInfo* a, size_t n, Pred pred) {
        size_t split_idx = 0, lo = 0, hi = n;
        for (;;) {
            long front = -1, back = -1;
            while (lo < hi) { size_t f = lo++; if (!pred(a[f])) { front = (long)f; break; } else split_idx++; }
            while (lo < hi) { size_t b = --hi; if (pred(a[b])) { back = (long)b; break; } }
            if (front >= 0 && back >= 0) { std::swap(a[front], a[back]); split_idx++; }
            else break;
        }
        return split_idx;
    }

    /* bvh.rs:29-59 / 61-78 */
    void build(uint32_t max_geom_, const std::vector<BBox>& geom_bounds) {
        max_geom = max_geom_;
        tree.clear(); ordered_geom.clear();
        std::vector<GeomInfo> info(geom_bounds.size());
        for (size_t i = 0; i < geom_bounds.size(); ++i) {
            info[i].geom_idx = (uint32_t)i;
            info[i].bounds = geom_bounds[i];
            info[i].center = geom_bounds[i].lerp(0.5f, 0.5f, 0.5f); /* bvh.rs:321 */
        }
        size_t total = 0;
        BuildNode* root = build_rec(info.data(), info.size(), total);
        tree.reserve(total);
        flatten(root);
        delete root;
    }

    BuildNode* make_leaf(GeomInfo* g, size_t n, const BBox& bounds) {
        BuildNode* b = new BuildNode;
        b->bounds = bounds; b->leaf = true; b->axis = 0; b->ngeom = (uint32_t)n; b->geom_offset = (uint32_t)ordered_geom.size();
        for (size_t i = 0; i < n; ++i) ordered_geom.push_back(g[i].geom_idx);
        return b;
    }
    BuildNode* make_interior(BuildNode* l, BuildNode* r, int axis) {
        BuildNode* b = new BuildNode;
        b->bounds = l->bounds.box_union(r->bounds); /* bvh.rs:358-362 */
        b->leaf = false; b->axis = axis; b->c[0] = l; b->c[1] = r; b->ngeom = 0; b->geom_offset = 0;
        return b;
    }
    /* bvh.rs:139-232 */
    BuildNode* build_rec(GeomInfo* g, size_t ngeom, size_t& total) {
        total += 1;
        BBox bounds;
        for (size_t i = 0; i < ngeom; ++i) bounds = bounds.box_union(g[i].bounds);
        if (ngeom == 1) return make_leaf(g, ngeom, bounds);
        BBox centroids;
        for (size_t i = 0; i < ngeom; ++i) centroids = centroids.point_union(g[i].center);
        int axis = centroids.max_extent();
        size_t mid = ngeom / 2;
        if (fabsf(centroids.max[axis] - centroids.min[axis]) < F32_EPSILON) {
            if (ngeom < max_geom) return make_leaf(g, ngeom, bounds);
            BuildNode* l = build_rec(g, mid, total);
            BuildNode* r = build_rec(g + mid, ngeom - mid, total);
            return make_interior(l, r, axis);
        }
        if (ngeom < 5) {
            std::stable_sort(g, g + ngeom, [axis](const GeomInfo& a, const GeomInfo& b) { return a.center[axis] < b.center[axis]; });
        } else {
            struct Bucket { size_t count = 0; BBox bounds; };
            Bucket buckets[12];
            const float cmin = centroids.min[axis], cmax = centroids.max[axis];
            auto bucket_of = [&](const GeomInfo& gi) -> uint32_t {
                uint32_t b = f2u((gi.center[axis] - cmin) / (cmax - cmin) * 12.0f);
                return b == 12 ? 11 : b;
            };
            for (size_t i = 0; i < ngeom; ++i) {
                uint32_t b = bucket_of(g[i]);
                if (b > 11) b = 11; /* unreachable for finite input; Rust would panic on OOB */
                buckets[b].count += 1;
                buckets[b].bounds = buckets[b].bounds.box_union(g[i].bounds);
            }
            float cost[11];
            for (int i = 0; i < 11; ++i) {
                Bucket left, right;
                for (int k = 0; k <= i; ++k) { left.bounds = left.bounds.box_union(buckets[k].bounds); left.count += buckets[k].count; }
                for (int k = i + 1; k < 12; ++k) { right.bounds = right.bounds.box_union(buckets[k].bounds); right.count += buckets[k].count; }
                cost[i] = 0.125f + ((float)left.count * left.bounds.surface_area() + (float)right.count * right.bounds.surface_area()) / bounds.surface_area();
            }
            int min_bucket = 0; float min_cost = F32_INF;
            for (int i = 0; i < 11; ++i) if (cost[i] < min_cost) { min_bucket = i; min_cost = cost[i]; }
            if (ngeom > max_geom || min_cost < (float)ngeom) {
                mid = partition(g, ngeom, [&](const GeomInfo& gi) { return bucket_of(gi) <= (uint32_t)min_bucket; });
            } else {
                return make_leaf(g, ngeom, bounds);
            }
        }
        BuildNode* l = build_rec(g, mid, total);
        BuildNode* r = build_rec(g + mid, ngeom - mid, total);
        return make_interior(l, r, axis);
    }
    /* bvh.rs:248-267: pre-order, first child at index+1 */
    uint32_t flatten(const BuildNode* n) {
        uint32_t offset = (uint32_t)tree.size();
        if (!n->leaf) {
            tree.push_back(FlatNode{n->bounds, false, 0, (uint32_t)n->axis});
            flatten(n->c[0]);
            uint32_t second = flatten(n->c[1]);
            tree[offset].a = second;
        } else {
            tree.push_back(FlatNode{n->bounds, true, n->geom_offset, n->ngeom});
        }
        return offset;
    }

    /* bvh.rs:81-130. f(ray, geom_index) -> bool hit (and shrinks ray.max_t, records its result itself) */
    template <class F>
    bool intersect(Ray& ray, Counters& cnt, F f) const {
        bool result = false;
        V3 inv_dir(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
        int neg_dir[3] = {ray.d.x < 0.0f, ray.d.y < 0.0f, ray.d.z < 0.0f};
        uint32_t stack[64];
        int sp = 0;
        uint32_t current = 0;
        for (;;) {
            const FlatNode& node = tree[current];
            cnt.node_tests++;
            if (node.bounds.fast_intersect(ray, inv_dir, neg_dir)) {
                if (node.leaf) {
                    for (uint32_t i = node.a; i < node.a + node.b; ++i) {
                        if (f(ray, ordered_geom[i])) result = true; /* f(ray,o).or(result): last accepted wins */
                    }
                    if (sp == 0) break;
                    current = stack[--sp];
                } else {
                    if (neg_dir[node.b] != 0) { stack[sp] = current + 1; current = node.a; }
                    else { stack[sp] = node.a; current += 1; }
                    sp++;
                }
            } else {
                if (sp == 0) break;
                current = stack[--sp];
            }
        }
        return result;
    }
};

/* ---- geometry::DifferentialGeometry (differential_geometry.rs) ------------------- */
struct DG {
    V3 p, n, ng;
    float u, v, time;
    V3 dp_du, dp_dv;
    /* ::new (differential_geometry.rs:31-47): n from cross(dp_du, dp_dv) */
    static DG make(V3 p, V3 ng, float u, float v, float time, V3 dp_du, V3 dp_dv) {
        DG d; d.p = p; d.n = normalized(cross(dp_du, dp_dv)); d.ng = normalized(ng);
        d.u = u; d.v = v; d.time = time; d.dp_du = dp_du; d.dp_dv = dp_dv; return d;
    }
    /* ::with_normal (:49-65) */
    static DG with_normal(V3 p, V3 n, float u, float v, float time, V3 dp_du, V3 dp_dv) {
        DG d; V3 nn = normalized(n); d.p = p; d.n = nn; d.ng = nn;
        d.u = u; d.v = v; d.time = time; d.dp_du = dp_du; d.dp_dv = dp_dv; return d;
    }
};

/* ---- shapes ----------------------------------------------------------------------- */
/* geometry::Sphere::intersect (sphere.rs:33-81) */
static inline bool sphere_intersect(float radius, Ray& ray, DG& out) {
    float a = length_sqr(ray.d);
    float b = 2.0f * dot(ray.d, ray.o);
    float c = dot(ray.o, ray.o) - radius * radius;
    float t0, t1;
    if (!solve_quadratic(a, b, c, t0, t1)) return false;
    if (t0 > ray.max_t || t1 < ray.min_t) return false;
    float t_hit = t0;
    if (t_hit < ray.min_t) { t_hit = t1; if (t_hit > ray.max_t) return false; }
    ray.max_t = t_hit;
    V3 p = ray.at(t_hit);
    V3 n(p.x, p.y, p.z);
    float theta = M_ACOS(clampf(p.z / radius, -1.0f, 1.0f));
    float inv_z = 1.0f / sqrtf(p.x * p.x + p.y * p.y);
    float cos_phi = p.x * inv_z, sin_phi = p.y * inv_z;
    float u = M_ATAN2(p.x, p.y) / (2.0f * PI);
    if (u < 0.0f) u = u + 1.0f;
    float v = theta / PI;
    V3 dp_du(-PI * 2.0f * p.y, PI * 2.0f * p.x, 0.0f);
    V3 dp_dv = V3(p.z * cos_phi, p.z * sin_phi, -radius * M_SIN(theta)) * PI;
    out = DG::with_normal(p, n, u, v, ray.time, dp_du, dp_dv);
    return true;
}
/* geometry::Disk::intersect (disk.rs:42-76) */
static inline bool disk_intersect(float radius, float inner_radius, Ray& ray, DG& out) {
    if (fabsf(ray.d.z) == 0.0f) return false;
    float t = -ray.o.z / ray.d.z;
    if (t < ray.min_t || t > ray.max_t) return false;
    V3 p = ray.at(t);
    float dist_sqr = p.x * p.x + p.y * p.y;
    if (dist_sqr > radius * radius || dist_sqr < inner_radius * inner_radius) return false;
    float phi = M_ATAN2(p.y, p.x);
    if (phi < 0.0f) phi += PI * 2.0f;
    if (phi > PI * 2.0f) return false;
    ray.max_t = t;
    float hit_radius = sqrtf(dist_sqr);
    float u = phi / (2.0f * PI);
    float v = 1.0f - (hit_radius - inner_radius) / (radius - inner_radius);
    V3 dp_du(-PI * 2.0f * p.y, PI * 2.0f * p.x, 0.0f);
    V3 dp_dv = ((inner_radius - radius) / hit_radius) * V3(p.x, p.y, 0.0f);
    out = DG::make(p, V3(0.0f, 0.0f, 1.0f), u, v, ray.time, dp_du, dp_dv);
    return true;
}
/* geometry::Rectangle::intersect (rectangle.rs:38-64) */
static inline bool rect_intersect(float width, float height, Ray& ray, DG& out) {
    if (fabsf(ray.d.z) < 1e-8f) return false;
    float t = -ray.o.z / ray.d.z;
    if (t < ray.min_t || t > ray.max_t) return false;
    V3 p = ray.at(t);
    float half_width = width / 2.0f, half_height = height / 2.0f;
    if (p.x >= -half_width && p.x <= half_width && p.y >= -half_height && p.y <= half_height) {
        ray.max_t = t;
        float u = (p.x + half_width) / (2.0f * half_width);
        float v = (p.y + half_height) / (2.0f * half_height);
        V3 dp_du(half_width * 2.0f, 0.0f, 0.0f), dp_dv(0.0f, half_height * 2.0f, 0.0f);
        out = DG::make(p, V3(0.0f, 0.0f, 1.0f), u, v, ray.time, dp_du, dp_dv);
        return true;
    }
    return false;
}
/* geometry::mesh::intersect_triangle (mesh.rs:136-198) */
static inline bool intersect_triangle(Ray& ray, V3 pa, V3 pb, V3 pc, V3 na, V3 nb, V3 nc, V3 ta, V3 tb, V3 tc, DG& out) {
    V3 e0 = pb - pa, e1 = pc - pa;
    V3 s0 = cross(ray.d, e1);
    float dd = dot(s0, e0);
    if (dd == 0.0f) return false;
    float div = 1.0f / dd;
    V3 d = ray.o - pa;
    float bary1 = dot(d, s0) * div;
    if (bary1 < 0.0f || bary1 > 1.0f) return false;
    V3 s1 = cross(d, e0);
    float bary2 = dot(ray.d, s1) * div;
    if (bary2 < 0.0f || bary1 + bary2 > 1.0f) return false;
    float t = dot(e1, s1) * div;
    if (t < ray.min_t || t > ray.max_t) return false;
    float bary0 = 1.0f - bary1 - bary2;
    ray.max_t = t;
    V3 p = ray.at(t);
    V3 n = normalized(bary0 * na + bary1 * nb + bary2 * nc);
    V3 texcoord = bary0 * ta + bary1 * tb + bary2 * tc;
    float du0 = ta.x - tc.x, du1 = tb.x - tc.x;
    float dv0 = ta.y - tc.y, dv1 = tb.y - tc.y;
    float det = du0 * dv1 - dv0 * du1;
    V3 dp_du, dp_dv;
    if (det == 0.0f) {
        coordinate_system(normalized(cross(e1, e0)), dp_du, dp_dv);
    } else {
        float idet = 1.0f / det;
        V3 dp0 = pa - pc, dp1 = pb - pc;
        dp_du = (dv1 * dp0 - dv0 * dp1) * idet;
        dp_dv = (-du1 * dp0 + du0 * dp1) * idet;
    }
    out = DG::with_normal(p, n, texcoord.x, texcoord.y, ray.time, dp_du, dp_dv);
    return true;
}

/* geometry::Mesh (mesh.rs:29-90): BVH<Triangle> with max_geom 16 */
struct Mesh {
    std::vector<V3> positions, normals, texcoords; /* texcoords as Point(u, v, 0) (mesh.rs:69-70) */
    std::vector<uint32_t> indices;
    BVH bvh;
    void build() {
        size_t nt = indices.size() / 3;
        std::vector<BBox> b(nt);
        for (size_t t = 0; t < nt; ++t) /* Triangle::bounds (mesh.rs:128-134) */
            b[t] = BBox(positions[indices[3 * t]], positions[indices[3 * t]]).point_union(positions[indices[3 * t + 1]]).point_union(positions[indices[3 * t + 2]]);
        bvh.build(16, b);
    }
    bool intersect(Ray& ray, DG& out, uint32_t& prim, Counters& cnt) const {
        return bvh.intersect(ray, cnt, [&](Ray& r, uint32_t t) {
            cnt.tri_tests++;
            uint32_t a = indices[3 * t], b = indices[3 * t + 1], c = indices[3 * t + 2];
            DG dg;
            if (intersect_triangle(r, positions[a], positions[b], positions[c], normals[a], normals[b], normals[c],
                                   texcoords[a], texcoords[b], texcoords[c], dg)) {
                out = dg; prim = t; return true;
            }
            return false;
        });
    }
    BBox bounds() const { return bvh.tree[0].bounds; }
};

/* ---- Sampleable (sphere.rs:91-141, disk.rs:84-110, rectangle.rs:74-105) + mc.rs ---- */
static inline void concentric_sample_disk(float u0, float u1, float& ox, float& oy) { /* mc.rs:22-51 */
    float s0 = 2.0f * u0 - 1.0f, s1 = 2.0f * u1 - 1.0f;
    float radius, theta;
    if (s0 == 0.0f && s1 == 0.0f) { ox = s0; oy = s1; return; }
    if (s0 >= -s1) {
        if (s0 > s1) { radius = s0; theta = s1 > 0.0f ? s1 / s0 : 8.0f + s1 / s0; }
        else { radius = s1; theta = 2.0f - s0 / s1; }
    } else if (s0 <= s1) { radius = -s0; theta = 4.0f + s1 / s0; }
    else { radius = -s1; theta = 6.0f - s0 / s1; }
    theta = theta * FRAC_PI_4;
    ox = radius * M_COS(theta); oy = radius * M_SIN(theta);
}
static inline V3 cos_sample_hemisphere(float u0, float u1) { /* mc.rs:11-16 */
    float dx, dy;
    concentric_sample_disk(u0, u1, dx, dy);
    return V3(dx, dy, sqrtf(fmaxf(0.0f, 1.0f - dx * dx - dy * dy)));
}
static inline float power_heuristic(float n_f, float pdf_f, float n_g, float pdf_g) { /* mc.rs:56-60 */
    float f = n_f * pdf_f, g = n_g * pdf_g;
    return (f * f) / (f * f + g * g);
}
static inline float uniform_cone_pdf(float cos_theta) { return 1.0f / (PI * 2.0f * (1.0f - cos_theta)); } /* mc.rs:62-64 */
static inline V3 uniform_sample_cone_frame(float u0, float u1, float cos_theta_max, V3 w_x, V3 w_y, V3 w_z) { /* mc.rs:77-83 */
    float cos_theta = lerpf(u0, cos_theta_max, 1.0f);
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float phi = u1 * PI * 2.0f;
    return M_COS(phi) * sin_theta * w_x + M_SIN(phi) * sin_theta * w_y + cos_theta * w_z;
}
static inline V3 uniform_sample_sphere(float u0, float u1) { /* mc.rs:85-90 */
    float z = 1.0f - 2.0f * u0;
    float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float phi = PI * 2.0f * u1;
    return V3(M_COS(phi) * r, M_SIN(phi) * r, z);
}

struct Shape {
    uint32_t kind = TRB_SHAPE_NONE;
    float p0 = 0, p1 = 0;
    const Mesh* mesh = nullptr;
    bool intersect(Ray& ray, DG& dg, uint32_t& prim, Counters& cnt) const {
        prim = 0;
        switch (kind) {
            case TRB_SHAPE_SPHERE: return sphere_intersect(p0, ray, dg);
            case TRB_SHAPE_DISK: return disk_intersect(p0, p1, ray, dg);
            case TRB_SHAPE_RECT: return rect_intersect(p0, p1, ray, dg);
            case TRB_SHAPE_MESH: return mesh->intersect(ray, dg, prim, cnt);
            default: return false;
        }
    }
    BBox bounds() const {
        switch (kind) {
            case TRB_SHAPE_SPHERE: return BBox(V3(-p0, -p0, -p0), V3(p0, p0, p0));              /* sphere.rs:84-88 */
            case TRB_SHAPE_DISK: return BBox(V3(-p0, -p0, -0.1f), V3(p0, p0, 0.1f));             /* disk.rs:79-81 */
            case TRB_SHAPE_RECT: { float hw = p0 / 2.0f, hh = p1 / 2.0f; return BBox(V3(-hw, -hh, 0.0f), V3(hw, hh, 0.0f)); } /* rectangle.rs:67-71 */
            case TRB_SHAPE_MESH: return mesh->bounds();
            default: return BBox(V3(0.0f), V3(0.0f)); /* point light: BBox::singular(origin) (emitter.rs:152) */
        }
    }
    float surface_area() const {
        switch (kind) {
            case TRB_SHAPE_SPHERE: return 4.0f * PI * p0;             /* sic (Q3), sphere.rs:126-128 */
            case TRB_SHAPE_DISK: return PI * (p0 * p0 - p1 * p1);    /* disk.rs:94-96 */
            default: return p0 * p1;                                  /* rectangle.rs:85-87 */
        }
    }
    void sample_uniform(float u0, float u1, V3& p, V3& n) const {
        switch (kind) {
            case TRB_SHAPE_SPHERE: { /* sphere.rs:92-95 */
                p = V3(0.0f) + p0 * uniform_sample_sphere(u0, u1);
                n = normalized(V3(p.x, p.y, p.z));
                break;
            }
            case TRB_SHAPE_DISK: { /* disk.rs:85-90, ignores inner_radius (Q4) */
                float dx, dy;
                concentric_sample_disk(u0, u1, dx, dy);
                p = V3(dx * p0, dy * p0, 0.0f); n = V3(0.0f, 0.0f, 1.0f);
                break;
            }
            default: /* rectangle.rs:77-80 */
                p = V3(u0 * p0 - p0 / 2.0f, u1 * p1 - p1 / 2.0f, 0.0f); n = V3(0.0f, 0.0f, 1.0f);
        }
    }
    /* Sampleable::sample(p, samples) */
    void sample(V3 pt, float u0, float u1, V3& p, V3& n) const {
        if (kind != TRB_SHAPE_SPHERE) { sample_uniform(u0, u1, p, n); return; }
        /* sphere.rs:99-124 */
        float dist_sqr = distance_sqr(pt, V3(0.0f));
        if (dist_sqr - p0 * p0 < 0.0001f) { sample_uniform(u0, u1, p, n); return; }
        V3 w_z = normalized(V3(0.0f) - pt);
        V3 w_x, w_y;
        coordinate_system(w_z, w_x, w_y);
        float cos_theta_max = sqrtf(fmaxf(0.0f, 1.0f - p0 * p0 / dist_sqr));
        Ray ray(pt, normalized(uniform_sample_cone_frame(u0, u1, cos_theta_max, w_x, w_y, w_z)), 0.0f);
        DG dg;
        if (sphere_intersect(p0, ray, dg)) { p = dg.p; n = dg.ng; return; }
        float t = dot(V3(0.0f) - pt, ray.d);
        p = ray.at(t);
        n = normalized(V3(p.x, p.y, p.z));
    }
    /* Sampleable::pdf(p, w_i) */
    float pdf(V3 pt, V3 w_i) const {
        if (kind == TRB_SHAPE_SPHERE) { /* sphere.rs:131-140 */
            float dist_sqr = distance_sqr(pt, V3(0.0f));
            if (dist_sqr - p0 * p0 < 0.0001f) return 1.0f / surface_area();
            float cos_theta_max = sqrtf(fmaxf(0.0f, 1.0f - p0 * p0 / dist_sqr));
            return uniform_cone_pdf(cos_theta_max);
        }
        /* disk.rs:97-110 / rectangle.rs:91-104 */
        Ray ray = Ray::segment(pt, w_i, 0.001f, F32_INF, 0.0f);
        DG d;
        bool hit = kind == TRB_SHAPE_DISK ? disk_intersect(p0, p1, ray, d) : rect_intersect(p0, p1, ray, d);
        if (!hit) return 0.0f;
        V3 w = -w_i;
        float pdf = distance_sqr(pt, ray.at(ray.max_t)) / (fabsf(dot(d.n, w)) * surface_area());
        return std::isfinite(pdf) ? pdf : 0.0f;
    }
};

/* film::AnimatedColor::color (animated_color.rs:52-78); rgb only is observable */
struct ColorKey { float c[4]; float time; };
struct AnimatedColor {
    std::vector<ColorKey> keys;
    void color(float time, float out[4]) const {
        if (keys.empty()) { out[0] = out[1] = out[2] = out[3] = 0.0f; return; }
        if (keys.size() == 1) { for (int i = 0; i < 4; ++i) out[i] = keys[0].c[i]; return; }
        const ColorKey *first = nullptr, *second = nullptr;
        for (const ColorKey& k : keys) { if (k.time < time) first = &k; else break; }
        for (const ColorKey& k : keys) { if (!(k.time < time)) { second = &k; break; } }
        if (!first) { for (int i = 0; i < 4; ++i) out[i] = keys.front().c[i]; return; }
        if (!second) { for (int i = 0; i < 4; ++i) out[i] = keys.back().c[i]; return; }
        float t = (time - first->time) / (second->time - first->time);
        for (int i = 0; i < 4; ++i) out[i] = lerpf(t, first->c[i], second->c[i]);
    }
};

/* geometry::Instance = Receiver | Emitter (instance.rs, receiver.rs, emitter.rs) */
struct Instance {
    uint32_t kind = TRB_INST_RECEIVER;
    Shape shape;
    uint32_t material = 0;
    AnimatedTransform transform;
    AnimatedColor emission;
    bool is_emitter() const { return kind != TRB_INST_RECEIVER; }
    /* Boundable::bounds (receiver.rs:55-59, emitter.rs:149-158) */
    BBox bounds(float start, float end) const { return transform.animation_bounds(shape.bounds(), start, end); }
};

struct Hit {
    DG dg;
    uint32_t inst = TRB_MISS;
    uint32_t prim = 0;
};

/* Receiver::intersect / Emitter::intersect (receiver.rs:29-43, emitter.rs:118-137).
 * baseline==true recomputes AnimatedTransform::transform(ray.time) per call like the
 * reference (the cost profile BASELINE.md times); otherwise a per-frame cache is used
 * for static instances — identical bits, because the recomposition is time-independent
 * when every spline has one control point (animated_transform.rs:47-48). */
struct SceneGeom {
    std::vector<Instance> instances;
    std::vector<Transform> cached; /* valid for static instances */
    std::vector<uint8_t> is_static;
    BVH tlas;
    bool baseline = false;

    Transform xf(uint32_t i, float time) const {
        if (!baseline && is_static[i]) return cached[i];
        return instances[i].transform.transform(time);
    }
    bool instance_intersect(uint32_t i, Ray& ray, Hit& out, Counters& cnt) const {
        cnt.inst_tests++;
        const Instance& in = instances[i];
        if (in.kind == TRB_INST_EMITTER_POINT) return false; /* emitter.rs:119-120 */
        Transform t = xf(i, ray.time);
        Ray local = t.inv_ray(ray);
        DG dg; uint32_t prim;
        if (!in.shape.intersect(local, dg, prim, cnt)) return false;
        ray.max_t = local.max_t;
        dg.p = t.point(dg.p);
        dg.n = t.normal(dg.n);
        dg.ng = t.normal(dg.ng);
        dg.dp_du = t.vector(dg.dp_du);
        dg.dp_dv = t.vector(dg.dp_dv);
        out.dg = dg; out.inst = i; out.prim = prim;
        return true;
    }
    /* Scene::intersect (scene.rs:148-150) */
    bool intersect(Ray& ray, Hit& out, Counters& cnt) const {
        return tlas.intersect(ray, cnt, [&](Ray& r, uint32_t i) { return instance_intersect(i, r, out, cnt); });
    }
    /* BVH<Instance>::rebuild(start, end) (bvh.rs:61-78) with max_geom 4 (scene.rs:141) */
    void rebuild(float start, float end) {
        std::vector<BBox> b(instances.size());
        cached.resize(instances.size()); is_static.resize(instances.size());
        for (size_t i = 0; i < instances.size(); ++i) {
            bool st = true;
            for (const Spline& s : instances[i].transform.keyframes) st = st && s.ctrl.size() == 1;
            is_static[i] = st;
            if (st) cached[i] = instances[i].transform.transform(start);
            b[i] = instances[i].bounds(start, end);
        }
        tlas.build(4, b);
    }
};

} // namespace orc
