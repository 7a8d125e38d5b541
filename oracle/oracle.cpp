/* oracle/oracle.cpp — TEST INFRASTRUCTURE (parity oracle + timed CPU baseline).
 *
 * A seeded CPU restatement of tray_rust's render hot path. It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it. The product (tray_rust_b200/csrc) shares no code
 * with it; both implement the written contract in DESIGN.md.
 *
 * Parity status: the reference (Rust) cannot be built here (no cargo/rustc) and its
 * own tests hold no vectors for this path => "parity unpinned":
 * this oracle is pinned only by (a) the reference's 17 linalg/partition unit tests,
 * restated in tests/test_oracle_linalg.py, (b) analytic checks (furnace, pdf
 * normalisation, stratification), (c) hand-derived known answers.
 *
 * Follows, in order: src/linalg/*, src/geometry/*, src/partition.rs (orc_linalg.h,
 * orc_geom.h); src/bxdf/**, src/material/*, src/light/mod.rs, src/integrator/{mod,path}.rs,
 * src/sampler/ld.rs (orc_shade.h); and here: src/film/{camera,render_target,color}.rs,
 * src/film/filter/*.rs, src/sampler/{block_queue,morton}.rs,
 * src/exec/multithreaded.rs:55-114, src/scene.rs:141-176.
 */
#include "../include/trb.h"
#include "orc_shade.h"
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <string>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

/* ---- film::filter (filter/mitchell_netravali.rs, filter/gaussian.rs) --------------- */
struct Filter {
    uint32_t type; float w, h, inv_w, inv_h, b, c, alpha, exp_x, exp_y;
    float mn_weight_1d(float x) const { /* mitchell_netravali.rs:35-50; tests the SIGNED x */
        float abs_x = fabsf(x);
        if (x >= 2.0f) return 0.0f;
        if (x >= 1.0f)
            return 1.0f / 6.0f * ((-b - 6.0f * c) * powf(abs_x, 3.0f) + (6.0f * b + 30.0f * c) * powf(abs_x, 2.0f) +
                                  (-12.0f * b - 48.0f * c) * abs_x + (8.0f * b + 24.0f * c));
        return 1.0f / 6.0f * ((12.0f - 9.0f * b - 6.0f * c) * powf(abs_x, 3.0f) + (-18.0f + 12.0f * b + 6.0f * c) * powf(abs_x, 2.0f) +
                              (6.0f - 2.0f * b));
    }
    float g_weight_1d(float x, float e) const { return fmaxf(0.0f, expf(-alpha * x * x) - e); } /* gaussian.rs:27-29 */
    float weight(float x, float y) const {
        if (type == TRB_FILTER_MITCHELL_NETRAVALI) return mn_weight_1d(2.0f * x * inv_w) * mn_weight_1d(2.0f * y * inv_h);
        return g_weight_1d(x, exp_x) * g_weight_1d(y, exp_y);
    }
};

/* ---- film::RenderTarget (render_target.rs) ----------------------------------------- */
struct ImageSample { float x, y; Col color; };
struct RenderTarget {
    int width, height;
    Filter filter;
    float table[256];
    int fpw[2];
    void init(const trb_film& f) {
        width = (int)f.width; height = (int)f.height;
        filter.type = f.filter_type; filter.w = f.filter_w; filter.h = f.filter_h;
        filter.inv_w = 1.0f / f.filter_w; filter.inv_h = 1.0f / f.filter_h;
        filter.b = clampf(f.filter_b, 0.0f, 1.0f); filter.c = clampf(f.filter_c, 0.0f, 1.0f);
        filter.alpha = f.filter_b;
        filter.exp_x = expf(-filter.alpha * f.filter_w * f.filter_w); filter.exp_y = expf(-filter.alpha * f.filter_h * f.filter_h);
        fpw[0] = (int)floorf(filter.w / 0.5f); fpw[1] = (int)floorf(filter.h / 0.5f); /* render_target.rs:48-49 */
        for (int y = 0; y < 16; ++y) { /* :52-58 */
            float fy = ((float)y + 0.5f) * filter.h / 16.0f;
            for (int x = 0; x < 16; ++x) {
                float fx = ((float)x + 0.5f) * filter.w / 16.0f;
                table[y * 16 + x] = filter.weight(fx, fy);
            }
        }
    }
    /* RenderTarget::write (render_target.rs:77-165), lock blocks are 2x2; `film` is the
     * row-major RGBW buffer (get_renderf32 layout). */
    void write(const std::vector<ImageSample>& samples, int rx0, int ry0, int rx1, int ry1, float* film, bool atomic) const {
        const int ls = 2;
        int x_range[2] = {std::max(rx0 - fpw[0], 0), std::min(rx1 + fpw[0], width - 1)};
        int y_range[2] = {std::max(ry0 - fpw[1], 0), std::min(ry1 + fpw[1], height - 1)};
        if (x_range[1] - x_range[0] < 0 || y_range[1] - y_range[0] < 0) return;
        int bxr[2] = {x_range[0] / ls, x_range[1] / ls}, byr[2] = {y_range[0] / ls, y_range[1] / ls};
        float filtered[4][4];
        for (int y = byr[0]; y <= byr[1]; ++y)
            for (int x = bxr[0]; x <= bxr[1]; ++x) {
                int bxs = x * ls, bys = y * ls;
                int xw[2] = {std::max(x_range[0], bxs), std::min(x_range[1] + 1, bxs + ls)};
                int yw[2] = {std::max(y_range[0], bys), std::min(y_range[1] + 1, bys + ls)};
                for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) filtered[i][k] = 0.0f;
                for (const ImageSample& c : samples) {
                    if (!(c.x >= (float)(xw[0] - fpw[0]) && c.x < (float)(xw[1] + fpw[0]) && c.y >= (float)(yw[0] - fpw[1]) &&
                          c.y < (float)(yw[1] + fpw[1])))
                        continue;
                    float img_x = c.x - 0.5f, img_y = c.y - 0.5f;
                    for (int iy = yw[0]; iy < yw[1]; ++iy) {
                        float fy = fabsf((float)iy - img_y) * filter.inv_h;
                        if (fy > filter.h) continue;
                        uint32_t fy_idx = std::min(f2u(fy * 16.0f), 15u);
                        for (int ix = xw[0]; ix < xw[1]; ++ix) {
                            float fx = fabsf((float)ix - img_x) * filter.inv_w;
                            if (fx > filter.w) continue;
                            uint32_t fx_idx = std::min(f2u(fx * 16.0f), 15u);
                            float weight = table[fy_idx * 16 + fx_idx];
                            int px = (iy - bys) * ls + ix - bxs;
                            filtered[px][0] += weight * c.color.r;
                            filtered[px][1] += weight * c.color.g;
                            filtered[px][2] += weight * c.color.b;
                            filtered[px][3] += weight;
                        }
                    }
                }
                for (int iy = yw[0]; iy < yw[1]; ++iy)
                    for (int ix = xw[0]; ix < xw[1]; ++ix) {
                        int px = (iy - bys) * ls + ix - bxs;
                        float* dst = film + ((size_t)iy * width + ix) * 4;
                        for (int k = 0; k < 4; ++k) {
                            if (atomic) {
#pragma omp atomic
                                dst[k] += filtered[px][k];
                            } else dst[k] += filtered[px][k];
                        }
                    }
            }
    }
};

/* ---- film::Camera (camera.rs) ------------------------------------------------------- */
struct Camera {
    AnimatedTransform cam_world;
    Transform raster_screen, proj_div_inv, px_to_cam;
    float shutter_open = 0, shutter_close = 0, shutter_size = 0.5f, fov = 30;
    /* CameraFov::Animated(BSpline<f32>) (camera.rs:23-30,95-125) */
    bool fov_animated = false;
    uint32_t fov_degree = 0;
    std::vector<float> fovs, fov_knots;
    V3 scaling;
    uint32_t active_at = 0;
    void init(float fov_, uint32_t w, uint32_t h) { /* camera.rs:64-91 */
        fov = fov_;
        float aspect_ratio = (float)w / (float)h;
        float screen[4];
        if (aspect_ratio > 1.0f) { screen[0] = -aspect_ratio; screen[1] = aspect_ratio; screen[2] = -1.0f; screen[3] = 1.0f; }
        else { screen[0] = -1.0f; screen[1] = 1.0f; screen[2] = -1.0f / aspect_ratio; screen[3] = 1.0f / aspect_ratio; }
        Transform screen_raster = Transform::scale(V3((float)w, (float)h, 1.0f)) *
                                  Transform::scale(V3(1.0f / (screen[1] - screen[0]), 1.0f / (screen[2] - screen[3]), 1.0f)) *
                                  Transform::translate(V3(-screen[0], -screen[3], 0.0f));
        raster_screen = screen_raster.inverse();
        float far = 1.0f, near = 1000.0f;
        M4 proj_div = M4::identity();
        proj_div.at(2, 2) = far / (far - near); proj_div.at(2, 3) = -far * near / (far - near);
        proj_div.at(3, 2) = 1.0f; proj_div.at(3, 3) = 0.0f;
        proj_div_inv = Transform::from_mat(proj_div).inverse();
        px_to_cam = proj_div_inv * raster_screen; /* left-assoc product in generate_ray (camera.rs:152), ray-invariant */
        float tan_fov = tanf(to_radians(fov) / 2.0f);
        scaling = V3(tan_fov, tan_fov, 1.0f);
    }
    /* bspline 0.2.2 BSpline<f32>::point: de Boor with Interpolate for f32 = a * (1 - t) + b * t */
    float fov_spline_point(float t) const {
        size_t n = fov_knots.size(), degree = fov_degree;
        size_t ub = n;
        for (size_t i = 0; i < n; ++i) if (fov_knots[i] > t) { ub = i; break; }
        size_t i0;
        if (ub == n) i0 = n - degree - 1;
        else if (ub == 0) i0 = degree;
        else if (ub >= n - degree - 1) i0 = n - degree - 1;
        else i0 = ub;
        std::vector<float> tmp(degree + 1);
        for (size_t j = 0; j <= degree; ++j) tmp[j] = fovs[j + i0 - degree - 1];
        for (size_t lvl = 0; lvl < degree; ++lvl) {
            size_t k = lvl + 1;
            for (size_t j = 0; j < degree - lvl; ++j) {
                size_t i = j + k + i0 - degree;
                float alpha = (t - fov_knots[i - 1]) / (fov_knots[i + degree - k] - fov_knots[i - 1]);
                tmp[j] = tmp[j] * (1.0f - alpha) + tmp[j + 1] * alpha;
            }
        }
        return tmp[0];
    }
    void update_frame(float start, float end) { /* camera.rs:127-144 */
        shutter_open = start;
        shutter_close = start + shutter_size * (end - start);
        if (fov_animated) { /* the spline is sampled once per frame, at the clamped mid-frame time (camera.rs:134-141) */
            float lo = fov_knots[fov_degree], hi = fov_knots[fov_knots.size() - 1 - fov_degree];
            float t = (start + end) / 2.0f;
            t = t < lo ? lo : (t > hi ? hi : t);
            fov = fov_spline_point(t);
        }
        float tan_fov = tanf(to_radians(fov) / 2.0f);
        scaling = V3(tan_fov, tan_fov, 1.0f);
    }
    Ray generate_ray(float px, float py, float time) const { /* camera.rs:150-157 */
        V3 px_pos = scaling * px_to_cam.point(V3(px, py, 0.0f));
        V3 d = normalized(V3(px_pos.x, px_pos.y, px_pos.z));
        float frame_time = (shutter_close - shutter_open) * time + shutter_open;
        return cam_world.transform(frame_time).ray(Ray(V3(0.0f), d, frame_time));
    }
};

/* sampler::morton (morton.rs) */
static inline uint32_t part1_by1(uint32_t x) {
    x &= 0x0000ffffu; x = (x ^ (x << 8)) & 0x00ff00ffu; x = (x ^ (x << 4)) & 0x0f0f0f0fu;
    x = (x ^ (x << 2)) & 0x33333333u; return (x ^ (x << 1)) & 0x55555555u;
}
static inline uint32_t morton2(uint32_t x, uint32_t y) { return (part1_by1(y) << 1) + part1_by1(x); }

thread_local std::string g_err;

} // namespace

struct orc_scene {
    trb_film film;
    SceneGeom geom;
    SceneShade shade;
    std::vector<Mesh*> meshes;
    std::vector<std::vector<float>> merl;
    TextureSet textures;
    std::vector<Camera> cameras;
    int active_camera = -1;
    RenderTarget rt;
    uint32_t spp_pow2 = 1;
    ~orc_scene() { for (Mesh* m : meshes) delete m; }

    /* BlockQueue::new (block_queue.rs:28-46) */
    std::vector<std::pair<uint32_t, uint32_t>> block_list(uint32_t start, uint32_t count) const {
        uint32_t nbx = film.width / 8, nby = film.height / 8;
        std::vector<std::pair<uint32_t, uint32_t>> blocks(nbx * nby);
        for (uint32_t i = 0; i < nbx * nby; ++i) blocks[i] = {i % nbx, i / nbx};
        std::stable_sort(blocks.begin(), blocks.end(), [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) {
            return morton2(a.first, a.second) < morton2(b.first, b.second);
        });
        if (count > 0) {
            std::vector<std::pair<uint32_t, uint32_t>> sel;
            for (size_t i = start; i < blocks.size() && sel.size() < count; ++i) sel.push_back(blocks[i]);
            return sel;
        }
        return blocks;
    }
    /* Scene::update_frame (scene.rs:152-176) */
    void update_frame(uint32_t frame, float start, float end) {
        int cam;
        if (active_camera >= 0) {
            cam = active_camera;
            if (cam != (int)cameras.size() - 1 && cameras[cam + 1].active_at == frame) cam = cam + 1;
        } else {
            int c = 0;
            for (const Camera& x : cameras) { if (x.active_at <= frame) c++; else break; }
            cam = c - 1;
        }
        active_camera = cam;
        cameras[cam].update_frame(start, end);
        geom.rebuild(cameras[cam].shutter_open, cameras[cam].shutter_close);
    }
};

static AnimatedTransform load_xf(const trb_scene_desc* d, uint32_t first, uint32_t n) {
    AnimatedTransform at;
    for (uint32_t s = first; s < first + n; ++s) {
        const trb_spline& sp = d->splines[s];
        Spline o; o.degree = sp.degree;
        for (uint32_t k = 0; k < sp.n_ctrl; ++k) {
            const trb_keyframe& kf = d->keyframes[sp.ctrl_first + k];
            Keyframe kk;
            kk.translation = V3(kf.translation[0], kf.translation[1], kf.translation[2]);
            kk.rotation = Quat{V3(kf.rotation[0], kf.rotation[1], kf.rotation[2]), kf.rotation[3]};
            kk.scaling = V3(kf.scaling[0], kf.scaling[1], kf.scaling[2]);
            o.ctrl.push_back(kk);
        }
        for (uint32_t k = 0; k < sp.n_knots; ++k) o.knots.push_back(d->knots[sp.knot_first + k]);
        std::stable_sort(o.knots.begin(), o.knots.end()); /* BSpline::new sorts the knots (bspline 0.2.2: knots.sort_by(partial_cmp)) */
        at.keyframes.push_back(o);
    }
    return at;
}

static uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

/* which libm this build evaluates transcendentals with: 0 = detmath, 1 = system */
int orc_libm_kind(void) {
#ifdef ORC_SYSTEM_LIBM
    return 1;
#else
    return 0;
#endif
}

int orc_scene_create(const trb_scene_desc* d, orc_scene** out) {
    if (!d || !out || d->abi_version != TRB_ABI_VERSION) { g_err = "bad desc"; return TRB_INVALID_ARG; }
    if (d->film.width % 8 || d->film.height % 8 || d->film.width == 0 || d->film.height == 0) { g_err = "image not divisible into 8x8 blocks"; return TRB_INVALID_ARG; }
    if (d->n_instances == 0) { g_err = "the scene does not have any objects"; return TRB_INVALID_ARG; }
    if (d->integrator.type > TRB_INTEGRATOR_NORMALS_DEBUG) { g_err = "Unrecognized integrator type"; return TRB_INVALID_ARG; }
    orc_scene* s = new orc_scene;
    s->film = d->film;
    s->spp_pow2 = next_pow2(std::max(1u, d->film.samples));
    s->rt.init(d->film);
    for (uint32_t m = 0; m < d->n_meshes; ++m) {
        const trb_mesh& tm = d->meshes[m];
        Mesh* mesh = new Mesh;
        for (uint32_t v = 0; v < tm.n_verts; ++v) {
            mesh->positions.push_back(V3(tm.positions[3 * v], tm.positions[3 * v + 1], tm.positions[3 * v + 2]));
            mesh->normals.push_back(V3(tm.normals[3 * v], tm.normals[3 * v + 1], tm.normals[3 * v + 2]));
            mesh->texcoords.push_back(V3(tm.texcoords[2 * v], tm.texcoords[2 * v + 1], 0.0f));
        }
        mesh->indices.assign(tm.indices, tm.indices + 3 * (size_t)tm.n_tris);
        mesh->build();
        s->meshes.push_back(mesh);
    }
    for (uint32_t m = 0; m < d->n_merl; ++m) s->merl.emplace_back(d->merl_tables[m], d->merl_tables[m] + TRB_MERL_TABLE_FLOATS);
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const trb_material& tm = d->materials[m];
        Material mat; mat.type = tm.type; mat.c0 = Col(tm.c0[0], tm.c0[1], tm.c0[2]); mat.c1 = Col(tm.c1[0], tm.c1[1], tm.c1[2]);
        mat.roughness = tm.roughness; mat.eta = tm.eta;
        if (tm.type == TRB_MAT_MERL) mat.merl = s->merl[tm.merl].data();
        for (int k = 0; k < 4; ++k) {
            mat.tex[k] = tm.tex[k];
            if (tm.tex[k] > d->n_textures) { delete s; g_err = "texture index out of range"; return TRB_INVALID_ARG; }
        }
        mat.textures = &s->textures;
        s->shade.materials.push_back(mat);
    }
    for (uint32_t t = 0; t < d->n_textures; ++t) {
        const trb_texture& tt = d->textures[t];
        if (tt.n_images == 0 || (uint64_t)tt.first_image + tt.n_images > d->n_images) { delete s; g_err = "texture image range out of bounds"; return TRB_INVALID_ARG; }
        s->textures.tex.push_back(tt);
    }
    for (uint32_t i = 0; i < d->n_images; ++i) {
        const trb_image& ti = d->images[i];
        if (ti.width == 0 || ti.height == 0 || !ti.rgba8) { delete s; g_err = "empty image"; return TRB_INVALID_ARG; }
        TexImage im; im.w = ti.width; im.h = ti.height; im.time = ti.time;
        im.px.assign(ti.rgba8, ti.rgba8 + (size_t)ti.width * ti.height * 4);
        s->textures.img.push_back(std::move(im));
    }
    for (uint32_t i = 0; i < d->n_instances; ++i) {
        const trb_instance& ti = d->instances[i];
        Instance in; in.kind = ti.kind; in.shape.kind = ti.shape; in.shape.p0 = ti.p0; in.shape.p1 = ti.p1;
        if (ti.shape == TRB_SHAPE_MESH) in.shape.mesh = s->meshes[ti.mesh];
        in.material = ti.material;
        in.transform = load_xf(d, ti.spline_first, ti.n_splines);
        for (uint32_t k = 0; k < ti.n_emission; ++k) {
            const trb_color_key& ck = d->color_keys[ti.emission_first + k];
            ColorKey c; memcpy(c.c, ck.rgba, 16); c.time = ck.time; in.emission.keys.push_back(c);
        }
        if (in.is_emitter()) s->shade.lights.push_back(i);
        s->geom.instances.push_back(in);
    }
    if (s->shade.lights.empty()) { delete s; g_err = "At least one light is required"; return TRB_INVALID_ARG; } /* multithreaded.rs:39 */
    for (uint32_t c = 0; c < d->n_cameras; ++c) {
        const trb_camera& tc = d->cameras[c];
        Camera cam; cam.cam_world = load_xf(d, tc.spline_first, tc.n_splines);
        cam.shutter_size = tc.shutter_size; cam.active_at = tc.active_at;
        if (tc.n_fov_ctrl > 0) { /* Camera::animated_fov (camera.rs:95-125): starts from fovs[0] */
            cam.fov_animated = true; cam.fov_degree = tc.fov_degree;
            cam.fovs.assign(d->fov_floats + tc.fov_ctrl_first, d->fov_floats + tc.fov_ctrl_first + tc.n_fov_ctrl);
            cam.fov_knots.assign(d->fov_floats + tc.fov_knot_first, d->fov_floats + tc.fov_knot_first + tc.n_fov_knots);
            std::stable_sort(cam.fov_knots.begin(), cam.fov_knots.end()); /* BSpline::new sorts the knots */
            if (cam.fov_knots.size() != cam.fovs.size() + cam.fov_degree + 1) { delete s; g_err = "Invalid B-spline: knots.len() != control_points.len() + degree + 1"; return TRB_INVALID_ARG; }
            cam.init(cam.fovs[0], d->film.width, d->film.height);
        } else cam.init(tc.fov, d->film.width, d->film.height);
        s->cameras.push_back(cam);
    }
    if (s->cameras.empty()) { delete s; g_err = "A camera is required"; return TRB_INVALID_ARG; }
    s->shade.geom = &s->geom;
    s->shade.min_depth = d->integrator.min_depth; s->shade.max_depth = d->integrator.max_depth; s->shade.integrator = d->integrator.type;
    /* Scene::load_file builds the TLAS for [0, scene_time] (scene.rs:141); update_frame rebuilds it */
    s->geom.rebuild(0.0f, d->film.scene_time);
    *out = s;
    return TRB_OK;
}
void orc_scene_destroy(orc_scene* s) { delete s; }

int orc_scene_update_frame(orc_scene* s, uint32_t frame, float start, float end) { s->update_frame(frame, start, end); return TRB_OK; }

/* baseline != 0: recompute AnimatedTransform::transform per ray like the reference */
void orc_set_baseline_mode(orc_scene* s, int baseline) { s->geom.baseline = baseline != 0; }

int orc_block_list(const orc_scene* s, uint32_t start, uint32_t count, uint32_t* n_out, uint32_t* xy, uint32_t cap) {
    auto bl = s->block_list(start, count);
    *n_out = (uint32_t)bl.size();
    if (xy) for (size_t i = 0; i < bl.size() && i < cap; ++i) { xy[2 * i] = bl[i].first; xy[2 * i + 1] = bl[i].second; }
    return TRB_OK;
}

int orc_scene_get_bvh(const orc_scene* s, int which, uint32_t* n_nodes, trb_bvh_node* nodes, uint32_t* n_ordered, uint32_t* ordered) {
    const BVH& b = which < 0 ? s->geom.tlas : s->meshes[which]->bvh;
    *n_nodes = (uint32_t)b.tree.size(); *n_ordered = (uint32_t)b.ordered_geom.size();
    if (nodes) for (size_t i = 0; i < b.tree.size(); ++i) {
        const FlatNode& f = b.tree[i];
        nodes[i].bmin[0] = f.bounds.min.x; nodes[i].bmin[1] = f.bounds.min.y; nodes[i].bmin[2] = f.bounds.min.z;
        nodes[i].bmax[0] = f.bounds.max.x; nodes[i].bmax[1] = f.bounds.max.y; nodes[i].bmax[2] = f.bounds.max.z;
        nodes[i].a = f.a; nodes[i].b = f.leaf ? (TRB_BVH_LEAF | f.b) : f.b;
    }
    if (ordered) memcpy(ordered, b.ordered_geom.data(), 4 * b.ordered_geom.size());
    return TRB_OK;
}
int orc_scene_get_transform(const orc_scene* s, uint32_t inst, float* mat16, float* inv16) {
    float t0 = s->active_camera >= 0 ? s->cameras[s->active_camera].shutter_open : 0.0f;
    Transform t = s->geom.instances[inst].transform.transform(t0);
    memcpy(mat16, t.mat.m, 64); memcpy(inv16, t.inv.m, 64);
    return TRB_OK;
}
int orc_scene_get_filter_table(const orc_scene* s, float* t) { memcpy(t, s->rt.table, 1024); return TRB_OK; }

/* Scene::intersect over a batch (scene.rs:148-150) */
int orc_intersect(orc_scene* s, size_t n, const trb_ray* rays, trb_hit* hits, trb_stats* stats) {
    Counters total;
    /* batch rays carry no time: they are traced at the frame's shutter-open time */
    const float t0 = s->active_camera >= 0 ? s->cameras[s->active_camera].shutter_open : 0.0f;
#pragma omp parallel
    {
        Counters cnt;
#pragma omp for schedule(dynamic, 1024)
        for (long i = 0; i < (long)n; ++i) {
            Ray r = Ray::segment(V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].min_t, rays[i].max_t, t0);
            Hit h;
            bool hit = s->geom.intersect(r, h, cnt);
            hits[i].t = r.max_t; hits[i].inst = hit ? h.inst : TRB_MISS; hits[i].prim = hit ? h.prim : 0; hits[i].pad = 0;
        }
#pragma omp critical
        total.add(cnt);
    }
    if (stats) { memset(stats, 0, sizeof *stats); stats->node_tests = total.node_tests; stats->tri_tests = total.tri_tests; stats->inst_tests = total.inst_tests; }
    return TRB_OK;
}

struct PixelStreams { uint32_t scr0, scr1, kpos, scrt, ktime; };
static inline PixelStreams pixel_streams(uint32_t seed, uint32_t pixel) {
    PixelStreams p;
    p.scr0 = dm_scramble(dm_rng(seed, pixel, DM_PIXEL_STREAM, DM_PX_POS0));
    p.scr1 = dm_scramble(dm_rng(seed, pixel, DM_PIXEL_STREAM, DM_PX_POS1));
    p.kpos = dm_rng(seed, pixel, DM_PIXEL_STREAM, DM_PX_POS_PERM);
    p.scrt = dm_scramble(dm_rng(seed, pixel, DM_PIXEL_STREAM, DM_PX_TIME));
    p.ktime = dm_rng(seed, pixel, DM_PIXEL_STREAM, DM_PX_TIME_PERM);
    return p;
}

/* The body of thread_work (multithreaded.rs:72-114). mode 0: splat to film; 1: dump samples; 2: dump camera rays */
static int render_impl(orc_scene* s, const trb_render_cfg* cfg, int mode, float* film, trb_sample* out_samples, trb_ray* out_rays,
                       float* out_xy, size_t n_out, trb_stats* stats, int threads) {
    if (s->active_camera < 0) { g_err = "update_frame must be called before rendering"; return TRB_INVALID_ARG; }
    uint32_t spp = cfg->spp ? next_pow2(cfg->spp) : s->spp_pow2;
    uint32_t s_first = cfg->sample_first, s_count = cfg->sample_count ? cfg->sample_count : spp - std::min(spp, s_first);
    if (s_first + s_count > spp) { g_err = "sample range exceeds spp"; return TRB_INVALID_ARG; }
    auto blocks = s->block_list(cfg->block_start, cfg->block_count);
    if (mode != 0 && n_out != blocks.size() * 64 * (size_t)s_count) { g_err = "output size mismatch"; return TRB_INVALID_ARG; }
    const Camera& camera = s->cameras[s->active_camera];
    const uint32_t width = s->film.width;
    Counters total;
    auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    (void)threads;
#pragma omp parallel
    {
        Counters cnt;
        std::vector<ImageSample> block_samples;
#pragma omp for schedule(dynamic, 1)
        for (long bi = 0; bi < (long)blocks.size(); ++bi) {
            uint32_t bx = blocks[bi].first * 8, by = blocks[bi].second * 8;
            block_samples.clear();
            size_t o = (size_t)bi * 64 * s_count;
            for (uint32_t py = by; py < by + 8; ++py)
                for (uint32_t px = bx; px < bx + 8; ++px) {
                    uint32_t pixel = py * width + px;
                    PixelStreams st = pixel_streams(cfg->seed, pixel);
                    for (uint32_t si = s_first; si < s_first + s_count; ++si, ++o) {
                        /* get_samples + get_samples_1d (ld.rs:33-64) */
                        uint32_t ip = dm_permute(si, spp, st.kpos);
                        float sx = van_der_corput(ip, st.scr0) + (float)px;
                        float sy = sobol(ip, st.scr1) + (float)py;
                        float tm = van_der_corput(dm_permute(si, spp, st.ktime), st.scrt);
                        Ray ray = camera.generate_ray(sx, sy, tm);
                        cnt.camera_samples++;
                        if (mode == 2) {
                            out_rays[o].o[0] = ray.o.x; out_rays[o].o[1] = ray.o.y; out_rays[o].o[2] = ray.o.z;
                            out_rays[o].d[0] = ray.d.x; out_rays[o].d[1] = ray.d.y; out_rays[o].d[2] = ray.d.z;
                            out_rays[o].min_t = ray.min_t; out_rays[o].max_t = ray.max_t;
                            out_xy[2 * o] = sx; out_xy[2 * o + 1] = sy;
                            continue;
                        }
                        Hit hit;
                        Col c(0.0f);
                        cnt.rays[0]++;
                        if (s->geom.intersect(ray, hit, cnt)) {
                            PathSamples ps{cfg->seed, pixel, si, s->shade.max_depth + 1};
                            c = s->shade.illumination(ray, hit, ps, cnt).clamp(); /* multithreaded.rs:98-99 */
                        }
                        if (mode == 1) { out_samples[o].x = sx; out_samples[o].y = sy; out_samples[o].r = c.r; out_samples[o].g = c.g; out_samples[o].b = c.b; }
                        else block_samples.push_back(ImageSample{sx, sy, c});
                    }
                }
            if (mode == 0) s->rt.write(block_samples, (int)bx, (int)by, (int)bx + 8, (int)by + 8, film, true);
        }
#pragma omp critical
        total.add(cnt);
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->camera_samples = total.camera_samples;
        stats->rays_primary = total.rays[0]; stats->rays_shadow = total.rays[1]; stats->rays_mis = total.rays[2]; stats->rays_continuation = total.rays[3];
        stats->node_tests = total.node_tests; stats->tri_tests = total.tri_tests; stats->inst_tests = total.inst_tests;
        stats->kernel_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
    }
    return TRB_OK;
}

/* Exec::render (multithreaded.rs:55-70); threads <= 0: all cores */
int orc_render(orc_scene* s, const trb_render_cfg* cfg, float* film_rgbw, trb_stats* stats, int threads) {
    if (!(cfg->flags & TRB_RENDER_NO_UPDATE)) {
        float time_step = s->film.scene_time / (float)s->film.frames;
        s->update_frame(cfg->current_frame, (float)cfg->current_frame * time_step, ((float)cfg->current_frame + 1.0f) * time_step);
    }
    return render_impl(s, cfg, 0, film_rgbw, nullptr, nullptr, nullptr, 0, stats, threads);
}
int orc_render_samples(orc_scene* s, const trb_render_cfg* cfg, size_t n, trb_sample* samples, trb_stats* stats, int threads) {
    return render_impl(s, cfg, 1, nullptr, samples, nullptr, nullptr, n, stats, threads);
}
int orc_camera_rays(orc_scene* s, const trb_render_cfg* cfg, size_t n, trb_ray* rays, float* xy) {
    return render_impl(s, cfg, 2, nullptr, nullptr, rays, xy, n, nullptr, 1);
}

/* RenderTarget::get_render (render_target.rs:185-210) + Colorf::to_srgb (color.rs:59-72) */
int orc_film_to_srgb8(const orc_scene* s, const float* film, uint8_t* rgb8) {
    size_t n = (size_t)s->film.width * s->film.height;
    for (size_t i = 0; i < n; ++i) {
        const float* c = film + 4 * i;
        uint8_t* o = rgb8 + 3 * i;
        o[0] = o[1] = o[2] = 0;
        if (c[3] > 0.0f) {
            for (int k = 0; k < 3; ++k) {
                float v = clampf(c[k] / c[3], 0.0f, 1.0f);
                float sr = v <= 0.0031308f ? 12.92f * v : (1.0f + 0.055f) * M_POW(v, 1.0f / 2.4f) - 0.055f;
                o[k] = (uint8_t)f2u(sr * 255.0f > 255.0f ? 255.0f : sr * 255.0f);
            }
        }
    }
    return TRB_OK;
}

/* ---- known-answer probes for tests ---------------------------------------------------- */
/* op: 0 sin 1 cos 2 acos 3 atan2(a,b) 4 exp 5 log 6 pow(a,b) — detmath regardless of build */
void orc_detmath(int op, size_t n, const float* a, const float* b, float* out) {
    for (size_t i = 0; i < n; ++i) switch (op) {
        case 0: out[i] = dm_sinf(a[i]); break; case 1: out[i] = dm_cosf(a[i]); break; case 2: out[i] = dm_acosf(a[i]); break;
        case 3: out[i] = dm_atan2f(a[i], b[i]); break; case 4: out[i] = dm_expf(a[i]); break; case 5: out[i] = dm_logf(a[i]); break;
        default: out[i] = dm_powf(a[i], b[i]);
    }
}
uint32_t orc_rng(uint32_t seed, uint32_t a, uint32_t b, uint32_t c) { return dm_rng(seed, a, b, c); }
uint32_t orc_permute(uint32_t i, uint32_t l, uint32_t p) { return dm_permute(i, l, p); }
void orc_sample_02(uint32_t n, uint32_t scr0, uint32_t scr1, float* out2) { out2[0] = van_der_corput(n, scr0); out2[1] = sobol(n, scr1); }
uint32_t orc_morton2(uint32_t x, uint32_t y) { return morton2(x, y); }

/* BSDF probe: builds Material::bsdf for a canonical frame (n = +z, dp_du = +x) and
 * evaluates eval / pdf / sample. out: f_eval[3], pdf, f_sample[3], wi[3], pdf_sample, sampled_type */
int orc_bsdf_probe(const trb_material* m, const float* merl_table, const float* wo, const float* wi, uint32_t flags,
                   const float* u3, float* out12) {
    Material mat; mat.type = m->type; mat.c0 = Col(m->c0[0], m->c0[1], m->c0[2]); mat.c1 = Col(m->c1[0], m->c1[1], m->c1[2]);
    mat.roughness = m->roughness; mat.eta = m->eta; mat.merl = merl_table;
    DG dg = DG::with_normal(V3(0.0f), V3(0.0f, 0.0f, 1.0f), 0.0f, 0.0f, 0.0f, V3(1.0f, 0.0f, 0.0f), V3(0.0f, 1.0f, 0.0f));
    BSDF b; mat.bsdf(dg, b);
    V3 o(wo[0], wo[1], wo[2]), i(wi[0], wi[1], wi[2]);
    Col f = b.eval(o, i, flags);
    out12[0] = f.r; out12[1] = f.g; out12[2] = f.b; out12[3] = b.pdf(o, i, flags);
    Col fs; V3 ws; float ps; uint32_t st;
    b.sample(o, flags, u3[0], u3[1], u3[2], fs, ws, ps, st);
    out12[4] = fs.r; out12[5] = fs.g; out12[6] = fs.b; out12[7] = ws.x; out12[8] = ws.y; out12[9] = ws.z; out12[10] = ps; out12[11] = (float)st;
    return TRB_OK;
}

/* linalg probes for the restated reference unit tests (transform.rs:284-379, matrix4.rs:265-305) */
void orc_m4_mul(const float* a, const float* b, float* out) { M4 x, y; memcpy(x.m, a, 64); memcpy(y.m, b, 64); M4 r = x * y; memcpy(out, r.m, 64); }
void orc_m4_inverse(const float* a, float* out) { M4 x; memcpy(x.m, a, 64); M4 r = x.inverse(); memcpy(out, r.m, 64); }
void orc_keyframe_transform(const trb_keyframe* kf, float* mat16, float* inv16) {
    Keyframe k; k.translation = V3(kf->translation[0], kf->translation[1], kf->translation[2]);
    k.rotation = Quat{V3(kf->rotation[0], kf->rotation[1], kf->rotation[2]), kf->rotation[3]};
    k.scaling = V3(kf->scaling[0], kf->scaling[1], kf->scaling[2]);
    Transform t = k.transform(); memcpy(mat16, t.mat.m, 64); memcpy(inv16, t.inv.m, 64);
}
/* partition.rs test_partition over u32 with predicate "even" */
size_t orc_partition_even(uint32_t* v, size_t n) {
    std::vector<BVH::GeomInfo> g(n);
    for (size_t i = 0; i < n; ++i) g[i].geom_idx = v[i];
    size_t r = BVH::partition(g.data(), n, [](const BVH::GeomInfo& x) { return x.geom_idx % 2 == 0; });
    for (size_t i = 0; i < n; ++i) v[i] = g[i].geom_idx;
    return r;
}

} // extern "C"

extern "C" {
/* linalg::cross / dot (src/linalg/mod.rs:38-45) for the restated reference unit tests */
void orc_cross_dot(const float* a, const float* b, float* out4) {
    V3 c = cross(V3(a[0], a[1], a[2]), V3(b[0], b[1], b[2]));
    out4[0] = c.x; out4[1] = c.y; out4[2] = c.z; out4[3] = dot(V3(a[0], a[1], a[2]), V3(b[0], b[1], b[2]));
}
/* Transform * {Point, Vector, Normal} with a keyframe's transform (transform.rs:199-242) */
void orc_xf_apply(const trb_keyframe* kf, const float* v, float* out9) {
    Keyframe k; k.translation = V3(kf->translation[0], kf->translation[1], kf->translation[2]);
    k.rotation = Quat{V3(kf->rotation[0], kf->rotation[1], kf->rotation[2]), kf->rotation[3]};
    k.scaling = V3(kf->scaling[0], kf->scaling[1], kf->scaling[2]);
    Transform t = k.transform();
    V3 p = t.point(V3(v[0], v[1], v[2])), w = t.vector(V3(v[0], v[1], v[2])), n = t.normal(V3(v[0], v[1], v[2]));
    out9[0] = p.x; out9[1] = p.y; out9[2] = p.z; out9[3] = w.x; out9[4] = w.y; out9[5] = w.z; out9[6] = n.x; out9[7] = n.y; out9[8] = n.z;
}
}
