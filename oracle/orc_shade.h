/* oracle/orc_shade.h — TEST INFRASTRUCTURE (parity oracle), not product code.
 * CPU restatement of /root/reference/src/bxdf/**, src/material/*.rs, src/light/mod.rs,
 * src/geometry/emitter.rs (Light impl), src/integrator/{mod,path}.rs, src/sampler/ld.rs.
 * Colours carry rgb only: alpha never reaches the film. */
#pragma once
#include "orc_geom.h"

namespace orc {

struct Col {
    float r, g, b;
    Col() : r(0), g(0), b(0) {}
    Col(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {}
    explicit Col(float v) : r(v), g(v), b(v) {}
    bool is_black() const { return r == 0.0f && g == 0.0f && b == 0.0f; } /* color.rs:49-51 */
    float luminance() const { return 0.2126f * r + 0.7152f * g + 0.0722f * b; } /* color.rs:45-47 */
    Col clamp() const { return Col(clampf(r, 0.0f, 1.0f), clampf(g, 0.0f, 1.0f), clampf(b, 0.0f, 1.0f)); }
};
static inline Col operator+(Col a, Col b) { return Col(a.r + b.r, a.g + b.g, a.b + b.b); }
static inline Col operator-(Col a, Col b) { return Col(a.r - b.r, a.g - b.g, a.b - b.b); }
static inline Col operator*(Col a, Col b) { return Col(a.r * b.r, a.g * b.g, a.b * b.b); }
static inline Col operator*(Col a, float s) { return Col(a.r * s, a.g * s, a.b * s); }
static inline Col operator/(Col a, Col b) { return Col(a.r / b.r, a.g / b.g, a.b / b.b); }
static inline Col operator/(Col a, float s) { return Col(a.r / s, a.g / s, a.b / s); }

/* BxDFType bitset (bxdf/mod.rs:37-82) */
enum { BX_REFLECTION = 1, BX_TRANSMISSION = 2, BX_DIFFUSE = 4, BX_GLOSSY = 8, BX_SPECULAR = 16 };
static const uint32_t BX_ALL = 31;
static const uint32_t BX_NON_SPECULAR = BX_DIFFUSE | BX_GLOSSY | BX_REFLECTION | BX_TRANSMISSION;

/* trig helpers (bxdf/mod.rs:125-166) */
static inline float cos_theta(V3 v) { return v.z; }
static inline float cos_theta_sqr(V3 v) { return v.z * v.z; }
static inline float sin_theta_sqr(V3 v) { return fmaxf(0.0f, 1.0f - v.z * v.z); }
static inline float sin_theta(V3 v) { return sqrtf(sin_theta_sqr(v)); }
static inline float tan_theta(V3 v) { float s2 = sin_theta_sqr(v); return s2 <= 0.0f ? 0.0f : sqrtf(s2) / cos_theta(v); }
static inline float tan_theta_sqr(V3 v) { return sin_theta_sqr(v) / cos_theta_sqr(v); }
static inline float cos_phi(V3 v) { float s = sin_theta(v); return s == 0.0f ? 1.0f : clampf(v.x / s, -1.0f, 1.0f); }
static inline float sin_phi(V3 v) { float s = sin_theta(v); return s == 0.0f ? 0.0f : clampf(v.y / s, -1.0f, 1.0f); }
static inline bool same_hemisphere(V3 a, V3 b) { return a.z * b.z > 0.0f; }

/* fresnel.rs */
struct Fresnel {
    bool conductor = false;
    float eta_i = 1.0f, eta_t = 1.0f; /* dielectric */
    Col eta, k;                       /* conductor */
    Col eval(float cos_i) const {
        if (conductor) { /* fresnel.rs:19-28, 85-87 */
            float c = fabsf(cos_i);
            Col a = (eta * eta + k * k) * c * c;
            Col one(1.0f);
            Col r_par = (a - eta * c * 2.0f + one) / (a + eta * c * 2.0f + one);
            Col b = eta * eta + k * k;
            Col cc(c * c);
            Col r_perp = (b - eta * c * 2.0f + cc) / (b + eta * c * 2.0f + cc);
            return (r_par + r_perp) * 0.5f;
        }
        /* fresnel.rs:48-66 */
        float ci = clampf(cos_i, -1.0f, 1.0f);
        float ei = ci > 0.0f ? eta_i : eta_t, et = ci > 0.0f ? eta_t : eta_i;
        float sin_t = ei / et * sqrtf(fmaxf(0.0f, 1.0f - ci * ci));
        if (sin_t >= 1.0f) return Col(1.0f);
        float ct = sqrtf(fmaxf(0.0f, 1.0f - sin_t * sin_t));
        float aci = fabsf(ci);
        /* fresnel.rs:10-14 */
        float r_par = (et * aci - ei * ct) / (et * aci + ei * ct);
        float r_perp = (ei * aci - et * ct) / (ei * aci + et * ct);
        return Col(0.5f * (r_par * r_par + r_perp * r_perp));
    }
};

/* microfacet/beckmann.rs */
struct Beckmann {
    float width = 0;
    static Beckmann make(float w) { Beckmann b; b.width = fmaxf(w, 0.000001f); return b; }
    float normal_distribution(V3 w_h) const {
        float tan_sqr = tan_theta_sqr(w_h);
        if (std::isinf(tan_sqr)) return 0.0f;
        float c2 = cos_theta_sqr(w_h);
        float cos_theta_4 = c2 * c2;
        float width_sqr = width * width;
        return M_EXP(-tan_sqr / width_sqr) / (PI * width_sqr * cos_theta_4);
    }
    V3 sample(float u0, float u1) const {
        float log_sample = M_LOG(1.0f - u0);
        if (std::isinf(log_sample)) log_sample = 0.0f;
        float tan_theta_sqr_ = -(width * width) * log_sample;
        float phi = 2.0f * PI * u1;
        float cos_t = 1.0f / sqrtf(1.0f + tan_theta_sqr_);
        float sin_t = sqrtf(fmaxf(0.0f, 1.0f - cos_t * cos_t));
        return spherical_dir(sin_t, cos_t, phi);
    }
    float pdf(V3 w_h) const { return fabsf(w_h.z) * normal_distribution(w_h); }
    float monodir_shadowing(V3 v) const {
        float a = 1.0f / (width * fabsf(tan_theta(v)));
        if (a < 1.6f) { float a2 = a * a; return (3.535f * a + 2.181f * a2) / (1.0f + 2.276f * a + 2.577f * a2); }
        return 1.0f;
    }
    float shadowing_masking(V3 w_i, V3 w_o) const { return monodir_shadowing(w_i) * monodir_shadowing(w_o); }
};

enum LobeKind { L_LAMBERT, L_OREN_NAYAR, L_SPEC_REFL, L_SPEC_TRANS, L_TORRANCE_SPARROW, L_MICROFACET_TRANS, L_MERL };

struct Lobe {
    LobeKind kind;
    uint32_t type; /* BxDFType set */
    Col c;         /* reflectance / albedo / transmission */
    float a = 0, b = 0; /* Oren-Nayar */
    Fresnel fresnel;
    Beckmann mf;
    const float* merl = nullptr;

    bool matches(uint32_t flags) const { return (type & ~flags) == 0; } /* is_subset */

    /* ---- eval ---- */
    Col eval(V3 w_o, V3 w_i) const {
        switch (kind) {
            case L_LAMBERT: return c * FRAC_1_PI; /* lambertian.rs:32-34 */
            case L_OREN_NAYAR: { /* oren_nayar.rs:43-61 */
                float sin_theta_o = sin_theta(w_o), sin_theta_i = sin_theta(w_i);
                float max_cos = 0.0f;
                if (sin_theta_i > 1e-4f && sin_theta_o > 1e-4f)
                    max_cos = fmaxf(0.0f, cos_phi(w_i) * cos_phi(w_o) + sin_phi(w_i) * sin_phi(w_o));
                float sin_alpha, tan_beta;
                if (fabsf(cos_theta(w_i)) > fabsf(cos_theta(w_o))) { sin_alpha = sin_theta_o; tan_beta = sin_theta_i / fabsf(cos_theta(w_i)); }
                else { sin_alpha = sin_theta_i; tan_beta = sin_theta_o / fabsf(cos_theta(w_o)); }
                return c * FRAC_1_PI * (a + b * max_cos * sin_alpha * tan_beta);
            }
            case L_SPEC_REFL: case L_SPEC_TRANS: return Col(0.0f);
            case L_TORRANCE_SPARROW: { /* torrance_sparrow.rs:40-56 */
                float cos_to = fabsf(cos_theta(w_o)), cos_ti = fabsf(cos_theta(w_i));
                if (cos_to == 0.0f || cos_ti == 0.0f) return Col(0.0f);
                V3 w_h = w_i + w_o;
                if (w_h.x == 0.0f && w_h.y == 0.0f && w_h.z == 0.0f) return Col(0.0f);
                w_h = normalized(w_h);
                float d = mf.normal_distribution(w_h);
                Col f = fresnel.eval(dot(w_i, w_h));
                float g = mf.shadowing_masking(w_i, w_o);
                return c * f * d * g / (4.0f * cos_ti * cos_to);
            }
            case L_MICROFACET_TRANS: { /* microfacet_transmission.rs:65-82 */
                if (same_hemisphere(w_o, w_i)) return Col(0.0f);
                float cos_to = cos_theta(w_o), cos_ti = cos_theta(w_i);
                if (cos_to == 0.0f || cos_ti == 0.0f) return Col(0.0f);
                float e0, e1;
                eta_for_interaction(w_o, e0, e1);
                V3 w_h = mt_half_vector(w_o, w_i, e0, e1);
                float d = mf.normal_distribution(w_h);
                Col f = Col(1.0f) - fresnel.eval(dot(w_i, w_h));
                float g = mf.shadowing_masking(w_i, w_o);
                float wi_dot_h = dot(w_i, w_h);
                float jac = mt_jacobian(w_o, w_i, w_h, e0, e1);
                return c * (fabsf(wi_dot_h) / (fabsf(w_i.z) * fabsf(w_o.z))) * (f * g * d) * jac;
            }
            case L_MERL: return merl_eval(w_o, w_i);
        }
        return Col(0.0f);
    }
    void eta_for_interaction(V3 w_o, float& e0, float& e1) const { /* microfacet_transmission.rs:33-39 */
        if (cos_theta(w_o) > 0.0f) { e0 = fresnel.eta_i; e1 = fresnel.eta_t; } else { e0 = fresnel.eta_t; e1 = fresnel.eta_i; }
    }
    static float mt_jacobian(V3 w_o, V3 w_i, V3 w_h, float e0, float e1) { /* :40-49 */
        float wi_dot_h = dot(w_i, w_h), wo_dot_h = dot(w_o, w_h);
        float s = e1 * wi_dot_h + e0 * wo_dot_h;
        float denom = s * s;
        if (denom != 0.0f) return fabsf(e0 * e0 * fabsf(wo_dot_h) / denom);
        return 0.0f;
    }
    static V3 mt_half_vector(V3 w_o, V3 w_i, float e0, float e1) { return normalized(-e1 * w_i - e0 * w_o); } /* :50-52 */
    /* bxdf/merl.rs:47-82 */
    static uint32_t map_index(float val, float mx, uint32_t n_vals) {
        uint32_t i = f2u(val / mx * (float)n_vals);
        return i > n_vals - 1 ? n_vals - 1 : i;
    }
    Col merl_eval(V3 w_oi, V3 w_ii) const {
        V3 w_i = w_ii;
        V3 w_h = w_oi + w_i;
        if (w_h.z < 0.0f) { w_i = -w_i; w_h = -w_h; }
        if (length_sqr(w_h) == 0.0f) return Col(0.0f);
        w_h = normalized(w_h);
        float theta_h = spherical_theta(w_h);
        float cos_phi_h = cos_phi(w_h), sin_phi_h = sin_phi(w_h);
        float cos_theta_h = cos_theta(w_h), sin_theta_h = sin_theta(w_h);
        V3 w_hx(cos_phi_h * cos_theta_h, sin_phi_h * cos_theta_h, -sin_theta_h);
        V3 w_hy(-sin_phi_h, cos_phi_h, 0.0f);
        V3 w_d(dot(w_i, w_hx), dot(w_i, w_hy), dot(w_i, w_h));
        float theta_d = spherical_theta(w_d);
        float phi_d = spherical_phi(w_d);
        if (phi_d > PI) phi_d = phi_d - PI;
        uint32_t theta_h_idx = map_index(sqrtf(fmaxf(0.0f, 2.0f * theta_h / PI)), 1.0f, TRB_MERL_N_THETA_H);
        uint32_t theta_d_idx = map_index(theta_d, PI / 2.0f, TRB_MERL_N_THETA_D);
        uint32_t phi_d_idx = map_index(phi_d, PI, TRB_MERL_N_PHI_D);
        uint32_t i = phi_d_idx + TRB_MERL_N_PHI_D * (theta_d_idx + theta_h_idx * TRB_MERL_N_THETA_D);
        return Col(merl[3 * i], merl[3 * i + 1], merl[3 * i + 2]);
    }

    /* ---- pdf ---- */
    float pdf(V3 w_o, V3 w_i) const {
        switch (kind) {
            case L_TORRANCE_SPARROW: { /* torrance_sparrow.rs:73-81 */
                if (!same_hemisphere(w_o, w_i)) return 0.0f;
                V3 w_h = normalized(w_o + w_i);
                float jacobian = 1.0f / (4.0f * fabsf(dot(w_o, w_h)));
                return mf.pdf(w_h) * jacobian;
            }
            case L_MICROFACET_TRANS: { /* microfacet_transmission.rs:100-108 */
                if (same_hemisphere(w_o, w_i)) return 0.0f;
                float e0, e1;
                eta_for_interaction(w_o, e0, e1);
                V3 w_h = mt_half_vector(w_o, w_i, e0, e1);
                return mf.pdf(w_h) * mt_jacobian(w_o, w_i, w_h, e0, e1);
            }
            default: /* BxDF::pdf default (bxdf/mod.rs:114-121) — also what the specular lobes inherit */
                return same_hemisphere(w_o, w_i) ? fabsf(cos_theta(w_i)) * FRAC_1_PI : 0.0f;
        }
    }

    /* ---- sample: returns f, w_i, pdf ---- */
    void sample(V3 w_o, float u0, float u1, Col& f, V3& w_i, float& pdf_out) const {
        switch (kind) {
            case L_SPEC_REFL: { /* specular_reflection.rs:39-50 */
                w_i = V3(-w_o.x, -w_o.y, w_o.z);
                if (w_i.z != 0.0f) { f = fresnel.eval(cos_theta(w_o)) * c / fabsf(cos_theta(w_i)); pdf_out = 1.0f; }
                else { f = Col(0.0f); pdf_out = 0.0f; }
                return;
            }
            case L_SPEC_TRANS: { /* specular_transmission.rs:39-56 */
                bool entering = cos_theta(w_o) > 0.0f;
                float ei = entering ? fresnel.eta_i : fresnel.eta_t, et = entering ? fresnel.eta_t : fresnel.eta_i;
                V3 n = entering ? V3(0.0f, 0.0f, 1.0f) : V3(0.0f, 0.0f, -1.0f);
                V3 r;
                if (refract(w_o, n, ei / et, r)) {
                    w_i = r;
                    Col fr = Col(1.0f) - fresnel.eval(cos_theta(w_i));
                    f = fr * c / fabsf(cos_theta(w_i)); pdf_out = 1.0f;
                } else { f = Col(0.0f); w_i = V3(0.0f); pdf_out = 0.0f; }
                return;
            }
            case L_TORRANCE_SPARROW: { /* torrance_sparrow.rs:57-72 */
                if (w_o.z == 0.0f) { f = Col(0.0f); w_i = V3(0.0f); pdf_out = 0.0f; return; }
                V3 w_h = mf.sample(u0, u1);
                if (!same_hemisphere(w_o, w_h)) w_h = -w_h;
                w_i = reflect(w_o, w_h);
                if (!same_hemisphere(w_o, w_i)) { f = Col(0.0f); w_i = V3(0.0f); pdf_out = 0.0f; }
                else { f = eval(w_o, w_i); pdf_out = pdf(w_o, w_i); }
                return;
            }
            case L_MICROFACET_TRANS: { /* microfacet_transmission.rs:83-99 */
                V3 w_h = mf.sample(u0, u1);
                if (!same_hemisphere(w_o, w_h)) w_h = -w_h;
                float e0, e1;
                eta_for_interaction(w_o, e0, e1);
                V3 r;
                if (refract(w_o, w_h, e0 / e1, r)) {
                    if (same_hemisphere(w_o, r)) { f = Col(0.0f); w_i = V3(0.0f); pdf_out = 0.0f; }
                    else { w_i = r; f = eval(w_o, w_i); pdf_out = pdf(w_o, w_i); }
                } else { f = Col(0.0f); w_i = V3(0.0f); pdf_out = 0.0f; }
                return;
            }
            default: { /* BxDF::sample default (bxdf/mod.rs:102-108): Lambertian, Oren-Nayar, Merl */
                w_i = cos_sample_hemisphere(u0, u1);
                if (w_o.z < 0.0f) w_i.z *= -1.0f;
                f = eval(w_o, w_i); pdf_out = pdf(w_o, w_i);
                return;
            }
        }
    }
};

/* bxdf::BSDF (bsdf.rs) */
struct BSDF {
    V3 p, n, ng, tan, bitan;
    float eta = 1.0f;
    Lobe lobes[2];
    int n_lobes = 0;

    void init_frame(const DG& dg) { /* bsdf.rs:38-44 */
        n = normalized(dg.n);
        V3 bt = normalized(dg.dp_du);
        tan = cross(n, bt);
        bitan = cross(tan, n);
        p = dg.p; ng = dg.ng;
    }
    int num_matching(uint32_t flags) const { int c = 0; for (int i = 0; i < n_lobes; ++i) c += lobes[i].matches(flags); return c; }
    V3 to_shading(V3 v) const { return V3(dot(v, bitan), dot(v, tan), dot(v, n)); }
    V3 from_shading(V3 v) const {
        return V3(bitan.x * v.x + tan.x * v.y + n.x * v.z, bitan.y * v.x + tan.y * v.y + n.y * v.z, bitan.z * v.x + tan.z * v.y + n.z * v.z);
    }
    Col eval(V3 wo_world, V3 wi_world, uint32_t flags) const { /* bsdf.rs:66-78 */
        V3 w_o = normalized(to_shading(wo_world)), w_i = normalized(to_shading(wi_world));
        if (w_o.z * w_i.z > 0.0f) flags &= ~(uint32_t)BX_TRANSMISSION; else flags &= ~(uint32_t)BX_REFLECTION;
        Col acc(0.0f);
        for (int i = 0; i < n_lobes; ++i) if (lobes[i].matches(flags)) acc = acc + lobes[i].eval(w_o, w_i);
        return acc;
    }
    float pdf(V3 wo_world, V3 wi_world, uint32_t flags) const { /* bsdf.rs:114-125 */
        V3 w_o = normalized(to_shading(wo_world)), w_i = normalized(to_shading(wi_world));
        float pdf_val = 0.0f; int n_comps = 0;
        for (int i = 0; i < n_lobes; ++i) if (lobes[i].matches(flags)) { pdf_val = pdf_val + lobes[i].pdf(w_o, w_i); n_comps++; }
        return n_comps > 0 ? pdf_val / (float)n_comps : 0.0f;
    }
    /* bsdf.rs:85-112 */
    void sample(V3 wo_world, uint32_t flags, float u0, float u1, float one_d, Col& f, V3& wi_world, float& pdf_out, uint32_t& sampled) const {
        int n_matching = num_matching(flags);
        if (n_matching == 0) { f = Col(0.0f); wi_world = V3(0.0f); pdf_out = 0.0f; sampled = 0; return; }
        uint32_t comp = f2u(one_d * (float)n_matching);
        if (comp > (uint32_t)n_matching - 1) comp = n_matching - 1;
        const Lobe* bx = nullptr;
        for (int i = 0, k = 0; i < n_lobes; ++i) if (lobes[i].matches(flags)) { if ((uint32_t)k == comp) { bx = &lobes[i]; break; } k++; }
        V3 w_o = normalized(to_shading(wo_world));
        V3 w_i;
        bx->sample(w_o, u0, u1, f, w_i, pdf_out);
        if (length_sqr(w_i) == 0.0f) { f = Col(0.0f); wi_world = V3(0.0f); pdf_out = 0.0f; sampled = 0; return; }
        wi_world = normalized(from_shading(w_i));
        bool spec = (bx->type & BX_SPECULAR) != 0;
        if (!spec && n_matching > 1) pdf_out = pdf(wo_world, wi_world, flags);
        if (!spec) f = eval(wo_world, wi_world, flags);
        sampled = bx->type;
    }
};

/* ---- textures (texture/mod.rs:21-41 bilinear_interpolate, image.rs:14-48, animated_image.rs:18-60) ------------------ */
struct TexImage { uint32_t w = 0, h = 0; std::vector<uint8_t> px; float time = 0.0f; };
struct TextureSet {
    std::vector<trb_texture> tex;
    std::vector<TexImage> img;
    static uint32_t clampu(uint32_t x, uint32_t hi) { return x > hi ? hi : x; }
    static float get_float(const TexImage& im, uint32_t x, uint32_t y) { /* Image::get_float: channel 0 */
        x = clampu(x, im.w - 1); y = clampu(y, im.h - 1);
        return (float)im.px[4 * ((size_t)y * im.w + x)] / 255.0f;
    }
    static Col get_color(const TexImage& im, uint32_t x, uint32_t y) { /* Image::get_color (alpha is carried by Colorf but never read by a BxDF) */
        x = clampu(x, im.w - 1); y = clampu(y, im.h - 1);
        const uint8_t* p = &im.px[4 * ((size_t)y * im.w + x)];
        return Col((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f);
    }
    static float image_f32(const TexImage& im, float u, float v) { /* Image::sample_f32 */
        float x = u * (float)im.w, y = v * (float)im.h;
        uint32_t x0 = f2u(x), y0 = f2u(y);
        float s00 = get_float(im, x0, y0), s10 = get_float(im, x0 + 1, y0), s01 = get_float(im, x0, y0 + 1), s11 = get_float(im, x0 + 1, y0 + 1);
        float sx = x - (float)x0, sy = y - (float)y0;
        return s00 * (1.0f - sx) * (1.0f - sy) + s10 * sx * (1.0f - sy) + s01 * (1.0f - sx) * sy + s11 * sx * sy;
    }
    static Col image_color(const TexImage& im, float u, float v) { /* Image::sample_color */
        float x = u * (float)im.w, y = v * (float)im.h;
        uint32_t x0 = f2u(x), y0 = f2u(y);
        Col s00 = get_color(im, x0, y0), s10 = get_color(im, x0 + 1, y0), s01 = get_color(im, x0, y0 + 1), s11 = get_color(im, x0 + 1, y0 + 1);
        float sx = x - (float)x0, sy = y - (float)y0;
        return s00 * (1.0f - sx) * (1.0f - sy) + s10 * sx * (1.0f - sy) + s01 * (1.0f - sx) * sy + s11 * sx * sy;
    }
    /* AnimatedImage::active_keyframes (animated_image.rs:23-37): binary search over the keyframe times */
    void active(const trb_texture& t, float time, uint32_t& lo, int& hi) const {
        uint32_t a = 0, b = t.n_images; /* first index whose time is not < `time` */
        while (a < b) { uint32_t m = (a + b) / 2; if (img[t.first_image + m].time < time) a = m + 1; else b = m; }
        if (a < t.n_images && img[t.first_image + a].time == time) { lo = a; hi = -1; }
        else if (a == t.n_images) { lo = a - 1; hi = -1; }
        else if (a == 0) { lo = 0; hi = -1; }
        else { lo = a - 1; hi = (int)a; }
    }
    float sample_f32(uint32_t ti, float u, float v, float time) const {
        const trb_texture& t = tex[ti];
        if (t.n_images == 1) return image_f32(img[t.first_image], u, v);
        uint32_t lo; int hi;
        active(t, time, lo, hi);
        const TexImage& a = img[t.first_image + lo];
        if (hi < 0) return image_f32(a, u, v);
        const TexImage& b = img[t.first_image + hi];
        float x = (time - a.time) / (b.time - a.time);
        return image_f32(a, u, v) * (1.0f - x) + image_f32(b, u, v) * x; /* linalg::lerp */
    }
    Col sample_color(uint32_t ti, float u, float v, float time) const {
        const trb_texture& t = tex[ti];
        if (t.n_images == 1) return image_color(img[t.first_image], u, v);
        uint32_t lo; int hi;
        active(t, time, lo, hi);
        const TexImage& a = img[t.first_image + lo];
        if (hi < 0) return image_color(a, u, v);
        const TexImage& b = img[t.first_image + hi];
        float x = (time - a.time) / (b.time - a.time);
        return image_color(a, u, v) * (1.0f - x) + image_color(b, u, v) * x;
    }
};

struct Material {
    uint32_t type = 0;
    Col c0, c1;
    float roughness = 0, eta = 1;
    const float* merl = nullptr;
    uint32_t tex[4] = {0, 0, 0, 0};      /* 1 + texture index bound to c0 / c1 / roughness / eta, 0 = constant */
    const TextureSet* textures = nullptr;

    /* Material::bsdf (material/{matte:52,plastic:59,metal:56,specular_metal:49,glass:51,rough_glass:57,merl:88}.rs): every parameter is
     * texture.sample_color / sample_f32(hit.dg.u, hit.dg.v, hit.dg.time); constants return themselves */
    void bsdf(const DG& dg, BSDF& out) const {
        if (tex[0] | tex[1] | tex[2] | tex[3]) {
            Material m = *this;
            m.tex[0] = m.tex[1] = m.tex[2] = m.tex[3] = 0;
            if (tex[0]) m.c0 = textures->sample_color(tex[0] - 1, dg.u, dg.v, dg.time);
            if (tex[1]) m.c1 = textures->sample_color(tex[1] - 1, dg.u, dg.v, dg.time);
            if (tex[2]) m.roughness = textures->sample_f32(tex[2] - 1, dg.u, dg.v, dg.time);
            if (tex[3]) m.eta = textures->sample_f32(tex[3] - 1, dg.u, dg.v, dg.time);
            m.bsdf(dg, out);
            return;
        }
        out.n_lobes = 0; out.eta = 1.0f;
        auto push = [&](const Lobe& l) { out.lobes[out.n_lobes++] = l; };
        switch (type) {
            case TRB_MAT_MATTE: {
                Lobe l; l.c = c0; l.type = BX_DIFFUSE | BX_REFLECTION;
                if (roughness == 0.0f) l.kind = L_LAMBERT;
                else { /* OrenNayar::new (oren_nayar.rs:26-34): roughness in degrees (Q17) */
                    l.kind = L_OREN_NAYAR;
                    float sigma = to_radians(roughness);
                    sigma *= sigma;
                    l.a = 1.0f - 0.5f * sigma / (sigma + 0.33f);
                    l.b = 0.45f * sigma / (sigma + 0.09f);
                }
                push(l); break;
            }
            case TRB_MAT_PLASTIC: {
                if (!c0.is_black()) { Lobe l; l.kind = L_LAMBERT; l.c = c0; l.type = BX_DIFFUSE | BX_REFLECTION; push(l); }
                if (!c1.is_black()) { /* Dielectric(1.0, 1.5) hard-wired (Q16, plastic.rs:83) */
                    Lobe l; l.kind = L_TORRANCE_SPARROW; l.c = c1; l.type = BX_GLOSSY | BX_REFLECTION;
                    l.fresnel.conductor = false; l.fresnel.eta_i = 1.0f; l.fresnel.eta_t = 1.5f;
                    l.mf = Beckmann::make(roughness); push(l);
                }
                break;
            }
            case TRB_MAT_METAL: {
                Lobe l; l.kind = L_TORRANCE_SPARROW; l.c = Col(1.0f); l.type = BX_GLOSSY | BX_REFLECTION;
                l.fresnel.conductor = true; l.fresnel.eta = c0; l.fresnel.k = c1; l.mf = Beckmann::make(roughness); push(l); break;
            }
            case TRB_MAT_SPECULAR_METAL: {
                Lobe l; l.kind = L_SPEC_REFL; l.c = Col(1.0f); l.type = BX_SPECULAR | BX_REFLECTION;
                l.fresnel.conductor = true; l.fresnel.eta = c0; l.fresnel.k = c1; push(l); break;
            }
            case TRB_MAT_GLASS: {
                Fresnel fr; fr.conductor = false; fr.eta_i = 1.0f; fr.eta_t = eta;
                if (!c0.is_black()) { Lobe l; l.kind = L_SPEC_REFL; l.c = c0; l.type = BX_SPECULAR | BX_REFLECTION; l.fresnel = fr; push(l); }
                if (!c1.is_black()) { Lobe l; l.kind = L_SPEC_TRANS; l.c = c1; l.type = BX_SPECULAR | BX_TRANSMISSION; l.fresnel = fr; push(l); }
                out.eta = eta; break;
            }
            case TRB_MAT_ROUGH_GLASS: {
                Fresnel fr; fr.conductor = false; fr.eta_i = 1.0f; fr.eta_t = eta;
                Beckmann mf = Beckmann::make(roughness);
                if (!c0.is_black()) { Lobe l; l.kind = L_TORRANCE_SPARROW; l.c = c0; l.type = BX_GLOSSY | BX_REFLECTION; l.fresnel = fr; l.mf = mf; push(l); }
                if (!c1.is_black()) { Lobe l; l.kind = L_MICROFACET_TRANS; l.c = c1; l.type = BX_GLOSSY | BX_TRANSMISSION; l.fresnel = fr; l.mf = mf; push(l); }
                out.eta = eta; break;
            }
            case TRB_MAT_MERL: {
                Lobe l; l.kind = L_MERL; l.type = BX_GLOSSY | BX_REFLECTION; l.merl = merl; push(l); break;
            }
        }
        out.init_frame(dg);
    }
};

/* ---- sampler::ld (ld.rs:91-119) --------------------------------------------------- */
static inline float van_der_corput(uint32_t n, uint32_t scramble) {
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
    n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
    n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
    n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
    n ^= scramble;
    return fminf((float)((n >> 8) & 0xffffffu) / (float)(1 << 24), 1.0f - F32_EPSILON);
}
static inline float sobol(uint32_t n, uint32_t scramble) {
    uint32_t i = 1u << 31;
    while (n != 0) {
        if (n & 1u) scramble ^= i;
        n >>= 1;
        i ^= i >> 1;
    }
    return fminf((float)((scramble >> 8) & 0xffffffu) / (float)(1 << 24), 1.0f - F32_EPSILON);
}

/* The six per-path sample arrays of Path::illumination (path.rs:48-60), evaluated lazily
 * per bounce through the counter RNG (DESIGN.md "RNG"): entry [b] of an array that was
 * filled with sample_02(i, scramble) and shuffled == sample_02(perm(b), scramble). */
struct PathSamples {
    uint32_t seed, pixel, sample, len;
    uint32_t draw(uint32_t dim) const { return dm_rng(seed, pixel, sample, dim); }
    void two_d(uint32_t b, uint32_t d0, uint32_t d1, uint32_t dperm, float& x, float& y) const {
        uint32_t i = dm_permute(b, len, draw(dperm));
        x = van_der_corput(i, dm_scramble(draw(d0)));
        y = sobol(i, dm_scramble(draw(d1)));
    }
    float one_d(uint32_t b, uint32_t d0, uint32_t dperm) const {
        uint32_t i = dm_permute(b, len, draw(dperm));
        return van_der_corput(i, dm_scramble(draw(d0)));
    }
    float rr(uint32_t bounce) const { return dm_next_f32(draw(DM_S_RR + bounce)); }
};

/* ---- lights: impl Light for Emitter (emitter.rs:140-204) -------------------------- */
struct SceneShade {
    const SceneGeom* geom = nullptr;
    std::vector<Material> materials;
    std::vector<uint32_t> lights; /* instance indices of emitters in object order (Q20) */
    uint32_t min_depth = 0, max_depth = 0;
    uint32_t integrator = TRB_INTEGRATOR_PATH; /* Whitted: max_depth is its recursion limit (scene.rs:305-309 reads it from "min_depth") */

    /* Emitter::radiance (emitter.rs:140-142) */
    Col radiance(const Instance& e, V3 w, V3 n, float time) const {
        if (dot(w, n) > 0.0f) { float c[4]; e.emission.color(time, c); return Col(c[0], c[1], c[2]); }
        return Col(0.0f);
    }
    bool occluded(const Ray& r, Counters& cnt) const { /* light/mod.rs:30-37: full closest-hit (Q6) */
        Ray ray = r; Hit h;
        cnt.rays[1]++;
        return geom->intersect(ray, h, cnt);
    }
    void sample_incident(uint32_t li, V3 p, float u0, float u1, float time, Col& rad, V3& w_i, float& pdf, Ray& occl) const {
        const Instance& e = geom->instances[li];
        Transform t = geom->xf(li, time);
        if (e.kind == TRB_INST_EMITTER_POINT) { /* emitter.rs:169-174 */
            V3 pos = t.point(V3(0.0f));
            w_i = normalized(pos - p);
            float c[4]; e.emission.color(time, c);
            rad = Col(c[0], c[1], c[2]) / distance_sqr(pos, p);
            pdf = 1.0f;
            occl = Ray::segment(p, pos - p, 0.001f, 0.999f, time);
            return;
        }
        /* emitter.rs:175-185 (object-space pdf, not corrected for scale: Q5) */
        V3 p_l = t.inv_point(p);
        V3 p_sampled, normal;
        e.shape.sample(p_l, u0, u1, p_sampled, normal);
        V3 w_il = normalized(p_sampled - p_l);
        pdf = e.shape.pdf(p_l, w_il);
        rad = radiance(e, -w_il, normal, time);
        V3 p_w = t.point(p_sampled);
        w_i = t.vector(w_il);
        occl = Ray::segment(p, p_w - p, 0.001f, 0.999f, time); /* OcclusionTester::test_points (light/mod.rs:21-23) */
    }
    float light_pdf(uint32_t li, V3 p, V3 w_i, float time) const { /* emitter.rs:193-203 */
        const Instance& e = geom->instances[li];
        if (e.kind == TRB_INST_EMITTER_POINT) return 0.0f;
        Transform t = geom->xf(li, time);
        V3 p_l = t.inv_point(p);
        V3 w = normalized(t.inv_vector(w_i));
        return e.shape.pdf(p_l, w);
    }

    /* Integrator::estimate_direct (integrator/mod.rs:122-169) */
    Col estimate_direct(V3 w_o, V3 p, const BSDF& bsdf, float l0, float l1, float b0, float b1, float bc, uint32_t li,
                        uint32_t flags, float time, Counters& cnt) const {
        Col direct_light(0.0f);
        const Instance& light = geom->instances[li];
        bool delta = light.kind == TRB_INST_EMITTER_POINT;
        Col lrad; V3 w_i; float pdf_light; Ray occl;
        sample_incident(li, bsdf.p, l0, l1, time, lrad, w_i, pdf_light, occl);
        if (pdf_light > 0.0f && !lrad.is_black() && !occluded(occl, cnt)) {
            Col f = bsdf.eval(w_o, w_i, flags);
            if (!f.is_black()) {
                if (delta) direct_light = f * lrad * fabsf(dot(w_i, bsdf.n)) / pdf_light;
                else {
                    float pdf_bsdf = bsdf.pdf(w_o, w_i, flags);
                    float w = power_heuristic(1.0f, pdf_light, 1.0f, pdf_bsdf);
                    direct_light = f * lrad * fabsf(dot(w_i, bsdf.n)) * w / pdf_light;
                }
            }
        }
        if (!delta) {
            Col f; V3 wi2; float pdf_bsdf; uint32_t sampled;
            bsdf.sample(w_o, flags, b0, b1, bc, f, wi2, pdf_bsdf, sampled);
            if (pdf_bsdf > 0.0f && !f.is_black()) {
                float w = 1.0f;
                if (!(sampled & BX_SPECULAR)) {
                    float pl = light_pdf(li, p, wi2, time);
                    if (pl == 0.0f) return direct_light; /* Q7 */
                    w = power_heuristic(1.0f, pdf_bsdf, 1.0f, pl);
                }
                Ray ray = Ray::segment(p, wi2, 0.001f, F32_INF, time);
                Col lr(0.0f);
                Hit h;
                cnt.rays[2]++;
                if (geom->intersect(ray, h, cnt)) {
                    if (h.inst == li) lr = radiance(light, -wi2, h.dg.ng, time);
                }
                if (!lr.is_black()) direct_light = direct_light + f * lr * fabsf(dot(wi2, bsdf.n)) * w / pdf_bsdf;
            }
        }
        return direct_light;
    }

    /* NormalsDebug::illumination (integrator/normals_debug.rs:28-36) */
    Col normals_debug(const Hit& hit) const {
        BSDF bsdf;
        materials[geom->instances[hit.inst].material].bsdf(hit.dg, bsdf);
        return (Col(bsdf.n.x, bsdf.n.y, bsdf.n.z) + Col(1.0f)) / 2.0f;
    }
    /* One 1-element sampler array: get_samples_2d / get_samples_1d with a fresh scramble and index 0 (ld.rs:55-64, 91-119) */
    static void whitted_2d(const PathSamples& ps, uint32_t node, uint32_t slot, float& x, float& y) {
        x = van_der_corput(0, dm_scramble(ps.draw(DM_S_WHITTED + 8 * node + slot)));
        y = sobol(0, dm_scramble(ps.draw(DM_S_WHITTED + 8 * node + slot + 1)));
    }
    static float whitted_1d(const PathSamples& ps, uint32_t node, uint32_t slot) { return van_der_corput(0, dm_scramble(ps.draw(DM_S_WHITTED + 8 * node + slot))); }
    /* Integrator::specular_reflection / specular_transmission (integrator/mod.rs:41-103); `flags` = {Specular, Reflection | Transmission} */
    Col whitted_specular(const Ray& ray, uint32_t depth, const BSDF& bsdf, uint32_t flags, uint32_t child, uint32_t slot, uint32_t node, const PathSamples& ps,
                         Counters& cnt) const {
        V3 w_o = -ray.d;
        float u0, u1;
        whitted_2d(ps, node, slot, u0, u1);
        float uc = whitted_1d(ps, node, slot + 2);
        Col f; V3 w_i; float pdf; uint32_t sampled;
        bsdf.sample(w_o, flags, u0, u1, uc, f, w_i, pdf, sampled);
        Col out(0.0f);
        if (pdf > 0.0f && !f.is_black() && fabsf(dot(w_i, bsdf.n)) != 0.0f) {
            Ray r2(bsdf.p, w_i, ray.time); /* ray.child(&bsdf.p, &w_i): direction as sampled */
            r2.min_t = 0.001f;
            Hit h;
            cnt.rays[3]++;
            if (geom->intersect(r2, h, cnt)) {
                Col li = whitted(r2, depth + 1, h, child, ps, cnt);
                out = f * li * fabsf(dot(w_i, bsdf.n)) / pdf;
            }
        }
        return out;
    }
    /* Whitted::illumination (integrator/whitted.rs:41-70) */
    Col whitted(const Ray& ray, uint32_t depth, const Hit& hit, uint32_t node, const PathSamples& ps, Counters& cnt) const {
        const Instance& inst = geom->instances[hit.inst];
        BSDF bsdf;
        materials[inst.material].bsdf(hit.dg, bsdf);
        V3 w_o = -ray.d;
        float u0, u1;
        whitted_2d(ps, node, 0, u0, u1);
        Col illum(0.0f);
        if (depth == 0 && inst.is_emitter()) illum = illum + radiance(inst, -ray.d, hit.dg.ng, ray.time);
        for (uint32_t li : lights) {
            Col lrad; V3 w_i; float pdf; Ray occl;
            sample_incident(li, hit.dg.p, u0, u1, ray.time, lrad, w_i, pdf, occl);
            Col f = bsdf.eval(w_o, w_i, BX_ALL);
            if (!lrad.is_black() && !f.is_black() && !occluded(occl, cnt)) illum = illum + f * lrad * fabsf(dot(w_i, bsdf.n)) / pdf;
        }
        if (depth < max_depth) {
            illum = illum + whitted_specular(ray, depth, bsdf, BX_SPECULAR | BX_REFLECTION, 2 * node, 2, node, ps, cnt);
            illum = illum + whitted_specular(ray, depth, bsdf, BX_SPECULAR | BX_TRANSMISSION, 2 * node + 1, 5, node, ps, cnt);
        }
        return illum;
    }

    /* Path::illumination (integrator/path.rs:45-119) */
    Col illumination(const Ray& r, const Hit& hit, const PathSamples& ps, Counters& cnt) const {
        if (integrator == TRB_INTEGRATOR_NORMALS_DEBUG) return normals_debug(hit);
        if (integrator == TRB_INTEGRATOR_WHITTED) return whitted(r, 0, hit, 1, ps, cnt);
        Col illum(0.0f), path_throughput(1.0f);
        bool specular_bounce = false;
        Hit current_hit = hit;
        Ray ray = r;
        uint32_t bounce = 0;
        for (;;) {
            const Instance& inst = geom->instances[current_hit.inst];
            if (bounce == 0 || specular_bounce) {
                if (inst.is_emitter()) {
                    V3 w = -ray.d;
                    illum = illum + path_throughput * radiance(inst, w, hit.dg.ng, ray.time); /* hit, not current_hit: Q1 */
                }
            }
            BSDF bsdf;
            materials[inst.material].bsdf(current_hit.dg, bsdf);
            V3 w_o = -ray.d;
            float l0, l1, b0, b1;
            ps.two_d(bounce, DM_S_L0, DM_S_L1, DM_S_L_PERM, l0, l1);
            ps.two_d(bounce, DM_S_B0, DM_S_B1, DM_S_B_PERM, b0, b1);
            float lc = ps.one_d(bounce, DM_S_LC, DM_S_LC_PERM);
            float bc = ps.one_d(bounce, DM_S_BC, DM_S_BC_PERM);
            /* sample_one_light (integrator/mod.rs:106-111): no xN weight (Q2) */
            uint32_t nl = (uint32_t)lights.size();
            uint32_t l = f2u(lc * (float)nl);
            if (l > nl - 1) l = nl - 1;
            Col li = estimate_direct(w_o, current_hit.dg.p, bsdf, l0, l1, b0, b1, bc, lights[l], BX_NON_SPECULAR, ray.time, cnt);
            illum = illum + path_throughput * li;

            float p0, p1;
            ps.two_d(bounce, DM_S_P0, DM_S_P1, DM_S_P_PERM, p0, p1);
            float pc = ps.one_d(bounce, DM_S_PC, DM_S_PC_PERM);
            Col f; V3 w_i; float pdf; uint32_t sampled;
            bsdf.sample(w_o, BX_ALL, p0, p1, pc, f, w_i, pdf, sampled);
            if (f.is_black() || pdf == 0.0f) break;
            specular_bounce = (sampled & BX_SPECULAR) != 0;
            path_throughput = path_throughput * f * fabsf(dot(w_i, bsdf.n)) / pdf;
            if (bounce > min_depth) {
                float cont_prob = fmaxf(0.5f, path_throughput.luminance());
                if (ps.rr(bounce) > cont_prob) break; /* Q8 */
                path_throughput = path_throughput / cont_prob;
            }
            if (bounce == max_depth) break;
            ray = Ray(bsdf.p, normalized(w_i), ray.time); /* ray.child */
            ray.min_t = 0.001f;
            Hit h;
            cnt.rays[3]++;
            if (!geom->intersect(ray, h, cnt)) break;
            current_hit = h;
            bounce += 1;
        }
        return illum;
    }
};

} // namespace orc
