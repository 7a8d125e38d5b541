// trb_loader.cpp — Scene::load_file restated (src/scene.rs:101-146 and the loaders at :185-850):
// JSON scene -> flattened trb_scene_desc. Also OBJ (the `tobj` crate's behaviour, mesh.rs:49-76) and MERL
// binary tables (material/merl.rs:51-84). Host-only; nothing here runs per ray.
//
// Every transform becomes a TRS keyframe via Keyframe::decompose (keyframe.rs:32-58): polar
// decomposition in f64 (the reference uses la::SVD<f64>; any correct f64 polar factorisation agrees with it to
// ~1e-15, i.e. to the same f32 after rounding except in rare ties — "parity unpinned").
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "../../include/trb.h"
#include "trb_host.h"

using namespace trbh;

namespace {

thread_local std::string g_lerr;
struct LoadError { trb_status st; std::string msg; };
[[noreturn]] void die(trb_status st, const std::string& m) { throw LoadError{st, m}; }

// ---------------------------------------------------------------------------------------------
// minimal JSON (RFC 8259) reader
// ---------------------------------------------------------------------------------------------
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    bool is_int = false;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* key) const {
        if (kind != Obj) return nullptr;
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const JVal& expect(const char* key, const char* msg) const { const JVal* v = get(key); if (!v) die(TRB_INVALID_ARG, msg); return *v; }
    double f64(const char* msg) const { if (kind != Num) die(TRB_INVALID_ARG, msg); return num; }
    uint64_t u64(const char* msg) const { if (kind != Num || !is_int || num < 0) die(TRB_INVALID_ARG, msg); return (uint64_t)num; }
    const std::string& s(const char* msg) const { if (kind != Str) die(TRB_INVALID_ARG, msg); return str; }
    const std::vector<JVal>& a(const char* msg) const { if (kind != Arr) die(TRB_INVALID_ARG, msg); return arr; }
};
struct JParser {
    const char* p; const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    [[noreturn]] void bad(const char* m) { die(TRB_INVALID_ARG, std::string("JSON parsing error: ") + m); }
    JVal value() {
        ws();
        if (p >= end) bad("unexpected end");
        JVal v;
        char c = *p;
        if (c == '{') {
            v.kind = JVal::Obj; ++p; ws();
            if (p < end && *p == '}') { ++p; return v; }
            for (;;) {
                ws();
                if (p >= end || *p != '"') bad("expected key");
                std::string k = string();
                ws();
                if (p >= end || *p != ':') bad("expected ':'");
                ++p;
                v.obj.emplace_back(k, value());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                bad("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = JVal::Arr; ++p; ws();
            if (p < end && *p == ']') { ++p; return v; }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                bad("expected ',' or ']'");
            }
        } else if (c == '"') { v.kind = JVal::Str; v.str = string(); }
        else if (c == 't' && end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JVal::Bool; v.b = true; p += 4; }
        else if (c == 'f' && end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JVal::Bool; p += 5; }
        else if (c == 'n' && end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; }
        else {
            const char* s0 = p;
            bool integral = true;
            if (p < end && (*p == '-' || *p == '+')) ++p;
            while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) {
                if (*p == '.' || *p == 'e' || *p == 'E') integral = false;
                ++p;
            }
            if (p == s0) bad("unexpected character");
            v.kind = JVal::Num; v.num = strtod(std::string(s0, p).c_str(), nullptr); v.is_int = integral;
        }
        return v;
    }
    std::string string() {
        ++p;
        std::string o;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                ++p;
                switch (*p) {
                    case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break;
                    case 'b': o += '\b'; break; case 'f': o += '\f'; break;
                    case 'u': { if (end - p < 5) bad("bad \\u"); unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16); o += (char)(cp < 128 ? cp : '?'); p += 4; break; }
                    default: o += *p;
                }
                ++p;
            } else o += *p++;
        }
        if (p >= end) bad("unterminated string");
        ++p;
        return o;
    }
};

// ---------------------------------------------------------------------------------------------
// Transform construction (transform.rs:41-135) on the host Xf type
// ---------------------------------------------------------------------------------------------
Mat4 transpose(const Mat4& a) { Mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[4 * i + j] = a.m[4 * j + i]; return r; }
float to_rad(float d) { return kPi / 180.0f * d; }
Xf rot_axis(int axis, float deg) { // rotate_x / rotate_y / rotate_z (transform.rs:66-103)
    const float r = to_rad(deg), s = sinf(r), c = cosf(r);
    Mat4 m = mat_identity();
    if (axis == 0) { m.m[5] = c; m.m[6] = -s; m.m[9] = s; m.m[10] = c; }
    else if (axis == 1) { m.m[0] = c; m.m[2] = s; m.m[8] = -s; m.m[10] = c; }
    else { m.m[0] = c; m.m[1] = -s; m.m[4] = s; m.m[5] = c; }
    return Xf{m, transpose(m)};
}
Xf rot_general(const float ax[3], float deg) { // Transform::rotate (transform.rs:105-123)
    const float len = sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    const float x = ax[0] / len, y = ax[1] / len, z = ax[2] / len;
    const float r = to_rad(deg), s = sinf(r), c = cosf(r);
    Mat4 m = mat_identity();
    m.m[0] = x * x + (1.0f - x * x) * c; m.m[1] = x * y * (1.0f - c) - z * s; m.m[2] = x * z * (1.0f - c) + y * s;
    m.m[4] = x * y * (1.0f - c) + z * s; m.m[5] = y * y + (1.0f - y * y) * c; m.m[6] = y * z * (1.0f - c) - x * s;
    m.m[8] = x * z * (1.0f - c) - y * s; m.m[9] = y * z * (1.0f - c) + x * s; m.m[10] = z * z + (1.0f - z * z) * c;
    return Xf{m, transpose(m)};
}
Xf look_at(const float pos[3], const float center[3], const float up[3]) { // transform.rs:124-137
    auto norm = [](float* v) { float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); v[0] /= l; v[1] /= l; v[2] /= l; };
    auto cross = [](const float* a, const float* b, float* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
    float dir[3] = {center[0] - pos[0], center[1] - pos[1], center[2] - pos[2]};
    norm(dir);
    float left[3]; cross(up, dir, left); norm(left);
    float u[3]; cross(dir, left, u); norm(u);
    Mat4 m = mat_identity();
    for (int i = 0; i < 3; ++i) { m.m[4 * i] = -left[i]; m.m[4 * i + 1] = u[i]; m.m[4 * i + 2] = dir[i]; m.m[4 * i + 3] = pos[i]; }
    return Xf{m, mat_inverse(m)};
}

void load_vec3(const JVal& e, float o[3], const char* msg) { // load_vector / load_point (scene.rs:677-711)
    if (e.kind != JVal::Arr || e.arr.size() != 3) die(TRB_INVALID_ARG, msg);
    for (int i = 0; i < 3; ++i) o[i] = (float)e.arr[i].f64(msg);
}
// load_color (scene.rs:713-733): rgb, optionally scaled by a 4th component (which also scales alpha, Q15)
void load_color(const JVal& e, float o[4], const char* msg) {
    if (e.kind != JVal::Arr || (e.arr.size() != 3 && e.arr.size() != 4)) die(TRB_INVALID_ARG, msg);
    float v[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < e.arr.size(); ++i) v[i] = (float)e.arr[i].f64(msg);
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = 1.0f;
    if (e.arr.size() == 4) for (int i = 0; i < 4; ++i) o[i] = o[i] * v[3];
}
// load_transform (scene.rs:761-830): each entry is applied on the left
Xf load_transform(const JVal& e) {
    Xf t = xf_identity();
    for (const JVal& x : e.a("Invalid transform specified")) {
        const std::string& ty = x.expect("type", "A type is required for a transform").s("Transform type must be a string");
        if (ty == "translate") {
            float v[3]; load_vec3(x.expect("translation", "A translation vector is required for translate"), v, "Invalid vector specified for translation direction");
            t = xf_compose(xf_translate(v), t);
        } else if (ty == "scale") {
            const JVal& s = x.expect("scaling", "A scaling value or vector is required for scale");
            float v[3];
            if (s.kind == JVal::Arr) load_vec3(s, v, "Invalid vector specified for scaling vector");
            else if (s.kind == JVal::Num) v[0] = v[1] = v[2] = (float)s.num;
            else die(TRB_INVALID_ARG, "Scaling value should be an array of 3 floats or a single float");
            t = xf_compose(xf_scale(v), t);
        } else if (ty == "rotate_x" || ty == "rotate_y" || ty == "rotate_z") {
            const float r = (float)x.expect("rotation", "A rotation in degrees is required").f64("rotation must be a number");
            t = xf_compose(rot_axis(ty[7] - 'x', r), t);
        } else if (ty == "rotate") {
            const float r = (float)x.expect("rotation", "A rotation in degrees is required for rotate").f64("rotation for rotate must be a number");
            float ax[3]; load_vec3(x.expect("axis", "An axis vector is required for rotate"), ax, "Invalid vector specified for rotation axis");
            t = xf_compose(rot_general(ax, r), t);
        } else if (ty == "matrix") {
            Mat4 m = mat_identity();
            const auto& rows = x.expect("matrix", "The rows of the matrix are required for matrix transform").a("The rows should be an array");
            size_t k = 0;
            for (const JVal& r : rows) {
                const auto& row = r.a("Each row of the matrix transform must be an array, specifying the row");
                if (row.size() != 4) die(TRB_INVALID_ARG, "Each row of the transformation matrix must contain 4 elements");
                for (const JVal& el : row) { if (k < 16) m.m[k] = (float)el.f64("Each element of a matrix row must be a float"); ++k; }
            }
            t = xf_compose(xf_from_mat(m), t);
        } else die(TRB_INVALID_ARG, "Unrecognized transform type '" + ty + "'");
    }
    return t;
}

// ---- Keyframe::decompose (keyframe.rs:32-58) --------------------------------------------------
void jacobi_eigen3(double A[3][3], double V[3][3]) { // symmetric A -> eigenvalues on the diagonal, eigenvectors in V's columns
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = i == j;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
            }
    }
}
// Quaternion::from_matrix (quaternion.rs:28-62)
void quat_from_matrix(const Mat4& m, float q[4]) {
    auto at = [&](int i, int j) { return m.m[4 * i + j]; };
    const float trace = at(0, 0) + at(1, 1) + at(2, 2);
    if (trace > 0.0f) {
        float s = sqrtf(trace + 1.0f);
        const float w = s / 2.0f;
        s = 0.5f / s;
        q[0] = s * (at(2, 1) - at(1, 2)); q[1] = s * (at(0, 2) - at(2, 0)); q[2] = s * (at(1, 0) - at(0, 1)); q[3] = w;
    } else {
        const int next[3] = {1, 2, 0};
        float v[3] = {0, 0, 0};
        int i = at(1, 1) > at(0, 0) ? 1 : (at(2, 2) > at(0, 0) ? 2 : 0);
        const int j = next[i], k = next[j];
        float s = sqrtf((at(i, i) - (at(j, j) + at(k, k))) + 1.0f);
        v[i] = s * 0.5f;
        if (s != 0.0f) s = 0.5f / s;
        const float w = (at(k, j) - at(j, k)) * s;
        v[j] = (at(j, i) + at(i, j)) * s;
        v[k] = (at(k, i) + at(i, k)) * s;
        q[0] = v[0]; q[1] = v[1]; q[2] = v[2]; q[3] = w;
    }
}
trb_keyframe decompose(const Xf& t) {
    trb_keyframe kf;
    const Mat4& m = t.fwd;
    for (int i = 0; i < 3; ++i) kf.translation[i] = m.m[4 * i + 3];
    double M[3][3], A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] = (double)m.m[4 * i + j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = 0; for (int k = 0; k < 3; ++k) A[i][j] += M[k][i] * M[k][j]; }
    jacobi_eigen3(A, V);
    double sv[3];
    for (int i = 0; i < 3; ++i) sv[i] = sqrt(A[i][i] > 0 ? A[i][i] : 0);
    double P[3][3], Pinv[3][3], Q[3][3]; // M = Q P, P = V S V^T
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        P[i][j] = 0; Pinv[i][j] = 0;
        for (int k = 0; k < 3; ++k) { P[i][j] += V[i][k] * sv[k] * V[j][k]; if (sv[k] > 0) Pinv[i][j] += V[i][k] / sv[k] * V[j][k]; }
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Q[i][j] = 0; for (int k = 0; k < 3; ++k) Q[i][j] += M[i][k] * Pinv[k][j]; }
    const double det = Q[0][0] * (Q[1][1] * Q[2][2] - Q[1][2] * Q[2][1]) - Q[0][1] * (Q[1][0] * Q[2][2] - Q[1][2] * Q[2][0]) + Q[0][2] * (Q[1][0] * Q[2][1] - Q[1][1] * Q[2][0]);
    if (det < 0.0) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Q[i][j] = -Q[i][j]; P[i][j] = -P[i][j]; }
    Mat4 qm = mat_identity();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) qm.m[4 * i + j] = (float)Q[i][j];
    quat_from_matrix(qm, kf.rotation);
    for (int i = 0; i < 3; ++i) kf.scaling[i] = (float)P[i][i]; // off-diagonal stretch is dropped, as in the reference
    return kf;
}

// ---------------------------------------------------------------------------------------------
// OBJ (tobj 0.1.6 as used by Mesh::load_obj): one model per `o`/`g`, vertices unified per (v,vt,vn) in
// first-seen order, polygons fan-triangulated.
// ---------------------------------------------------------------------------------------------
struct ObjModel { std::string name; std::vector<float> pos, nrm, uv; std::vector<uint32_t> idx; };
std::vector<ObjModel> load_obj(const std::string& path) {
    std::ifstream f(path);
    if (!f) die(TRB_IO, "Failed to load " + path);
    std::vector<float> P, N, T;
    std::vector<ObjModel> models;
    ObjModel cur; cur.name = "unnamed_object";
    std::map<std::tuple<long, long, long>, uint32_t> remap;
    bool has_faces = false;
    auto flush = [&]() {
        if (has_faces) models.push_back(cur);
        cur = ObjModel(); remap.clear(); has_faces = false;
    };
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::string tag;
        if (!(ss >> tag)) continue;
        if (tag == "v") { float a, b, c; ss >> a >> b >> c; P.push_back(a); P.push_back(b); P.push_back(c); }
        else if (tag == "vn") { float a, b, c; ss >> a >> b >> c; N.push_back(a); N.push_back(b); N.push_back(c); }
        else if (tag == "vt") { float a = 0, b = 0; ss >> a >> b; T.push_back(a); T.push_back(b); }
        else if (tag == "o" || tag == "g") { flush(); std::string nm; std::getline(ss, nm); size_t s0 = nm.find_first_not_of(" \t"); cur.name = s0 == std::string::npos ? "unnamed_object" : nm.substr(s0); while (!cur.name.empty() && (cur.name.back() == '\r' || cur.name.back() == ' ')) cur.name.pop_back(); }
        else if (tag == "f") {
            std::vector<uint32_t> corner;
            std::string tok;
            while (ss >> tok) {
                long v = 0, vt = 0, vn = 0;
                size_t s1 = tok.find('/');
                v = atol(tok.substr(0, s1).c_str());
                if (s1 != std::string::npos) {
                    size_t s2 = tok.find('/', s1 + 1);
                    std::string a = tok.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
                    if (!a.empty()) vt = atol(a.c_str());
                    if (s2 != std::string::npos) { std::string b = tok.substr(s2 + 1); if (!b.empty()) vn = atol(b.c_str()); }
                }
                auto fix = [](long i, size_t n) -> long { return i > 0 ? i - 1 : (i < 0 ? (long)n + i : -1); };
                const long iv = fix(v, P.size() / 3), it = fix(vt, T.size() / 2), in = fix(vn, N.size() / 3);
                if (iv < 0 || (size_t)iv >= P.size() / 3) die(TRB_INVALID_ARG, "OBJ face references a missing vertex in " + path);
                auto key = std::make_tuple(iv, it, in);
                auto fnd = remap.find(key);
                uint32_t id;
                if (fnd != remap.end()) id = fnd->second;
                else {
                    id = (uint32_t)remap.size();
                    remap[key] = id;
                    for (int k = 0; k < 3; ++k) cur.pos.push_back(P[3 * iv + k]);
                    if (it >= 0 && (size_t)it < T.size() / 2) { cur.uv.push_back(T[2 * it]); cur.uv.push_back(T[2 * it + 1]); }
                    if (in >= 0 && (size_t)in < N.size() / 3) for (int k = 0; k < 3; ++k) cur.nrm.push_back(N[3 * in + k]);
                }
                corner.push_back(id);
            }
            for (size_t k = 2; k < corner.size(); ++k) { cur.idx.push_back(corner[0]); cur.idx.push_back(corner[k - 1]); cur.idx.push_back(corner[k]); }
            has_faces = has_faces || corner.size() >= 3;
        }
    }
    flush();
    return models;
}

// material::Merl::load_file (material/merl.rs:51-84)
std::vector<float> load_merl(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) die(TRB_IO, "material::Merl::load_file - failed to open " + path);
    int32_t dims[3];
    if (fread(dims, 4, 3, f) != 3 || dims[0] != 90 || dims[1] != 90 || dims[2] != 180) { fclose(f); die(TRB_INVALID_ARG, "material::Merl::load_file - Invalid MERL file header, aborting"); }
    const size_t n = 90u * 90u * 180u;
    std::vector<double> plane(n);
    std::vector<float> brdf(3 * n, 0.0f);
    const double scaling[3] = {1.0 / 1500.0, 1.0 / 1500.0, 1.66 / 1500.0};
    for (int c = 0; c < 3; ++c) {
        if (fread(plane.data(), 8, n, f) != n) { fclose(f); die(TRB_IO, "MERL file truncated: " + path); }
        for (size_t i = 0; i < n; ++i) brdf[3 * i + c] = fmaxf(0.0f, (float)(plane[i] * scaling[c]));
    }
    fclose(f);
    return brdf;
}

// ---------------------------------------------------------------------------------------------
struct DescOwner {
    trb_scene_desc desc; // must stay first: trb_desc_free recovers the owner from the desc pointer
    std::vector<trb_camera> cameras;
    std::vector<trb_instance> instances;
    std::vector<trb_spline> splines;
    std::vector<trb_keyframe> keyframes;
    std::vector<float> knots;
    std::vector<trb_color_key> color_keys;
    std::vector<trb_mesh> meshes;
    std::vector<std::unique_ptr<ObjModel>> mesh_data;
    std::vector<trb_material> materials;
    std::vector<std::vector<float>> merl;
    std::vector<const float*> merl_ptrs;
    std::vector<float> fov_floats;
    std::vector<trb_texture> textures;
    std::vector<trb_image> images;
    std::vector<std::vector<uint8_t>> image_data;
    std::map<std::string, uint32_t> texture_names;
    std::map<std::string, uint32_t> material_names;
    std::map<std::string, std::map<std::string, uint32_t>> mesh_cache; // file -> model -> mesh index
    std::string dir;
};

std::string join(const std::string& dir, const std::string& p) { return (!p.empty() && p[0] == '/') ? p : dir + "/" + p; }

// AnimatedTransform::unanimated / load_keyframes (scene.rs:832-850, animated_transform.rs:22-37)
std::pair<uint32_t, uint32_t> add_xf(DescOwner& o, const JVal& obj, const char* who) {
    const uint32_t first = (uint32_t)o.splines.size();
    if (const JVal* k = obj.get("keyframes")) {
        trb_spline sp{};
        sp.ctrl_first = (uint32_t)o.keyframes.size(); sp.knot_first = (uint32_t)o.knots.size();
        std::vector<trb_keyframe> kfs;
        for (const JVal& p : k->expect("control_points", "Control points are required for bspline keyframes").a("Invalid keyframes specified"))
            kfs.push_back(decompose(load_transform(p.expect("transform", "A transform is required for a keyframe"))));
        for (size_t i = 1; i < kfs.size(); ++i) { // with_keyframes: keep quaternions in one hemisphere
            float d = 0; for (int c = 0; c < 4; ++c) d += kfs[i - 1].rotation[c] * kfs[i].rotation[c];
            if (d < 0.0f) for (int c = 0; c < 4; ++c) kfs[i].rotation[c] = -kfs[i].rotation[c];
        }
        for (const auto& kf : kfs) o.keyframes.push_back(kf);
        for (const JVal& kn : k->expect("knots", "knots are required for bspline keyframes").a("Invalid keyframes specified")) o.knots.push_back((float)kn.f64("Knots must be numbers"));
        sp.n_ctrl = (uint32_t)kfs.size(); sp.n_knots = (uint32_t)o.knots.size() - sp.knot_first;
        sp.degree = k->get("degree") ? (uint32_t)k->get("degree")->u64("Curve degree must be a positive integer") : 3;
        o.splines.push_back(sp);
    } else {
        const JVal* t = obj.get("transform");
        if (!t) die(TRB_INVALID_ARG, std::string("No keyframes or transform specified for ") + who);
        trb_spline sp{};
        sp.degree = 0; sp.n_ctrl = 1; sp.ctrl_first = (uint32_t)o.keyframes.size(); sp.n_knots = 2; sp.knot_first = (uint32_t)o.knots.size();
        o.keyframes.push_back(decompose(load_transform(*t)));
        o.knots.push_back(0.0f); o.knots.push_back(1.0f);
        o.splines.push_back(sp);
    }
    return {first, (uint32_t)o.splines.size() - first};
}

// ---- image textures (scene.rs:317-394 load_textures; image::open) -------------------------------------------------------
// PNG decoding (what `image::open` does for the formats this build reads): 8-bit grey / grey+alpha / RGB / palette / RGBA,
// non-interlaced, expanded to RGBA8 like DynamicImage::get_pixel (grey -> l,l,l,255; RGB -> alpha 255). Own inflate (RFC 1951).
struct BitReader {
    const uint8_t* p; size_t n, pos = 0; uint32_t bitbuf = 0; int bitcnt = 0;
    uint32_t bits(int k) { while (bitcnt < k) { if (pos >= n) die(TRB_IO, "truncated deflate stream"); bitbuf |= (uint32_t)p[pos++] << bitcnt; bitcnt += 8; } const uint32_t v = bitbuf & ((1u << k) - 1u); bitbuf >>= k; bitcnt -= k; return v; }
};
struct Huffman {
    uint16_t count[16] = {0}, symbol[320] = {0};
    void build(const uint8_t* len, int n) {
        for (int i = 0; i < 16; ++i) count[i] = 0;
        for (int i = 0; i < n; ++i) count[len[i]]++;
        count[0] = 0;
        uint16_t offs[16]; offs[1] = 0;
        for (int i = 1; i < 15; ++i) offs[i + 1] = offs[i] + count[i];
        for (int i = 0; i < n; ++i) if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len <= 15; ++len) {
            code |= (int)br.bits(1);
            const int c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        die(TRB_IO, "bad Huffman code in deflate stream");
        return -1;
    }
};
std::vector<uint8_t> inflate_zlib(const uint8_t* data, size_t n) {
    if (n < 6) die(TRB_IO, "truncated zlib stream");
    BitReader br{data + 2, n - 2};
    std::vector<uint8_t> out;
    static const uint16_t lbase[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
    static const uint16_t lext[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
    static const uint16_t dbase[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
    static const uint16_t dext[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
    for (bool last = false; !last;) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (type == 0) { // stored
            br.bitbuf = 0; br.bitcnt = 0;
            if (br.pos + 4 > br.n) die(TRB_IO, "truncated deflate stream");
            const uint32_t len = br.p[br.pos] | (br.p[br.pos + 1] << 8);
            br.pos += 4;
            if (br.pos + len > br.n) die(TRB_IO, "truncated deflate stream");
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
            continue;
        }
        if (type == 3) die(TRB_IO, "bad deflate block type");
        Huffman lit, dist;
        uint8_t lens[320];
        if (type == 1) {
            for (int i = 0; i < 144; ++i) lens[i] = 8; for (int i = 144; i < 256; ++i) lens[i] = 9; for (int i = 256; i < 280; ++i) lens[i] = 7; for (int i = 280; i < 288; ++i) lens[i] = 8;
            lit.build(lens, 288);
            for (int i = 0; i < 30; ++i) lens[i] = 5;
            dist.build(lens, 30);
        } else {
            const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
            static const uint8_t order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
            uint8_t cl[19] = {0};
            for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.bits(3);
            Huffman lc; lc.build(cl, 19);
            for (int i = 0; i < nlen + ndist;) {
                const int sym = lc.decode(br);
                if (sym < 16) lens[i++] = (uint8_t)sym;
                else {
                    int rep; uint8_t v = 0;
                    if (sym == 16) { if (i == 0) die(TRB_IO, "bad deflate lengths"); v = lens[i - 1]; rep = 3 + (int)br.bits(2); }
                    else if (sym == 17) rep = 3 + (int)br.bits(3);
                    else rep = 11 + (int)br.bits(7);
                    if (i + rep > nlen + ndist) die(TRB_IO, "bad deflate lengths");
                    while (rep--) lens[i++] = v;
                }
            }
            lit.build(lens, nlen); dist.build(lens + nlen, ndist);
        }
        for (;;) {
            const int sym = lit.decode(br);
            if (sym < 256) out.push_back((uint8_t)sym);
            else if (sym == 256) break;
            else {
                if (sym > 285) die(TRB_IO, "bad deflate symbol");
                const uint32_t len = lbase[sym - 257] + br.bits(lext[sym - 257]);
                const int ds = dist.decode(br);
                if (ds > 29) die(TRB_IO, "bad deflate distance");
                const uint32_t d = dbase[ds] + br.bits(dext[ds]);
                if (d > out.size()) die(TRB_IO, "deflate distance too far back");
                for (uint32_t k = 0; k < len; ++k) out.push_back(out[out.size() - d]);
            }
        }
    }
    return out;
}
void load_png(const std::string& path, uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgba) {
    std::ifstream f(path, std::ios::binary);
    if (!f) die(TRB_IO, "Failed to load image file " + path);
    std::vector<uint8_t> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (raw.size() < 8 || std::memcmp(raw.data(), sig, 8) != 0) die(TRB_UNSUPPORTED, "only PNG image textures are read by this build: " + path);
    auto be32 = [&](size_t o) { return ((uint32_t)raw[o] << 24) | ((uint32_t)raw[o + 1] << 16) | ((uint32_t)raw[o + 2] << 8) | raw[o + 3]; };
    uint32_t depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    w = h = 0;
    for (size_t o = 8; o + 12 <= raw.size();) {
        const uint32_t n = be32(o);
        if (o + 12 + (size_t)n > raw.size()) die(TRB_IO, "truncated PNG " + path);
        const std::string ty(reinterpret_cast<const char*>(&raw[o + 4]), 4);
        const uint8_t* d = &raw[o + 8];
        if (ty == "IHDR") { if (n < 13) die(TRB_IO, "bad IHDR"); w = be32(o + 8); h = be32(o + 12); depth = d[8]; ctype = d[9]; interlace = d[12]; }
        else if (ty == "PLTE") plte.assign(d, d + n);
        else if (ty == "tRNS") trns.assign(d, d + n);
        else if (ty == "IDAT") idat.insert(idat.end(), d, d + n);
        else if (ty == "IEND") break;
        o += 12 + (size_t)n;
    }
    if (w == 0 || h == 0 || w > 32768 || h > 32768) die(TRB_IO, "bad PNG dimensions in " + path);
    if (depth != 8 || interlace != 0) die(TRB_UNSUPPORTED, "only 8-bit non-interlaced PNG textures are read by this build: " + path);
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) die(TRB_IO, "bad PNG colour type in " + path);
    const std::vector<uint8_t> data = inflate_zlib(idat.data(), idat.size());
    const size_t stride = (size_t)w * ch;
    if (data.size() < (stride + 1) * h) die(TRB_IO, "truncated PNG image data in " + path);
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; ++y) { // undo the scanline filters (PNG spec 9.2)
        const uint8_t ft = data[(stride + 1) * y];
        const uint8_t* in = &data[(stride + 1) * y + 1];
        uint8_t* cur = &img[stride * y];
        const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)ch ? cur[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
            int v = in[x];
            if (ft == 1) v += a;
            else if (ft == 2) v += b;
            else if (ft == 3) v += (a + b) >> 1;
            else if (ft == 4) { const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
            else if (ft != 0) die(TRB_IO, "bad PNG filter type in " + path);
            cur[x] = (uint8_t)v;
        }
    }
    rgba.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        uint8_t* o = &rgba[4 * i];
        const uint8_t* p = &img[i * ch];
        if (ctype == 0) { o[0] = o[1] = o[2] = p[0]; o[3] = 255; }
        else if (ctype == 2) { o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = 255; }
        else if (ctype == 3) { const size_t k = p[0]; if (3 * k + 2 >= plte.size()) die(TRB_IO, "PNG palette index out of range in " + path); o[0] = plte[3 * k]; o[1] = plte[3 * k + 1]; o[2] = plte[3 * k + 2]; o[3] = k < trns.size() ? trns[k] : 255; }
        else if (ctype == 4) { o[0] = o[1] = o[2] = p[0]; o[3] = p[1]; }
        else { o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3]; }
    }
}
uint32_t add_image(DescOwner& o, const std::string& file, float time) {
    trb_image im{};
    std::vector<uint8_t> px;
    load_png(join(o.dir, file), im.width, im.height, px);
    im.time = time;
    o.image_data.push_back(std::move(px));
    o.images.push_back(im); // rgba8 pointers are set once every image is loaded (vectors may still move)
    return (uint32_t)o.images.size() - 1;
}
void load_textures(DescOwner& o, const JVal& e) { // scene.rs:317-394
    for (const JVal& t : e.a("The 'textures' must be an array of textures to load")) {
        const std::string& name = t.expect("name", "Error loading texture: A name is required").s("name must be a string");
        const std::string& ty = t.expect("type", "A texture type is required").s("Texture type must be a string");
        if (o.texture_names.count(name)) die(TRB_INVALID_ARG, "Error loading texture '" + name + "': name conflicts with an existing entry");
        trb_texture tex{};
        tex.first_image = (uint32_t)o.images.size();
        if (ty == "image") add_image(o, t.expect("file", "Image textures must specify an image file").s("Image file name must be a string"), 0.0f);
        else if (ty == "animated_image") {
            const auto& frames = t.expect("keyframes", "animated_image requires keyframes").a("animated_image keyframes must be an array");
            if (frames.size() < 2) die(TRB_INVALID_ARG, "animated_image must have at least 2 frames");
            for (const JVal& f : frames)
                add_image(o, f.expect("file", "Image textures must specify an image file").s("Image file name must be a string"),
                          (float)f.expect("time", "animated_image keyframe requires time").f64("animated_image keyframe time must be a number"));
        } else if (ty == "movie") { // keyframes generated from a file name pattern and a frame rate
            const std::string& prefix = t.expect("file_prefix", "A file_prefix for movie is required").s("file_prefix for movie must be a string");
            const std::string& suffix = t.expect("file_suffix", "A file_suffix for movie is required").s("file_suffix for movie must be a string");
            const uint64_t total = t.expect("frames", "# of frames for movie texture is required").u64("frames for movie texture must be an int");
            const uint64_t rate = t.expect("framerate", "A framerate for movie is required").u64("framerate for movie must be an int");
            if (total < 2) die(TRB_INVALID_ARG, "animated_image must have at least 2 frames"); // AnimatedImage::new asserts frames.len() >= 2
            for (uint64_t fr = 0; fr < total; ++fr) {
                char num[32]; std::snprintf(num, sizeof num, "%05llu", (unsigned long long)fr);
                add_image(o, prefix + num + suffix, (float)fr / (float)rate);
            }
        } else die(TRB_INVALID_ARG, "Unrecognized texture type '" + ty + "' for texture '" + name + "'");
        tex.n_images = (uint32_t)o.images.size() - tex.first_image;
        o.texture_names[name] = (uint32_t)o.textures.size();
        o.textures.push_back(tex);
    }
}
uint32_t named_texture(DescOwner& o, const JVal& e, const char* what) { // a string value names a loaded texture (find_color / find_scalar return None otherwise)
    auto it = o.texture_names.find(e.s(what));
    if (it == o.texture_names.end()) die(TRB_INVALID_ARG, std::string("Invalid texture name specified for ") + what);
    return it->second + 1;
}
void color_tex(DescOwner& o, const JVal& e, float out[3], uint32_t& tex, const char* what) { // LoadedTextures::find_color (scene.rs:45-69)
    if (e.kind == JVal::Str) { tex = named_texture(o, e, what); return; }
    float c[4]; load_color(e, c, what);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
}
float scalar_tex(DescOwner& o, const JVal& e, uint32_t& tex, const char* what) { // find_scalar (scene.rs:70-87)
    if (e.kind == JVal::Str) { tex = named_texture(o, e, what); return 0.0f; }
    return (float)e.f64(what);
}

void load_materials(DescOwner& o, const JVal& e) { // scene.rs:433-541
    for (const JVal& m : e.a("The materials must be an array of materials used")) {
        const std::string& name = m.expect("name", "Error loading material: A name is required").s("name must be a string");
        const std::string& ty = m.expect("type", "a type is required").s("type must be a string");
        if (o.material_names.count(name)) die(TRB_INVALID_ARG, "Error loading material '" + name + "': name conflicts with an existing entry");
        trb_material t{};
        t.eta = 1.0f;
        if (ty == "glass" || ty == "rough_glass") {
            t.type = ty == "glass" ? TRB_MAT_GLASS : TRB_MAT_ROUGH_GLASS;
            color_tex(o, m.expect("reflect", "reflect color/texture name is required for glass"), t.c0, t.tex[0], "reflect");
            color_tex(o, m.expect("transmit", "transmit color/texture name is required for glass"), t.c1, t.tex[1], "transmit");
            t.eta = scalar_tex(o, m.expect("eta", "eta color/texture name is required for glass"), t.tex[3], "eta");
            if (ty == "rough_glass") t.roughness = scalar_tex(o, m.expect("roughness", "roughness is required for rough glass"), t.tex[2], "roughness");
        } else if (ty == "matte") {
            t.type = TRB_MAT_MATTE;
            color_tex(o, m.expect("diffuse", "diffuse color/texture name is required for matte"), t.c0, t.tex[0], "diffuse");
            t.roughness = scalar_tex(o, m.expect("roughness", "roughness color/texture is required for matte"), t.tex[2], "roughness");
        } else if (ty == "merl") {
            t.type = TRB_MAT_MERL;
            t.merl = (uint32_t)o.merl.size();
            o.merl.push_back(load_merl(join(o.dir, m.expect("file", "A filename containing the MERL material data is required").s("The MERL file must be a string"))));
        } else if (ty == "metal" || ty == "specular_metal") {
            t.type = ty == "metal" ? TRB_MAT_METAL : TRB_MAT_SPECULAR_METAL;
            color_tex(o, m.expect("refractive_index", "refractive_index color/texture name is required for metal"), t.c0, t.tex[0], "refractive_index");
            color_tex(o, m.expect("absorption_coefficient", "absorption_coefficient color/texture name is required for metal"), t.c1, t.tex[1], "absorption_coefficient");
            if (ty == "metal") t.roughness = scalar_tex(o, m.expect("roughness", "roughness color/texture is required for metal"), t.tex[2], "roughness");
        } else if (ty == "plastic") {
            t.type = TRB_MAT_PLASTIC;
            color_tex(o, m.expect("diffuse", "diffuse color/texture name is required for plastic"), t.c0, t.tex[0], "diffuse");
            color_tex(o, m.expect("gloss", "gloss color/texture name is required for plastic"), t.c1, t.tex[1], "gloss");
            t.roughness = scalar_tex(o, m.expect("roughness", "roughness color/texture is required for plastic"), t.tex[2], "roughness");
        } else die(TRB_INVALID_ARG, "Error parsing material '" + name + "': unrecognized type '" + ty + "'");
        o.material_names[name] = (uint32_t)o.materials.size();
        o.materials.push_back(t);
    }
}

uint32_t find_material(DescOwner& o, const JVal& obj) {
    const std::string& n = obj.expect("material", "A material is required for an object").s("Object material name must be a string");
    auto it = o.material_names.find(n);
    if (it == o.material_names.end()) die(TRB_INVALID_ARG, "Material " + n + " was not found in the material list");
    return it->second;
}

// load_geometry / load_sampleable_geometry (scene.rs:608-675)
void load_geometry(DescOwner& o, const JVal& g, bool sampleable, trb_instance& in) {
    const std::string& ty = g.expect("type", "A type is required for geometry").s("Geometry type must be a string");
    if (ty == "sphere") { in.shape = TRB_SHAPE_SPHERE; in.p0 = (float)g.expect("radius", "A radius is required for a sphere").f64("radius must be a number"); }
    else if (ty == "disk") {
        in.shape = TRB_SHAPE_DISK;
        in.p0 = (float)g.expect("radius", "A radius is required for a disk").f64("radius must be a number");
        in.p1 = (float)g.expect("inner_radius", "An inner radius is required for a disk").f64("inner radius must be a number");
    } else if (ty == "plane" && !sampleable) { in.shape = TRB_SHAPE_RECT; in.p0 = 2.0f; in.p1 = 2.0f; }
    else if (ty == "rectangle") {
        in.shape = TRB_SHAPE_RECT;
        in.p0 = (float)g.expect("width", "A width is required for a rectangle").f64("width must be a number");
        in.p1 = (float)g.expect("height", "A height is required for a rectangle").f64("height must be a number");
    } else if (ty == "mesh" && !sampleable) {
        const std::string file = join(o.dir, g.expect("file", "An OBJ file is required for meshes").s("OBJ filename must be a string"));
        const std::string& model = g.expect("model", "A model name is required for geometry").s("Model name type must be a string");
        if (!o.mesh_cache.count(file)) {
            auto& per_file = o.mesh_cache[file];
            for (ObjModel& m : load_obj(file)) {
                if (m.nrm.size() != m.pos.size() || m.uv.size() / 2 != m.pos.size() / 3) continue; // "Normals and texture coordinates are required! Skipping" (mesh.rs:57-61)
                per_file[m.name] = (uint32_t)o.meshes.size();
                o.mesh_data.emplace_back(new ObjModel(std::move(m)));
                const ObjModel& d = *o.mesh_data.back();
                trb_mesh tm{};
                tm.n_verts = (uint32_t)(d.pos.size() / 3); tm.n_tris = (uint32_t)(d.idx.size() / 3);
                tm.positions = d.pos.data(); tm.normals = d.nrm.data(); tm.texcoords = d.uv.data(); tm.indices = d.idx.data();
                o.meshes.push_back(tm);
            }
        }
        auto it = o.mesh_cache[file].find(model);
        if (it == o.mesh_cache[file].end()) die(TRB_INVALID_ARG, "Requested model '" + model + "' was not found in '" + file + "'");
        in.shape = TRB_SHAPE_MESH; in.mesh = it->second;
    } else if (sampleable) die(TRB_INVALID_ARG, "Geometry of type '" + ty + "' is not sampleable and can't be used for area light geometry");
    else die(TRB_INVALID_ARG, "Unrecognized geometry type '" + ty + "'");
}

// load_animated_color (scene.rs:735-759)
void load_emission(DescOwner& o, const JVal& e, trb_instance& in) {
    const auto& arr = e.a("Emitter emission must be a color");
    if (arr.empty()) die(TRB_INVALID_ARG, "Emitter emission must be a color");
    in.emission_first = (uint32_t)o.color_keys.size();
    if (arr[0].kind == JVal::Num) {
        trb_color_key k{}; load_color(e, k.rgba, "Emitter emission must be a color"); k.time = 0.0f;
        o.color_keys.push_back(k);
    } else {
        std::vector<trb_color_key> ks;
        for (const JVal& c : arr) {
            trb_color_key k{};
            k.time = (float)c.expect("time", "A time must be specified for a color keyframe").f64("Time for color keyframe must be a number");
            load_color(c.expect("color", "A color must be specified for a color keyframe"), k.rgba, "A valid color is required for a color keyframe");
            ks.push_back(k);
        }
        std::stable_sort(ks.begin(), ks.end(), [](const trb_color_key& a, const trb_color_key& b) { return a.time < b.time; });
        for (auto& k : ks) o.color_keys.push_back(k);
    }
    in.n_emission = (uint32_t)o.color_keys.size() - in.emission_first;
}

// load_objects (scene.rs:543-606). Group members get the group's splines appended after their own (AnimatedTransform::mul).
struct PendingInstance { trb_instance in; std::vector<trb_spline> levels; };
void load_objects(DescOwner& o, const JVal& e, std::vector<PendingInstance>& out) {
    for (const JVal& ob : e.a("The objects must be an array of objects used")) {
        const std::string& name = ob.expect("name", "A name is required for an object").s("Object name must be a string");
        const std::string& ty = ob.expect("type", "A type is required for an object").s("Object type must be a string");
        auto range = add_xf(o, ob, name.c_str());
        std::vector<trb_spline> mine(o.splines.begin() + range.first, o.splines.begin() + range.first + range.second);
        o.splines.resize(range.first); // re-emitted contiguously per instance at the end
        PendingInstance pi{};
        pi.levels = mine;
        if (ty == "emitter") {
            const std::string& et = ob.expect("emitter", "An emitter type is required for emitters").s("Emitter type must be a string");
            load_emission(o, ob.expect("emission", "An emission color is required for emitters"), pi.in);
            if (et == "point") { pi.in.kind = TRB_INST_EMITTER_POINT; pi.in.shape = TRB_SHAPE_NONE; }
            else if (et == "area") {
                pi.in.kind = TRB_INST_EMITTER_AREA;
                pi.in.material = find_material(o, ob);
                load_geometry(o, ob.expect("geometry", "Geometry is required for area lights"), true, pi.in);
            } else die(TRB_INVALID_ARG, "Invalid emitter type specified: " + et);
            out.push_back(pi);
        } else if (ty == "receiver") {
            pi.in.kind = TRB_INST_RECEIVER;
            pi.in.material = find_material(o, ob);
            load_geometry(o, ob.expect("geometry", "Geometry is required for receivers"), false, pi.in);
            out.push_back(pi);
        } else if (ty == "group") {
            std::vector<PendingInstance> members;
            load_objects(o, ob.expect("objects", "A group must specify an array of objects in the group"), members);
            for (PendingInstance& m : members) { for (const trb_spline& l : mine) m.levels.push_back(l); out.push_back(m); }
        } else die(TRB_INVALID_ARG, "Error parsing object '" + name + "': unrecognized type '" + ty + "'");
    }
}

void load_camera(DescOwner& o, const JVal& e) { // scene.rs:291-333
    trb_camera c{};
    c.shutter_size = e.get("shutter_size") ? (float)e.get("shutter_size")->f64("Shutter size should be a float from 0 to 1") : 0.5f;
    c.active_at = e.get("active_at") ? (uint32_t)e.get("active_at")->u64("The camera activation frame 'active_at' must be an unsigned int") : 0;
    if (e.get("keyframes") || e.get("transform")) {
        auto r = add_xf(o, e, "camera");
        c.spline_first = r.first; c.n_splines = r.second;
    } else { // deprecated position/target/up
        float pos[3], target[3], up[3];
        load_vec3(e.expect("position", "The camera must specify a position"), pos, "position must be an array of 3 floats");
        load_vec3(e.expect("target", "The camera must specify a target"), target, "target must be an array of 3 floats");
        load_vec3(e.expect("up", "The camera must specify an up vector"), up, "up must be an array of 3 floats");
        trb_spline sp{};
        sp.degree = 0; sp.n_ctrl = 1; sp.ctrl_first = (uint32_t)o.keyframes.size(); sp.n_knots = 2; sp.knot_first = (uint32_t)o.knots.size();
        o.keyframes.push_back(decompose(look_at(pos, target, up)));
        o.knots.push_back(0.0f); o.knots.push_back(1.0f);
        c.spline_first = (uint32_t)o.splines.size(); c.n_splines = 1;
        o.splines.push_back(sp);
    }
    const JVal& fov = e.expect("fov", "The camera must specify a field of view");
    if (fov.kind == JVal::Arr) {
        c.fov_ctrl_first = (uint32_t)o.fov_floats.size();
        for (const JVal& f : fov.arr) o.fov_floats.push_back((float)f.f64("fovs must be a number"));
        c.n_fov_ctrl = (uint32_t)fov.arr.size();
        c.fov_knot_first = (uint32_t)o.fov_floats.size();
        for (const JVal& f : e.expect("fov_knots", "Animated field of view must specify spline knots").a("Fov spline knots must be an array")) o.fov_floats.push_back((float)f.f64("fov knots must be a number"));
        c.n_fov_knots = (uint32_t)o.fov_floats.size() - c.fov_knot_first;
        c.fov_degree = (uint32_t)e.expect("fov_spline_degree", "Animated fov spline must have degree").u64("Animated fov spline degree must be a u64");
        c.fov = o.fov_floats[c.fov_ctrl_first];
    } else c.fov = (float)fov.f64("Camera fov must be a number");
    o.cameras.push_back(c);
}

DescOwner* load_scene(const char* path, uint32_t w, uint32_t h, uint32_t spp) {
    std::ifstream f(path, std::ios::binary);
    if (!f) die(TRB_IO, std::string("Failed to open scene file: ") + path);
    std::stringstream buf; buf << f.rdbuf();
    const std::string text = buf.str();
    JParser jp{text.data(), text.data() + text.size()};
    const JVal root = jp.value();
    if (root.kind != JVal::Obj) die(TRB_INVALID_ARG, "Expected a root JSON object. See example scenes");
    std::unique_ptr<DescOwner> o(new DescOwner);
    std::string p(path);
    size_t slash = p.find_last_of('/');
    o->dir = slash == std::string::npos ? "." : p.substr(0, slash);

    // load_film / load_filter (scene.rs:185-272)
    const JVal& film = root.expect("film", "The scene must specify a film to write to");
    trb_film& tf = o->desc.film;
    tf.width = (uint32_t)film.expect("width", "The film must specify the image width").u64("Image width must be a number");
    tf.height = (uint32_t)film.expect("height", "The film must specify the image height").u64("Image height must be a number");
    tf.samples = (uint32_t)film.expect("samples", "The film must specify the number of samples per pixel").u64("Samples per pixel must be a number");
    tf.start_frame = (uint32_t)film.expect("start_frame", "The film must specify the starting frame").u64("Start frame must be a number");
    tf.end_frame = (uint32_t)film.expect("end_frame", "The film must specify the frame to end on").u64("End frame must be a number");
    if (tf.end_frame < tf.start_frame) die(TRB_INVALID_ARG, "End frame must be greater or equal to the starting frame");
    tf.frames = (uint32_t)film.expect("frames", "The film must specify the total number of frames").u64("Frames must be a number");
    tf.scene_time = (float)film.expect("scene_time", "The film must specify the overall scene time").f64("Scene time must be a number");
    const JVal& flt = film.expect("filter", "The film must specify a reconstruction filter");
    tf.filter_w = (float)flt.expect("width", "The filter must specify the filter width").f64("Filter width must be a number");
    tf.filter_h = (float)flt.expect("height", "The filter must specify the filter height").f64("Filter height must be a number");
    const std::string& fty = flt.expect("type", "A type is required for the filter").s("Filter type must be a string");
    if (fty == "mitchell_netravali") {
        tf.filter_type = TRB_FILTER_MITCHELL_NETRAVALI;
        tf.filter_b = (float)flt.expect("b", "A b parameter is required for the Mitchell-Netravali filter").f64("b must be a number");
        tf.filter_c = (float)flt.expect("c", "A c parameter is required for the Mitchell-Netravali filter").f64("c must be a number");
    } else if (fty == "gaussian") {
        tf.filter_type = TRB_FILTER_GAUSSIAN;
        tf.filter_b = (float)flt.expect("alpha", "An alpha parameter is required for the Gaussian filter").f64("alpha must be a number");
    } else die(TRB_INVALID_ARG, "Unrecognized filter type " + fty + "!");
    if (w) tf.width = w; if (h) tf.height = h; if (spp) tf.samples = spp;

    // load_cameras (scene.rs:274-290): stable sort by active_at
    if (const JVal* cams = root.get("cameras")) { for (const JVal& c : cams->a("cameras listing must be an array of cameras")) load_camera(*o, c); }
    else load_camera(*o, root.expect("camera", "Error: A camera is required!"));
    std::stable_sort(o->cameras.begin(), o->cameras.end(), [](const trb_camera& a, const trb_camera& b) { return a.active_at < b.active_at; });

    // load_integrator (scene.rs:335-353)
    const JVal& integ = root.expect("integrator", "The scene must specify the integrator to render with");
    const std::string& ity = integ.expect("type", "Integrator must specify a type").s("Integrator type must be a string");
    if (ity == "pathtracer") {
        o->desc.integrator.type = TRB_INTEGRATOR_PATH;
        o->desc.integrator.min_depth = (uint32_t)integ.expect("min_depth", "The integrator must specify the minimum ray depth").u64("min_depth must be a number");
        o->desc.integrator.max_depth = (uint32_t)integ.expect("max_depth", "The integrator must specify the maximum ray depth").u64("max_depth must be a number");
    } else if (ity == "whitted") { // scene.rs:305-309: Whitted::new(min_depth) — the recursion limit is read from the key "min_depth" (sic)
        o->desc.integrator.type = TRB_INTEGRATOR_WHITTED;
        o->desc.integrator.min_depth = 0;
        o->desc.integrator.max_depth = (uint32_t)integ.expect("min_depth", "The integrator must specify the minimum ray depth").u64("min_depth must be a number");
    } else if (ity == "normals_debug") { // scene.rs:310-311
        o->desc.integrator.type = TRB_INTEGRATOR_NORMALS_DEBUG;
        o->desc.integrator.min_depth = o->desc.integrator.max_depth = 0;
    }
    else die(TRB_INVALID_ARG, "Unrecognized integrator type '" + ity + "'");

    if (const JVal* tx = root.get("textures")) load_textures(*o, *tx); // scene.rs:118-121: loaded before the materials that name them
    load_materials(*o, root.expect("materials", "An array of materials is required"));
    std::vector<PendingInstance> pend;
    load_objects(*o, root.expect("objects", "The scene must specify a list of objects"), pend);
    if (pend.empty()) die(TRB_INVALID_ARG, "Aborting: the scene does not have any objects!");
    for (PendingInstance& pi : pend) {
        pi.in.spline_first = (uint32_t)o->splines.size(); pi.in.n_splines = (uint32_t)pi.levels.size();
        for (const trb_spline& l : pi.levels) o->splines.push_back(l);
        o->instances.push_back(pi.in);
    }
    for (const auto& t : o->merl) o->merl_ptrs.push_back(t.data());

    trb_scene_desc& d = o->desc;
    d.abi_version = TRB_ABI_VERSION;
    d.n_cameras = (uint32_t)o->cameras.size(); d.cameras = o->cameras.data();
    d.n_instances = (uint32_t)o->instances.size(); d.instances = o->instances.data();
    d.n_splines = (uint32_t)o->splines.size(); d.splines = o->splines.data();
    d.n_keyframes = (uint32_t)o->keyframes.size(); d.keyframes = o->keyframes.data();
    d.n_knots = (uint32_t)o->knots.size(); d.knots = o->knots.data();
    d.n_color_keys = (uint32_t)o->color_keys.size(); d.color_keys = o->color_keys.data();
    d.n_meshes = (uint32_t)o->meshes.size(); d.meshes = o->meshes.data();
    d.n_materials = (uint32_t)o->materials.size(); d.materials = o->materials.data();
    d.n_merl = (uint32_t)o->merl.size(); d.merl_tables = o->merl_ptrs.data();
    d.n_fov_floats = (uint32_t)o->fov_floats.size(); d.fov_floats = o->fov_floats.data();
    for (size_t i = 0; i < o->images.size(); ++i) o->images[i].rgba8 = o->image_data[i].data();
    d.n_textures = (uint32_t)o->textures.size(); d.textures = o->textures.data();
    d.n_images = (uint32_t)o->images.size(); d.images = o->images.data();
    return o.release();
}

} // namespace

extern "C" void trb_internal_set_error(const char* msg); // trb_api.cu

extern "C" {

const char* trb_loader_last_error(void) { return g_lerr.c_str(); }

trb_status trb_desc_load_json(const char* path, uint32_t w, uint32_t h, uint32_t spp, trb_scene_desc** out) {
    if (!path || !out) { g_lerr = "null argument"; return TRB_INVALID_ARG; }
    *out = nullptr;
    try {
        DescOwner* o = load_scene(path, w, h, spp);
        *out = &o->desc;
        return TRB_OK;
    } catch (const LoadError& e) { g_lerr = e.msg; trb_internal_set_error(e.msg.c_str()); return e.st; }
    catch (const std::exception& e) { g_lerr = e.what(); trb_internal_set_error(e.what()); return TRB_INVALID_ARG; }
}

void trb_desc_free(trb_scene_desc* d) {
    if (d) delete reinterpret_cast<DescOwner*>(d); // desc is the first member
}

trb_status trb_scene_load_json(const char* path, uint32_t w, uint32_t h, uint32_t spp, int device, trb_scene** out) {
    trb_scene_desc* d = nullptr;
    trb_status r = trb_desc_load_json(path, w, h, spp, &d);
    if (r != TRB_OK) return r;
    r = trb_scene_create(d, device, out); // deep-copies
    trb_desc_free(d);
    return r;
}

} // extern "C"


// ------------------------------------------------------------------------------------------------------------
// PNG output (the output half): what `image::save_buffer(path, &img, w, h, image::RGB(8))` does for the frames
// of main.rs:95-103 / master.rs:137-142 — 8-bit RGB, no interlace, filter type 0 on every scanline. The zlib stream uses
// stored (uncompressed) deflate blocks: any PNG reader accepts it, and the library needs no compression dependency.
// ------------------------------------------------------------------------------------------------------------
namespace {
uint32_t png_crc(const uint8_t* p, size_t n, uint32_t crc) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}
void png_be32(std::vector<uint8_t>& b, uint32_t v) { b.push_back(v >> 24); b.push_back(v >> 16); b.push_back(v >> 8); b.push_back(v); }
void png_chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
    png_be32(out, (uint32_t)data.size());
    const size_t o = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    png_be32(out, png_crc(&out[o], out.size() - o, 0xffffffffu) ^ 0xffffffffu);
}
} // namespace

extern "C" trb_status trb_write_png(const char* path, const uint8_t* rgb8, uint32_t width, uint32_t height) {
    if (!path || !rgb8 || width == 0 || height == 0) { trb_internal_set_error("null argument"); return TRB_INVALID_ARG; }
    std::vector<uint8_t> raw; // scanlines, each prefixed with filter type 0
    raw.reserve((size_t)height * (3 * (size_t)width + 1));
    for (uint32_t y = 0; y < height; ++y) { raw.push_back(0); raw.insert(raw.end(), rgb8 + (size_t)y * width * 3, rgb8 + (size_t)(y + 1) * width * 3); }
    std::vector<uint8_t> z = {0x78, 0x01}; // zlib header, then stored deflate blocks of <= 65535 bytes
    uint32_t a = 1, b = 0;                 // Adler-32 of the raw data
    for (size_t o = 0; o < raw.size();) {
        const size_t n = std::min<size_t>(65535, raw.size() - o);
        z.push_back(o + n == raw.size() ? 1 : 0);
        z.push_back(n & 0xff); z.push_back(n >> 8); z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
        z.insert(z.end(), raw.begin() + o, raw.begin() + o + n);
        for (size_t i = o; i < o + n; ++i) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
        o += n;
    }
    png_be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    png_be32(ihdr, width); png_be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); // 8 bits, colour type 2 (RGB), deflate, adaptive filtering, no interlace
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", z);
    png_chunk(out, "IEND", {});
    FILE* f = std::fopen(path, "wb");
    if (!f) { trb_internal_set_error((std::string("cannot open ") + path).c_str()); return TRB_IO; }
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    if (!ok) { trb_internal_set_error("short write"); return TRB_IO; }
    return TRB_OK;
}
