// trb_api.cu — the C ABI of include/trb.h: scene upload, per-frame update and kernel launches.
// Everything that touches device memory lives here; there is no CPU rendering path in this library.
#include <chrono>
#include <dlfcn.h>
#include <nccl.h>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>
#include "trb_host.h"
#include "trb_kernels.cuh"

using namespace trbh;

namespace {

thread_local std::string g_error;
unsigned long long g_launches = 0; // kernels launched by this library (bench.py reports it as gpu_launches)
trb_status fail(trb_status s, const std::string& msg) { g_error = msg; return s; }

#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? TRB_OOM : TRB_CUDA,                   \
                                           std::string(#call) + ": " + cudaGetErrorString(e_));                    \
    } while (0)

struct DeviceArena { // owns every cudaMalloc of a scene
    std::vector<void*> ptrs;
    ~DeviceArena() { for (void* p : ptrs) cudaFree(p); }
    template <class T>
    cudaError_t upload(const T* host, size_t n, T** out) {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, std::max<size_t>(1, n) * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(d);
        if (n) e = cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice);
        *out = static_cast<T*>(d);
        return e;
    }
    template <class T>
    cudaError_t alloc(size_t n, T** out) {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, std::max<size_t>(1, n) * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(d);
        *out = static_cast<T*>(d);
        return e;
    }
};

struct HostMesh {
    std::vector<float> pos, nrm, uv;
    std::vector<uint32_t> idx;
    std::vector<trb_bvh_node> nodes;
    std::vector<uint32_t> order;
    Box3 bounds;
};

uint32_t pow2_ceil(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

// Re-layout of the reference-order flat tree into child-pair records (trb_device.h DPair). Pure layout: no box,
// child order or primitive order changes. Returns false if a leaf does not fit the 25-bit slot / 5-bit count fields.
float bits_f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
constexpr uint32_t QUAD_EMPTY_HOST = 0xffffffffu;
bool pack_pairs(const std::vector<trb_bvh_node>& in, std::vector<trb::DPair>& out, trb::DBvh& hdr) {
    std::vector<uint32_t> rec_of(in.size(), 0);
    uint32_t n_rec = 0;
    for (size_t i = 0; i < in.size(); ++i) if (!(in[i].b & TRB_BVH_LEAF)) rec_of[i] = n_rec++;
    bool ok = n_rec < (1u << 30);
    auto ref_of = [&](uint32_t i) -> uint32_t {
        if (in[i].b & TRB_BVH_LEAF) {
            const uint32_t cnt = in[i].b & ~TRB_BVH_LEAF, first = in[i].a;
            if (cnt > 31 || first >= (1u << 25)) ok = false;
            return trb::REF_LEAF | (cnt << 25) | first;
        }
        return trb::REF_INTERIOR | rec_of[i];
    };
    out.resize(n_rec);
    for (size_t i = 0; i < in.size(); ++i) {
        if (in[i].b & TRB_BVH_LEAF) continue;
        const trb_bvh_node& l = in[i + 1];
        const trb_bvh_node& r = in[in[i].a];
        trb::DPair& p = out[rec_of[i]];
        p.l_lo = make_float4(l.bmin[0], l.bmin[1], l.bmin[2], bits_f(ref_of((uint32_t)i + 1)));
        p.l_hi = make_float4(l.bmax[0], l.bmax[1], l.bmax[2], bits_f(ref_of(in[i].a)));
        p.r_lo = make_float4(r.bmin[0], r.bmin[1], r.bmin[2], bits_f(in[i].b));
        p.r_hi = make_float4(r.bmax[0], r.bmax[1], r.bmax[2], 0.f);
    }
    hdr.root_lo = make_float4(in[0].bmin[0], in[0].bmin[1], in[0].bmin[2], bits_f(ref_of(0)));
    hdr.root_hi = make_float4(in[0].bmax[0], in[0].bmax[1], in[0].bmax[2], 0.f);
    return ok;
}

// Collapse pairs of levels of the reference-order tree into DQuad records (trb_device.h). Layout only: boxes, child
// order and primitive order are the reference's. A child whose box is not inside its parent's (cannot happen for boxes
// built as unions, checked anyway) is kept as a one-slot half so the containment argument never has to be trusted.
// Returns the root reference in DQuad index space.
bool pack_quads(const std::vector<trb_bvh_node>& in, std::vector<trb::DQuad>& out, uint32_t& root_ref) {
    out.clear();
    bool ok = true;
    auto is_leaf = [&](uint32_t i) { return (in[i].b & TRB_BVH_LEAF) != 0; };
    auto leaf_ref = [&](uint32_t i) -> uint32_t {
        const uint32_t cnt = in[i].b & ~TRB_BVH_LEAF, first = in[i].a;
        if (cnt > 31 || first >= (1u << 25)) ok = false;
        return trb::REF_LEAF | (cnt << 25) | first;
    };
    auto inside = [&](uint32_t c, uint32_t p) {
        for (int k = 0; k < 3; ++k) if (!(in[c].bmin[k] >= in[p].bmin[k] && in[c].bmax[k] <= in[p].bmax[k])) return false;
        return true;
    };
    if (in.empty()) { root_ref = QUAD_EMPTY_HOST; return true; }
    if (is_leaf(0)) { root_ref = leaf_ref(0); return ok; }
    // quad roots in depth-first order (a record is followed by the records below its first slots: locality for near-first descent)
    std::vector<uint32_t> quad_of(in.size(), 0xffffffffu), order, todo{0};
    while (!todo.empty()) {
        const uint32_t p = todo.back(); todo.pop_back();
        quad_of[p] = (uint32_t)order.size(); order.push_back(p);
        uint32_t kids[4]; int nk = 0;
        for (uint32_t c : {p + 1, in[p].a}) {
            if (is_leaf(c)) continue;
            const uint32_t g0 = c + 1, g1 = in[c].a;
            if (inside(g0, c) && inside(g1, c)) { if (!is_leaf(g0)) kids[nk++] = g0; if (!is_leaf(g1)) kids[nk++] = g1; }
            else kids[nk++] = c;
        }
        for (int k = nk; k-- > 0;) todo.push_back(kids[k]);
    }
    if (order.size() >= (1u << 30)) return false;
    out.resize(order.size());
    auto slot = [&](trb::DQuad& q, int k, uint32_t node, bool empty) {
        if (empty) { q.q[2 * k] = make_float4(0.f, 0.f, 0.f, bits_f(0xffffffffu)); q.q[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
        const uint32_t ref = is_leaf(node) ? leaf_ref(node) : (trb::REF_INTERIOR | quad_of[node]);
        q.q[2 * k] = make_float4(in[node].bmin[0], in[node].bmin[1], in[node].bmin[2], bits_f(ref));
        q.q[2 * k + 1] = make_float4(in[node].bmax[0], in[node].bmax[1], in[node].bmax[2], 0.f);
    };
    for (size_t qi = 0; qi < order.size(); ++qi) {
        const uint32_t p = order[qi];
        trb::DQuad& q = out[qi];
        uint32_t axes[2] = {0, 0};
        int h = 0;
        for (uint32_t c : {p + 1, in[p].a}) {
            bool split = false;
            if (!is_leaf(c)) { const uint32_t g0 = c + 1, g1 = in[c].a; split = inside(g0, c) && inside(g1, c); }
            if (split) { slot(q, 2 * h, c + 1, false); slot(q, 2 * h + 1, in[c].a, false); axes[h] = in[c].b & 3u; }
            else { slot(q, 2 * h, c, false); slot(q, 2 * h + 1, 0, true); }
            ++h;
        }
        q.q[1].w = bits_f((in[p].b & 3u) | axes[0] << 2 | axes[1] << 4);
    }
    root_ref = trb::REF_INTERIOR | quad_of[0];
    return ok;
}

} // namespace

namespace {
// Launch-shape knobs of the wavefront pipeline. Read ONCE from the environment when the scene is created (developer
// sweeps), changed afterwards only through trb_scene_set_option: the launch path never touches getenv.
struct Tuning {
    int refill = 8;            // trace: idle lanes that trigger a warp refill
    unsigned trace_grid = 0;   // trace: CTAs per SM launched (0 = twice what the chosen variant keeps resident)
    uint32_t sched = 6;        // trace: quorum of the phased loop (0 = flat state machine)
    int quads = 0;             // trace: DQuad two-level records (measured slower on C4)
    int exact_box = 0;         // trace: test option — every ray takes the literal BBox::fast_intersect transcription (box_hit) instead of box_hit_finite
    int pipe = 36;             // trace: kernel variant. 0 = round-1 kernel; 1 = + box_hit_finite; 33 = + RayHome + fused non-node chains at 7 CTAs per SM; 34 / 35 / 36 / 37 = the same at 8 / 8 / 9 / 9 CTAs with 16 / 12 / 12 / 8 stack entries in shared memory
    int film_v2 = 1;           // film: per-warp private tiles (0 = shared-memory atomics)
    int sort = 0;              // ray queues: 0 = path order; 1 / 2 = counting sort by (octant, origin cell) / (cell, octant) before each trace round (measured: -1.5 % trace time, +10 % step time on C4)
    int sort_bits = 5;         // bits per axis of the origin cell grid
    int sort_min_round = 1;    // first bounce round whose queues are sorted (round 0 = primary rays, already coherent)
    int shade_anim_occ = 4;    // keyframed shade kernel variant: resident CTAs per SM it is compiled for (3 or 4)
    int frame_device = 1;      // Scene::update_frame on the device (instance transforms, animation bounds, TLAS SAH build); 0 = on the host
    int anim_table = 2;        // keyframed scenes: evaluate each keyframed instance's transform once per path (0 = per ray per instance, like the reference; 1 = one thread per (path, instance) evaluates the whole stack; 2 = each distinct keyframed spline once per path, then the stacks)
    int shade_split = -1;      // shading as three kernels (surface | direct light | BSDF sample) instead of one: 1 / 0, -1 = per scene — split
                               // when the scene mixes material kinds or uses MERL (tr15: 274 -> 337 Mrays/s, tr15-like 388 -> 577), fused for
                               // one-material scenes like C4 (134.4 vs 134.9 ms per step)
    int shade_sort = 1;        // split shading: bucket the paths by material kind between k_wf_shade_a and _b / _c
    int shade_kind = 1;        // split shading with buckets: the matte bucket goes through _b / _c instantiations compiled for matte alone
    uint64_t pass_paths = 1ull << 24; // camera samples per wavefront pass (the frame is rendered in additive passes)
};
int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? (int)strtol(v, nullptr, 0) : dflt; }
void tuning_from_env(Tuning& t) {
    t.refill = env_int("TRB_REFILL", t.refill); t.trace_grid = (unsigned)env_int("TRB_TRACE_GRID", (int)t.trace_grid);
    t.sched = (uint32_t)env_int("TRB_TRACE_SCHED", (int)t.sched); t.quads = env_int("TRB_TRACE_QUADS", t.quads); t.pipe = env_int("TRB_TRACE_PIPE", t.pipe);
    t.film_v2 = env_int("TRB_FILM_V2", t.film_v2); t.sort = env_int("TRB_SORT", t.sort); t.sort_bits = env_int("TRB_SORT_BITS", t.sort_bits);
    t.sort_min_round = env_int("TRB_SORT_MIN_ROUND", t.sort_min_round); t.shade_split = env_int("TRB_SHADE_SPLIT", t.shade_split); t.anim_table = env_int("TRB_ANIM_TABLE", t.anim_table); t.frame_device = env_int("TRB_FRAME_DEVICE", t.frame_device);
    if (getenv("TRB_PASS_PATHS")) t.pass_paths = strtoull(getenv("TRB_PASS_PATHS"), nullptr, 0);
}

} // namespace

struct BlockList { // one selection of the Morton block list, resident on the device (never overwritten in place: kernels in flight may read it)
    uint32_t key[5];
    std::vector<uint32_t> host; // (bx, by) pairs
    uint2* dev = nullptr;
};

struct trb_scene {
    int device = 0;
    int sm_count = 148;
    Tuning tune;
    // deep copy of the description
    trb_film film{};
    trb_integrator integrator{};
    std::vector<trb_camera> cameras;
    std::vector<trb_instance> instances;
    std::vector<trb_spline> splines;
    std::vector<trb_keyframe> keyframes;
    std::vector<float> knots;
    std::vector<trb_color_key> color_keys;
    std::vector<float> fov_floats;
    std::vector<trb_material> materials;
    std::vector<HostMesh> meshes;
    uint32_t spp_pow2 = 1;
    uint32_t n_anim = 0;                 // instances whose transform stack is keyframed (evaluated per path into WfState::xf_tab)
    bool frame_set = false; uint32_t last_frame = 0; float last_start = 0, last_end = 0; // the arguments of the last update_frame (re-run when an option changes what it builds)
    uint32_t material_kinds = 0;         // bit k: some hittable instance's material is of kind k (TRB_MAT_*)
    bool mixed_materials = false;        // the hittable instances use >= 2 material kinds or a MERL table: the split shade kernels with material buckets win (Tuning::shade_split = -1)
    uint32_t* d_anim_instances = nullptr;
    // per-frame host state
    int active_camera = -1;
    float shutter_open = 0, shutter_close = 0;
    std::vector<Xf> world; // instance world transforms of the current frame
    std::vector<trb_bvh_node> tlas_nodes;
    std::vector<uint32_t> tlas_order;
    float table[256];
    // device
    DeviceArena arena;
    trb::DScene ds{};
    trb::DInstance* d_instances = nullptr;
    trb::DPair* d_tlas = nullptr;
    trb::DQuad* d_tlas_quads = nullptr;
    trb::DBvh* d_tlas_hdr = nullptr;
    uint32_t* d_tlas_order = nullptr;
    size_t tlas_capacity = 0;
    // device-side update_frame (k_frame_instances / k_tlas_build): reference-order nodes, instance bounds, builder scratch
    trb_bvh_node* d_tlas_nodes = nullptr;
    float* d_bounds = nullptr;           // n x Box3 (6 floats)
    float* d_build_f = nullptr;
    uint32_t* d_build_u = nullptr;
    uint32_t* d_build_counts = nullptr;
    uint32_t tlas_n_nodes = 0;
    bool instances_static_uploaded = false, host_frame_stale = false;
    std::vector<BlockList> block_lists;
    uint32_t* d_counter = nullptr;
    int* d_error = nullptr;
    trb::DStats* d_stats = nullptr;
    float4* d_film = nullptr;
    float* h_film_staging = nullptr; // pinned
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool frame_ready = false;
    // wavefront path state (grown on demand, owned outside the arena so it can be re-sized)
    trb::WfState wf{};
    size_t wf_capacity = 0;
    std::vector<void*> wf_allocs;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> trace_events; // TRB_RENDER_TIME_TRACE
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> event_pool;
    ~trb_scene() {
        for (auto& b : block_lists) cudaFree(b.dev);
        for (void* p : wf_allocs) cudaFree(p);
        for (auto& e : trace_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto& e : event_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        if (h_film_staging) cudaFreeHost(h_film_staging);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

namespace {

Box3 shape_bounds(const trb_scene& s, const trb_instance& in) {
    Box3 b;
    switch (in.shape) {
        case TRB_SHAPE_SPHERE: for (int i = 0; i < 3; ++i) { b.lo[i] = -in.p0; b.hi[i] = in.p0; } break;          // sphere.rs:84-88
        case TRB_SHAPE_DISK: b.lo[0] = b.lo[1] = -in.p0; b.hi[0] = b.hi[1] = in.p0; b.lo[2] = -0.1f; b.hi[2] = 0.1f; break; // disk.rs:79-81
        case TRB_SHAPE_RECT: { const float hw = in.p0 / 2.0f, hh = in.p1 / 2.0f; b.lo[0] = -hw; b.lo[1] = -hh; b.hi[0] = hw; b.hi[1] = hh; b.lo[2] = b.hi[2] = 0.0f; break; } // rectangle.rs:67-71
        case TRB_SHAPE_MESH: b = s.meshes[in.mesh].bounds; break;                                                   // mesh.rs:87-90
        default: for (int i = 0; i < 3; ++i) b.lo[i] = b.hi[i] = 0.0f;                                            // point light (emitter.rs:152)
    }
    return b;
}

trb_status validate(const trb_scene_desc* d) {
    if (!d) return fail(TRB_INVALID_ARG, "null scene description");
    if (d->abi_version != TRB_ABI_VERSION) return fail(TRB_INVALID_ARG, "trb_scene_desc.abi_version mismatch");
    if (d->film.width == 0 || d->film.height == 0 || d->film.width % 8 || d->film.height % 8)
        return fail(TRB_INVALID_ARG, "Image not evenly divided by blocks of (8, 8)"); // block_queue.rs:29-31
    if (d->film.frames == 0) return fail(TRB_INVALID_ARG, "film.frames must be >= 1");
    if (d->n_instances == 0) return fail(TRB_INVALID_ARG, "Aborting: the scene does not have any objects!"); // scene.rs:134
    if (d->n_cameras == 0) return fail(TRB_INVALID_ARG, "Error: A camera is required!");
    if (d->integrator.type > TRB_INTEGRATOR_NORMALS_DEBUG) return fail(TRB_INVALID_ARG, "Unrecognized integrator type"); // scene.rs:313
    if (d->integrator.type == TRB_INTEGRATOR_PATH && d->integrator.max_depth > 57u) return fail(TRB_UNSUPPORTED, "max_depth > 57");
    if (d->integrator.type == TRB_INTEGRATOR_WHITTED && d->integrator.max_depth > 24u) return fail(TRB_UNSUPPORTED, "whitted max_depth > 24 (device recursion stack)");
    if (!(d->film.filter_w > 0.0f && d->film.filter_h > 0.0f)) return fail(TRB_INVALID_ARG, "filter width/height must be positive");
    if (floorf(d->film.filter_w / 0.5f) > 8.0f || floorf(d->film.filter_h / 0.5f) > 8.0f) return fail(TRB_UNSUPPORTED, "filter wider than 4 pixels");
    bool light = false;
    for (uint32_t i = 0; i < d->n_instances; ++i) {
        const trb_instance& in = d->instances[i];
        if (in.kind > TRB_INST_EMITTER_POINT || in.shape > TRB_SHAPE_MESH) return fail(TRB_INVALID_ARG, "unknown instance kind/shape");
        if (in.kind != TRB_INST_EMITTER_POINT && in.shape == TRB_SHAPE_NONE) return fail(TRB_INVALID_ARG, "instance without geometry");
        if (in.kind == TRB_INST_EMITTER_AREA && in.shape == TRB_SHAPE_MESH)
            return fail(TRB_INVALID_ARG, "Geometry of type 'mesh' is not sampleable and can't be used for area light geometry"); // scene.rs:577-579
        if (in.shape == TRB_SHAPE_MESH && in.mesh >= d->n_meshes) return fail(TRB_INVALID_ARG, "mesh index out of range");
        if (in.kind != TRB_INST_EMITTER_POINT && in.material >= d->n_materials) return fail(TRB_INVALID_ARG, "material index out of range");
        if ((uint64_t)in.spline_first + in.n_splines > d->n_splines) return fail(TRB_INVALID_ARG, "spline range out of bounds");
        if (in.kind != TRB_INST_RECEIVER) {
            light = true;
            if (in.n_emission == 0 || (uint64_t)in.emission_first + in.n_emission > d->n_color_keys) return fail(TRB_INVALID_ARG, "An emission color is required for emitters");
        }
    }
    for (uint32_t k = 0; k < d->n_splines; ++k) { // BSpline::new's invariants (bspline 0.2.2) + the device evaluator's degree cap
        const trb_spline& sp = d->splines[k];
        if (sp.n_ctrl == 0 || (uint64_t)sp.ctrl_first + sp.n_ctrl > d->n_keyframes) return fail(TRB_INVALID_ARG, "spline control points out of bounds");
        if (sp.n_ctrl > 1) {
            if ((uint64_t)sp.knot_first + sp.n_knots > d->n_knots) return fail(TRB_INVALID_ARG, "spline knots out of bounds");
            if (sp.n_ctrl <= sp.degree) return fail(TRB_INVALID_ARG, "Too few control points for curve"); // BSpline::new panics with this message
            if ((uint64_t)sp.n_knots != (uint64_t)sp.n_ctrl + sp.degree + 1) return fail(TRB_INVALID_ARG, "Invalid B-spline: knots.len() != control_points.len() + degree + 1");
            if (sp.degree > (uint32_t)trbh::kMaxSplineDegree) return fail(TRB_UNSUPPORTED, "B-spline degree above 5");
            for (uint32_t i = 0; i < sp.n_knots; ++i) if (d->knots[sp.knot_first + i] != d->knots[sp.knot_first + i]) return fail(TRB_INVALID_ARG, "NaN knot in B-spline"); // BSpline::new sorts with partial_cmp().unwrap(): panics on NaN
        }
    }
    if (!light) return fail(TRB_INVALID_ARG, "At least one light is required"); // multithreaded.rs:39
    for (uint32_t i = 0; i < d->n_cameras; ++i) {
        const trb_camera& c = d->cameras[i];
        if (c.n_fov_ctrl) { // CameraFov::Animated (camera.rs:95-125)
            if ((uint64_t)c.fov_ctrl_first + c.n_fov_ctrl > d->n_fov_floats || (uint64_t)c.fov_knot_first + c.n_fov_knots > d->n_fov_floats) return fail(TRB_INVALID_ARG, "fov spline out of bounds");
            if (c.n_fov_ctrl <= c.fov_degree) return fail(TRB_INVALID_ARG, "Too few control points for curve");
            if ((uint64_t)c.n_fov_knots != (uint64_t)c.n_fov_ctrl + c.fov_degree + 1) return fail(TRB_INVALID_ARG, "Invalid B-spline: knots.len() != control_points.len() + degree + 1");
            for (uint32_t i = 0; i < c.n_fov_knots; ++i) if (d->fov_floats[c.fov_knot_first + i] != d->fov_floats[c.fov_knot_first + i]) return fail(TRB_INVALID_ARG, "NaN knot in B-spline");
            if (c.fov_degree > (uint32_t)trbh::kMaxSplineDegree) return fail(TRB_UNSUPPORTED, "B-spline degree above 5");
        }
        if ((uint64_t)c.spline_first + c.n_splines > d->n_splines) return fail(TRB_INVALID_ARG, "camera spline range out of bounds");
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        if (d->materials[i].type > TRB_MAT_MERL) return fail(TRB_INVALID_ARG, "unrecognized material type");
        if (d->materials[i].type == TRB_MAT_MERL && d->materials[i].merl >= d->n_merl) return fail(TRB_INVALID_ARG, "merl table index out of range");
        for (int k = 0; k < 4; ++k) if (d->materials[i].tex[k] > d->n_textures) return fail(TRB_INVALID_ARG, "texture index out of range");
    }
    uint64_t texels = 0;
    for (uint32_t i = 0; i < d->n_textures; ++i)
        if (d->textures[i].n_images == 0 || (uint64_t)d->textures[i].first_image + d->textures[i].n_images > d->n_images) return fail(TRB_INVALID_ARG, "texture image range out of bounds");
    for (uint32_t i = 0; i < d->n_images; ++i) {
        if (d->images[i].width == 0 || d->images[i].height == 0 || !d->images[i].rgba8) return fail(TRB_INVALID_ARG, "empty image");
        texels += (uint64_t)d->images[i].width * d->images[i].height;
    }
    if (texels >= (1ull << 32)) return fail(TRB_UNSUPPORTED, "more than 2^32 texels of image textures");
    for (uint32_t i = 0; i < d->n_meshes; ++i) {
        const trb_mesh& m = d->meshes[i];
        if (m.n_tris == 0 || m.n_verts == 0) return fail(TRB_INVALID_ARG, "empty mesh");
        if (!m.positions || !m.normals || !m.texcoords || !m.indices) return fail(TRB_INVALID_ARG, "Normals and texture coordinates are required!"); // mesh.rs:57-61
        for (size_t k = 0; k < 3 * (size_t)m.n_tris; ++k) if (m.indices[k] >= m.n_verts) return fail(TRB_INVALID_ARG, "mesh index out of range");
    }
    return TRB_OK;
}

// The selected, sharded block list on the device. Lists are cached per selection and never overwritten in place, so a
// pass still in flight on a caller stream keeps reading valid memory when the next call selects other blocks.
trb_status ensure_blocks(trb_scene* s, const trb_render_cfg* cfg, const uint2** d_blocks, uint32_t* n_blocks) {
    if (cfg->shard_count > 1 && cfg->shard_index >= cfg->shard_count) return fail(TRB_INVALID_ARG, "shard_index must be < shard_count");
    const uint32_t key[5] = {cfg->block_start, cfg->block_count, cfg->shard_index, cfg->shard_count, cfg->shard_chunk};
    for (const BlockList& b : s->block_lists)
        if (std::memcmp(b.key, key, sizeof key) == 0) { *d_blocks = b.dev; *n_blocks = (uint32_t)(b.host.size() / 2); return TRB_OK; }
    if (s->block_lists.size() >= 16) { // bounded cache: retire everything once nothing can be reading it any more
        CU(cudaDeviceSynchronize());
        for (BlockList& b : s->block_lists) cudaFree(b.dev);
        s->block_lists.clear();
    }
    BlockList bl;
    std::memcpy(bl.key, key, sizeof key);
    bl.host = morton_blocks(s->film.width, s->film.height, cfg->block_start, cfg->block_count, cfg->shard_index, cfg->shard_count, cfg->shard_chunk);
    const size_t n = bl.host.size() / 2;
    CU(cudaMalloc(reinterpret_cast<void**>(&bl.dev), std::max<size_t>(1, n) * sizeof(uint2)));
    if (n) { cudaError_t e = cudaMemcpy(bl.dev, bl.host.data(), n * sizeof(uint2), cudaMemcpyHostToDevice); if (e != cudaSuccess) { cudaFree(bl.dev); CU(e); } }
    s->block_lists.push_back(std::move(bl));
    *d_blocks = s->block_lists.back().dev; *n_blocks = (uint32_t)n;
    return TRB_OK;
}

trb_status resolve_samples(const trb_scene* s, const trb_render_cfg* cfg, uint32_t& spp, uint32_t& first, uint32_t& count) {
    if (cfg->spp > (1u << 31)) return fail(TRB_INVALID_ARG, "spp too large");
    spp = cfg->spp ? pow2_ceil(cfg->spp) : s->spp_pow2; // ld.rs:22-26
    first = cfg->sample_first;
    if (first > spp) return fail(TRB_INVALID_ARG, "sample_first exceeds spp");
    count = cfg->sample_count ? cfg->sample_count : spp - first;
    if ((uint64_t)first + count > spp) return fail(TRB_INVALID_ARG, "sample range exceeds spp");
    return TRB_OK;
}

template <bool STATS, int MODE, bool ANIM>
trb_status launch_render_t(trb_scene* s, const trb::RenderParams& rp, uint32_t flags, cudaStream_t st) {
    const int T = 9 + 2 * std::max(s->ds.fpw_x, s->ds.fpw_y);
    const size_t smem = (size_t)T * T * sizeof(float4);
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trb::k_render<STATS, MODE, ANIM>, trb::RENDER_THREADS, smem));
    const uint32_t grid = std::max(1u, std::min<uint32_t>(rp.n_blocks, (uint32_t)(std::max(1, per_sm) * s->sm_count)));
    trb::k_render<STATS, MODE, ANIM><<<grid, trb::RENDER_THREADS, smem, st>>>(s->ds, rp, flags);
    g_launches++;
    CU(cudaGetLastError());
    return TRB_OK;
}
trb_status ensure_wavefront(trb_scene* s, size_t n_paths) {
    if (n_paths <= s->wf_capacity) return TRB_OK;
    CU(cudaDeviceSynchronize()); // a pass still in flight owns the old buffers
    for (void* p : s->wf_allocs) cudaFree(p);
    s->wf_allocs.clear(); s->wf_capacity = 0;
    const size_t cap = n_paths;
    cudaError_t err = cudaSuccess;
    auto grab = [&](size_t bytes, void** out) {
        if (err != cudaSuccess) return;
        err = cudaMalloc(out, bytes);
        if (err == cudaSuccess) s->wf_allocs.push_back(*out);
    };
    trb::WfState& w = s->wf;
    float4** f4[] = {&w.org, &w.cont, &w.shadow, &w.mis, &w.a, &w.b, &w.tprev, &w.thr, &w.illum, &w.ng, &w.rad, &w.f_p, &w.f_n, &w.f_t, &w.f_b};
    for (float4** q : f4) grab(cap * sizeof(float4), reinterpret_cast<void**>(q));
    grab(cap * sizeof(uint4), reinterpret_cast<void**>(&w.hit));
    uint32_t** u1[] = {&w.q_active[0], &w.q_active[1], &w.q_ending[0], &w.q_ending[1], &w.q_cont, &w.q_shadow, &w.q_mis};
    for (uint32_t** q : u1) grab(cap * sizeof(uint32_t), reinterpret_cast<void**>(q));
    grab(cap * trb::WF_MID_BUCKETS * sizeof(uint32_t), reinterpret_cast<void**>(&w.q_mid)); // one list per material kind (split shading)
    grab(64 * trb::WF_CNT * sizeof(uint32_t), reinterpret_cast<void**>(&w.counters));
    // ray sorting: up to three rays per path and round; bins for the finest grid the options allow
    const size_t max_bins = (size_t)3 * (8u << (3 * trb::WF_SORT_MAX_BITS));
    uint32_t** u3[] = {&w.q_sorted, &w.sort_key, &w.sort_rank};
    for (uint32_t** q : u3) grab(3 * cap * sizeof(uint32_t), reinterpret_cast<void**>(q));
    grab(max_bins * sizeof(uint32_t), reinterpret_cast<void**>(&w.sort_hist));
    grab(max_bins * sizeof(uint32_t), reinterpret_cast<void**>(&w.sort_offs));
    grab(64 * 8 * sizeof(uint32_t), reinterpret_cast<void**>(&w.bounds));
    w.xf_tab = nullptr; w.n_anim = s->n_anim;
    if (s->n_anim) grab(cap * (size_t)s->n_anim * 32 * sizeof(float), reinterpret_cast<void**>(&w.xf_tab)); // per-path keyframed transforms (128 B per path and keyframed instance)
    if (err == cudaSuccess) err = cudaMemset(w.sort_hist, 0, max_bins * sizeof(uint32_t)); // invariant: all zero between sorts (the scan clears what it reads)
    if (err != cudaSuccess) {
        for (void* p : s->wf_allocs) cudaFree(p);
        s->wf_allocs.clear();
        cudaGetLastError();
        return fail(err == cudaErrorMemoryAllocation ? TRB_OOM : TRB_CUDA, std::string("wavefront state: ") + cudaGetErrorString(err));
    }
    s->wf_capacity = cap;
    return TRB_OK;
}

// One wavefront pass over rp's blocks x samples: generate, then (trace, shade) per bounce round, then the film
// (DESIGN.md "Execution shape"). n_paths = blocks * 64 * sample_count must fit the allocated path state.
trb_status launch_wavefront(trb_scene* s, const trb::RenderParams& rp, uint32_t flags, int mode, cudaStream_t st) {
    const size_t n_paths = (size_t)rp.n_blocks * 64 * rp.sample_count;
    if (n_paths > s->wf_capacity || n_paths >= (1ull << 30)) return fail(TRB_INVALID_ARG, "pass larger than the wavefront state");
    const Tuning& tu = s->tune;
    trb::WfState wf = s->wf;
    wf.n_paths = (uint32_t)n_paths;
    wf.mid_keyed = s->tune.shade_sort ? 1u : 0u;
    if (s->integrator.type != TRB_INTEGRATOR_PATH) { // Whitted / NormalsDebug: one thread per camera sample, then the same film kernel
        const unsigned grid = (unsigned)std::min<size_t>((n_paths + 127) / 128, (size_t)s->sm_count * 8);
        const bool anim = s->ds.has_anim != 0;
        if (anim) { if (mode == 0) trb::k_simple_integrator<0, true><<<grid, 128, 0, st>>>(s->ds, rp, wf.rad, wf.n_paths, s->integrator.type, flags);
                    else trb::k_simple_integrator<1, true><<<grid, 128, 0, st>>>(s->ds, rp, wf.rad, wf.n_paths, s->integrator.type, flags); }
        else if (mode == 0) trb::k_simple_integrator<0, false><<<grid, 128, 0, st>>>(s->ds, rp, wf.rad, wf.n_paths, s->integrator.type, flags);
        else trb::k_simple_integrator<1, false><<<grid, 128, 0, st>>>(s->ds, rp, wf.rad, wf.n_paths, s->integrator.type, flags);
        g_launches++;
        if (mode == 0) {
            const int T = 9 + 2 * std::max(s->ds.fpw_x, s->ds.fpw_y);
            const unsigned film_grid = std::min<unsigned>(rp.n_blocks, (unsigned)s->sm_count * 8);
            if (tu.film_v2) trb::k_wf_film_v2<<<film_grid, trb::RENDER_THREADS, (size_t)4 * T * T * sizeof(float4), st>>>(s->ds, rp, wf);
            else trb::k_wf_film<<<film_grid, trb::RENDER_THREADS, (size_t)T * T * sizeof(float4), st>>>(s->ds, rp, wf);
            g_launches++;
        }
        CU(cudaGetLastError());
        return TRB_OK;
    }
    const bool stats = (flags & TRB_RENDER_STATS) != 0;
    const uint32_t rounds = s->integrator.max_depth + 2; // bounces 0..max_depth, plus the round that only resolves
    CU(cudaMemsetAsync(wf.counters, 0, 64 * trb::WF_CNT * sizeof(uint32_t), st));
    const unsigned gen_grid = (unsigned)std::min<size_t>((n_paths + 255) / 256, (size_t)s->sm_count * 8);
    const bool anim = s->ds.has_anim != 0; // static scenes run kernels with no animation code in them at all
    if (anim) trb::k_wf_generate<true><<<gen_grid, 256, 0, st>>>(s->ds, rp, wf);
    else trb::k_wf_generate<false><<<gen_grid, 256, 0, st>>>(s->ds, rp, wf);
    g_launches++;
    if (anim && wf.xf_tab && s->tune.anim_table) { // AnimatedTransform::transform(ray.time) once per (path, keyframed instance)
        const size_t items = n_paths * wf.n_anim;
        const uint32_t nu = s->ds.n_uniq_splines;
        if (s->tune.anim_table >= 2 && nu > 0 && nu <= 256) { // each distinct keyframed spline once per path, in shared memory; then the stacks
            const uint32_t per_iter = std::max(1u, std::min(32u, 128u / nu));
            const size_t smem = (size_t)per_iter * nu * 32 * sizeof(float);
            const unsigned grid = (unsigned)std::min<size_t>((n_paths + per_iter - 1) / per_iter, (size_t)s->sm_count * 16);
            trb::k_wf_anim_table2<<<grid, 128, smem, st>>>(s->ds, wf, per_iter);
        } else
            trb::k_wf_anim_table<<<(unsigned)std::min<size_t>((items + 127) / 128, (size_t)s->sm_count * 16), 128, 0, st>>>(s->ds, wf);
        g_launches++;
    } else wf.xf_tab = nullptr;
    const unsigned shade_grid = (unsigned)s->sm_count * 4;
    const int refill = tu.refill;
    // persistent CTAs: two rounds of what is resident per SM (9 for the default variant, 7 keyframed, 4 with counters) unless set
    const unsigned resident = (flags & TRB_RENDER_STATS) ? 4u : (tu.pipe == 0 || tu.pipe == 1 || tu.pipe == 33 ? 7u : (tu.pipe == 34 || tu.pipe == 35 || s->ds.has_anim ? 8u : 9u));
    const unsigned tgrid = (unsigned)s->sm_count * (tu.trace_grid ? tu.trace_grid : 2u * resident);
    const uint32_t sched = tu.sched;   // 0 = flat state machine; else the quorum of the phased loop (see k_wf_trace)
    const bool quads = tu.quads != 0;  // DQuad two-level records (never in the STATS variants: their counters are the reference's)
    const uint32_t tflags = flags | (tu.exact_box ? trb::WF_TRACE_FORCE_EXACT_BOX : 0u);
    for (uint32_t round = 0; round < rounds; ++round) {
        const uint32_t* q_sorted = nullptr;
        if (tu.sort && (int)round >= tu.sort_min_round) { // counting sort of this round's rays by (type, octant, origin cell): DESIGN.md "Ray sorting"
            const uint32_t bits = (uint32_t)std::min(trb::WF_SORT_MAX_BITS, std::max(1, tu.sort_bits));
            const unsigned sgrid = (unsigned)s->sm_count * 8;
            trb::k_wf_sort_count<<<sgrid, 256, 0, st>>>(wf, round, bits, tu.sort == 2 ? 1u : 0u);
            trb::k_wf_sort_scan<<<1, 1024, 0, st>>>(wf, 3u * (8u << (3u * bits)));
            trb::k_wf_sort_scatter<<<sgrid, 256, 0, st>>>(wf, round);
            g_launches += 3;
            q_sorted = wf.q_sorted;
        }
        std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
        if (flags & TRB_RENDER_TIME_TRACE) {
            if (!s->event_pool.empty()) { ev = s->event_pool.back(); s->event_pool.pop_back(); }
            else { CU(cudaEventCreate(&ev.first)); CU(cudaEventCreate(&ev.second)); }
            CU(cudaEventRecord(ev.first, st));
        }
#define TRB_TRACE_LAUNCH(ST, MB, SS, AN, PH, QD, PIPE) trb::k_wf_trace<ST, MB, SS, AN, PH, QD, PIPE><<<tgrid, 128, 0, st>>>(s->ds, rp, wf, round, tflags, refill, sched, q_sorted)
        // PIPE 33 = box_hit_finite + RayHome + fused non-node chains (trb_kernels.cuh); tu.pipe picks the variant and how many
        // CTAs per SM it is compiled for: 36 (default) = 9 CTAs / 12 stack entries in shared memory. The STATS and keyframed
        // variants run the same code at their own occupancy, so the parity tests' counters cover it.
        const bool v2 = tu.pipe != 0;
        if (anim) {
            if (stats) { if (v2) TRB_TRACE_LAUNCH(true, 4, 16, true, true, false, 33); else TRB_TRACE_LAUNCH(true, 4, 16, true, true, false, 0); }
            else if (!v2) TRB_TRACE_LAUNCH(false, 7, 16, true, true, false, 0);
            else if (tu.pipe == 33) TRB_TRACE_LAUNCH(false, 7, 16, true, true, false, 33);
            else if (tu.pipe == 37) TRB_TRACE_LAUNCH(false, 9, 12, true, true, false, 33);
            else TRB_TRACE_LAUNCH(false, 8, 12, true, true, false, 33); // keyframed default: 8 CTAs per SM (353 vs 352 / 346 Mrays/s at 7 / 9 on tr15.json)
        } else if (stats) {
            if (sched == 0) TRB_TRACE_LAUNCH(true, 4, 16, false, false, false, 0);
            else if (v2) TRB_TRACE_LAUNCH(true, 4, 16, false, true, false, 33); else TRB_TRACE_LAUNCH(true, 4, 16, false, true, false, 0);
        } else if (sched == 0) TRB_TRACE_LAUNCH(false, 7, 16, false, false, false, 0); // flat state machine (no phases)
        else if (quads) TRB_TRACE_LAUNCH(false, 7, 16, false, true, true, 0);
        else switch (tu.pipe) {
            case 0: TRB_TRACE_LAUNCH(false, 7, 16, false, true, false, 0); break;  // round-1 kernel
            case 1: TRB_TRACE_LAUNCH(false, 7, 16, false, true, false, 1); break;  // + box_hit_finite only
            case 33: TRB_TRACE_LAUNCH(false, 7, 16, false, true, false, 33); break;
            case 34: TRB_TRACE_LAUNCH(false, 8, 16, false, true, false, 33); break;
            case 35: TRB_TRACE_LAUNCH(false, 8, 12, false, true, false, 33); break;
            case 37: TRB_TRACE_LAUNCH(false, 9, 8, false, true, false, 33); break;
            default: TRB_TRACE_LAUNCH(false, 9, 12, false, true, false, 33); break; // 36
        }
#undef TRB_TRACE_LAUNCH
        if (ev.first) { CU(cudaEventRecord(ev.second, st)); s->trace_events.push_back(ev); }
        // per scene (shade_split < 0): split when the scene mixes material kinds, and also when it is all matte — the matte instantiations of
        // _b / _c beat the fused kernel (C4: 1066 vs 1027 Mrays/s, profiles/r02_c26_matte_instantiation.log); other one-kind scenes stay fused
        const bool split_auto = s->mixed_materials || (tu.shade_sort && tu.shade_kind && s->material_kinds == (1u << TRB_MAT_MATTE));
        if (tu.shade_split > 0 || (tu.shade_split < 0 && split_auto)) { // three kernels with fewer live values each (DESIGN.md "Split shading"); same device functions, same results
            // with the material buckets, _b and _c run once per kind that has an instantiation of its own (matte: the commonest), and once
            // for the other kinds the scene uses
            const uint32_t all = 0xffu, own = (tu.shade_sort && tu.shade_kind) ? (s->material_kinds & (1u << TRB_MAT_MATTE)) : 0u;
            const uint32_t rest = tu.shade_sort ? (s->material_kinds & ~own) : all;
            if (anim) {
                if (mode == 0) trb::k_wf_shade_a<0, true, 3><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round);
                else trb::k_wf_shade_a<1, true, 3><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round);
                if (own) trb::k_wf_shade_b<true, 4, TRB_MAT_MATTE><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round, own);
                if (rest) trb::k_wf_shade_b<true, 3><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round, rest);
                if (mode == 0) {
                    if (own) trb::k_wf_shade_c<0, true, 6, TRB_MAT_MATTE><<<(unsigned)s->sm_count * 6, 128, 0, st>>>(s->ds, rp, wf, round, own);
                    if (rest) trb::k_wf_shade_c<0, true, 4><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round, rest);
                } else {
                    if (own) trb::k_wf_shade_c<1, true, 6, TRB_MAT_MATTE><<<(unsigned)s->sm_count * 6, 128, 0, st>>>(s->ds, rp, wf, round, own);
                    if (rest) trb::k_wf_shade_c<1, true, 4><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round, rest);
                }
            } else {
                const unsigned ga = (unsigned)s->sm_count * 6, gb = (unsigned)s->sm_count * 5, gc = (unsigned)s->sm_count * 6;
                if (mode == 0) trb::k_wf_shade_a<0, false, 6><<<ga, 128, 0, st>>>(s->ds, rp, wf, round);
                else trb::k_wf_shade_a<1, false, 6><<<ga, 128, 0, st>>>(s->ds, rp, wf, round);
                if (own) trb::k_wf_shade_b<false, 5, TRB_MAT_MATTE><<<gb, 128, 0, st>>>(s->ds, rp, wf, round, own);
                if (rest) trb::k_wf_shade_b<false, 5><<<gb, 128, 0, st>>>(s->ds, rp, wf, round, rest);
                if (mode == 0) {
                    if (own) trb::k_wf_shade_c<0, false, 6, TRB_MAT_MATTE><<<gc, 128, 0, st>>>(s->ds, rp, wf, round, own);
                    if (rest) trb::k_wf_shade_c<0, false, 6><<<gc, 128, 0, st>>>(s->ds, rp, wf, round, rest);
                } else {
                    if (own) trb::k_wf_shade_c<1, false, 6, TRB_MAT_MATTE><<<gc, 128, 0, st>>>(s->ds, rp, wf, round, own);
                    if (rest) trb::k_wf_shade_c<1, false, 6><<<gc, 128, 0, st>>>(s->ds, rp, wf, round, rest);
                }
            }
            g_launches += 4;
            continue;
        }
        // (occupancy 5 / 6 variants of the shade kernel were measured 1-2 % slower: spills outweigh the extra warps)
        if (anim) { // keyframed variant: 168 registers at 3 CTAs per SM, or capped to 128 (some spills) at 4
            if (mode != 0) trb::k_wf_shade<1, true, 3><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round);
            else if (tu.shade_anim_occ >= 4) trb::k_wf_shade<0, true, 4><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round);
            else trb::k_wf_shade<0, true, 3><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round);
        } else if (mode == 0) trb::k_wf_shade<0, false, 4><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round);
        else trb::k_wf_shade<1, false, 4><<<shade_grid, 128, 0, st>>>(s->ds, rp, wf, round);
        g_launches += 2;
    }
    if (mode == 0) {
        const int T = 9 + 2 * std::max(s->ds.fpw_x, s->ds.fpw_y);
        const unsigned film_grid = std::min<unsigned>(rp.n_blocks, (unsigned)s->sm_count * 8);
        if (tu.film_v2) trb::k_wf_film_v2<<<film_grid, trb::RENDER_THREADS, (size_t)4 * T * T * sizeof(float4), st>>>(s->ds, rp, wf);
        else trb::k_wf_film<<<film_grid, trb::RENDER_THREADS, (size_t)T * T * sizeof(float4), st>>>(s->ds, rp, wf);
        g_launches++;
    }
    CU(cudaGetLastError());
    return TRB_OK;
}

// Exec::render renders every selected block at the full spp in one call (multithreaded.rs:55-114). The wavefront keeps
// 212 B of path state per camera sample in HBM, so a frame is cut into additive passes of at most `pass_paths` camera
// samples (fewer if device memory is short): all selected blocks x a sample sub-range, or — for images with more than
// pass_paths / 64 blocks — a block sub-range x one sample. A camera sample's radiance is a pure function of
// (scene, seed, pixel, sample index), so the split changes nothing but the order of the film's float additions.
trb_status render_passes(trb_scene* s, trb::RenderParams rp, uint32_t flags, int mode, cudaStream_t st) {
    const uint32_t nb = rp.n_blocks, first = rp.sample_first, count = rp.sample_count;
    const uint2* blocks = rp.blocks;
    const uint64_t total = (uint64_t)nb * 64 * count;
    uint64_t want = std::min<uint64_t>(total, std::min<uint64_t>(std::max<uint64_t>(s->tune.pass_paths, 64), (1ull << 30) - 64));
    if (mode == 1) { // per-sample records are indexed by the path number of ONE pass
        if (total >= (1ull << 30)) return fail(TRB_INVALID_ARG, "sample buffer too large: blocks*64*sample_count must be < 2^30; select fewer blocks or samples");
        want = total;
    }
    want = ((want + 63) / 64) * 64;
    if (want > s->wf_capacity) {
        trb_status r;
        while ((r = ensure_wavefront(s, (size_t)want)) == TRB_OOM && want > (1u << 16) && mode == 0) want = ((want / 2 + 63) / 64) * 64;
        if (r != TRB_OK) return r;
    }
    const uint64_t cap = std::max<uint64_t>(want, 64);  // paths per pass actually used (a larger state left by an earlier call is not required)
    uint32_t bp, sp;
    if ((uint64_t)nb * 64 <= cap) { bp = nb; sp = (uint32_t)std::min<uint64_t>(count, cap / ((uint64_t)nb * 64)); }
    else { bp = (uint32_t)(cap / 64); sp = 1; }
    for (uint32_t b0 = 0; b0 < nb; b0 += bp)
        for (uint32_t s0 = 0; s0 < count; s0 += sp) {
            rp.blocks = blocks + b0; rp.n_blocks = std::min(bp, nb - b0);
            rp.sample_first = first + s0; rp.sample_count = std::min(sp, count - s0);
            trb_status r = launch_wavefront(s, rp, flags, mode, st);
            if (r != TRB_OK) return r;
        }
    return TRB_OK;
}

trb_status launch_render(trb_scene* s, const trb::RenderParams& rp, uint32_t flags, int mode, cudaStream_t st) {
    if (!(flags & TRB_RENDER_MEGAKERNEL) || s->integrator.type != TRB_INTEGRATOR_PATH) return render_passes(s, rp, flags, mode, st);
    const bool stats = (flags & TRB_RENDER_STATS) != 0;
    CU(cudaMemsetAsync(rp.work_counter, 0, sizeof(uint32_t), st));
    if (s->ds.has_anim) {
        if (mode == 0) return stats ? launch_render_t<true, 0, true>(s, rp, flags, st) : launch_render_t<false, 0, true>(s, rp, flags, st);
        return stats ? launch_render_t<true, 1, true>(s, rp, flags, st) : launch_render_t<false, 1, true>(s, rp, flags, st);
    }
    if (mode == 0) return stats ? launch_render_t<true, 0, false>(s, rp, flags, st) : launch_render_t<false, 0, false>(s, rp, flags, st);
    return stats ? launch_render_t<true, 1, false>(s, rp, flags, st) : launch_render_t<false, 1, false>(s, rp, flags, st);
}

void stats_out(const trb::DStats& d, trb_stats* o) {
    o->camera_samples = d.camera_samples; o->rays_primary = d.rays_primary; o->rays_shadow = d.rays_shadow; o->rays_mis = d.rays_mis;
    o->rays_continuation = d.rays_continuation; o->node_tests = d.node_tests; o->tri_tests = d.tri_tests; o->inst_tests = d.inst_tests;
}

trb_status check_error_flag(trb_scene* s) {
    int e = 0;
    CU(cudaMemcpy(&e, s->d_error, sizeof e, cudaMemcpyDeviceToHost));
    if (e) { CU(cudaMemset(s->d_error, 0, sizeof(int))); return fail(TRB_CUDA, "BVH traversal stack overflow (depth > 64; the reference would panic)"); }
    return TRB_OK;
}

} // namespace

extern "C" {

const char* trb_last_error(void) { return g_error.c_str(); }
void trb_internal_set_error(const char* msg) { g_error = msg ? msg : ""; } // used by trb_loader.cpp
uint32_t trb_abi_version(void) { return TRB_ABI_VERSION; }
unsigned long long trb_launch_count(void) { return g_launches; }

trb_status trb_scene_set_option(trb_scene* s, const char* name, long long value) {
    if (!s || !name) return fail(TRB_INVALID_ARG, "null argument");
    Tuning& t = s->tune;
    const std::string k(name);
    if (k == "trace.refill") t.refill = (int)value;
    else if (k == "trace.occupancy" || k == "trace.smem_stack") {} // round-1 knobs: the variant (trace.pipe) now fixes both
    else if (k == "trace.grid") t.trace_grid = (unsigned)std::max<long long>(0, value);
    else if (k == "trace.sched") t.sched = (uint32_t)value;
    else if (k == "trace.quads" || k == "frame.device") { // both decide what update_frame builds (the DQuad TLAS records exist on the host path only): rebuild the current frame
        int& field = k == "trace.quads" ? t.quads : t.frame_device;
        const bool changed = (field != 0) != (value != 0);
        field = (int)value;
        if (changed && s->frame_set) return trb_scene_update_frame(s, s->last_frame, s->last_start, s->last_end);
    }
    else if (k == "trace.pipe") t.pipe = (int)value;
    else if (k == "trace.exact_box") t.exact_box = (int)value;
    else if (k == "film.v2") t.film_v2 = (int)value;
    else if (k == "sort.mode") t.sort = (int)value;
    else if (k == "sort.bits") t.sort_bits = (int)std::min<long long>(6, std::max<long long>(1, value));
    else if (k == "sort.min_round") t.sort_min_round = (int)value;
    else if (k == "shade.split") t.shade_split = (int)value;
    else if (k == "shade.sort") t.shade_sort = (int)value;
    else if (k == "shade.kind") t.shade_kind = (int)value;
    else if (k == "anim.table") t.anim_table = (int)value;
    else if (k == "shade.anim_occupancy") t.shade_anim_occ = (int)value;
    else if (k == "pass.paths") { if (value < 64) return fail(TRB_INVALID_ARG, "pass.paths must be >= 64"); t.pass_paths = (uint64_t)value; }
    else return fail(TRB_INVALID_ARG, "unknown option: " + k);
    return TRB_OK;
}

trb_status trb_scene_check_error(trb_scene* s) {
    if (!s) return fail(TRB_INVALID_ARG, "null scene");
    CU(cudaSetDevice(s->device));
    CU(cudaDeviceSynchronize());
    return check_error_flag(s);
}

trb_status trb_scene_trace_time(trb_scene* s, float* total_ms, uint32_t* n_launches) {
    if (!s || !total_ms || !n_launches) return fail(TRB_INVALID_ARG, "null argument");
    CU(cudaSetDevice(s->device));
    float total = 0.f;
    for (auto& e : s->trace_events) {
        CU(cudaEventSynchronize(e.second));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e.first, e.second));
        total += ms;
        s->event_pool.push_back(e);
    }
    *total_ms = total; *n_launches = (uint32_t)s->trace_events.size();
    s->trace_events.clear();
    return TRB_OK;
}

trb_status trb_scene_create(const trb_scene_desc* d, int device, trb_scene** out) {
    if (!out) return fail(TRB_INVALID_ARG, "null out pointer");
    *out = nullptr;
    trb_status v = validate(d);
    if (v != TRB_OK) return v;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TRB_NO_DEVICE, "no CUDA device: tray_rust_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(TRB_INVALID_ARG, "device ordinal out of range");
    CU(cudaSetDevice(device));
    std::unique_ptr<trb_scene> s(new trb_scene);
    s->device = device;
    tuning_from_env(s->tune);
    if (d->integrator.type == TRB_INTEGRATOR_WHITTED) { // the reference's recursion is kept as device recursion: one frame per ray depth
        size_t have = 0;
        CU(cudaDeviceGetLimit(&have, cudaLimitStackSize));
        const size_t need = 4096 + (size_t)2048 * (d->integrator.max_depth + 2);
        if (have < need) CU(cudaDeviceSetLimit(cudaLimitStackSize, need));
    }
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    s->sm_count = prop.multiProcessorCount;
    s->film = d->film; s->integrator = d->integrator;
    s->spp_pow2 = pow2_ceil(std::max(1u, d->film.samples));
    s->cameras.assign(d->cameras, d->cameras + d->n_cameras);
    s->instances.assign(d->instances, d->instances + d->n_instances);
    s->splines.assign(d->splines, d->splines + d->n_splines);
    s->keyframes.assign(d->keyframes, d->keyframes + d->n_keyframes);
    s->knots.assign(d->knots, d->knots + d->n_knots);
    for (const trb_spline& sp : s->splines) // BSpline::new sorts its knots (bspline 0.2.2); ranges were bounds-checked by validate()
        if (sp.n_ctrl > 1) std::stable_sort(s->knots.begin() + sp.knot_first, s->knots.begin() + sp.knot_first + sp.n_knots);
    s->color_keys.assign(d->color_keys, d->color_keys + d->n_color_keys);
    if (d->n_fov_floats) s->fov_floats.assign(d->fov_floats, d->fov_floats + d->n_fov_floats);
    for (const trb_camera& c : s->cameras)
        if (c.n_fov_ctrl) std::stable_sort(s->fov_floats.begin() + c.fov_knot_first, s->fov_floats.begin() + c.fov_knot_first + c.n_fov_knots);
    s->materials.assign(d->materials, d->materials + d->n_materials);

    // meshes: BVH<Triangle> with max_geom 16 (mesh.rs:44), then leaf-ordered triangle records
    std::vector<trb::DMesh> dmeshes(d->n_meshes);
    s->meshes.resize(d->n_meshes);
    for (uint32_t mi = 0; mi < d->n_meshes; ++mi) {
        const trb_mesh& m = d->meshes[mi];
        HostMesh& hm = s->meshes[mi];
        hm.pos.assign(m.positions, m.positions + 3 * (size_t)m.n_verts);
        hm.nrm.assign(m.normals, m.normals + 3 * (size_t)m.n_verts);
        hm.uv.assign(m.texcoords, m.texcoords + 2 * (size_t)m.n_verts);
        hm.idx.assign(m.indices, m.indices + 3 * (size_t)m.n_tris);
        std::vector<Box3> tb(m.n_tris);
        for (uint32_t t = 0; t < m.n_tris; ++t) { // Triangle::bounds (mesh.rs:128-134)
            Box3 b;
            const float* pa = &hm.pos[3 * hm.idx[3 * t]];
            for (int k = 0; k < 3; ++k) b.lo[k] = b.hi[k] = pa[k];
            box_grow_pt(b, &hm.pos[3 * hm.idx[3 * t + 1]]);
            box_grow_pt(b, &hm.pos[3 * hm.idx[3 * t + 2]]);
            tb[t] = b;
        }
        BvhBuilder bb;
        bb.build(tb, 16);
        hm.nodes = bb.nodes; hm.order = bb.order;
        for (int k = 0; k < 3; ++k) { hm.bounds.lo[k] = hm.nodes[0].bmin[k]; hm.bounds.hi[k] = hm.nodes[0].bmax[k]; }
        std::vector<trb::DPair> pn;
        trb::DBvh hdr{};
        if (!pack_pairs(hm.nodes, pn, hdr)) return fail(TRB_UNSUPPORTED, "mesh too large for the leaf encoding (2^25 triangles)");
        std::vector<trb::DTri> tris(m.n_tris);
        for (uint32_t slot = 0; slot < m.n_tris; ++slot) {
            const uint32_t t = hm.order[slot];
            const float* pa = &hm.pos[3 * hm.idx[3 * t]];
            const float* pb = &hm.pos[3 * hm.idx[3 * t + 1]];
            const float* pc = &hm.pos[3 * hm.idx[3 * t + 2]];
            float tid; std::memcpy(&tid, &t, 4);
            tris[slot].v0 = make_float4(pa[0], pa[1], pa[2], tid);
            tris[slot].e0 = make_float4(pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2], 0.f);
            tris[slot].e1 = make_float4(pc[0] - pa[0], pc[1] - pa[1], pc[2] - pa[2], 0.f);
            tris[slot].pad = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        std::vector<trb::DQuad> qn;
        uint32_t qroot = 0;
        if (!pack_quads(hm.nodes, qn, qroot)) return fail(TRB_UNSUPPORTED, "mesh too large for the leaf encoding (2^25 triangles)");
        hdr.root_hi.w = bits_f(qroot);
        trb::DQuad* dquads;
        CU(s->arena.upload(qn.data(), qn.size(), &dquads));
        hdr.quads = dquads;
        trb::DMesh& dm = dmeshes[mi];
        float *dp, *dn, *dt; uint32_t* di; trb::DPair* dnodes; trb::DTri* dtris;
        CU(s->arena.upload(hm.pos.data(), hm.pos.size(), &dp));
        CU(s->arena.upload(hm.nrm.data(), hm.nrm.size(), &dn));
        CU(s->arena.upload(hm.uv.data(), hm.uv.size(), &dt));
        CU(s->arena.upload(hm.idx.data(), hm.idx.size(), &di));
        CU(s->arena.upload(pn.data(), pn.size(), &dnodes));
        CU(s->arena.upload(tris.data(), tris.size(), &dtris));
        hdr.pairs = dnodes;
        dm.positions = dp; dm.normals = dn; dm.texcoords = dt; dm.indices = di; dm.bvh = hdr; dm.tris = dtris;
        dm.n_nodes = (uint32_t)hm.nodes.size(); dm.n_tris = m.n_tris;
    }
    trb::DMesh* d_meshes;
    CU(s->arena.upload(dmeshes.data(), dmeshes.size(), &d_meshes));

    // materials (precompute what Material::bsdf recomputes per hit from constant textures)
    std::vector<trb::DMaterial> dmats(d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const trb_material& m = d->materials[i];
        trb::DMaterial& o = dmats[i];
        std::memset(&o, 0, sizeof o);
        o.type = m.type;
        for (int k = 0; k < 3; ++k) { o.c0[k] = m.c0[k]; o.c1[k] = m.c1[k]; }
        o.roughness = m.roughness; o.eta = m.eta;
        o.width = fmaxf(m.roughness, 0.000001f);         // Beckmann::new (beckmann.rs:19-22)
        float sigma = kPi / 180.0f * m.roughness;        // OrenNayar::new (oren_nayar.rs:26-34), roughness in degrees
        sigma *= sigma;
        o.on_a = 1.0f - 0.5f * sigma / (sigma + 0.33f);
        o.on_b = 0.45f * sigma / (sigma + 0.09f);
        o.merl_off = m.type == TRB_MAT_MERL ? m.merl * TRB_MERL_TABLE_FLOATS : 0;
        for (int k = 0; k < 4; ++k) o.tex[k] = m.tex[k];
    }
    trb::DMaterial* d_mats;
    CU(s->arena.upload(dmats.data(), dmats.size(), &d_mats));
    float* d_merl = nullptr;
    CU(s->arena.alloc((size_t)d->n_merl * TRB_MERL_TABLE_FLOATS, &d_merl));
    for (uint32_t i = 0; i < d->n_merl; ++i)
        CU(cudaMemcpy(d_merl + (size_t)i * TRB_MERL_TABLE_FLOATS, d->merl_tables[i], sizeof(float) * TRB_MERL_TABLE_FLOATS, cudaMemcpyHostToDevice));

    { // image textures: every frame's RGBA8 texels in one array (texture/image.rs)
        std::vector<trb::DImage> dimg(d->n_images);
        std::vector<trb::DTexture> dtex(d->n_textures);
        std::vector<uchar4> tx;
        for (uint32_t i = 0; i < d->n_images; ++i) {
            const trb_image& im = d->images[i];
            dimg[i].width = im.width; dimg[i].height = im.height; dimg[i].offset = (uint32_t)tx.size(); dimg[i].time = im.time;
            const size_t n = (size_t)im.width * im.height;
            tx.resize(tx.size() + n);
            std::memcpy(tx.data() + dimg[i].offset, im.rgba8, n * 4);
        }
        for (uint32_t i = 0; i < d->n_textures; ++i) { dtex[i].first_image = d->textures[i].first_image; dtex[i].n_images = d->textures[i].n_images; }
        trb::DImage* d_img = nullptr; trb::DTexture* d_tex = nullptr; uchar4* d_tx = nullptr;
        CU(s->arena.upload(dimg.data(), dimg.size(), &d_img));
        CU(s->arena.upload(dtex.data(), dtex.size(), &d_tex));
        CU(s->arena.upload(tx.data(), tx.size(), &d_tx));
        s->ds.images = d_img; s->ds.textures = d_tex; s->ds.texels = d_tx; s->ds.n_textures = d->n_textures;
    }
    std::vector<uint32_t> lights;
    for (uint32_t i = 0; i < d->n_instances; ++i) if (d->instances[i].kind != TRB_INST_RECEIVER) lights.push_back(i); // multithreaded.rs:33-38
    uint32_t* d_lights;
    CU(s->arena.upload(lights.data(), lights.size(), &d_lights));

    filter_table(d->film, s->table);
    float* d_table;
    CU(s->arena.upload(s->table, 256, &d_table));

    CU(s->arena.alloc(d->n_instances, &s->d_instances));
    CU(s->arena.alloc(d->n_instances, &s->d_anim_instances));
    for (const trb_instance& in : s->instances) if (!trbh::xf_is_static(s->splines.data(), in.spline_first, in.n_splines)) s->n_anim++;
    {
        uint32_t kinds = 0;
        for (const trb_instance& in : s->instances) if (in.kind != TRB_INST_EMITTER_POINT && in.material < s->materials.size()) kinds |= 1u << s->materials[in.material].type;
        // two kinds are enough: with the material buckets the split kernels run one kind's code at a time, the fused kernel runs every
        // kind a warp holds (C3, matte + plastic: 857 -> 2157 Mrays/s; cornell_box.json: 1099 -> 2087; profiles/r02_c25_split_two_kinds.log)
        s->mixed_materials = __builtin_popcount(kinds) >= 2 || (kinds & (1u << TRB_MAT_MERL)) != 0;
        s->material_kinds = kinds;
    }
    CU(s->arena.alloc(1, &s->d_counter));
    CU(s->arena.alloc(1, &s->d_error));
    CU(cudaMemset(s->d_error, 0, sizeof(int)));
    CU(s->arena.alloc(1, &s->d_stats));
    CU(s->arena.alloc((size_t)d->film.width * d->film.height, &s->d_film));
    CU(cudaMallocHost(&s->h_film_staging, (size_t)d->film.width * d->film.height * 4 * sizeof(float)));
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));

    trb::DScene& ds = s->ds;
    ds.instances = s->d_instances; ds.meshes = d_meshes; ds.materials = d_mats; ds.merl = d_merl; ds.lights = d_lights;
    ds.n_instances = d->n_instances; ds.n_lights = (uint32_t)lights.size();
    ds.width = d->film.width; ds.height = d->film.height;
    ds.min_depth = d->integrator.min_depth; ds.max_depth = d->integrator.max_depth;
    ds.filter_w = d->film.filter_w; ds.filter_h = d->film.filter_h;
    ds.filter_inv_w = 1.0f / d->film.filter_w; ds.filter_inv_h = 1.0f / d->film.filter_h;
    ds.fpw_x = (int)floorf(d->film.filter_w / 0.5f); ds.fpw_y = (int)floorf(d->film.filter_h / 0.5f); // render_target.rs:48-49
    // the per-pixel test accepts |d| <= w / inv_w; the lock-block filter of render_target.rs:104-109 can only reject beyond fpw - 0.5
    ds.film_block_filter = (d->film.filter_w / ds.filter_inv_w <= (float)ds.fpw_x && d->film.filter_h / ds.filter_inv_h <= (float)ds.fpw_y) ? 0u : 1u;
    ds.filter_table = d_table;
    { // animation tables (evaluated per ray for keyframed instances / camera / emission)
        trb_spline* d_sp = nullptr; trb_keyframe* d_kf = nullptr; float* d_kn = nullptr; trb_color_key* d_ck = nullptr;
        if (!s->splines.empty()) CU(s->arena.upload(s->splines.data(), s->splines.size(), &d_sp));
        if (!s->keyframes.empty()) CU(s->arena.upload(s->keyframes.data(), s->keyframes.size(), &d_kf));
        if (!s->knots.empty()) CU(s->arena.upload(s->knots.data(), s->knots.size(), &d_kn));
        if (!s->color_keys.empty()) CU(s->arena.upload(s->color_keys.data(), s->color_keys.size(), &d_ck));
        std::vector<Xf> level(s->splines.size(), xf_identity());
        for (size_t k = 0; k < s->splines.size(); ++k)
            if (s->splines[k].n_ctrl == 1) level[k] = keyframe_xf(s->keyframes[s->splines[k].ctrl_first]);
        Xf* d_lv = nullptr;
        if (!level.empty()) CU(s->arena.upload(level.data(), level.size(), &d_lv));
        // distinct keyframed splines by content (degree, knots, control keyframes compared bit for bit)
        std::vector<uint32_t> uniq_of(s->splines.size(), 0xffffffffu), uniq_list;
        auto same = [&](const trb_spline& a, const trb_spline& b) {
            return a.degree == b.degree && a.n_ctrl == b.n_ctrl && a.n_knots == b.n_knots &&
                   memcmp(&s->keyframes[a.ctrl_first], &s->keyframes[b.ctrl_first], a.n_ctrl * sizeof(trb_keyframe)) == 0 &&
                   memcmp(&s->knots[a.knot_first], &s->knots[b.knot_first], a.n_knots * sizeof(float)) == 0;
        };
        for (size_t k = 0; k < s->splines.size(); ++k) {
            if (s->splines[k].n_ctrl <= 1) continue;
            for (size_t u = 0; u < uniq_list.size() && uniq_of[k] == 0xffffffffu; ++u)
                if (same(s->splines[k], s->splines[uniq_list[u]])) uniq_of[k] = (uint32_t)u;
            if (uniq_of[k] == 0xffffffffu) { uniq_of[k] = (uint32_t)uniq_list.size(); uniq_list.push_back((uint32_t)k); }
        }
        uint32_t* d_uo = nullptr; uint32_t* d_ul = nullptr;
        if (!uniq_of.empty()) CU(s->arena.upload(uniq_of.data(), uniq_of.size(), &d_uo));
        if (!uniq_list.empty()) CU(s->arena.upload(uniq_list.data(), uniq_list.size(), &d_ul));
        ds.spline_uniq = d_uo; ds.uniq_splines = d_ul; ds.n_uniq_splines = (uint32_t)uniq_list.size();
        ds.splines = d_sp; ds.keyframes = d_kf; ds.knots = d_kn; ds.color_keys = d_ck; ds.level_xf = d_lv; ds.has_anim = 0;
    }
    // Scene::load_file builds the BVH<Instance> for [0, scene_time] (scene.rs:141); the first render rebuilds it
    *out = s.release();
    return TRB_OK;
}

void trb_scene_destroy(trb_scene* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    delete s;
}

trb_status trb_scene_info(const trb_scene* s, uint32_t* w, uint32_t* h, uint32_t* spp, uint32_t* n_blocks, uint32_t* n_inst, uint32_t* n_lights) {
    if (!s) return fail(TRB_INVALID_ARG, "null scene");
    if (w) *w = s->film.width; if (h) *h = s->film.height; if (spp) *spp = s->spp_pow2;
    if (n_blocks) *n_blocks = (s->film.width / 8) * (s->film.height / 8);
    if (n_inst) *n_inst = (uint32_t)s->instances.size(); if (n_lights) *n_lights = s->ds.n_lights;
    return TRB_OK;
}

trb_status trb_scene_update_frame(trb_scene* s, uint32_t frame, float start, float end) {
    if (!s) return fail(TRB_INVALID_ARG, "null scene");
    CU(cudaSetDevice(s->device));
    // A frame boundary: passes enqueued with trb_render_device may still be reading the instances / TLAS this call
    // overwrites, so the device is drained first (the reference's update_frame likewise runs between renders, scene.rs:152).
    CU(cudaDeviceSynchronize());
    s->frame_set = true; s->last_frame = frame; s->last_start = start; s->last_end = end;
    // camera selection (scene.rs:153-166)
    int cam;
    if (s->active_camera >= 0) {
        cam = s->active_camera;
        if (cam != (int)s->cameras.size() - 1 && s->cameras[cam + 1].active_at == frame) cam += 1;
    } else {
        int c = 0;
        for (const trb_camera& x : s->cameras) { if (x.active_at <= frame) c++; else break; }
        if (c == 0) return fail(TRB_INVALID_ARG, "no camera is active at this frame");
        cam = c - 1;
    }
    s->active_camera = cam;
    const trb_camera& c = s->cameras[cam];
    s->shutter_open = start;                                  // camera.rs:127-129
    s->shutter_close = start + c.shutter_size * (end - start);
    Mat4 px_to_cam; float scaling[3];
    float fov = c.fov;
    if (c.n_fov_ctrl) { // sampled once per frame at the clamped mid-frame time (camera.rs:134-141)
        const float* kn = s->fov_floats.data() + c.fov_knot_first;
        const float lo = kn[c.fov_degree], hi = kn[c.n_fov_knots - 1 - c.fov_degree];
        float t = (start + end) / 2.0f;
        t = t < lo ? lo : (t > hi ? hi : t);
        fov = trbh::spline_point_f32(c.fov_degree, s->fov_floats.data() + c.fov_ctrl_first, kn, c.n_fov_knots, t);
    }
    camera_setup(fov, s->film.width, s->film.height, px_to_cam, scaling);
    // cam_world.transform(frame_time): a keyframed camera is evaluated per ray on the device (camera.rs:156)
    const bool cam_static = trbh::xf_is_static(s->splines.data(), c.spline_first, c.n_splines);
    const Xf cam_world = trbh::animated_xf(s->splines.data(), c.spline_first, c.n_splines, s->keyframes.data(), s->knots.data(), s->shutter_open);
    s->ds.cam.animated = cam_static ? 0u : 1u; s->ds.cam.spline_first = c.spline_first; s->ds.cam.n_splines = c.n_splines;
    bool any_anim = !cam_static;
    std::memcpy(s->ds.cam.px_to_cam, px_to_cam.m, 64);
    std::memcpy(s->ds.cam.cam_mat, cam_world.fwd.m, 64);
    std::memcpy(s->ds.cam.scaling, scaling, 12);
    s->ds.cam.shutter_open = s->shutter_open; s->ds.cam.shutter_close = s->shutter_close;

    // instance transforms + bounds, then BVH<Instance>::rebuild(shutter_open, shutter_close) (scene.rs:175, bvh.rs:61-78)
    const size_t n = s->instances.size();
    // static part of the device instance records (everything but the matrices); keyframed / animated flags are scene properties
    std::vector<trb::DInstance> di(n);
    std::vector<uint32_t> anim_list;
    for (size_t i = 0; i < n; ++i) {
        const trb_instance& in = s->instances[i];
        trb::DInstance& o = di[i];
        std::memset(&o, 0, sizeof o);
        o.kind = in.kind; o.shape = in.shape; o.p0 = in.p0; o.p1 = in.p1; o.mesh = in.mesh; o.material = in.material;
        o.xf_first = in.spline_first; o.xf_count = in.n_splines;
        if (!trbh::xf_is_static(s->splines.data(), in.spline_first, in.n_splines)) {
            o.flags |= trb::DI_ANIM_XF; o.spline_first = in.spline_first; o.n_splines = in.n_splines; any_anim = true;
            o.anim_slot = (uint32_t)anim_list.size(); anim_list.push_back((uint32_t)i);
        }
        if (in.kind != TRB_INST_RECEIVER) {
            for (int k = 0; k < 3; ++k) o.emission[k] = s->color_keys[in.emission_first].rgba[k];
            if (in.n_emission > 1) { o.flags |= trb::DI_ANIM_EMISSION; o.emission_first = in.emission_first; o.n_emission = in.n_emission; any_anim = true; }
        }
    }
    s->ds.has_anim = any_anim ? 1u : 0u;
    if (n + 1 > s->tlas_capacity) { // a tree over n instances has < n interior records and < 2n nodes
        CU(s->arena.alloc(n + 1, &s->d_tlas_quads));
        CU(s->arena.alloc(n + 1, &s->d_tlas));
        CU(s->arena.alloc(n, &s->d_tlas_order));
        CU(s->arena.alloc(2 * n + 2, &s->d_tlas_nodes));
        CU(s->arena.alloc(6 * n, &s->d_bounds));
        CU(s->arena.alloc(3 * n, &s->d_build_f));
        CU(s->arena.alloc(7 * n + 8, &s->d_build_u));
        CU(s->arena.alloc(4, &s->d_build_counts));
        s->tlas_capacity = n + 1;
    }
    if (!s->d_tlas_hdr) CU(s->arena.alloc(1, &s->d_tlas_hdr));
    if (!anim_list.empty()) CU(cudaMemcpy(s->d_anim_instances, anim_list.data(), anim_list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    s->ds.anim_instances = s->d_anim_instances; s->ds.n_anim_instances = (uint32_t)anim_list.size();
    s->ds.tlas = s->d_tlas_hdr; s->ds.tlas_pairs = s->d_tlas; s->ds.tlas_quads = s->d_tlas_quads; s->ds.tlas_order = s->d_tlas_order;

    if (s->tune.frame_device && !s->tune.quads) {
        // ---- device path: transforms, animation bounds, SAH build and record packing all on the GPU (k_frame_instances, k_tlas_build)
        if (!s->instances_static_uploaded) { CU(cudaMemcpy(s->d_instances, di.data(), n * sizeof(trb::DInstance), cudaMemcpyHostToDevice)); s->instances_static_uploaded = true; }
        trb::FrameBuild fb{};
        fb.instances = s->d_instances; fb.bounds = reinterpret_cast<trbh::Box3*>(s->d_bounds); fb.n = (uint32_t)n;
        fb.shutter_open = s->shutter_open; fb.shutter_close = s->shutter_close;
        fb.cx = s->d_build_f; fb.cy = s->d_build_f + n; fb.cz = s->d_build_f + 2 * n;
        fb.idx = s->d_build_u; fb.task = s->d_build_u + n; fb.rec_of = s->d_build_u + 4 * n + 4;
        fb.nodes = s->d_tlas_nodes; fb.order = s->d_tlas_order; fb.counts = s->d_build_counts; fb.pairs = s->d_tlas; fb.hdr = s->d_tlas_hdr;
        trb::k_frame_instances<<<(unsigned)((n + 63) / 64), 64>>>(s->ds, fb);
        trb::k_tlas_build<<<1, 1>>>(fb);
        g_launches += 2;
        CU(cudaGetLastError());
        uint32_t counts[3] = {0, 0, 0};
        CU(cudaMemcpy(counts, s->d_build_counts, sizeof counts, cudaMemcpyDeviceToHost)); // also the frame's only synchronisation point
        if (!counts[2]) return fail(TRB_UNSUPPORTED, "too many instances for the leaf encoding");
        s->tlas_n_nodes = counts[0];
        s->host_frame_stale = true; // world transforms / TLAS nodes are fetched from the device when a getter asks for them
        s->frame_ready = true;
        return TRB_OK;
    }

    // ---- host path (kept for the DQuad experiment and as an independent check of the device path)
    s->world.resize(n);
    std::vector<Box3> bounds(n);
    for (size_t i = 0; i < n; ++i) {
        const trb_instance& in = s->instances[i];
        // world[i] = transform(shutter_open): exact for static instances; keyframed ones are re-evaluated per ray / per path on the device
        s->world[i] = trbh::animated_xf(s->splines.data(), in.spline_first, in.n_splines, s->keyframes.data(), s->knots.data(), s->shutter_open);
        const Box3 local = shape_bounds(*s, in);
        if (!trbh::xf_is_animated(s->splines.data(), in.spline_first, in.n_splines)) { // animation_bounds (animated_transform.rs:57-70, Q22)
            bounds[i] = arvo_bounds(s->world[i].fwd, local);
        } else {
            Box3 acc = box_empty();
            for (int k = 0; k < 128; ++k) {
                const float u = (float)k / 127.0f;
                const float time = s->shutter_open * (1.0f - u) + s->shutter_close * u; // linalg::lerp
                const Xf x = trbh::animated_xf(s->splines.data(), in.spline_first, in.n_splines, s->keyframes.data(), s->knots.data(), time);
                box_grow(acc, arvo_bounds(x.fwd, local));
            }
            bounds[i] = acc;
        }
        std::memcpy(di[i].inv, s->world[i].inv.m, 64);
        std::memcpy(di[i].mat, s->world[i].fwd.m, 64);
    }
    BvhBuilder bb;
    bb.build(bounds, 4);
    s->tlas_nodes = bb.nodes; s->tlas_order = bb.order;
    s->host_frame_stale = false; s->instances_static_uploaded = false;
    std::vector<trb::DPair> pn;
    trb::DBvh hdr{};
    if (!pack_pairs(s->tlas_nodes, pn, hdr)) return fail(TRB_UNSUPPORTED, "too many instances for the leaf encoding");
    std::vector<trb::DQuad> qn;
    uint32_t qroot = 0;
    if (!pack_quads(s->tlas_nodes, qn, qroot)) return fail(TRB_UNSUPPORTED, "too many instances for the leaf encoding");
    hdr.root_hi.w = bits_f(qroot);
    if (!pn.empty()) CU(cudaMemcpy(s->d_tlas, pn.data(), pn.size() * sizeof(trb::DPair), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_tlas_order, s->tlas_order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_instances, di.data(), n * sizeof(trb::DInstance), cudaMemcpyHostToDevice));
    if (!qn.empty()) CU(cudaMemcpy(s->d_tlas_quads, qn.data(), qn.size() * sizeof(trb::DQuad), cudaMemcpyHostToDevice));
    hdr.pairs = s->d_tlas; hdr.quads = s->d_tlas_quads;
    CU(cudaMemcpy(s->d_tlas_hdr, &hdr, sizeof hdr, cudaMemcpyHostToDevice));
    s->frame_ready = true;
    return TRB_OK;
}

trb_status trb_render_device(trb_scene* s, const trb_render_cfg* cfg, float* d_film, trb_stats* d_stats, void* stream) {
    if (!s || !cfg || !d_film) return fail(TRB_INVALID_ARG, "null argument");
    if (!s->frame_ready) return fail(TRB_INVALID_ARG, "Update frame must be called before rendering"); // scene.rs:179
    CU(cudaSetDevice(s->device));
    uint32_t spp, first, count, nb;
    const uint2* d_blocks = nullptr;
    trb_status r = resolve_samples(s, cfg, spp, first, count);
    if (r != TRB_OK) return r;
    r = ensure_blocks(s, cfg, &d_blocks, &nb);
    if (r != TRB_OK) return r;
    if (nb == 0 || count == 0) return TRB_OK; // "Warning: This block queue is empty!" (block_queue.rs:42-44)
    trb::RenderParams rp{};
    rp.blocks = d_blocks; rp.n_blocks = nb; rp.spp = spp; rp.sample_first = first; rp.sample_count = count; rp.seed = cfg->seed;
    rp.work_counter = s->d_counter; rp.film = reinterpret_cast<float4*>(d_film); rp.stats = reinterpret_cast<trb::DStats*>(d_stats);
    rp.error_flag = s->d_error;
    return launch_render(s, rp, cfg->flags, 0, static_cast<cudaStream_t>(stream));
}

trb_status trb_render(trb_scene* s, const trb_render_cfg* cfg, float* film, trb_stats* stats) {
    if (!s || !cfg || !film) return fail(TRB_INVALID_ARG, "null argument");
    CU(cudaSetDevice(s->device));
    float update_ms = 0.f;
    if (!(cfg->flags & TRB_RENDER_NO_UPDATE)) { // Exec::render: scene.update_frame first (multithreaded.rs:57-60)
        auto t0 = std::chrono::steady_clock::now();
        const float step = s->film.scene_time / (float)s->film.frames;
        trb_status r = trb_scene_update_frame(s, cfg->current_frame, (float)cfg->current_frame * step, ((float)cfg->current_frame + 1.0f) * step);
        if (r != TRB_OK) return r;
        update_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    const size_t npx = (size_t)s->film.width * s->film.height;
    CU(cudaMemsetAsync(s->d_film, 0, npx * sizeof(float4), 0));
    CU(cudaMemsetAsync(s->d_stats, 0, sizeof(trb::DStats), 0));
    CU(cudaEventRecord(s->ev0, 0));
    trb_status r = trb_render_device(s, cfg, reinterpret_cast<float*>(s->d_film), reinterpret_cast<trb_stats*>(s->d_stats), nullptr);
    if (r != TRB_OK) return r;
    CU(cudaEventRecord(s->ev1, 0));
    CU(cudaMemcpyAsync(s->h_film_staging, s->d_film, npx * sizeof(float4), cudaMemcpyDeviceToHost, 0));
    CU(cudaStreamSynchronize(0));
    r = check_error_flag(s);
    if (r != TRB_OK) return r;
    { // additive, like film::Image::add_pixels (image.rs:21-33); a 33 MB read-modify-write: split over a few host threads
        const size_t n = npx * 4;
        const unsigned nt = n >= (1u << 20) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
        const float* src = s->h_film_staging;
        auto add = [film, src](size_t a, size_t b) { for (size_t i = a; i < b; ++i) film[i] += src[i]; };
        std::vector<std::thread> th;
        for (unsigned k = 1; k < nt; ++k) th.emplace_back(add, n * k / nt, n * (k + 1) / nt);
        add(0, n / nt);
        for (auto& t : th) t.join();
    }
    if (stats) {
        trb::DStats h;
        CU(cudaMemcpy(&h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost));
        std::memset(stats, 0, sizeof *stats);
        stats_out(h, stats);
        CU(cudaEventElapsedTime(&stats->kernel_ms, s->ev0, s->ev1));
        stats->update_ms = update_ms;
    }
    return TRB_OK;
}

trb_status trb_render_samples(trb_scene* s, const trb_render_cfg* cfg, size_t n, trb_sample* samples, trb_stats* stats) {
    if (!s || !cfg || !samples) return fail(TRB_INVALID_ARG, "null argument");
    if (!s->frame_ready) return fail(TRB_INVALID_ARG, "Update frame must be called before rendering");
    CU(cudaSetDevice(s->device));
    uint32_t spp, first, count, nb;
    const uint2* d_blocks = nullptr;
    trb_status r = resolve_samples(s, cfg, spp, first, count);
    if (r != TRB_OK) return r;
    r = ensure_blocks(s, cfg, &d_blocks, &nb);
    if (r != TRB_OK) return r;
    if (n != (size_t)nb * 64 * count) return fail(TRB_INVALID_ARG, "sample buffer size must be blocks*64*sample_count");
    if (n == 0) return TRB_OK;
    trb_sample* d_out = nullptr;
    CU(cudaMalloc(&d_out, n * sizeof(trb_sample)));
    CU(cudaMemsetAsync(s->d_stats, 0, sizeof(trb::DStats), 0));
    trb::RenderParams rp{};
    rp.blocks = d_blocks; rp.n_blocks = nb; rp.spp = spp; rp.sample_first = first; rp.sample_count = count; rp.seed = cfg->seed;
    rp.work_counter = s->d_counter; rp.film = nullptr; rp.samples_out = d_out; rp.stats = s->d_stats; rp.error_flag = s->d_error;
    CU(cudaEventRecord(s->ev0, 0));
    r = launch_render(s, rp, cfg->flags, 1, 0);
    if (r != TRB_OK) { cudaFree(d_out); return r; }
    CU(cudaEventRecord(s->ev1, 0));
    cudaError_t e = cudaMemcpy(samples, d_out, n * sizeof(trb_sample), cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    CU(e);
    r = check_error_flag(s);
    if (r != TRB_OK) return r;
    if (stats) {
        trb::DStats h;
        CU(cudaMemcpy(&h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost));
        std::memset(stats, 0, sizeof *stats);
        stats_out(h, stats);
        CU(cudaEventElapsedTime(&stats->kernel_ms, s->ev0, s->ev1));
    }
    return TRB_OK;
}

trb_status trb_camera_rays(trb_scene* s, const trb_render_cfg* cfg, size_t n, trb_ray* rays, float* xy) {
    if (!s || !cfg || !rays || !xy) return fail(TRB_INVALID_ARG, "null argument");
    if (!s->frame_ready) return fail(TRB_INVALID_ARG, "Update frame must be called before rendering");
    CU(cudaSetDevice(s->device));
    uint32_t spp, first, count, nb;
    const uint2* d_blocks = nullptr;
    trb_status r = resolve_samples(s, cfg, spp, first, count);
    if (r != TRB_OK) return r;
    r = ensure_blocks(s, cfg, &d_blocks, &nb);
    if (r != TRB_OK) return r;
    if (n != (size_t)nb * 64 * count) return fail(TRB_INVALID_ARG, "ray buffer size must be blocks*64*sample_count");
    if (n == 0) return TRB_OK;
    trb_ray* d_rays = nullptr; float* d_xy = nullptr;
    CU(cudaMalloc(&d_rays, n * sizeof(trb_ray)));
    cudaError_t e = cudaMalloc(&d_xy, n * 2 * sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_rays); CU(e); }
    trb::RenderParams rp{};
    rp.blocks = d_blocks; rp.n_blocks = nb; rp.spp = spp; rp.sample_first = first; rp.sample_count = count; rp.seed = cfg->seed;
    if (s->ds.has_anim) trb::k_camera_rays<true><<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256>>>(s->ds, rp, d_rays, d_xy);
    else trb::k_camera_rays<false><<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256>>>(s->ds, rp, d_rays, d_xy);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(rays, d_rays, n * sizeof(trb_ray), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(xy, d_xy, n * 2 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_rays); cudaFree(d_xy);
    CU(e);
    return TRB_OK;
}

trb_status trb_intersect_device(trb_scene* s, size_t n, const trb_ray* d_rays, trb_hit* d_hits, trb_stats* d_stats, void* stream) {
    if (!s || (n && (!d_rays || !d_hits))) return fail(TRB_INVALID_ARG, "null argument");
    if (!s->frame_ready) return fail(TRB_INVALID_ARG, "Update frame must be called before intersecting");
    if (n == 0) return TRB_OK;
    CU(cudaSetDevice(s->device));
    const unsigned grid = (unsigned)std::min<size_t>((n + 127) / 128, (size_t)s->sm_count * 16);
    g_launches++;
    if (s->ds.has_anim) trb::k_intersect<false, true><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(s->ds, n, d_rays, d_hits, reinterpret_cast<trb::DStats*>(d_stats), s->d_error);
    else trb::k_intersect<false, false><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(s->ds, n, d_rays, d_hits, reinterpret_cast<trb::DStats*>(d_stats), s->d_error);
    CU(cudaGetLastError());
    return TRB_OK;
}

trb_status trb_intersect(trb_scene* s, size_t n, const trb_ray* rays, trb_hit* hits, trb_stats* stats) {
    if (!s || (n && (!rays || !hits))) return fail(TRB_INVALID_ARG, "null argument");
    if (!s->frame_ready) return fail(TRB_INVALID_ARG, "Update frame must be called before intersecting");
    if (n == 0) return TRB_OK;
    CU(cudaSetDevice(s->device));
    trb_ray* d_rays = nullptr; trb_hit* d_hits = nullptr;
    CU(cudaMalloc(&d_rays, n * sizeof(trb_ray)));
    cudaError_t e = cudaMalloc(&d_hits, n * sizeof(trb_hit));
    if (e != cudaSuccess) { cudaFree(d_rays); CU(e); }
    e = cudaMemcpy(d_rays, rays, n * sizeof(trb_ray), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_stats, 0, sizeof(trb::DStats), 0);
    if (e == cudaSuccess) {
        cudaEventRecord(s->ev0, 0);
        const unsigned grid = (unsigned)std::min<size_t>((n + 127) / 128, (size_t)s->sm_count * 16);
        if (s->ds.has_anim) trb::k_intersect<true, true><<<grid, 128>>>(s->ds, n, d_rays, d_hits, s->d_stats, s->d_error);
        else trb::k_intersect<true, false><<<grid, 128>>>(s->ds, n, d_rays, d_hits, s->d_stats, s->d_error); // host variant always counts tests
        cudaEventRecord(s->ev1, 0);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(hits, d_hits, n * sizeof(trb_hit), cudaMemcpyDeviceToHost);
    cudaFree(d_rays); cudaFree(d_hits);
    CU(e);
    trb_status r = check_error_flag(s);
    if (r != TRB_OK) return r;
    if (stats) {
        trb::DStats h;
        CU(cudaMemcpy(&h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost));
        std::memset(stats, 0, sizeof *stats);
        stats_out(h, stats);
        CU(cudaEventElapsedTime(&stats->kernel_ms, s->ev0, s->ev1));
    }
    return TRB_OK;
}

trb_status trb_film_to_srgb8(trb_scene* s, const float* film, uint8_t* rgb8) {
    if (!s || !film || !rgb8) return fail(TRB_INVALID_ARG, "null argument");
    CU(cudaSetDevice(s->device));
    const size_t npx = (size_t)s->film.width * s->film.height;
    uint8_t* d_out = nullptr;
    CU(cudaMalloc(&d_out, npx * 3));
    cudaError_t e = cudaMemcpy(s->d_film, film, npx * sizeof(float4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        trb::k_srgb8<<<(unsigned)std::min<size_t>((npx + 255) / 256, 148 * 16), 256>>>(npx, s->d_film, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(rgb8, d_out, npx * 3, cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    CU(e);
    return TRB_OK;
}

trb_status trb_block_list(const trb_scene* s, uint32_t start, uint32_t count, uint32_t* n_out, uint32_t* xy, uint32_t cap) {
    if (!s || !n_out) return fail(TRB_INVALID_ARG, "null argument");
    std::vector<uint32_t> b = morton_blocks(s->film.width, s->film.height, start, count);
    *n_out = (uint32_t)(b.size() / 2);
    if (xy) std::memcpy(xy, b.data(), sizeof(uint32_t) * std::min<size_t>(b.size(), 2 * (size_t)cap));
    return TRB_OK;
}

trb_status trb_scene_get_bvh(const trb_scene* s, int which, uint32_t* n_nodes, trb_bvh_node* nodes, uint32_t* n_ordered, uint32_t* ordered) {
    if (!s || !n_nodes || !n_ordered) return fail(TRB_INVALID_ARG, "null argument");
    const std::vector<trb_bvh_node>* nn; const std::vector<uint32_t>* oo;
    std::vector<trb_bvh_node> dev_nodes; std::vector<uint32_t> dev_order;
    if (which < 0) {
        if (!s->frame_ready) return fail(TRB_INVALID_ARG, "update_frame first");
        if (s->host_frame_stale) { // the frame was prepared on the device: read its TLAS back
            CU(cudaSetDevice(s->device));
            dev_nodes.resize(s->tlas_n_nodes); dev_order.resize(s->instances.size());
            CU(cudaMemcpy(dev_nodes.data(), s->d_tlas_nodes, dev_nodes.size() * sizeof(trb_bvh_node), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(dev_order.data(), s->d_tlas_order, dev_order.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            nn = &dev_nodes; oo = &dev_order;
        } else { nn = &s->tlas_nodes; oo = &s->tlas_order; }
    }
    else { if ((size_t)which >= s->meshes.size()) return fail(TRB_INVALID_ARG, "mesh index out of range"); nn = &s->meshes[which].nodes; oo = &s->meshes[which].order; }
    *n_nodes = (uint32_t)nn->size(); *n_ordered = (uint32_t)oo->size();
    if (nodes) std::memcpy(nodes, nn->data(), nn->size() * sizeof(trb_bvh_node));
    if (ordered) std::memcpy(ordered, oo->data(), oo->size() * sizeof(uint32_t));
    return TRB_OK;
}

trb_status trb_scene_get_transform(const trb_scene* s, uint32_t inst, float* mat16, float* inv16) {
    if (!s || !s->frame_ready || inst >= s->instances.size()) return fail(TRB_INVALID_ARG, "bad instance / update_frame first");
    if (s->host_frame_stale) { // prepared on the device
        trb::DInstance di;
        CU(cudaSetDevice(s->device));
        CU(cudaMemcpy(&di, s->d_instances + inst, sizeof di, cudaMemcpyDeviceToHost));
        std::memcpy(mat16, di.mat, 64); std::memcpy(inv16, di.inv, 64);
        return TRB_OK;
    }
    std::memcpy(mat16, s->world[inst].fwd.m, 64); std::memcpy(inv16, s->world[inst].inv.m, 64);
    return TRB_OK;
}

trb_status trb_scene_get_filter_table(const trb_scene* s, float* t) {
    if (!s || !t) return fail(TRB_INVALID_ARG, "null argument");
    std::memcpy(t, s->table, sizeof s->table);
    return TRB_OK;
}

trb_status trb_host_build_bvh(const float* boxes6, uint32_t n, uint32_t max_geom, uint32_t* n_nodes, trb_bvh_node* nodes, uint32_t* ordered) {
    if (!boxes6 || n == 0 || !n_nodes) return fail(TRB_INVALID_ARG, "empty geometry"); // bvh.rs:35 assert!(!geometry.is_empty())
    std::vector<Box3> b(n);
    for (uint32_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) { b[i].lo[k] = boxes6[6 * i + k]; b[i].hi[k] = boxes6[6 * i + 3 + k]; }
    BvhBuilder bb;
    bb.build(b, max_geom);
    *n_nodes = (uint32_t)bb.nodes.size();
    if (nodes) std::memcpy(nodes, bb.nodes.data(), bb.nodes.size() * sizeof(trb_bvh_node));
    if (ordered) std::memcpy(ordered, bb.order.data(), bb.order.size() * sizeof(uint32_t));
    return TRB_OK;
}

trb_status trb_host_keyframe_transform(const trb_keyframe* kf, float* mat16, float* inv16) {
    if (!kf || !mat16 || !inv16) return fail(TRB_INVALID_ARG, "null argument");
    const Xf x = keyframe_xf(*kf);
    std::memcpy(mat16, x.fwd.m, 64); std::memcpy(inv16, x.inv.m, 64);
    return TRB_OK;
}

// Layout check without a GPU: every ray is traversed (a) literally like bvh.rs:81-130 over the reference-order nodes and
// (b) through the DQuad records with quad_visit — the function the trace kernel runs — and the two must visit the same
// leaves in the same order with the same max_t history. Leaves "hit" pseudo-randomly (a hash of leaf and ray decides a
// distance) so that max_t shrinks during the walk. Returns the number of rays whose walks differ.
trb_status trb_selftest_box(uint32_t n_cases, uint32_t seed, uint64_t out[4]) {
    if (!out || n_cases == 0) return fail(TRB_INVALID_ARG, "null argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return fail(TRB_NO_DEVICE, "no CUDA device"); }
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, 4 * sizeof(unsigned long long)));
    cudaMemset(d, 0, 4 * sizeof(unsigned long long));
    trb::k_selftest_box<<<(n_cases + 255) / 256, 256>>>(n_cases, seed, d);
    g_launches++;
    unsigned long long h[4] = {0, 0, 0, 0};
    const cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(TRB_CUDA, cudaGetErrorString(e));
    for (int k = 0; k < 4; ++k) out[k] = h[k];
    return TRB_OK;
}

trb_status trb_host_quad_check(const trb_bvh_node* nodes, uint32_t n_nodes, const trb_ray* rays, uint32_t n_rays, uint32_t* mismatches,
                               uint64_t* leaf_visits, uint64_t* quad_visits) {
    if (!nodes || !rays || !mismatches || n_nodes == 0) return fail(TRB_INVALID_ARG, "null argument");
    const std::vector<trb_bvh_node> in(nodes, nodes + n_nodes);
    std::vector<trb::DQuad> quads;
    uint32_t qroot = 0;
    if (!pack_quads(in, quads, qroot)) return fail(TRB_UNSUPPORTED, "leaf encoding");
    uint32_t bad = 0; uint64_t nl = 0, nq = 0;
    auto leaf_hit = [](uint32_t first, uint32_t ray, float tmin, float& tmax) {
        uint32_t h = (first * 0x9E3779B1u) ^ (ray * 0x85EBCA77u); h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
        if ((h & 3u) == 0) { const float tc = tmin + (float)(h >> 8) * (1.0f / 16777216.0f) * 40.0f; if (tc >= tmin && tc <= tmax) tmax = tc; }
    };
    for (uint32_t ri = 0; ri < n_rays; ++ri) {
        const trb_ray& r = rays[ri];
        const trb::f3 o = trb::mk(r.o[0], r.o[1], r.o[2]);
        const trb::f3 inv = trb::mk(1.0f / r.d[0], 1.0f / r.d[1], 1.0f / r.d[2]);
        const bool nx = r.d[0] < 0.0f, ny = r.d[1] < 0.0f, nz = r.d[2] < 0.0f;
        const bool neg[3] = {nx, ny, nz};
        if (!(std::isfinite(inv.x) && std::isfinite(inv.y) && std::isfinite(inv.z))) continue; // these rays take the DPair path
        // (a) the reference loop
        std::vector<uint32_t> seq_a, seq_b;
        float tmax_a = r.max_t, tmax_b = r.max_t;
        {
            uint32_t stack[64]; int sp = 0; uint32_t cur = 0;
            for (;;) {
                const trb_bvh_node& n = in[cur];
                float te;
                const float4 lo = make_float4(n.bmin[0], n.bmin[1], n.bmin[2], 0.f), hi = make_float4(n.bmax[0], n.bmax[1], n.bmax[2], 0.f);
                if (trb::box_hit(lo, hi, o, inv, nx, ny, nz, r.min_t, tmax_a, te)) {
                    if (n.b & TRB_BVH_LEAF) {
                        seq_a.push_back(n.a); leaf_hit(n.a, ri, r.min_t, tmax_a);
                        if (sp == 0) break;
                        cur = stack[--sp];
                    } else {
                        if (neg[n.b & 3u]) { stack[sp++] = cur + 1; cur = n.a; } else { stack[sp++] = n.a; cur += 1; }
                    }
                } else { if (sp == 0) break; cur = stack[--sp]; }
            }
        }
        // (b) root box, then DQuad records with a stack of (entry distance, reference)
        {
            const trb_bvh_node& n = in[0];
            float te;
            const float4 lo = make_float4(n.bmin[0], n.bmin[1], n.bmin[2], 0.f), hi = make_float4(n.bmax[0], n.bmax[1], n.bmax[2], 0.f);
            std::vector<unsigned long long> stack;
            uint32_t cur = trb::box_hit(lo, hi, o, inv, nx, ny, nz, r.min_t, tmax_b, te) ? qroot : 0xffffffffu;
            for (;;) {
                if (cur == 0xffffffffu) { // pop
                    bool got = false;
                    while (!stack.empty()) {
                        const unsigned long long e = stack.back(); stack.pop_back();
                        float tent; const uint32_t tb = (uint32_t)(e >> 32); std::memcpy(&tent, &tb, 4);
                        if (tent < tmax_b) { cur = (uint32_t)e; got = true; break; }
                    }
                    if (!got) break;
                }
                if ((cur & trb::REF_TAG) == trb::REF_LEAF) {
                    const uint32_t first = cur & 0x01ffffffu;
                    seq_b.push_back(first); leaf_hit(first, ri, r.min_t, tmax_b);
                    cur = 0xffffffffu;
                } else {
                    const float4* q = quads[cur].q;
                    trb::QuadOut qo;
                    trb::quad_visit(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], o, inv, nx, ny, nz, r.min_t, tmax_b, qo);
                    ++nq;
                    if (qo.p0) stack.push_back(qo.e0);
                    if (qo.p1) stack.push_back(qo.e1);
                    if (qo.p2) stack.push_back(qo.e2);
                    cur = qo.next;
                }
            }
        }
        nl += seq_a.size();
        if (seq_a != seq_b || std::memcmp(&tmax_a, &tmax_b, 4) != 0) ++bad;
    }
    *mismatches = bad;
    if (leaf_visits) *leaf_visits = nl;
    if (quad_visits) *quad_visits = nq;
    return TRB_OK;
}

trb_status trb_host_animated_transform(const trb_scene_desc* d, uint32_t first, uint32_t count, float time, float* mat16, float* inv16) {
    if (!d || !mat16 || !inv16) return fail(TRB_INVALID_ARG, "null argument");
    const trb_status r = validate(d);
    if (r != TRB_OK) return r;
    if ((uint64_t)first + count > d->n_splines) return fail(TRB_INVALID_ARG, "spline range out of bounds");
    std::vector<float> knots(d->knots, d->knots + d->n_knots); // BSpline::new sorts its knots
    for (uint32_t k = first; k < first + count; ++k)
        if (d->splines[k].n_ctrl > 1) std::stable_sort(knots.begin() + d->splines[k].knot_first, knots.begin() + d->splines[k].knot_first + d->splines[k].n_knots);
    const Xf x = trbh::animated_xf(d->splines, first, count, d->keyframes, knots.data(), time);
    std::memcpy(mat16, x.fwd.m, 64); std::memcpy(inv16, x.inv.m, 64);
    return TRB_OK;
}

trb_status trb_host_animated_color(const trb_scene_desc* d, uint32_t first, uint32_t count, float time, float* rgb3) {
    if (!d || !rgb3) return fail(TRB_INVALID_ARG, "null argument");
    if (count == 0 || (uint64_t)first + count > d->n_color_keys) return fail(TRB_INVALID_ARG, "colour key range out of bounds");
    trbh::animated_color(d->color_keys, first, count, time, rgb3);
    return TRB_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU: tile sharding + one film SUM-reduce per frame with NCCL called directly (no torch).
// libnccl is resolved at run time so that the single-GPU library has no link-time dependency on it.
// ---------------------------------------------------------------------------------------------------------------
} // extern "C"

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
trb_status nccl_load() {
    if (g_nccl.lib) return TRB_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // the copy the process already uses (e.g. PyTorch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(TRB_NCCL, std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"));
    NcclApi a; a.lib = h;
#define TRB_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name)); if (!a.field) return fail(TRB_NCCL, "libnccl lacks " name)
    TRB_SYM(GetUniqueId, "ncclGetUniqueId"); TRB_SYM(CommInitRank, "ncclCommInitRank"); TRB_SYM(CommInitAll, "ncclCommInitAll");
    TRB_SYM(CommDestroy, "ncclCommDestroy"); TRB_SYM(Reduce, "ncclReduce"); TRB_SYM(GroupStart, "ncclGroupStart"); TRB_SYM(GroupEnd, "ncclGroupEnd");
    TRB_SYM(GetErrorString, "ncclGetErrorString");
#undef TRB_SYM
    g_nccl = a;
    return TRB_OK;
}
#define NC(call)                                                                                                   \
    do {                                                                                                           \
        ncclResult_t r_ = (call);                                                                                  \
        if (r_ != ncclSuccess) return fail(TRB_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_));       \
    } while (0)
} // namespace

struct trb_comm { ncclComm_t comm = nullptr; int n_ranks = 1, rank = 0, device = 0; };
struct trb_group { std::vector<trb_scene*> scenes; std::vector<ncclComm_t> comms; std::vector<int> devices; };

namespace {
// this rank's shard of the selected block list inside cfg (interleaved chunks unless the caller asked for the reference's contiguous ranges)
trb_status shard_cfg(const trb_scene* s, const trb_render_cfg* in, int rank, int n_ranks, trb_render_cfg* out, bool* empty) {
    *out = *in; *empty = false;
    if (n_ranks <= 1) { out->shard_index = out->shard_count = out->shard_chunk = 0; return TRB_OK; }
    if (in->shard_count == 0xffffffffu) { // exec/distrib/master.rs:91-93,218-224: floor(B / W) blocks each, the remainder to the last worker
        const uint32_t all = (s->film.width / 8) * (s->film.height / 8);
        const uint32_t sel0 = in->block_count ? std::min(in->block_start, all) : 0u;
        const uint32_t sel = in->block_count ? (uint32_t)std::min<uint64_t>(all - sel0, in->block_count) : all;
        const uint32_t per = sel / (uint32_t)n_ranks, start = sel0 + (uint32_t)rank * per;
        const uint32_t count = rank == n_ranks - 1 ? sel0 + sel - start : per;
        out->block_start = start; out->block_count = count; out->shard_index = out->shard_count = out->shard_chunk = 0;
        *empty = count == 0; // block_count 0 would mean "all blocks" (block_queue.rs:39-41): an idle rank must not render
        return TRB_OK;
    }
    out->shard_index = (uint32_t)rank; out->shard_count = (uint32_t)n_ranks; out->shard_chunk = in->shard_chunk ? in->shard_chunk : 32u;
    return TRB_OK;
}
// Exec::render on one replica with the film left on the device: update_frame, clear, all passes of this shard.
trb_status render_to_device_film(trb_scene* s, const trb_render_cfg* cfg, bool empty, cudaStream_t st) {
    CU(cudaSetDevice(s->device));
    if (!(cfg->flags & TRB_RENDER_NO_UPDATE)) {
        const float step = s->film.scene_time / (float)s->film.frames;
        trb_status r = trb_scene_update_frame(s, cfg->current_frame, (float)cfg->current_frame * step, ((float)cfg->current_frame + 1.0f) * step);
        if (r != TRB_OK) return r;
    }
    const size_t npx = (size_t)s->film.width * s->film.height;
    CU(cudaMemsetAsync(s->d_film, 0, npx * sizeof(float4), st));
    CU(cudaMemsetAsync(s->d_stats, 0, sizeof(trb::DStats), st));
    if (empty) return TRB_OK;
    return trb_render_device(s, cfg, reinterpret_cast<float*>(s->d_film), reinterpret_cast<trb_stats*>(s->d_stats), st);
}
trb_status film_to_host_add(trb_scene* s, float* film, cudaStream_t st) { // additive, like film::Image::add_pixels (image.rs:21-33)
    const size_t npx = (size_t)s->film.width * s->film.height;
    CU(cudaMemcpyAsync(s->h_film_staging, s->d_film, npx * sizeof(float4), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const float* src = s->h_film_staging;
    const size_t n = npx * 4;
    const unsigned nt = n >= (1u << 20) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
    auto add = [film, src](size_t a, size_t b) { for (size_t i = a; i < b; ++i) film[i] += src[i]; };
    std::vector<std::thread> th;
    for (unsigned k = 1; k < nt; ++k) th.emplace_back(add, n * k / nt, n * (k + 1) / nt);
    add(0, n / nt);
    for (auto& t : th) t.join();
    return TRB_OK;
}
trb_status stats_to_host(trb_scene* s, trb_stats* stats, bool accumulate) {
    trb::DStats h;
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpy(&h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost));
    trb_stats one; std::memset(&one, 0, sizeof one);
    stats_out(h, &one);
    if (!accumulate) { *stats = one; return TRB_OK; }
    stats->camera_samples += one.camera_samples; stats->rays_primary += one.rays_primary; stats->rays_shadow += one.rays_shadow; stats->rays_mis += one.rays_mis;
    stats->rays_continuation += one.rays_continuation; stats->node_tests += one.node_tests; stats->tri_tests += one.tri_tests; stats->inst_tests += one.inst_tests;
    return TRB_OK;
}
} // namespace

extern "C" {

trb_status trb_nccl_unique_id(void* id128) {
    if (!id128) return fail(TRB_INVALID_ARG, "null argument");
    trb_status r = nccl_load();
    if (r != TRB_OK) return r;
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == TRB_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    std::memcpy(id128, &id, sizeof id);
    return TRB_OK;
}

trb_status trb_comm_create(const void* id128, int n_ranks, int rank, int device, trb_comm** out) {
    if (!id128 || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(TRB_INVALID_ARG, "bad communicator arguments");
    *out = nullptr;
    trb_status r = nccl_load();
    if (r != TRB_OK) return r;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TRB_NO_DEVICE, "no CUDA device: tray_rust_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(TRB_INVALID_ARG, "device ordinal out of range");
    CU(cudaSetDevice(device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    std::unique_ptr<trb_comm> c(new trb_comm);
    c->n_ranks = n_ranks; c->rank = rank; c->device = device;
    NC(g_nccl.CommInitRank(&c->comm, n_ranks, id, rank));
    *out = c.release();
    return TRB_OK;
}

void trb_comm_destroy(trb_comm* c) {
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) { cudaSetDevice(c->device); g_nccl.CommDestroy(c->comm); }
    delete c;
}

trb_status trb_comm_info(const trb_comm* c, int* n_ranks, int* rank) {
    if (!c) return fail(TRB_INVALID_ARG, "null communicator");
    if (n_ranks) *n_ranks = c->n_ranks;
    if (rank) *rank = c->rank;
    return TRB_OK;
}

trb_status trb_comm_reduce_film(trb_comm* c, float* d_film, size_t n, int root, void* stream) {
    if (!c || !d_film || root < 0 || root >= c->n_ranks) return fail(TRB_INVALID_ARG, "bad reduce arguments");
    if (c->n_ranks == 1) return TRB_OK;
    CU(cudaSetDevice(c->device));
    NC(g_nccl.Reduce(d_film, d_film, n, ncclFloat, ncclSum, root, c->comm, static_cast<cudaStream_t>(stream)));
    return TRB_OK;
}

trb_status trb_render_sharded(trb_scene* s, trb_comm* c, const trb_render_cfg* cfg, int root, float* film, trb_stats* stats) {
    if (!s || !c || !cfg || root < 0 || root >= c->n_ranks) return fail(TRB_INVALID_ARG, "null or bad argument");
    if (c->rank == root && !film) return fail(TRB_INVALID_ARG, "the root rank needs a film buffer");
    if (c->device != s->device) return fail(TRB_INVALID_ARG, "scene and communicator live on different devices");
    trb_render_cfg mine; bool empty;
    trb_status r = shard_cfg(s, cfg, c->rank, c->n_ranks, &mine, &empty);
    if (r != TRB_OK) return r;
    auto t0 = std::chrono::steady_clock::now();
    CU(cudaSetDevice(s->device));
    CU(cudaEventRecord(s->ev0, 0));
    r = render_to_device_film(s, &mine, empty, nullptr);
    if (r != TRB_OK) return r;
    CU(cudaEventRecord(s->ev1, 0));
    r = trb_comm_reduce_film(c, reinterpret_cast<float*>(s->d_film), (size_t)s->film.width * s->film.height * 4, root, nullptr); // ONE reduce per frame
    if (r != TRB_OK) return r;
    if (c->rank == root) { r = film_to_host_add(s, film, nullptr); if (r != TRB_OK) return r; }
    else CU(cudaStreamSynchronize(nullptr));
    r = check_error_flag(s);
    if (r != TRB_OK) return r;
    if (stats) {
        r = stats_to_host(s, stats, false);
        if (r != TRB_OK) return r;
        CU(cudaEventElapsedTime(&stats->kernel_ms, s->ev0, s->ev1));
        stats->update_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count() - stats->kernel_ms;
    }
    return TRB_OK;
}

trb_status trb_group_create(const trb_scene_desc* desc, const int* devices, int n, trb_group** out) {
    if (!out || !devices || n < 1) return fail(TRB_INVALID_ARG, "bad group arguments");
    *out = nullptr;
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) if (devices[i] == devices[j]) return fail(TRB_INVALID_ARG, "duplicate device in group");
    std::unique_ptr<trb_group> g(new trb_group);
    g->devices.assign(devices, devices + n);
    auto cleanup = [&]() { for (trb_scene* s : g->scenes) trb_scene_destroy(s); g->scenes.clear(); };
    // replicas are built concurrently (the per-mesh SAH build is host work)
    std::vector<trb_scene*> scenes(n, nullptr);
    std::vector<trb_status> rc(n, TRB_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i) th.emplace_back([&, i]() { rc[i] = trb_scene_create(desc, devices[i], &scenes[i]); if (rc[i] != TRB_OK) msg[i] = trb_last_error(); });
    for (auto& t : th) t.join();
    g->scenes = scenes;
    for (int i = 0; i < n; ++i) if (rc[i] != TRB_OK) { g->scenes.erase(std::remove(g->scenes.begin(), g->scenes.end(), nullptr), g->scenes.end()); cleanup(); return fail(rc[i], msg[i]); }
    if (n > 1) {
        trb_status r = nccl_load();
        if (r != TRB_OK) { cleanup(); return r; }
        g->comms.resize(n);
        ncclResult_t e = g_nccl.CommInitAll(g->comms.data(), n, devices);
        if (e != ncclSuccess) { cleanup(); return fail(TRB_NCCL, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(e)); }
    }
    *out = g.release();
    return TRB_OK;
}

trb_status trb_group_load_json(const char* path, uint32_t w, uint32_t h, uint32_t spp, const int* devices, int n, trb_group** out) {
    trb_scene_desc* d = nullptr;
    trb_status r = trb_desc_load_json(path, w, h, spp, &d);
    if (r != TRB_OK) return r;
    r = trb_group_create(d, devices, n, out);
    const std::string keep = g_error;
    trb_desc_free(d);
    g_error = keep;
    return r;
}

trb_scene* trb_group_scene(trb_group* g, int i) { return (g && i >= 0 && i < (int)g->scenes.size()) ? g->scenes[i] : nullptr; }

void trb_group_destroy(trb_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->comms.size(); ++i) if (g->comms[i]) { cudaSetDevice(g->devices[i]); g_nccl.CommDestroy(g->comms[i]); }
    for (trb_scene* s : g->scenes) trb_scene_destroy(s);
    delete g;
}

trb_status trb_group_render(trb_group* g, const trb_render_cfg* cfg, float* film, trb_stats* stats) {
    if (!g || !cfg || !film) return fail(TRB_INVALID_ARG, "null argument");
    const int n = (int)g->scenes.size();
    if (n == 1) return trb_render(g->scenes[0], cfg, film, stats);
    auto t0 = std::chrono::steady_clock::now();
    // enqueue every replica's shard (update_frame is host work per replica; the kernels of all devices then run concurrently)
    std::vector<trb_status> rc(n, TRB_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i) th.emplace_back([&, i]() {
        trb_render_cfg mine; bool empty;
        rc[i] = shard_cfg(g->scenes[i], cfg, i, n, &mine, &empty);
        if (rc[i] == TRB_OK) { cudaSetDevice(g->scenes[i]->device); cudaEventRecord(g->scenes[i]->ev0, 0); rc[i] = render_to_device_film(g->scenes[i], &mine, empty, nullptr); cudaEventRecord(g->scenes[i]->ev1, 0); }
        if (rc[i] != TRB_OK) msg[i] = trb_last_error();
    });
    for (auto& t : th) t.join();
    for (int i = 0; i < n; ++i) if (rc[i] != TRB_OK) return fail(rc[i], msg[i]);
    const size_t nfl = (size_t)g->scenes[0]->film.width * g->scenes[0]->film.height * 4;
    NC(g_nccl.GroupStart()); // ONE reduce per frame, root = devices[0]
    for (int i = 0; i < n; ++i) {
        CU(cudaSetDevice(g->scenes[i]->device));
        NC(g_nccl.Reduce(g->scenes[i]->d_film, g->scenes[i]->d_film, nfl, ncclFloat, ncclSum, 0, g->comms[i], nullptr));
    }
    NC(g_nccl.GroupEnd());
    CU(cudaSetDevice(g->scenes[0]->device));
    trb_status r = film_to_host_add(g->scenes[0], film, nullptr);
    if (r != TRB_OK) return r;
    float kernel_ms = 0.f;
    if (stats) std::memset(stats, 0, sizeof *stats);
    for (int i = 0; i < n; ++i) {
        CU(cudaSetDevice(g->scenes[i]->device));
        CU(cudaDeviceSynchronize());
        r = check_error_flag(g->scenes[i]);
        if (r != TRB_OK) return r;
        if (stats) {
            r = stats_to_host(g->scenes[i], stats, true);
            if (r != TRB_OK) return r;
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, g->scenes[i]->ev0, g->scenes[i]->ev1));
            kernel_ms = std::max(kernel_ms, ms);
        }
    }
    if (stats) { stats->kernel_ms = kernel_ms; stats->update_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count() - kernel_ms; }
    return TRB_OK;
}

} // extern "C"
