/* trb.h — C ABI of the B200-native render path for tray_rust.
 *
 * This is the drop-in boundary for ONE path of the reference: the call
 *     Exec::render(&mut self, scene: &mut Scene, rt: &mut RenderTarget, config: &Config)
 * (/root/reference/src/exec/mod.rs:41-49; sole implementation
 * exec::MultiThreaded, src/exec/multithreaded.rs:54-114) and the data it consumes
 * (Scene, src/scene.rs:93-98) and produces (the RGBW f32 film in the layout of
 * RenderTarget::get_renderf32, src/film/render_target.rs:243-265).
 *
 * Everything here is plain C: POD structs, pointers and sizes. No torch types, no
 * C++ types, no callbacks into the host. The library owns all device memory behind
 * the opaque trb_scene handle; host buffers are borrowed for the duration of a call.
 *
 * All entry points return trb_status; on failure trb_last_error() (thread-local)
 * describes why. The conditions the reference panics on (image not a multiple of the
 * 8x8 block, no lights, empty scene, unknown types) are reported as
 * TRB_INVALID_ARG / TRB_UNSUPPORTED instead of aborting the process.
 */
#ifndef TRB_H
#define TRB_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRB_ABI_VERSION 4u

typedef enum trb_status {
    TRB_OK = 0,
    TRB_INVALID_ARG = 1, /* what the reference would panic!/assert! on */
    TRB_CUDA = 2,        /* a CUDA runtime call or kernel failed */
    TRB_OOM = 3,
    TRB_UNSUPPORTED = 4, /* valid in the reference, not in this build */
    TRB_IO = 5,
    TRB_NO_DEVICE = 6,   /* no CUDA device: this library has no CPU fallback */
    TRB_NCCL = 7         /* an NCCL call failed, or libnccl.so.2 could not be loaded */
} trb_status;

/* ---------------------------------------------------------------------------------
 * Scene description (source-level, nothing derived: no matrices, no BVH).
 * It is the FFI-safe flattening of `Scene` + `RenderTarget` construction arguments
 * (src/scene.rs:93-146). Every array is caller-owned and deep-copied by
 * trb_scene_create.
 * ------------------------------------------------------------------------------- */

/* linalg::Keyframe (src/linalg/keyframe.rs:14-21): translation, rotation quaternion
 * (v.x, v.y, v.z, w), scaling. The reference stores EVERY transform this way (even a
 * static JSON one, src/linalg/animated_transform.rs:34-37) and recomposes T*R*S per
 * ray (keyframe.rs:60-63), so the TRS triple — not a matrix — is the ABI. */
typedef struct trb_keyframe {
    float translation[3];
    float rotation[4];
    float scaling[3];
} trb_keyframe;

/* bspline::BSpline<Keyframe> (animated_transform.rs:22-37): control points
 * keyframes[ctrl_first .. ctrl_first+n_ctrl), knots[knot_first .. knot_first+n_knots). */
typedef struct trb_spline {
    uint32_t degree;
    uint32_t n_ctrl;
    uint32_t ctrl_first;
    uint32_t n_knots;
    uint32_t knot_first;
} trb_spline;

/* film::ColorKeyframe (src/film/animated_color.rs:12-17), sorted by time. */
typedef struct trb_color_key {
    float rgba[4];
    float time;
} trb_color_key;

enum { /* geometry::Instance / EmitterType (src/geometry/instance.rs:81-84, emitter.rs:74-79) */
    TRB_INST_RECEIVER = 0,
    TRB_INST_EMITTER_AREA = 1,
    TRB_INST_EMITTER_POINT = 2
};
enum { /* shapes reachable from scene.rs:513-580 */
    TRB_SHAPE_NONE = 0,   /* point light */
    TRB_SHAPE_SPHERE = 1, /* p0 = radius                       (geometry/sphere.rs)    */
    TRB_SHAPE_DISK = 2,   /* p0 = radius, p1 = inner_radius    (geometry/disk.rs)      */
    TRB_SHAPE_RECT = 3,   /* p0 = width,  p1 = height; "plane" = 2x2 (scene.rs:527-528) */
    TRB_SHAPE_MESH = 4    /* mesh = index into meshes          (geometry/mesh.rs)      */
};

/* One geometry::Instance in JSON object order (group members flattened in place,
 * scene.rs:496-505). The AnimatedTransform is splines[spline_first .. +n_splines) in
 * application order (animated_transform.rs:42-54: transform = t_i * transform). */
typedef struct trb_instance {
    uint32_t kind;
    uint32_t shape;
    float p0, p1;
    uint32_t mesh;
    uint32_t material;
    uint32_t spline_first, n_splines;
    uint32_t emission_first, n_emission; /* color keys; emitters only */
} trb_instance;

/* geometry::Mesh buffers as produced by Mesh::load_obj (mesh.rs:49-76): positions and
 * normals 3 floats per vertex, texcoords 2 floats per vertex, 3 indices per triangle. */
typedef struct trb_mesh {
    uint32_t n_verts;
    uint32_t n_tris;
    const float* positions;
    const float* normals;
    const float* texcoords;
    const uint32_t* indices;
} trb_mesh;

/* texture::Image / texture::AnimatedImage (src/texture/image.rs:36-48, animated_image.rs): what scene.rs:317-394 builds for
 * "image", "animated_image" and "movie" textures. A frame is RGBA8 with the semantics of image::DynamicImage::get_pixel
 * (grey -> (l, l, l, 255), RGB -> alpha 255), row-major from the top; one frame = Image, two or more = AnimatedImage whose
 * frames carry their keyframe times in the order given. */
typedef struct trb_image {
    uint32_t width, height;
    const uint8_t* rgba8; /* width * height * 4 bytes */
    float time;
    uint32_t pad;
} trb_image;
typedef struct trb_texture {
    uint32_t first_image, n_images; /* images[first_image .. first_image + n_images) */
} trb_texture;

enum { /* material types (src/material/); every colour / scalar parameter is a constant or a texture (texture/mod.rs) */
    TRB_MAT_MATTE = 0,          /* c0 = diffuse, roughness (degrees; 0 => Lambertian) */
    TRB_MAT_PLASTIC = 1,        /* c0 = diffuse, c1 = gloss, roughness                */
    TRB_MAT_METAL = 2,          /* c0 = refractive_index, c1 = absorption_coefficient, roughness */
    TRB_MAT_SPECULAR_METAL = 3, /* c0 = refractive_index, c1 = absorption_coefficient */
    TRB_MAT_GLASS = 4,          /* c0 = reflect, c1 = transmit, eta                   */
    TRB_MAT_ROUGH_GLASS = 5,    /* c0 = reflect, c1 = transmit, eta, roughness        */
    TRB_MAT_MERL = 6            /* merl = index into merl_tables                       */
};
typedef struct trb_material {
    uint32_t type;
    float c0[3];
    float c1[3];
    float roughness;
    float eta;
    uint32_t merl;
    /* LoadedTextures::find_color / find_scalar (scene.rs:45-87): a parameter given as a texture NAME in the JSON is sampled at the
     * hit's (u, v, time) in Material::bsdf. tex[k] = 1 + index into trb_scene_desc.textures, 0 = the constant above;
     * k: 0 = c0, 1 = c1 (sample_color), 2 = roughness, 3 = eta (sample_f32). */
    uint32_t tex[4];
} trb_material;

#define TRB_MERL_N_THETA_H 90u
#define TRB_MERL_N_THETA_D 90u
#define TRB_MERL_N_PHI_D 180u
#define TRB_MERL_TABLE_FLOATS (90u * 90u * 180u * 3u) /* interleaved RGB f32 as built by material::Merl::load_file (merl.rs:51-84) */

/* film::Camera construction arguments (src/film/camera.rs:64-91). Animated fov
 * (camera.rs:95-125): fov_ctrl/fov_knots non-empty. */
typedef struct trb_camera {
    uint32_t spline_first, n_splines; /* cam_world AnimatedTransform */
    float fov;                        /* degrees */
    float shutter_size;               /* default 0.5 (scene.rs:197-200) */
    uint32_t active_at;
    uint32_t fov_degree, n_fov_ctrl, fov_ctrl_first, n_fov_knots, fov_knot_first; /* into fov_floats */
} trb_camera;

enum { TRB_FILTER_MITCHELL_NETRAVALI = 0, TRB_FILTER_GAUSSIAN = 1 };
typedef struct trb_film {
    uint32_t width, height;  /* multiples of 8 (block_queue.rs:29-31) and of 2 (render_target.rs:43-45) */
    uint32_t samples;        /* spp; rounded up to a power of two (ld.rs:22-26) */
    uint32_t frames, start_frame, end_frame;
    float scene_time;
    uint32_t filter_type;
    float filter_w, filter_h;
    float filter_b, filter_c; /* Mitchell-Netravali b,c; Gaussian: filter_b = alpha */
} trb_film;

enum {
    TRB_INTEGRATOR_PATH = 0,          /* integrator/path.rs:35-43: min_depth, max_depth */
    TRB_INTEGRATOR_WHITTED = 1,       /* integrator/whitted.rs:29-38: max_depth = recursion limit (the JSON loader reads it from
                                         "min_depth", like scene.rs:305-309); min_depth unused */
    TRB_INTEGRATOR_NORMALS_DEBUG = 2  /* integrator/normals_debug.rs:25-36: (shading normal + 1) / 2 */
};
typedef struct trb_integrator {
    uint32_t type;
    uint32_t min_depth, max_depth;
} trb_integrator;

typedef struct trb_scene_desc {
    uint32_t abi_version; /* TRB_ABI_VERSION */
    trb_film film;
    trb_integrator integrator;
    uint32_t n_cameras;   const trb_camera* cameras;     /* sorted by active_at (scene.rs:190) */
    uint32_t n_instances; const trb_instance* instances; /* JSON object order == light order (Q20) */
    uint32_t n_splines;   const trb_spline* splines;
    uint32_t n_keyframes; const trb_keyframe* keyframes;
    uint32_t n_knots;     const float* knots;
    uint32_t n_color_keys; const trb_color_key* color_keys;
    uint32_t n_meshes;    const trb_mesh* meshes;
    uint32_t n_materials; const trb_material* materials;
    uint32_t n_merl;      const float* const* merl_tables; /* each TRB_MERL_TABLE_FLOATS floats */
    uint32_t n_fov_floats; const float* fov_floats;
    uint32_t n_textures;  const trb_texture* textures;
    uint32_t n_images;    const trb_image* images;
} trb_scene_desc;

/* ---------------------------------------------------------------------------------
 * Render configuration = exec::Config (src/exec/mod.rs:17-37) minus paths/threads,
 * plus what the reference leaves to the OS: the RNG seed (multithreaded.rs:79 seeds
 * StdRng from the OS; see DESIGN.md "RNG") and a sample sub-range so a frame can be
 * rendered in additive passes.
 * ------------------------------------------------------------------------------- */
enum {
    TRB_RENDER_STATS = 1u,            /* also count BVH node / triangle / instance tests */
    TRB_RENDER_NO_UPDATE = 2u,        /* skip Scene::update_frame (caller already did it) */
    TRB_RENDER_REFERENCE_SHADOW = 4u, /* trace shadow rays as full closest-hit like light/mod.rs:30-37 instead of
                                         stopping at the first accepted hit (same boolean, fewer tests) */
    TRB_RENDER_MEGAKERNEL = 8u,       /* one persistent kernel per pass instead of the wavefront pipeline (same results) */
    TRB_RENDER_TIME_TRACE = 16u       /* bracket every trace-kernel launch with CUDA events (read with trb_scene_trace_time) */
};
typedef struct trb_render_cfg {
    uint32_t spp;           /* Config.spp; 0 = film.samples. Rounded up to pow2 like ld.rs:22-26 */
    uint32_t sample_first;  /* render sample indices [sample_first, sample_first+sample_count) */
    uint32_t sample_count;  /* of each pixel's spp; 0 = all */
    uint32_t block_start;   /* Config.select_blocks.0 (index into the Morton-sorted 8x8 block list) */
    uint32_t block_count;   /* Config.select_blocks.1; 0 = all blocks (block_queue.rs:39-41) */
    uint32_t current_frame; /* Config.current_frame */
    uint32_t seed;
    uint32_t flags;
    /* Optional interleaved sharding of the selected block list for multi-GPU load balance: keep block j (index in the
     * list after select_blocks) iff (j / shard_chunk) % shard_count == shard_index. shard_count <= 1 disables it. The
     * reference shards by contiguous ranges only (master.rs:91-93); the union and the summed film are the same. */
    uint32_t shard_index, shard_count, shard_chunk;
} trb_render_cfg;

typedef struct trb_stats {
    uint64_t camera_samples;
    uint64_t rays_primary;      /* multithreaded.rs:97        */
    uint64_t rays_shadow;       /* light/mod.rs:30-37         */
    uint64_t rays_mis;          /* integrator/mod.rs:156-157  */
    uint64_t rays_continuation; /* integrator/path.rs:112     */
    uint64_t node_tests;        /* BBox::fast_intersect calls (TLAS + BLAS); TRB_RENDER_STATS only */
    uint64_t tri_tests;         /* intersect_triangle calls;                 TRB_RENDER_STATS only */
    uint64_t inst_tests;        /* Instance::intersect calls;                TRB_RENDER_STATS only */
    float kernel_ms;            /* device time of the render kernels of this call (CUDA events) */
    float update_ms;            /* host time of Scene::update_frame + upload */
} trb_stats;

/* linalg::Ray without depth/time (src/linalg/ray.rs:9-22): 32 bytes. */
typedef struct trb_ray {
    float o[3];
    float d[3];
    float min_t, max_t;
} trb_ray;

/* Result of Scene::intersect (scene.rs:148-150): ray.max_t after traversal, the
 * instance hit (index into instances) and, for meshes, the triangle index.
 * inst == TRB_MISS when nothing was hit. 16 bytes. */
#define TRB_MISS 0xffffffffu
typedef struct trb_hit {
    float t;
    uint32_t inst;
    uint32_t prim;
    uint32_t pad;
} trb_hit;

/* Per camera sample record for parity tests: film position and the clamped radiance
 * pushed as ImageSample (multithreaded.rs:98-102). */
typedef struct trb_sample {
    float x, y;
    float r, g, b;
} trb_sample;

/* Flattened BVH node in the reference's order (bvh.rs:248-267): interior: a =
 * second_child, b = axis (0,1,2); leaf: a = geom_offset, b = 0x80000000 | ngeom. */
typedef struct trb_bvh_node {
    float bmin[3];
    float bmax[3];
    uint32_t a;
    uint32_t b;
} trb_bvh_node;
#define TRB_BVH_LEAF 0x80000000u

typedef struct trb_scene trb_scene;

/* -- lifecycle -------------------------------------------------------------------- */

/* ≙ the construction half of Scene::load_file (scene.rs:101-146): builds each mesh's
 * BVH<Triangle> (mesh.rs:44, max_geom 16) and uploads the scene to `device`
 * (cudaSetDevice ordinal). */
trb_status trb_scene_create(const trb_scene_desc* desc, int device, trb_scene** out);

/* ≙ Scene::load_file(file) (scene.rs:101): JSON + OBJ + MERL loading, then
 * trb_scene_create. width/height/spp > 0 override film.width/height/samples (the
 * BASELINE.json configs do this). */
trb_status trb_scene_load_json(const char* path, uint32_t width, uint32_t height, uint32_t spp,
                               int device, trb_scene** out);

void trb_scene_destroy(trb_scene* scene);

/* film dimensions and rounded spp of a scene */
trb_status trb_scene_info(const trb_scene* scene, uint32_t* width, uint32_t* height, uint32_t* spp,
                          uint32_t* n_blocks, uint32_t* n_instances, uint32_t* n_lights);

/* ≙ Scene::update_frame(frame, start, end) (scene.rs:152-176): selects the camera,
 * sets the shutter interval, recomposes instance transforms, rebuilds the
 * BVH<Instance> (max_geom 4) for the shutter interval and uploads it. */
trb_status trb_scene_update_frame(trb_scene* scene, uint32_t frame, float start, float end);

/* -- the hot path ------------------------------------------------------------------ */

/* ≙ Exec::render (exec/mod.rs:48; multithreaded.rs:55-70). Renders the selected
 * blocks at the selected samples — with sample_count = 0 the WHOLE spp of the frame,
 * like MultiThreaded::render — on the scene's GPU and ADDS the RGBW film (row-major,
 * width*height*4 floats, layout of get_renderf32 render_target.rs:243-265) into the
 * HOST buffer `film_rgbw` — additive like film::Image::add_pixels (film/image.rs:21-33).
 * Internally the frame is rendered in additive passes sized to the free device memory
 * (212 B of path state per camera sample in flight); the film is accumulated on the
 * device and copied to the host once. Includes update_frame unless
 * TRB_RENDER_NO_UPDATE. Blocking. */
trb_status trb_render(trb_scene* scene, const trb_render_cfg* cfg, float* film_rgbw, trb_stats* stats);

/* Same, but the film is a DEVICE buffer on the scene's GPU (accumulated into), and the
 * work is enqueued on `cuda_stream` (a cudaStream_t; NULL = default stream) without
 * host synchronisation. Never calls update_frame. `stats` (may be NULL) is a DEVICE
 * pointer to a trb_stats the kernels accumulate ray counters into.
 * Synchronisation contract: passes of one scene must be enqueued on ONE stream at a time
 * (they share the scene's path-state buffers); trb_scene_update_frame drains the device
 * before it touches the instance / TLAS buffers; a traversal-stack overflow (the reference
 * panics) is latched on the device and reported by trb_scene_check_error or by the next
 * host-buffer call (trb_render / trb_intersect / trb_render_samples). */
trb_status trb_render_device(trb_scene* scene, const trb_render_cfg* cfg, float* d_film_rgbw,
                             trb_stats* d_stats, void* cuda_stream);

/* -- multi-GPU: tile sharding + ONE film SUM-reduce per frame -------------
 *
 * The reference's distributed mode gives worker r of W the blocks [r*floor(B/W), ...) of the Morton block list
 * (exec/distrib/master.rs:88-93,218-224), every worker renders its blocks at full spp with the scene replicated
 * (worker.rs:37-89), and the master adds the workers' films (film/image.rs:21-50, master.rs:124-163). Here the
 * exchange is one ncclReduce(SUM, fp32) of the RGBW film over NVLink at frame end. libnccl.so.2 is resolved at run
 * time (the copy already loaded in the process, else the system one): a single-GPU user never needs it.
 *
 * Two shapes:
 *   one process per GPU  trb_nccl_unique_id (rank 0; ship the 128 bytes to the other ranks by any means)
 *                        -> trb_comm_create on every rank -> trb_render_sharded per frame
 *   one process, n GPUs  trb_group_create / trb_group_load_json -> trb_group_render per frame
 */
#define TRB_NCCL_UNIQUE_ID_BYTES 128
typedef struct trb_comm trb_comm;
typedef struct trb_group trb_group;

trb_status trb_nccl_unique_id(void* id128);
/* ncclCommInitRank on `device`: collective over the n_ranks processes. */
trb_status trb_comm_create(const void* id128, int n_ranks, int rank, int device, trb_comm** out);
void trb_comm_destroy(trb_comm* comm);
trb_status trb_comm_info(const trb_comm* comm, int* n_ranks, int* rank);
/* SUM-reduce `n_floats` of a DEVICE film into rank `root`'s buffer (in place), enqueued on cuda_stream. */
trb_status trb_comm_reduce_film(trb_comm* comm, float* d_film_rgbw, size_t n_floats, int root, void* cuda_stream);

/* ≙ Exec::render of one rank of the distributed mode + the master's film sum: renders this rank's share of the
 * selected blocks (interleaved chunks of cfg->shard_chunk blocks, default 32, rank = shard index; cfg->shard_count
 * == 0xffffffff selects the reference's contiguous ranges instead), all samples, film kept on the device, then ONE
 * reduce to `root`; on the root the summed film is ADDED into the host buffer `film_rgbw` (ignored elsewhere, may
 * be NULL). `stats` receives this rank's counters. Blocking. */
trb_status trb_render_sharded(trb_scene* scene, trb_comm* comm, const trb_render_cfg* cfg, int root, float* film_rgbw,
                              trb_stats* stats);

/* One process driving n GPUs: a scene replica per device (trb_scene_create on each) plus communicators from
 * ncclCommInitAll. trb_group_render ≙ Exec::render on all of them: tile-sharded, one reduce to devices[0], film
 * ADDED into the host buffer; `stats` is the sum over the devices. */
trb_status trb_group_create(const trb_scene_desc* desc, const int* devices, int n_devices, trb_group** out);
trb_status trb_group_load_json(const char* path, uint32_t width, uint32_t height, uint32_t spp, const int* devices,
                               int n_devices, trb_group** out);
trb_status trb_group_render(trb_group* group, const trb_render_cfg* cfg, float* film_rgbw, trb_stats* stats);
trb_scene* trb_group_scene(trb_group* group, int index); /* borrowed: replica `index` (film_to_srgb8, info, options) */
void trb_group_destroy(trb_group* group);

/* ≙ Scene::intersect (scene.rs:148-150) for a batch of rays: closest hit through the
 * two-level BVH in the reference's traversal order. Host buffers. */
trb_status trb_intersect(trb_scene* scene, size_t n, const trb_ray* rays, trb_hit* hits, trb_stats* stats);

/* Device-buffer variant of trb_intersect, enqueued on cuda_stream. */
trb_status trb_intersect_device(trb_scene* scene, size_t n, const trb_ray* d_rays, trb_hit* d_hits,
                                trb_stats* d_stats, void* cuda_stream);

/* ≙ LowDiscrepancy::get_samples + get_samples_1d + Camera::generate_ray
 * (ld.rs:33-64, camera.rs:150-157) for the selected blocks/samples: writes one ray and
 * one film position per camera sample, in block-list order, pixel row-major within the
 * block, sample index minor. Host buffers sized n = blocks*64*sample_count. */
trb_status trb_camera_rays(trb_scene* scene, const trb_render_cfg* cfg, size_t n, trb_ray* rays, float* xy);

/* Parity/debug variant of trb_render: instead of splatting, writes the clamped
 * radiance of every camera sample (same order as trb_camera_rays). Host buffer. */
trb_status trb_render_samples(trb_scene* scene, const trb_render_cfg* cfg, size_t n, trb_sample* samples,
                              trb_stats* stats);

/* ≙ RenderTarget::get_render (render_target.rs:185-210): rgb/weight, clamp, sRGB,
 * (c*255) as u8; pixels with weight <= 0 stay 0. Host buffers; runs on the scene's GPU. */
trb_status trb_film_to_srgb8(trb_scene* scene, const float* film_rgbw, uint8_t* rgb8);

/* ≙ image::save_buffer(path, &img, w, h, image::RGB(8)) for the frames written by main.rs:95-103 and by the distributed
 * master (exec/distrib/master.rs:137-142): an 8-bit RGB PNG (stored deflate blocks; host only, no device needed). */
trb_status trb_write_png(const char* path, const uint8_t* rgb8, uint32_t width, uint32_t height);

/* -- introspection for parity tests ------------------------------------------------ */

/* The Morton-sorted 8x8 block list (block_queue.rs:28-46) after select_blocks: pairs (bx,by). */
trb_status trb_block_list(const trb_scene* scene, uint32_t block_start, uint32_t block_count,
                          uint32_t* n_out, uint32_t* xy_pairs, uint32_t capacity);

/* Flattened BVH in the reference's node order. which = -1: BVH<Instance> (after
 * update_frame); which >= 0: the BVH<Triangle> of mesh `which`. Pass NULL buffers to
 * query sizes. `ordered` is bvh.rs `ordered_geom`. */
trb_status trb_scene_get_bvh(const trb_scene* scene, int which, uint32_t* n_nodes, trb_bvh_node* nodes,
                             uint32_t* n_ordered, uint32_t* ordered);

/* World transform (mat, inv: 16 floats each, row-major) of instance i after update_frame. */
trb_status trb_scene_get_transform(const trb_scene* scene, uint32_t inst, float* mat16, float* inv16);

/* The 16x16 filter table (render_target.rs:50-57). */
trb_status trb_scene_get_filter_table(const trb_scene* scene, float* table256);

/* -- host-only helpers (no device needed; used by the CPU test-suite) ---------------- */

/* BVH::new over `n` boxes (6 floats each: min xyz, max xyz) with the reference's SAH
 * build (bvh.rs:139-267). Pass NULL buffers to query sizes. */
trb_status trb_host_build_bvh(const float* boxes6, uint32_t n, uint32_t max_geom, uint32_t* n_nodes,
                              trb_bvh_node* nodes, uint32_t* ordered);

/* Keyframe::transform (keyframe.rs:60-63): T * R * S and its inverse, row-major. */
trb_status trb_host_keyframe_transform(const trb_keyframe* kf, float* mat16, float* inv16);

/* Self-check of the two-level node records the trace kernel walks (csrc/trb_device.h DQuad): for every ray
 * with a finite 1/d, a literal BVH::intersect walk (bvh.rs:81-130) over `nodes` and a walk through the packed
 * records must visit the same leaves in the same order. Host only. */
trb_status trb_host_quad_check(const trb_bvh_node* nodes, uint32_t n_nodes, const trb_ray* rays, uint32_t n_rays,
                               uint32_t* mismatches, uint64_t* leaf_visits, uint64_t* quad_visits);

/* Device self-check of the trace kernel's short box test (csrc/trb_kernels.cuh box_hit_finite, used for rays whose origin,
 * direction and 1/direction are all finite) against the literal transcription of BBox::fast_intersect (bbox.rs:75-104) on
 * n_cases generated boxes and rays built from awkward values (signed zeros, denormals, huge / tiny magnitudes, origins on box
 * planes, flat boxes). out[0] = cases with an all-finite ray, out[1] = hits among them, out[2] = MISMATCHES (must be 0),
 * out[3] = cases that take the literal path. Needs a GPU (TRB_NO_DEVICE otherwise). */
trb_status trb_selftest_box(uint32_t n_cases, uint32_t seed, uint64_t out[4]);

/* AnimatedTransform::transform(time) (animated_transform.rs:40-56) of the transform stack
 * desc->splines[first .. first+count): the same code the device runs per ray for keyframed
 * instances, compiled for the host. AnimatedColor::color(time) (animated_color.rs:52-78) of
 * desc->color_keys[first .. first+count) likewise. */
trb_status trb_host_animated_transform(const trb_scene_desc* desc, uint32_t first, uint32_t count, float time,
                                       float* mat16, float* inv16);
trb_status trb_host_animated_color(const trb_scene_desc* desc, uint32_t first, uint32_t count, float time, float* rgb3);

/* The JSON loader alone (Scene::load_file up to the flattened description): the result is
 * owned by the library; release with trb_desc_free. */
trb_status trb_desc_load_json(const char* path, uint32_t width, uint32_t height, uint32_t spp,
                              trb_scene_desc** out);
void trb_desc_free(trb_scene_desc* desc);

/* Drains the scene's device and reports a latched traversal-stack overflow of passes enqueued with
 * trb_render_device / trb_intersect_device (TRB_CUDA), else TRB_OK. */
trb_status trb_scene_check_error(trb_scene* scene);

/* Launch-shape options of the wavefront pipeline (results never depend on them; DESIGN.md "Launch options"):
 * "pass.paths" camera samples per pass; "trace.pipe" trace-kernel variant (36 = default: 9 CTAs per SM; 0 = the round-1 kernel),
 * "trace.refill", "trace.grid", "trace.sched", "trace.quads", "trace.exact_box" (test: every ray takes the literal box test);
 * "shade.split" -1 per scene / 0 fused / 1 split, "shade.sort" 0/1 material buckets between the split kernels, "shade.kind" 0/1 the
 * matte instantiations of the split kernels, "shade.anim_occupancy" 3/4; "anim.table" 0/1/2, "frame.device" 0/1; "film.v2" 0/1;
 * "sort.mode" 0/1/2 ray-queue sorting, "sort.bits", "sort.min_round". "trace.quads" and "frame.device" re-run update_frame for the
 * current frame (they change what it builds). Defaults can also be preset by TRB_* environment variables, read once by
 * trb_scene_create. */
trb_status trb_scene_set_option(trb_scene* scene, const char* name, long long value);

/* Device time spent in the dominant kernel (k_wf_trace) by launches made with TRB_RENDER_TIME_TRACE since the last
 * call, measured with CUDA events on the launching stream; synchronises those events. */
trb_status trb_scene_trace_time(trb_scene* scene, float* total_ms, uint32_t* n_launches);

/* Number of CUDA kernels this library has launched in this process (monotonic). */
unsigned long long trb_launch_count(void);

const char* trb_last_error(void);
uint32_t trb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TRB_H */
