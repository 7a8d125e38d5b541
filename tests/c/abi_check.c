/* Plain-C caller of include/trb.h (the drop-in boundary): compiled by tests/test_abi_and_host.py with `gcc -std=c11`.
 *
 *   abi_check layout                  prints sizeof / offsetof of every ABI struct as "struct.field offset" lines; the Python test
 *                                     compares them with the ctypes mirrors (tray_rust_b200/_ffi.py) and with the Rust #[repr(C)]
 *                                     declarations documented in INTEGRATION.md. The _Static_asserts pin ABI v4's layout.
 *   abi_check render scene.json w h spp   the documented call sequence of INTEGRATION.md from C: trb_scene_load_json ->
 *                                     trb_render (sample_count = 0: the whole frame in one call) -> trb_film_to_srgb8.
 * What the reference does at this boundary: Exec::render(&mut Scene, &mut RenderTarget, &Config)
 * (/root/reference/src/exec/mod.rs:17-49, exec/multithreaded.rs:55-70) and RenderTarget::get_render (film/render_target.rs:185-210). */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "trb.h"

_Static_assert(TRB_ABI_VERSION == 4u, "layout below is ABI v4");
_Static_assert(sizeof(trb_keyframe) == 40, "trb_keyframe");
_Static_assert(sizeof(trb_spline) == 20, "trb_spline");
_Static_assert(sizeof(trb_color_key) == 20, "trb_color_key");
_Static_assert(sizeof(trb_instance) == 40, "trb_instance");
_Static_assert(sizeof(trb_mesh) == 40 && offsetof(trb_mesh, positions) == 8, "trb_mesh");
_Static_assert(sizeof(trb_material) == 56 && offsetof(trb_material, tex) == 40, "trb_material");
_Static_assert(sizeof(trb_image) == 24 && offsetof(trb_image, rgba8) == 8 && sizeof(trb_texture) == 8, "trb_image / trb_texture");
_Static_assert(sizeof(trb_camera) == 40, "trb_camera");
_Static_assert(sizeof(trb_film) == 48, "trb_film");
_Static_assert(sizeof(trb_integrator) == 12, "trb_integrator");
_Static_assert(sizeof(trb_render_cfg) == 44 && offsetof(trb_render_cfg, shard_index) == 32, "trb_render_cfg: 11 x u32");
_Static_assert(sizeof(trb_stats) == 72 && offsetof(trb_stats, kernel_ms) == 64, "trb_stats: 8 x u64 + 2 x f32");
_Static_assert(sizeof(trb_ray) == 32 && sizeof(trb_hit) == 16 && sizeof(trb_sample) == 20 && sizeof(trb_bvh_node) == 32, "ray / hit / sample / node records");
_Static_assert(offsetof(trb_scene_desc, film) == 4 && offsetof(trb_scene_desc, integrator) == 52 && offsetof(trb_scene_desc, n_cameras) == 64 &&
               offsetof(trb_scene_desc, cameras) == 72 && sizeof(trb_scene_desc) == 256, "trb_scene_desc");

#define S(T) printf(#T " sizeof %zu\n", sizeof(T))
#define F(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))

static int layout(void) {
    S(trb_keyframe); F(trb_keyframe, translation); F(trb_keyframe, rotation); F(trb_keyframe, scaling);
    S(trb_spline); F(trb_spline, degree); F(trb_spline, n_ctrl); F(trb_spline, ctrl_first); F(trb_spline, n_knots); F(trb_spline, knot_first);
    S(trb_color_key); F(trb_color_key, rgba); F(trb_color_key, time);
    S(trb_instance); F(trb_instance, kind); F(trb_instance, shape); F(trb_instance, p0); F(trb_instance, p1); F(trb_instance, mesh); F(trb_instance, material);
    F(trb_instance, spline_first); F(trb_instance, n_splines); F(trb_instance, emission_first); F(trb_instance, n_emission);
    S(trb_mesh); F(trb_mesh, n_verts); F(trb_mesh, n_tris); F(trb_mesh, positions); F(trb_mesh, normals); F(trb_mesh, texcoords); F(trb_mesh, indices);
    S(trb_material); F(trb_material, type); F(trb_material, c0); F(trb_material, c1); F(trb_material, roughness); F(trb_material, eta); F(trb_material, merl); F(trb_material, tex);
    S(trb_image); F(trb_image, width); F(trb_image, height); F(trb_image, rgba8); F(trb_image, time); F(trb_image, pad);
    S(trb_texture); F(trb_texture, first_image); F(trb_texture, n_images);
    S(trb_camera); F(trb_camera, spline_first); F(trb_camera, n_splines); F(trb_camera, fov); F(trb_camera, shutter_size); F(trb_camera, active_at);
    F(trb_camera, fov_degree); F(trb_camera, n_fov_ctrl); F(trb_camera, fov_ctrl_first); F(trb_camera, n_fov_knots); F(trb_camera, fov_knot_first);
    S(trb_film); F(trb_film, width); F(trb_film, height); F(trb_film, samples); F(trb_film, frames); F(trb_film, start_frame); F(trb_film, end_frame);
    F(trb_film, scene_time); F(trb_film, filter_type); F(trb_film, filter_w); F(trb_film, filter_h); F(trb_film, filter_b); F(trb_film, filter_c);
    S(trb_integrator); F(trb_integrator, type); F(trb_integrator, min_depth); F(trb_integrator, max_depth);
    S(trb_scene_desc); F(trb_scene_desc, abi_version); F(trb_scene_desc, film); F(trb_scene_desc, integrator);
    F(trb_scene_desc, n_cameras); F(trb_scene_desc, cameras); F(trb_scene_desc, n_instances); F(trb_scene_desc, instances); F(trb_scene_desc, n_splines);
    F(trb_scene_desc, splines); F(trb_scene_desc, n_keyframes); F(trb_scene_desc, keyframes); F(trb_scene_desc, n_knots); F(trb_scene_desc, knots);
    F(trb_scene_desc, n_color_keys); F(trb_scene_desc, color_keys); F(trb_scene_desc, n_meshes); F(trb_scene_desc, meshes); F(trb_scene_desc, n_materials);
    F(trb_scene_desc, materials); F(trb_scene_desc, n_merl); F(trb_scene_desc, merl_tables); F(trb_scene_desc, n_fov_floats); F(trb_scene_desc, fov_floats);
    F(trb_scene_desc, n_textures); F(trb_scene_desc, textures); F(trb_scene_desc, n_images); F(trb_scene_desc, images);
    S(trb_render_cfg); F(trb_render_cfg, spp); F(trb_render_cfg, sample_first); F(trb_render_cfg, sample_count); F(trb_render_cfg, block_start);
    F(trb_render_cfg, block_count); F(trb_render_cfg, current_frame); F(trb_render_cfg, seed); F(trb_render_cfg, flags); F(trb_render_cfg, shard_index);
    F(trb_render_cfg, shard_count); F(trb_render_cfg, shard_chunk);
    S(trb_stats); F(trb_stats, camera_samples); F(trb_stats, rays_primary); F(trb_stats, rays_shadow); F(trb_stats, rays_mis); F(trb_stats, rays_continuation);
    F(trb_stats, node_tests); F(trb_stats, tri_tests); F(trb_stats, inst_tests); F(trb_stats, kernel_ms); F(trb_stats, update_ms);
    S(trb_ray); F(trb_ray, o); F(trb_ray, d); F(trb_ray, min_t); F(trb_ray, max_t);
    S(trb_hit); F(trb_hit, t); F(trb_hit, inst); F(trb_hit, prim); F(trb_hit, pad);
    S(trb_sample); F(trb_sample, x); F(trb_sample, y); F(trb_sample, r); F(trb_sample, g); F(trb_sample, b);
    S(trb_bvh_node); F(trb_bvh_node, bmin); F(trb_bvh_node, bmax); F(trb_bvh_node, a); F(trb_bvh_node, b);
    printf("abi_version %u\n", trb_abi_version());
    return 0;
}

static int render(const char* path, unsigned w, unsigned h, unsigned spp) {
    trb_scene* scene = NULL;
    trb_status rc = trb_scene_load_json(path, w, h, spp, 0, &scene);
    if (rc != TRB_OK) { printf("load status %d: %s\n", (int)rc, trb_last_error()); return rc == TRB_NO_DEVICE ? 3 : 1; }
    uint32_t fw, fh, fspp, nb, ni, nl;
    if (trb_scene_info(scene, &fw, &fh, &fspp, &nb, &ni, &nl) != TRB_OK) return 1;
    const size_t npx = (size_t)fw * fh;
    float* film = (float*)calloc(npx * 4, sizeof(float));     /* RenderTarget::get_renderf32 layout, accumulated into */
    unsigned char* img = (unsigned char*)calloc(npx * 3, 1);
    trb_render_cfg cfg;
    memset(&cfg, 0, sizeof cfg);                               /* spp 0 = film.samples, sample_count 0 = all of them, block_count 0 = all blocks */
    cfg.seed = 7;
    trb_stats st;
    rc = trb_render(scene, &cfg, film, &st);                   /* == Exec::render: update_frame + every pass of the frame + film to host */
    if (rc != TRB_OK) { printf("render status %d: %s\n", (int)rc, trb_last_error()); return 1; }
    rc = trb_film_to_srgb8(scene, film, img);                  /* == RenderTarget::get_render */
    if (rc != TRB_OK) { printf("srgb status %d: %s\n", (int)rc, trb_last_error()); return 1; }
    double wsum = 0.0; unsigned long lit = 0;
    for (size_t i = 0; i < npx; ++i) { wsum += film[4 * i + 3]; lit += (img[3 * i] | img[3 * i + 1] | img[3 * i + 2]) != 0; }
    printf("rendered %ux%u spp %u blocks %u instances %u lights %u\n", fw, fh, fspp, nb, ni, nl);
    printf("camera_samples %llu rays %llu kernel_ms %.3f weight_sum %.1f lit_pixels %lu\n", (unsigned long long)st.camera_samples,
           (unsigned long long)(st.rays_primary + st.rays_shadow + st.rays_mis + st.rays_continuation), st.kernel_ms, wsum, lit);
    const int ok = st.camera_samples == (unsigned long long)npx * fspp && lit > npx / 2;
    trb_scene_destroy(scene);
    free(film); free(img);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "layout") == 0) return layout();
    if (argc >= 6 && strcmp(argv[1], "render") == 0) return render(argv[2], (unsigned)atoi(argv[3]), (unsigned)atoi(argv[4]), (unsigned)atoi(argv[5]));
    fprintf(stderr, "usage: %s layout | render scene.json width height spp\n", argv[0]);
    return 2;
}
