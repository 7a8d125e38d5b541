// The reference's `tray_rust scene.json` single-node path (src/main.rs:56-109) against the C++ host mirror:
//   g++ -std=c++17 -Iinclude examples/render_cornell.cpp -Ltray_rust_b200/lib -ltrb -Wl,-rpath,$PWD/tray_rust_b200/lib -o /tmp/render_cornell
//   /tmp/render_cornell tests/golden/scenes/c1_cornell_box.json 400 400 64
#include <cstdio>
#include <cstdlib>
#include "tray_exec.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s scene.json [width height spp]\n", argv[0]); return 2; }
    try {
        const uint32_t w = argc > 2 ? std::atoi(argv[2]) : 0, h = argc > 3 ? std::atoi(argv[3]) : 0, spp = argc > 4 ? std::atoi(argv[4]) : 0;
        tray::Scene scene = tray::Scene::load_file(argv[1], 0, w, h, spp);
        tray::RenderTarget rt = scene.make_render_target();
        tray::Config config;
        config.scene_file = argv[1];
        tray::B200 exec;
        exec.render(scene, rt, config);
        const trb_stats& st = exec.last_stats;
        const double rays = double(st.rays_primary + st.rays_shadow + st.rays_mis + st.rays_continuation);
        std::printf("Frame 0: rendering took %.4fs (%.1f Mrays/s, %llu camera samples)\n", st.kernel_ms * 1e-3, rays / st.kernel_ms / 1e3,
                    (unsigned long long)st.camera_samples);
        std::vector<uint8_t> img = tray::get_render(scene, rt);
        auto d = rt.dimensions();
        FILE* f = std::fopen("frame00000.ppm", "wb");
        std::fprintf(f, "P6\n%zu %zu\n255\n", d.first, d.second);
        std::fwrite(img.data(), 1, img.size(), f);
        std::fclose(f);
    } catch (const tray::Error& e) { std::fprintf(stderr, "error (status %d): %s\n", (int)e.status, e.what()); return 1; }
    return 0;
}
