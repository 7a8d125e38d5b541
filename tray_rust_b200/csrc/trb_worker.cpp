// trb_worker — a wire-compatible replacement for `tray_rust --worker` that renders on a B200 through the C ABI.
//
// It speaks the reference master's protocol unchanged, so an unmodified `tray_rust scene.json --master host...` can drive it:
//   * listens on exec::distrib::worker::PORT = 63234 (src/exec/distrib/worker.rs:16), accepts ONE connection (worker.rs:60-89);
//   * reads `Instructions` (src/exec/distrib/mod.rs:51-72): bincode 0.x "Infinite" encoding = little-endian, u64 lengths,
//     usize as u64, tuples and structs inline:   encoded_size u64 | scene: u64 len + utf-8 | frames (u64, u64) | block_start u64 | block_count u64
//   * Scene::load_file(instructions.scene), then for every frame of the inclusive range: Exec::render with
//     select_blocks = (block_start, block_count) (worker.rs:37-50, main.rs:148-166) == trb_render, and sends a `Frame`
//     (mod.rs:76-100):   encoded_size u64 | frame u64 | block_size (u64, u64) | blocks: u64 n + n x (u64, u64) | pixels: u64 n + n x f32
//     holding RenderTarget::get_rendered_blocks (film/render_target.rs:215-241): the 2x2 lock blocks whose four weights are all
//     non-zero, row-major over the block grid, 16 floats RGBW per block; then clears the film (main.rs:161).
//   * exits after the last frame, like the reference worker.
//
//   trb_worker [--port P] [--device D] [--seed S] [--spp N] [--once-host-check]
// No CPU fallback: without a CUDA device the scene load fails with TRB_NO_DEVICE and the worker exits non-zero.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/trb.h"

namespace {

bool read_all(int fd, void* dst, size_t n) {
    uint8_t* p = static_cast<uint8_t*>(dst);
    while (n) { const ssize_t r = ::read(fd, p, n); if (r <= 0) return false; p += r; n -= (size_t)r; }
    return true;
}
bool write_all(int fd, const void* src, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(src);
    while (n) { const ssize_t r = ::write(fd, p, n); if (r <= 0) return false; p += r; n -= (size_t)r; }
    return true;
}
uint64_t get_u64(const std::vector<uint8_t>& b, size_t& o) { uint64_t v = 0; if (o + 8 <= b.size()) std::memcpy(&v, &b[o], 8); o += 8; return v; } // x86-64: little-endian
void put_u64(std::vector<uint8_t>& b, uint64_t v) { const size_t o = b.size(); b.resize(o + 8); std::memcpy(&b[o], &v, 8); }

struct Instructions { uint64_t encoded_size = 0; std::string scene; uint64_t frame_start = 0, frame_end = 0, block_start = 0, block_count = 0; };

bool decode_instructions(const std::vector<uint8_t>& buf, Instructions& in) {
    size_t o = 0;
    in.encoded_size = get_u64(buf, o);
    const uint64_t len = get_u64(buf, o);
    if (o + len + 32 > buf.size()) return false;
    in.scene.assign(reinterpret_cast<const char*>(&buf[o]), (size_t)len); o += (size_t)len;
    in.frame_start = get_u64(buf, o); in.frame_end = get_u64(buf, o);
    in.block_start = get_u64(buf, o); in.block_count = get_u64(buf, o);
    return o == buf.size();
}

// RenderTarget::get_rendered_blocks + Frame::new + bincode::serialize
std::vector<uint8_t> encode_frame(uint64_t frame, const float* film, uint32_t w, uint32_t h) {
    std::vector<uint64_t> blocks;
    std::vector<float> pixels;
    for (uint32_t by = 0; by < h / 2; ++by)
        for (uint32_t bx = 0; bx < w / 2; ++bx) {
            bool all = true;
            for (uint32_t y = 0; y < 2; ++y) for (uint32_t x = 0; x < 2; ++x) all = all && film[4 * ((size_t)(2 * by + y) * w + 2 * bx + x) + 3] != 0.0f;
            if (!all) continue;
            blocks.push_back(2 * bx); blocks.push_back(2 * by);
            for (uint32_t y = 0; y < 2; ++y) for (uint32_t x = 0; x < 2; ++x) { const float* c = &film[4 * ((size_t)(2 * by + y) * w + 2 * bx + x)]; pixels.insert(pixels.end(), c, c + 4); }
        }
    std::vector<uint8_t> out;
    const uint64_t size = 8 + 8 + 16 + 8 + 8 * blocks.size() + 8 + 4 * pixels.size(); // == bincode::serialized_size(&frame) with encoded_size in place
    put_u64(out, size); put_u64(out, frame); put_u64(out, 2); put_u64(out, 2);
    put_u64(out, blocks.size() / 2);
    for (uint64_t v : blocks) put_u64(out, v);
    put_u64(out, pixels.size());
    const size_t o = out.size();
    out.resize(o + 4 * pixels.size());
    if (!pixels.empty()) std::memcpy(&out[o], pixels.data(), 4 * pixels.size());
    return out;
}

} // namespace

int main(int argc, char** argv) {
    int port = 63234, device = 0; // worker.rs:16
    uint32_t seed = 1, spp = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--port" && i + 1 < argc) port = std::atoi(argv[++i]);
        else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "--seed" && i + 1 < argc) seed = (uint32_t)std::strtoul(argv[++i], nullptr, 0);
        else if (a == "--spp" && i + 1 < argc) spp = (uint32_t)std::strtoul(argv[++i], nullptr, 0);
        else if (a == "--worker" || a == "-n") { if (a == "-n") ++i; } // accepted for command-line compatibility with `tray_rust --worker [-n threads]`
        else { std::fprintf(stderr, "usage: %s [--worker] [--port P] [--device D] [--seed S] [--spp N]\n", argv[0]); return 2; }
    }
    const int lfd = ::socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    ::setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in addr{};
    addr.sin_family = AF_INET; addr.sin_addr.s_addr = htonl(INADDR_ANY); addr.sin_port = htons((uint16_t)port);
    if (lfd < 0 || ::bind(lfd, reinterpret_cast<sockaddr*>(&addr), sizeof addr) != 0 || ::listen(lfd, 1) != 0) { std::perror("Worker failed to get port"); return 1; }
    std::printf("Worker listening for master on %d\n", port); std::fflush(stdout);
    const int fd = ::accept(lfd, nullptr, nullptr);
    if (fd < 0) { std::perror("Error accepting"); return 1; }
    std::vector<uint8_t> buf(8);
    if (!read_all(fd, buf.data(), 8)) { std::fprintf(stderr, "Failed to read from master\n"); return 1; }
    uint64_t expected = 0; std::memcpy(&expected, buf.data(), 8);
    if (expected < 48 || expected > (1u << 20)) { std::fprintf(stderr, "implausible instruction size %llu\n", (unsigned long long)expected); return 1; }
    buf.resize((size_t)expected);
    if (!read_all(fd, buf.data() + 8, (size_t)expected - 8)) { std::fprintf(stderr, "Failed to read from master\n"); return 1; }
    Instructions in;
    if (!decode_instructions(buf, in)) { std::fprintf(stderr, "malformed instructions\n"); return 1; }
    std::printf("Received instructions: Instructions { encoded_size: %llu, scene: \"%s\", frames: (%llu, %llu), block_start: %llu, block_count: %llu }\n",
                (unsigned long long)in.encoded_size, in.scene.c_str(), (unsigned long long)in.frame_start, (unsigned long long)in.frame_end,
                (unsigned long long)in.block_start, (unsigned long long)in.block_count);
    trb_scene* scene = nullptr;
    trb_status rc = trb_scene_load_json(in.scene.c_str(), 0, 0, spp, device, &scene); // Scene::load_file(&instructions.scene) (worker.rs:39)
    if (rc != TRB_OK) { std::fprintf(stderr, "trb_scene_load_json status %d: %s\n", (int)rc, trb_last_error()); return rc == TRB_NO_DEVICE ? 3 : 1; }
    uint32_t w = 0, h = 0;
    trb_scene_info(scene, &w, &h, nullptr, nullptr, nullptr, nullptr);
    std::vector<float> film((size_t)w * h * 4);
    for (uint64_t frame = in.frame_start; frame <= in.frame_end; ++frame) { // main.rs:157-163
        std::fill(film.begin(), film.end(), 0.0f);                            // render_target.clear()
        trb_render_cfg cfg{};
        cfg.block_start = (uint32_t)in.block_start; cfg.block_count = (uint32_t)in.block_count; cfg.current_frame = (uint32_t)frame; cfg.seed = seed;
        trb_stats st{};
        rc = trb_render(scene, &cfg, film.data(), &st);
        if (rc != TRB_OK) { std::fprintf(stderr, "trb_render status %d: %s\n", (int)rc, trb_last_error()); return 1; }
        const std::vector<uint8_t> bytes = encode_frame(frame, film.data(), w, h);
        if (!write_all(fd, bytes.data(), bytes.size())) { std::fprintf(stderr, "Failed to send frame to the master\n"); return 1; }
        std::printf("Frame %llu: rendering took %.4fs\n--------------------\n", (unsigned long long)frame, st.kernel_ms * 1e-3);
        std::fflush(stdout);
    }
    trb_scene_destroy(scene);
    ::close(fd); ::close(lfd);
    return 0;
}
