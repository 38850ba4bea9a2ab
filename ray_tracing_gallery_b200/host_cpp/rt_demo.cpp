// rt_demo — headless frame loop of the reference on libb200rt, driven from C++ (the reference's
// `main.rs` loop without the window: load assets, build the scene, render frames with two in flight).
//   rt_demo [--config c1|c2|c3|default] [--width W] [--height H] [--frames N] [--device D]
//           [--lib path/to/libb200rt.so] [--assets dir] [--out frame.ppm] [--heatmap] [--animate]
// --animate: every frame is one tick of the reference's loop (src/main.rs:845-948, src/scene.rs:162-204): the camera and
// sun velocities are integrated from a fixed key schedule, lain's rotation advances by 0.05 and its ONE 64-byte
// instance record is rewritten (default scene), the TLAS is refitted in place, frame_index += 1, two frames in flight.
// Prints one JSON line (rays, ms/frame, Mrays/s).  No CPU fallback: fails without a CUDA device.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "host.hpp"

using namespace b200rt_host;

static std::string dir_of(const std::string& path) {
    size_t p = path.find_last_of('/');
    return p == std::string::npos ? "." : path.substr(0, p);
}

int main(int argc, char** argv) {
    std::string self = dir_of(argv[0]);
    std::string lib = self + "/../csrc/libb200rt.so", assets = self + "/../../assets", config = "c2", out;
    uint32_t width = 0, height = 0, frames = 20;
    int device = 0;
    bool heatmap = false;  // the `H` key of the reference (src/main.rs:820)
    bool animate = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); }
            return argv[++i];
        };
        if (a == "--config") config = next();
        else if (a == "--width") width = (uint32_t)std::atoi(next().c_str());
        else if (a == "--height") height = (uint32_t)std::atoi(next().c_str());
        else if (a == "--frames") frames = (uint32_t)std::atoi(next().c_str());
        else if (a == "--device") device = std::atoi(next().c_str());
        else if (a == "--lib") lib = next();
        else if (a == "--assets") assets = next();
        else if (a == "--out") out = next();
        else if (a == "--heatmap") heatmap = true;
        else if (a == "--animate") animate = true;
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    Backend be;
    try {
        be.open_b200rt(lib, device);
        Host host(be, assets);
        SceneSetup s = build_scene(host, config, width, height);
        RtRenderParams p = s.params();
        size_t bytes = (size_t)s.width * s.height * 4;
        // page-locked read-back buffers (the reference's host-visible buffers): the copies are asynchronous
        uint8_t* fb[2] = {nullptr, nullptr};
        uint64_t* counts[2] = {nullptr, nullptr};
        for (int b = 0; b < 2; b++) {
            be.check(be.host_alloc(be.ctx, bytes, reinterpret_cast<void**>(&fb[b])), "host_alloc");
            be.check(be.host_alloc(be.ctx, 16, reinterpret_cast<void**>(&counts[b])), "host_alloc");
            counts[b][0] = counts[b][1] = 0;
        }
        uint64_t rays = 0;
        Controls controls;
        // warm-up frame (blocking), then `frames` frames with two in flight (src/main.rs:917-928)
        RtUniforms u0 = s.uniforms(1);
        RtFrameOutputs o0 = {fb[0], nullptr, nullptr, counts[0]};
        be.check(be.render(be.ctx, &u0, &p, &o0), "render");
        auto t0 = std::chrono::steady_clock::now();
        uint32_t slot_of[2] = {0, 0};
        for (uint32_t k = 0; k < frames; k++) {
            uint32_t b = k & 1;
            if (k >= 2) {
                be.check(be.wait_frame(be.ctx, slot_of[b]), "wait_frame");
                rays += counts[b][0] + counts[b][1];
            }
            if (animate) {
                integrate_controls(s.camera, s.sun, controls, scripted_keys(k));  // Event::MainEventsCleared
                if (config == "default") {  // DefaultScene::update + write_resources: lain's record only, then UPDATE in place
                    const float a150 = (float)(150.0 * 3.14159265358979323846 / 180.0);
                    RtInstance lain = make_instance(mat_translation(-2.0f, 0.0f, -1.0f) * mat_scale(0.5f) * mat_rotation_y(a150 + 0.05f * (float)(k + 1)), 0, 0, 0);
                    RtInstance rec = s.instances[2];
                    std::memcpy(rec.transform, lain.transform, sizeof(rec.transform));
                    be.check(be.update_instances(be.ctx, 2, 1, &rec), "update_instances");
                    be.check(be.update_tlas(be.ctx, RT_UPDATE_REFIT), "update_tlas");
                }
            }
            RtUniforms u = s.uniforms(1 + k);
            u.show_heatmap = heatmap ? 1 : 0;
            RtFrameOutputs o = {fb[b], nullptr, nullptr, counts[b]};
            be.check(be.render_async(be.ctx, &u, &p, &o, &slot_of[b]), "render_async");
        }
        for (uint32_t k = frames >= 2 ? frames - 2 : 0; k < frames; k++) {
            uint32_t b = k & 1;
            be.check(be.wait_frame(be.ctx, slot_of[b]), "wait_frame");
            rays += counts[b][0] + counts[b][1];
        }
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("{\"host\": \"c++\", \"config\": \"%s\", \"resolution\": \"%ux%u\", \"frames\": %u, \"animated\": %s, \"rays\": %llu, \"ms_per_frame\": %.4f, "
                    "\"mrays_per_s\": %.1f, \"eye\": [%.4f, %.4f, %.4f], \"sun\": [%.4f, %.4f]}\n",
                    config.c_str(), s.width, s.height, frames, animate ? "true" : "false", (unsigned long long)rays, sec / frames * 1e3, rays / sec / 1e6,
                    s.camera.eye[0], s.camera.eye[1], s.camera.eye[2], s.sun.pitch, s.sun.yaw);
        if (!out.empty()) {
            const uint8_t* last = fb[(frames - 1) & 1];
            FILE* f = std::fopen(out.c_str(), "wb");
            if (!f) throw std::runtime_error("cannot write " + out);
            std::fprintf(f, "P6\n%u %u\n255\n", s.width, s.height);
            for (size_t i = 0; i < (size_t)s.width * s.height; i++) std::fwrite(&last[4 * i], 1, 3, f);
            std::fclose(f);
        }
        for (int b = 0; b < 2; b++) { be.host_free(be.ctx, fb[b]); be.host_free(be.ctx, counts[b]); }
        be.close();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rt_demo: %s\n", e.what());
        return 1;
    }
    return 0;
}
