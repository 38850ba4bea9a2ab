// host_dump — CPU-only tool for the tests: runs the C++ loader / scene recipes against a recording
// backend (no GPU, no library) and writes what they produced as tagged binary sections, so that
// tests/test_host_cpp.py can compare it with the Python host byte for byte.
//   host_dump model <file.glb> <fallback_index> <first_image_index> <out.bin>
//   host_dump scene <config> <assets_dir> <out.bin> [width height]
//   host_dump png <file.png> <out.bin>
//   host_dump controls <ticks> <out.bin>      (tag 6000: per tick {eye[3], pitch, yaw, sun pitch, sun yaw} as 7 floats)
// Section = u32 tag, u64 byte count, payload.  Tags: 1 positions, 2 normals, 3 uvs, 100+g geometry header
// {opaque, diffuse, metal-rough, normal map}, 200+g indices, 1000+i image header {w, h, format, linear}, 2000+i texels,
// 5000 instances (64 B each), 5001 uniforms (176 B), 5002 {width, height, shadow_rays, max_segments}.
#include <cstdio>
#include <cstdlib>

#include "host.hpp"

using namespace b200rt_host;

static FILE* g_out = nullptr;
static void section(uint32_t tag, const void* data, uint64_t bytes) {
    std::fwrite(&tag, 4, 1, g_out);
    std::fwrite(&bytes, 8, 1, g_out);
    if (bytes) std::fwrite(data, 1, bytes, g_out);
}

// recording backend: indices are handed out densely, models get handle 1000 + id
static uint32_t g_images = 0, g_models = 0;
static std::vector<RtInstance> g_instances;
static int rec_push_image(void*, const void* texels, uint32_t w, uint32_t h, uint32_t format, int linear, uint32_t* out) {
    uint32_t hdr[4] = {w, h, format, (uint32_t)(linear != 0)};
    section(1000 + g_images, hdr, sizeof(hdr));
    section(2000 + g_images, texels, (uint64_t)w * h * (format == RT_FORMAT_RGBA32_SFLOAT ? 16 : 4));
    *out = g_images++;
    return 0;
}
static int rec_create_model(void*, const RtModelDesc*, uint32_t* id, uint64_t* handle) {
    *id = g_models;
    *handle = 1000 + g_models;
    g_models++;
    return 0;
}
static int rec_build_tlas(void*, const RtInstance* inst, uint32_t n) {
    g_instances.assign(inst, inst + n);
    return 0;
}
static const char* rec_last_error(const void*) { return ""; }

int main(int argc, char** argv) {
    try {
        if (argc >= 6 && std::string(argv[1]) == "model") {
            g_out = std::fopen(argv[5], "wb");
            g_images = (uint32_t)std::atoi(argv[4]);
            uint32_t first = g_images;
            (void)first;
            ModelArrays m = load_gltf(read_file(argv[2]), argv[2], (uint32_t)std::atoi(argv[3]),
                                      [](const void* t, uint32_t w, uint32_t h, uint32_t f, bool l) {
                                          uint32_t idx = 0;
                                          rec_push_image(nullptr, t, w, h, f, l, &idx);
                                          return idx;
                                      });
            section(1, m.positions.data(), m.positions.size() * 4);
            section(2, m.normals.data(), m.normals.size() * 4);
            section(3, m.uvs.data(), m.uvs.size() * 4);
            for (size_t g = 0; g < m.geometries.size(); g++) {
                const Geometry& ge = m.geometries[g];
                int32_t hdr[4] = {ge.opaque ? 1 : 0, (int32_t)ge.diffuse_image_index, (int32_t)ge.metallic_roughness_image_index, ge.normal_map_image_index};
                section(100 + (uint32_t)g, hdr, sizeof(hdr));
                section(200 + (uint32_t)g, ge.indices.data(), ge.indices.size() * 4);
            }
        } else if (argc >= 4 && std::string(argv[1]) == "png") {
            g_out = std::fopen(argv[3], "wb");
            std::vector<uint8_t> bytes = read_file(argv[2]);
            ImageRgba8 im = decode_png_rgba8(bytes.data(), bytes.size());
            uint32_t hdr[4] = {im.width, im.height, 0, 0};
            section(1000, hdr, sizeof(hdr));
            section(2000, im.texels.data(), im.texels.size());
        } else if (argc >= 4 && std::string(argv[1]) == "controls") {
            g_out = std::fopen(argv[3], "wb");
            Camera cam;
            Sun sun;
            Controls ctl;
            std::vector<float> rows;
            for (uint32_t t = 0; t < (uint32_t)std::atoi(argv[2]); t++) {
                integrate_controls(cam, sun, ctl, scripted_keys(t));
                const float r[7] = {cam.eye[0], cam.eye[1], cam.eye[2], cam.pitch, cam.yaw, sun.pitch, sun.yaw};
                rows.insert(rows.end(), r, r + 7);
            }
            section(6000, rows.data(), rows.size() * 4);
        } else if (argc >= 5 && std::string(argv[1]) == "scene") {
            g_out = std::fopen(argv[4], "wb");
            Backend be;
            be.push_image = rec_push_image;
            be.create_model = rec_create_model;
            be.build_tlas = rec_build_tlas;
            be.last_error = rec_last_error;
            Host host(be, argv[3]);
            uint32_t w = argc > 5 ? (uint32_t)std::atoi(argv[5]) : 0, h = argc > 6 ? (uint32_t)std::atoi(argv[6]) : 0;
            SceneSetup s = build_scene(host, argv[2], w, h);
            section(5000, g_instances.data(), g_instances.size() * sizeof(RtInstance));
            RtUniforms u = s.uniforms(1);
            section(5001, &u, sizeof(u));
            uint32_t hdr[4] = {s.width, s.height, s.shadow_rays, s.max_segments};
            section(5002, hdr, sizeof(hdr));
        } else {
            std::fprintf(stderr, "usage: host_dump model <glb> <fallback> <first_image> <out> | scene <config> <assets> <out> [w h]\n");
            return 2;
        }
        std::fclose(g_out);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "host_dump: %s\n", e.what());
        return 1;
    }
    return 0;
}
