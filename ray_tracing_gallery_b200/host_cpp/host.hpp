// host.hpp — the C++ host side above the C ABI of libb200rt.
//
// The reference's host is compiled Rust calling Vulkan; this image has no Rust toolchain, so the host side
// that mirrors it in a compiled language is C++ (the Python package next to this directory mirrors the same
// interface for the tests and the benchmark).  Names and argument meaning follow the reference:
//   FirstPersonCamera::as_view_matrix       src/main.rs:1206-1231      -> Camera::view_inverse
//   perspective_reversed_infinite_z_vk      src/main.rs:586-593        -> Camera::proj_inverse
//   Sun::as_normal                          src/main.rs:1189-1197      -> Sun::as_normal
//   Uniforms initial values                 src/main.rs:568-598        -> make_uniforms
//   AccelerationStructureInstance::new      src/gpu_structs.rs:28-58   -> make_instance
//   built-in textures 0..3                  src/main.rs:416-460        -> push_builtin_images
//   Model::load_gltf + Model::new           src/util_structs.rs:1049-1236 -> load_model
//   DefaultScene / LoadedModelScene         src/scene.rs:35-254        -> build_scene
// A `Backend` is a set of entry points with the signatures of include/b200rt.h resolved from a shared
// library by name prefix ("rt_" for libb200rt.so).
#pragma once
#include <dlfcn.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200rt.h"
#include "gltf.hpp"

namespace b200rt_host {

// ------------------------------------------------------------------------------------------ backend
struct Backend {
    void* handle = nullptr;
    void* ctx = nullptr;
    int (*push_image)(void*, const void*, uint32_t, uint32_t, uint32_t, int, uint32_t*) = nullptr;
    int (*create_model)(void*, const RtModelDesc*, uint32_t*, uint64_t*) = nullptr;
    int (*build_tlas)(void*, const RtInstance*, uint32_t) = nullptr;
    int (*update_instances)(void*, uint32_t, uint32_t, const RtInstance*) = nullptr;
    int (*update_tlas)(void*, uint32_t) = nullptr;
    int (*render)(void*, const RtUniforms*, const RtRenderParams*, const RtFrameOutputs*) = nullptr;
    const char* (*last_error)(const void*) = nullptr;
    void (*destroy)(void*) = nullptr;
    // libb200rt only (null for other prefixes)
    int (*render_async)(void*, const RtUniforms*, const RtRenderParams*, const RtFrameOutputs*, uint32_t*) = nullptr;
    int (*wait_frame)(void*, uint32_t) = nullptr;
    int (*sync)(void*) = nullptr;
    int (*get_stats)(void*, RtStats*) = nullptr;
    int (*host_alloc)(void*, size_t, void**) = nullptr;
    int (*host_free)(void*, void*) = nullptr;

    template <typename F>
    void bind(F& fn, const std::string& name, bool required = true) {
        fn = reinterpret_cast<F>(dlsym(handle, name.c_str()));
        if (!fn && required) throw std::runtime_error("missing symbol " + name);
    }
    // Open `path` and bind the `<prefix>*` entry points; does NOT create a context.
    void open(const std::string& path, const std::string& prefix) {
        handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!handle) throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
        bind(push_image, prefix + "push_image");
        bind(create_model, prefix + "create_model");
        bind(build_tlas, prefix + "build_tlas");
        bind(update_instances, prefix + "update_instances");
        bind(update_tlas, prefix + "update_tlas");
        bind(render, prefix + "render");
        bind(last_error, prefix + "last_error");
        bind(destroy, prefix + "destroy");
        bind(render_async, prefix + "render_async", false);
        bind(wait_frame, prefix + "wait_frame", false);
        bind(sync, prefix + "sync", false);
        bind(get_stats, prefix + "get_stats", false);
        bind(host_alloc, prefix + "host_alloc", false);
        bind(host_free, prefix + "host_free", false);
    }
    // The product: libb200rt.so on CUDA device `device`.  There is no fallback: this throws without a GPU.
    void open_b200rt(const std::string& lib_path, int device) {
        open(lib_path, "rt_");
        int (*create)(int, void**) = nullptr;
        bind(create, "rt_create");
        int rc = create(device, &ctx);
        if (rc != 0) {
            const char* msg = last_error(nullptr);
            throw std::runtime_error("rt_create failed (" + std::to_string(rc) + "): " + (msg ? msg : ""));
        }
    }
    void check(int rc, const char* what) const {
        if (rc == 0) return;
        const char* msg = last_error ? last_error(ctx) : nullptr;
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (msg ? msg : ""));
    }
    void close() {
        if (ctx && destroy) destroy(ctx);
        ctx = nullptr;
        if (handle) dlclose(handle);
        handle = nullptr;
    }
};

// ------------------------------------------------------------------------------------------ matrices (column vectors, M * v), float32
struct Mat4 {
    float m[4][4];
};
inline Mat4 mat_identity() {
    Mat4 r{};
    for (int i = 0; i < 4; i++) r.m[i][i] = 1.0f;
    return r;
}
inline Mat4 mat_scale(float s) {
    Mat4 r = mat_identity();
    r.m[0][0] = r.m[1][1] = r.m[2][2] = s;
    return r;
}
inline Mat4 mat_translation(float x, float y, float z) {
    Mat4 r = mat_identity();
    r.m[0][3] = x; r.m[1][3] = y; r.m[2][3] = z;
    return r;
}
// ultraviolet `Mat4::from_rotation_y`: columns (c,0,-s), (0,1,0), (s,0,c)
inline Mat4 mat_rotation_y(float angle) {
    float s = (float)std::sin((double)angle), c = (float)std::cos((double)angle);
    Mat4 r = mat_identity();
    r.m[0][0] = c; r.m[0][2] = s; r.m[2][0] = -s; r.m[2][2] = c;
    return r;
}
inline Mat4 operator*(const Mat4& a, const Mat4& b) {
    Mat4 r{};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float acc = 0.0f;
            for (int k = 0; k < 4; k++) acc += a.m[i][k] * b.m[k][j];
            r.m[i][j] = acc;
        }
    return r;
}

enum HitShader : uint32_t { Textured = RT_HIT_TEXTURED, Mirror = RT_HIT_MIRROR, Portal = RT_HIT_PORTAL };

// src/gpu_structs.rs:28-58: rows 0..2 of the object->world matrix, row-major 3x4; custom index = model id, mask 0xFF
inline RtInstance make_instance(const Mat4& transform, uint32_t model_id, uint64_t blas_handle, uint32_t hit_shader, bool double_sided = false) {
    RtInstance r;
    std::memset(&r, 0, sizeof(r));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) r.transform[4 * i + j] = transform.m[i][j];
    r.instance_custom_index_and_mask = (model_id & 0xFFFFFFu) | (0xFFu << 24);
    uint32_t flags = double_sided ? (uint32_t)RT_INSTANCE_TRIANGLE_FACING_CULL_DISABLE : 0u;
    r.sbt_record_offset_and_flags = (hit_shader & 0xFFFFFFu) | (flags << 24);
    r.acceleration_structure_device_address = blas_handle;
    return r;
}

// ------------------------------------------------------------------------------------------ camera / sun / uniforms
struct Camera {  // FirstPersonCamera; positive pitch looks up
    float eye[3] = {0.0f, 2.0f, -5.0f};
    float pitch = 0.0f, yaw = 3.14159265358979323846f, fov_deg = 59.0f, near_plane = 0.1f;

    void view_inverse(float out_colmajor[16]) const {
        float sp = (float)std::sin((double)pitch), cp = (float)std::cos((double)pitch);
        float sy = (float)std::sin((double)yaw), cy = (float)std::cos((double)yaw);
        float xa[3] = {cy, 0.0f, -sy}, ya[3] = {sy * sp, cp, cy * sp}, za[3] = {sy * cp, -sp, cp * cy};
        // view = [xa; ya; za | -axis.eye]; its inverse: rotation transposed, translation = eye
        float inv[4][4] = {{xa[0], ya[0], za[0], 0}, {xa[1], ya[1], za[1], 0}, {xa[2], ya[2], za[2], 0}, {0, 0, 0, 1}};
        float t[3] = {-(xa[0] * eye[0] + xa[1] * eye[1] + xa[2] * eye[2]), -(ya[0] * eye[0] + ya[1] * eye[1] + ya[2] * eye[2]),
                      -(za[0] * eye[0] + za[1] * eye[1] + za[2] * eye[2])};
        for (int i = 0; i < 3; i++) inv[i][3] = -(inv[i][0] * t[0] + inv[i][1] * t[1] + inv[i][2] * t[2]);
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) out_colmajor[4 * c + r] = inv[r][c];
    }
    // inverse of ultraviolet's perspective_reversed_infinite_z_vk(fov, aspect, near)
    void proj_inverse(uint32_t width, uint32_t height, float out_colmajor[16]) const {
        float fov = (float)((double)fov_deg * 3.14159265358979323846 / 180.0);
        float t = (float)std::tan((double)(fov / 2.0f));
        float sy = 1.0f / t, sx = sy / ((float)width / (float)height);
        float m[4][4] = {};
        m[0][0] = 1.0f / sx; m[1][1] = -1.0f / sy; m[2][3] = -1.0f; m[3][2] = 1.0f / near_plane;
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) out_colmajor[4 * c + r] = m[r][c];
    }
};

struct Sun {
    float pitch = 0.5f, yaw = 1.0f;
    void as_normal(float out[3]) const {
        double p = pitch, y = yaw;
        out[0] = (float)(std::cos(p) * std::sin(y)); out[1] = (float)std::sin(p); out[2] = (float)(std::cos(p) * std::cos(y));
    }
};

// ------------------------------------------------------------------------------------------ per-tick control integration
// `KbdState`, `camera_velocity`, `sun_velocity` and the Event::MainEventsCleared block of the reference (src/main.rs:845-910).
struct KbdState {
    bool forward = false, back = false, left = false, right = false, sun_up = false, sun_down = false, sun_cw = false, sun_ccw = false;
};
struct Controls {
    float camera_velocity[3] = {0.0f, 0.0f, 0.0f};
    float sun_velocity[2] = {0.0f, 0.0f};
};
inline void integrate_controls(Camera& camera, Sun& sun, Controls& ctl, const KbdState& kbd) {
    {
        const float acc = 0.005f, vmax = 0.2f;
        float lv[3] = {0.0f, 0.0f, 0.0f};
        const float cp = (float)std::cos((double)camera.pitch), sp = (float)std::sin((double)camera.pitch);
        if (kbd.forward) { lv[2] -= acc * cp; lv[1] += acc * sp; }
        if (kbd.back) { lv[2] += acc * cp; lv[1] -= acc * sp; }
        if (kbd.left) lv[0] -= acc;
        if (kbd.right) lv[0] += acc;
        // Mat3::from_rotation_y(yaw): columns (c,0,-s), (0,1,0), (s,0,c)
        const float s = (float)std::sin((double)camera.yaw), c = (float)std::cos((double)camera.yaw);
        const float wv[3] = {c * lv[0] + s * lv[2], lv[1], -s * lv[0] + c * lv[2]};
        for (int k = 0; k < 3; k++) ctl.camera_velocity[k] += wv[k];
        float* v = ctl.camera_velocity;
        const float mag = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (mag > vmax) { const float f = std::fmin(mag, vmax) / mag; for (int k = 0; k < 3; k++) v[k] *= f; }
        for (int k = 0; k < 3; k++) { camera.eye[k] += v[k]; v[k] *= 0.9f; }
    }
    {
        const float acc = 0.002f, vmax = 0.05f, half_pi = (float)(3.14159265358979323846 / 2.0);
        float* v = ctl.sun_velocity;
        if (kbd.sun_up) v[1] += acc;
        if (kbd.sun_down) v[1] -= acc;
        if (kbd.sun_cw) v[0] += acc;
        if (kbd.sun_ccw) v[0] -= acc;
        const float mag = std::sqrt(v[0] * v[0] + v[1] * v[1]);
        if (mag > vmax) { const float f = std::fmin(mag, vmax) / mag; v[0] *= f; v[1] *= f; }
        sun.yaw -= v[0];
        sun.pitch = std::fmax(std::fmin(sun.pitch + v[1], half_pi), 0.0f);
        v[0] *= 0.95f; v[1] *= 0.95f;
    }
}
// the fixed key schedule of headless animated runs (scene.py scripted_keys)
inline KbdState scripted_keys(uint32_t tick) {
    KbdState k;
    k.forward = tick < 20; k.right = tick >= 10 && tick < 30; k.sun_cw = tick >= 20 && tick < 40; k.sun_up = tick >= 30 && tick < 45;
    k.back = tick >= 40 && tick < 50; k.left = tick >= 50 && tick < 55; k.sun_ccw = tick >= 45 && tick < 50; k.sun_down = tick >= 50 && tick < 60;
    return k;
}

inline RtUniforms make_uniforms(const Camera& cam, const Sun& sun, uint32_t width, uint32_t height, float sun_radius, uint32_t frame_index) {
    RtUniforms u;
    std::memset(&u, 0, sizeof(u));
    cam.view_inverse(u.view_inverse);
    cam.proj_inverse(width, height, u.proj_inverse);
    sun.as_normal(u.sun_dir);
    u.sun_radius = sun_radius;
    u.blue_noise_texture_index = 2;  // src/main.rs:595
    u.ggx_lut_texture_index = 3;     // src/main.rs:596
    u.frame_index = frame_index;
    return u;
}

// ------------------------------------------------------------------------------------------ seeded streams (scene.py hash_uniform)
inline double hash_uniform(uint64_t seed, uint64_t stream, uint64_t i) {
    uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + seed * 0xD1342543DE82EF95ull + stream * 0xA24BAED4963EE407ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// translate * rotation_y * scale record (src/scene.rs:139-145), float32 like the Python host
inline RtInstance instance_from_trs(const double pos[3], float rot_y, float scale, uint32_t model_id, uint64_t handle, uint32_t hit_shader) {
    float c = std::cos(rot_y), s = std::sin(rot_y);
    Mat4 m = mat_identity();
    m.m[0][0] = c * scale; m.m[0][2] = s * scale; m.m[0][3] = (float)pos[0];
    m.m[1][1] = scale; m.m[1][3] = (float)pos[1];
    m.m[2][0] = -s * scale; m.m[2][2] = c * scale; m.m[2][3] = (float)pos[2];
    RtInstance r = make_instance(m, model_id, handle, hit_shader);
    r.sbt_record_offset_and_flags = hit_shader & 0xFFFFFFu;
    return r;
}

// ------------------------------------------------------------------------------------------ assets
inline std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

struct LoadedModel {
    uint32_t id = 0;
    uint64_t blas = 0;
    ModelArrays arrays;
};

struct Host {
    Backend& be;
    std::string asset_dir;
    Host(Backend& b, std::string assets) : be(b), asset_dir(std::move(assets)) {}

    uint32_t push_image(const void* texels, uint32_t w, uint32_t h, uint32_t format, bool linear) {
        uint32_t idx = 0;
        be.check(be.push_image(be.ctx, texels, w, h, format, linear ? 1 : 0, &idx), "push_image");
        return idx;
    }
    uint32_t push_png_file(const std::string& name, uint32_t format, bool linear) {
        std::vector<uint8_t> bytes = read_file(asset_dir + "/" + name);
        ImageRgba8 im = decode_png_rgba8(bytes.data(), bytes.size());
        return push_image(im.texels.data(), im.width, im.height, format, linear);
    }
    // texture indices 0..3, src/main.rs:416-460
    void push_builtin_images() {
        uint32_t i0 = push_png_file("green.png", RT_FORMAT_RGBA8_SRGB, false);
        uint32_t i1 = push_png_file("pink.png", RT_FORMAT_RGBA8_SRGB, false);
        uint32_t i2 = push_png_file("blue_noise_64x64.png", RT_FORMAT_RGBA8_UNORM, false);
        uint32_t i3 = push_png_file("flipped_ggx_lut.png", RT_FORMAT_RGBA8_UNORM, true);
        if (i0 != 0 || i1 != 1 || i2 != 2 || i3 != 3) throw std::runtime_error("built-in images must be pushed first");
    }
    // Model::new: upload + BLAS build
    void create_model(LoadedModel& lm) {
        const ModelArrays& a = lm.arrays;
        std::vector<RtGeometryDesc> geoms(a.geometries.size());
        for (size_t g = 0; g < geoms.size(); g++) {
            std::memset(&geoms[g], 0, sizeof(RtGeometryDesc));
            geoms[g].indices = a.geometries[g].indices.data();
            geoms[g].num_indices = (uint32_t)a.geometries[g].indices.size();
            geoms[g].opaque = a.geometries[g].opaque ? 1 : 0;
            geoms[g].images.diffuse_image_index = a.geometries[g].diffuse_image_index;
            geoms[g].images.metallic_roughness_image_index = a.geometries[g].metallic_roughness_image_index;
            geoms[g].images.normal_map_image_index = a.geometries[g].normal_map_image_index;
        }
        RtModelDesc d;
        std::memset(&d, 0, sizeof(d));
        d.positions = a.positions.data(); d.normals = a.normals.data(); d.uvs = a.uvs.data();
        d.num_vertices = (uint32_t)a.num_vertices();
        d.num_geometries = (uint32_t)geoms.size();
        d.geometries = geoms.data();
        be.check(be.create_model(be.ctx, &d, &lm.id, &lm.blas), "create_model");
    }
    LoadedModel load_model(const std::string& file, uint32_t fallback_image_index) {
        LoadedModel lm;
        lm.arrays = load_gltf(read_file(asset_dir + "/" + file), file, fallback_image_index,
                              [this](const void* t, uint32_t w, uint32_t h, uint32_t f, bool l) { return push_image(t, w, h, f, l); });
        create_model(lm);
        return lm;
    }
};

// ------------------------------------------------------------------------------------------ scenes (BASELINE.json configs + DefaultScene)
struct SceneSetup {
    std::string name;
    std::vector<RtInstance> instances;
    Camera camera;
    Sun sun;
    uint32_t width = 1280, height = 720, shadow_rays = 2, max_segments = 3;
    float sun_radius = 0.05f;

    RtUniforms uniforms(uint32_t frame_index = 1) const { return make_uniforms(camera, sun, width, height, sun_radius, frame_index); }
    RtRenderParams params() const {
        RtRenderParams p;
        std::memset(&p, 0, sizeof(p));
        p.width = width; p.height = height; p.max_segments = max_segments; p.shadow_rays = shadow_rays;
        return p;
    }
};

inline SceneSetup build_scene(Host& host, const std::string& config, uint32_t width = 0, uint32_t height = 0, uint32_t num_instances = 0) {
    host.push_builtin_images();
    SceneSetup s;
    s.name = config;
    auto field = [&](uint64_t seed, uint32_t n, const LoadedModel& tori, double mirror_fraction) {
        // distribution of src/scene.rs:138-155 with the seeded stream of the Python host (_mirror_field)
        for (uint32_t i = 0; i < n; i++) {
            double pos[3] = {hash_uniform(seed, 0, i) * 20.0 - 10.0, hash_uniform(seed, 1, i) * 2.0 + 0.5, hash_uniform(seed, 2, i) * 20.0 - 10.0};
            float rot = (float)(hash_uniform(seed, 3, i) * 100.0);
            float scale = (float)(hash_uniform(seed, 4, i) * (0.1 - 0.01) + 0.01);
            uint32_t kind = hash_uniform(seed, 5, i) < mirror_fraction ? Mirror : Textured;
            s.instances.push_back(instance_from_trs(pos, rot, scale, tori.id, tori.blas, kind));
        }
    };
    if (config == "c1") {  // src/scene.rs:96-109: plane scale(10), torus translate(0,1,0); hard shadow
        LoadedModel plane = host.load_model("plane.glb", 0), tori = host.load_model("tori.glb", 1);
        s.instances = {make_instance(mat_scale(10.0f), plane.id, plane.blas, Textured), make_instance(mat_translation(0, 1, 0), tori.id, tori.blas, Textured)};
        s.width = 1280; s.height = 720; s.shadow_rays = 1; s.sun_radius = 0.0f;
    } else if (config == "c2") {  // LoadedModelScene recipe on the plane; 4 soft-shadow rays
        LoadedModel plane = host.load_model("plane.glb", 0), lain = host.load_model("lain.glb", 1);
        s.instances = {make_instance(mat_scale(10.0f), plane.id, plane.blas, Textured), make_instance(mat_identity(), lain.id, lain.blas, Textured)};
        s.camera.eye[1] = 2.5f; s.camera.eye[2] = -7.0f;
        s.width = 1920; s.height = 1080; s.shadow_rays = 4;
    } else if (config == "c3") {
        LoadedModel plane = host.load_model("plane.glb", 0), tori = host.load_model("tori.glb", 1), fence = host.load_model("fence.glb", 0);
        s.instances = {make_instance(mat_scale(10.0f), plane.id, plane.blas, Textured),
                       make_instance(mat_translation(2, 0, 2), fence.id, fence.blas, Textured, true),
                       make_instance(mat_translation(-1, 0, 1.0f) * mat_rotation_y(0.5f), fence.id, fence.blas, Textured, true),
                       make_instance(mat_translation(-3.0f, 1.25f, 3.0f) * mat_rotation_y(0.6f), tori.id, tori.blas, Mirror),
                       make_instance(mat_translation(4.5f, 1.25f, 5.0f) * mat_rotation_y(2.2f), tori.id, tori.blas, Mirror),
                       make_instance(mat_translation(0.5f, 1.25f, 7.0f) * mat_rotation_y(1.3f) * mat_scale(1.5f), tori.id, tori.blas, Mirror)};
        field(0xC0FFEE, num_instances ? num_instances : 100, tori, 1.0);
        s.width = 1920; s.height = 1080; s.shadow_rays = 2;
    } else if (config == "default") {  // src/scene.rs:35-159 with the random tori drawn from the seeded stream
        LoadedModel plane = host.load_model("plane.glb", 0), tori = host.load_model("tori.glb", 1), lain = host.load_model("lain.glb", 1),
                    fence = host.load_model("fence.glb", 0);
        Mat4 lain_base = mat_translation(-2.0f, 0.0f, -1.0f) * mat_scale(0.5f);
        float a150 = (float)(150.0 * 3.14159265358979323846 / 180.0);
        s.instances = {make_instance(mat_scale(10.0f), plane.id, plane.blas, Textured),
                       make_instance(mat_translation(0, 1, 0), tori.id, tori.blas, Textured),
                       make_instance(lain_base * mat_rotation_y(a150), lain.id, lain.blas, Textured),
                       make_instance(mat_translation(0, 1, 0), plane.id, plane.blas, Portal, true),
                       make_instance(mat_translation(2, 0, 2), fence.id, fence.blas, Textured, true)};
        field(0xD5CE, num_instances ? num_instances : 100, tori, 0.5);
        s.width = 1280; s.height = 720; s.shadow_rays = 2;
    } else {
        throw std::runtime_error("unknown config " + config + " (c1, c2, c3, default)");
    }
    if (width) s.width = width;
    if (height) s.height = height;
    host.be.check(host.be.build_tlas(host.be.ctx, s.instances.data(), (uint32_t)s.instances.size()), "build_tlas");
    return s;
}

}  // namespace b200rt_host
