// gltf.hpp — GLB -> model arrays with the semantics of the reference's `Model::load_gltf`
// (src/util_structs.rs:903-1156) and its PNG path (src/util_functions.rs:239-265):
//   * GLB only; images must be bufferView PNGs (:957-968);
//   * one geometry per *material* (:1079-1088); a model without materials gets one default geometry:
//     diffuse = the caller's fallback image index, metallic-roughness = a 1x1 RGBA32F constant
//     (1, roughness 1, metallic 0, 1), opaque (:1090-1111);
//   * images are pushed in material order: diffuse (sRGB), metallic-roughness (sRGB — a reference quirk
//     that is kept), optional normal map (UNORM) (:994-1020); a missing texture becomes a 1x1 RGBA32F
//     constant, nearest-filtered (:938-950);
//   * linear filtering iff the sampler's magFilter != NEAREST (:954-955);
//   * mesh primitives are appended in order into shared vertex arrays, indices rebased by the running
//     vertex count and appended to the geometry of their material (`unwrap_or(0)`); node transforms are
//     ignored (:1113-1137).
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200rt.h"
#include "json.hpp"
#include "png.hpp"

namespace b200rt_host {

struct Geometry {
    std::vector<uint32_t> indices;  // multiple of 3, rebased to the model's vertex arrays
    bool opaque = true;
    uint32_t diffuse_image_index = 0, metallic_roughness_image_index = 0;
    int32_t normal_map_image_index = -1;
};

struct ModelArrays {
    std::string name;
    std::vector<float> positions, normals, uvs;  // V*3, V*3, V*2
    std::vector<Geometry> geometries;
    size_t num_vertices() const { return positions.size() / 3; }
    size_t num_triangles() const {
        size_t n = 0;
        for (const auto& g : geometries) n += g.indices.size() / 3;
        return n;
    }
};

// ImageManager::push_image: texels (w*h*4 bytes, or w*h*16 for RGBA32F), format, sampler choice -> dense index
using PushImage = std::function<uint32_t(const void* texels, uint32_t w, uint32_t h, uint32_t format, bool linear)>;

namespace gltf_detail {

struct Glb {
    Json doc;
    const uint8_t* blob = nullptr;
    size_t blob_size = 0;
};

inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

inline Glb split_glb(const std::vector<uint8_t>& data) {
    if (data.size() < 12 || le32(data.data()) != 0x46546C67u) throw std::runtime_error("not a GLB container");
    size_t length = le32(data.data() + 8);
    if (length > data.size()) length = data.size();
    Glb g;
    bool have_json = false;
    size_t off = 12;
    while (off + 8 <= length) {
        size_t clen = le32(data.data() + off);
        uint32_t ctype = le32(data.data() + off + 4);
        if (off + 8 + clen > data.size()) throw std::runtime_error("GLB: truncated chunk");
        const uint8_t* chunk = data.data() + off + 8;
        if (ctype == 0x4E4F534Au) {
            g.doc = parse_json(std::string(reinterpret_cast<const char*>(chunk), clen));
            have_json = true;
        } else if (ctype == 0x004E4942u && !g.blob) {
            g.blob = chunk;
            g.blob_size = clen;
        }
        off += 8 + clen;
    }
    if (!have_json) throw std::runtime_error("GLB without JSON chunk");
    if (!g.blob) throw std::runtime_error("GLB without binary chunk");  // `gltf.blob.as_ref().unwrap()`, :1069
    return g;
}

inline int component_size(long long ct) {
    switch (ct) {
        case 5120: case 5121: return 1;
        case 5122: case 5123: return 2;
        case 5125: case 5126: return 4;
    }
    throw std::runtime_error("glTF: unknown componentType");
}
inline int type_components(const std::string& t) {
    if (t == "SCALAR") return 1;
    if (t == "VEC2") return 2;
    if (t == "VEC3") return 3;
    if (t == "VEC4") return 4;
    if (t == "MAT4") return 16;
    throw std::runtime_error("glTF: unknown accessor type");
}

// accessor -> doubles are overkill; two typed readers cover what the loader needs
struct AccessorView {
    const uint8_t* base;
    size_t count, stride;
    long long component_type;
    int ncomp;
    bool normalized;
};
inline AccessorView view_accessor(const Glb& g, long long index) {
    const Json& acc = g.doc.at("accessors").at((size_t)index);
    if (acc.has("sparse")) throw std::runtime_error("sparse accessors are not supported");
    const Json& bv = g.doc.at("bufferViews").at((size_t)acc.at("bufferView").integer());
    if (bv.integer_or("buffer", 0) != 0) throw std::runtime_error("only buffer 0 (the GLB blob) is supported");
    AccessorView v;
    v.component_type = acc.at("componentType").integer();
    v.ncomp = type_components(acc.at("type").string());
    v.count = (size_t)acc.at("count").integer();
    size_t elem = (size_t)component_size(v.component_type) * v.ncomp;
    size_t stride = (size_t)bv.integer_or("byteStride", 0);
    v.stride = stride ? stride : elem;
    size_t start = (size_t)bv.integer_or("byteOffset", 0) + (size_t)acc.integer_or("byteOffset", 0);
    if (v.count && start + (v.count - 1) * v.stride + elem > g.blob_size) throw std::runtime_error("glTF: accessor outside the blob");
    v.base = g.blob + start;
    const Json* n = acc.find("normalized");
    v.normalized = n && n->type == Json::Bool && n->b;
    return v;
}
inline double read_component(const uint8_t* p, long long ct) {
    switch (ct) {
        case 5120: { int8_t v; std::memcpy(&v, p, 1); return v; }
        case 5121: return *p;
        case 5122: { int16_t v; std::memcpy(&v, p, 2); return v; }
        case 5123: { uint16_t v; std::memcpy(&v, p, 2); return v; }
        case 5125: { uint32_t v; std::memcpy(&v, p, 4); return v; }
        default: { float v; std::memcpy(&v, p, 4); return v; }
    }
}
inline float component_max(long long ct) {
    switch (ct) {
        case 5120: return 127.f;
        case 5121: return 255.f;
        case 5122: return 32767.f;
        case 5123: return 65535.f;
        default: return 4294967295.f;
    }
}
inline void read_floats(const Glb& g, long long index, int want_ncomp, std::vector<float>& out) {
    AccessorView v = view_accessor(g, index);
    if (v.ncomp != want_ncomp) throw std::runtime_error("glTF: unexpected accessor width");
    int cs = component_size(v.component_type);
    for (size_t i = 0; i < v.count; i++)
        for (int c = 0; c < v.ncomp; c++) {
            const uint8_t* p = v.base + i * v.stride + (size_t)c * cs;
            float f = (float)read_component(p, v.component_type);
            if (v.normalized && v.component_type != 5126) f = f / component_max(v.component_type);  // gltf crate `into_f32()`
            out.push_back(f);
        }
}
inline void read_indices(const Glb& g, long long index, uint32_t rebase, std::vector<uint32_t>& out) {
    AccessorView v = view_accessor(g, index);
    for (size_t i = 0; i < v.count; i++) out.push_back((uint32_t)read_component(v.base + i * v.stride, v.component_type) + rebase);
}

}  // namespace gltf_detail

inline ModelArrays load_gltf(const std::vector<uint8_t>& glb, const std::string& name, uint32_t fallback_image_index, const PushImage& push_image) {
    using namespace gltf_detail;
    Glb g = split_glb(glb);
    const Json& js = g.doc;

    auto constant_image = [&](float r, float gg, float b, float a) {
        float texel[4] = {r, gg, b, a};
        return push_image(texel, 1, 1, RT_FORMAT_RGBA32_SFLOAT, false);
    };
    auto image_from_texture = [&](const Json* tex_info, const float* backup_rgba, uint32_t format) -> uint32_t {
        if (!tex_info) return constant_image(backup_rgba[0], backup_rgba[1], backup_rgba[2], backup_rgba[3]);
        const Json& tex = js.at("textures").at((size_t)tex_info->at("index").integer());
        bool linear = true;
        if (tex.has("sampler")) {
            const Json& smp = js.at("samplers").at((size_t)tex.at("sampler").integer());
            linear = smp.integer_or("magFilter", 0) != 9728;  // NEAREST
        }
        const Json& img = js.at("images").at((size_t)tex.at("source").integer());
        if (!img.has("bufferView")) throw std::runtime_error("Image source is a uri which we don't support");
        const Json& view = js.at("bufferViews").at((size_t)img.at("bufferView").integer());
        size_t start = (size_t)view.integer_or("byteOffset", 0), len = (size_t)view.at("byteLength").integer();
        if (start + len > g.blob_size) throw std::runtime_error("glTF: image outside the blob");
        ImageRgba8 im = decode_png_rgba8(g.blob + start, len);
        return push_image(im.texels.data(), im.width, im.height, format, linear);
    };

    ModelArrays m;
    m.name = name;
    if (const Json* mats = js.find("materials")) {
        for (const Json& material : mats->arr) {
            static const Json empty_obj = [] { Json j; j.type = Json::Obj; return j; }();
            const Json* pbrp = material.find("pbrMetallicRoughness");
            const Json& pbr = pbrp ? *pbrp : empty_obj;
            float base[4] = {1.f, 1.f, 1.f, 1.f};
            if (const Json* bf = pbr.find("baseColorFactor"))
                for (size_t k = 0; k < 4 && k < bf->size(); k++) base[k] = (float)bf->at(k).number();
            float metallic = (float)pbr.number_or("metallicFactor", 1.0), roughness = (float)pbr.number_or("roughnessFactor", 1.0);
            Geometry geo;
            geo.diffuse_image_index = image_from_texture(pbr.find("baseColorTexture"), base, RT_FORMAT_RGBA8_SRGB);
            float mr_backup[4] = {1.f, roughness, metallic, 1.f};
            geo.metallic_roughness_image_index = image_from_texture(pbr.find("metallicRoughnessTexture"), mr_backup, RT_FORMAT_RGBA8_SRGB);
            const Json* nt = material.find("normalTexture");
            geo.normal_map_image_index = nt ? (int32_t)image_from_texture(nt, nullptr, RT_FORMAT_RGBA8_UNORM) : -1;
            geo.opaque = material.string_or("alphaMode", "OPAQUE") == "OPAQUE";
            m.geometries.push_back(std::move(geo));
        }
    }
    if (m.geometries.empty()) {
        Geometry geo;
        geo.metallic_roughness_image_index = constant_image(1.f, 1.f, 0.f, 1.f);
        geo.diffuse_image_index = fallback_image_index;
        m.geometries.push_back(std::move(geo));
    }
    uint32_t num_vertices = 0;
    if (const Json* meshes = js.find("meshes")) {
        for (const Json& mesh : meshes->arr)
            for (const Json& prim : mesh.at("primitives").arr) {
                size_t gi = (size_t)prim.integer_or("material", 0);
                if (gi >= m.geometries.size()) throw std::runtime_error("glTF: primitive refers to a missing material");
                const Json& attrs = prim.at("attributes");
                read_indices(g, prim.at("indices").integer(), num_vertices, m.geometries[gi].indices);
                size_t before = m.positions.size() / 3;
                read_floats(g, attrs.at("POSITION").integer(), 3, m.positions);
                read_floats(g, attrs.at("NORMAL").integer(), 3, m.normals);
                read_floats(g, attrs.at("TEXCOORD_0").integer(), 2, m.uvs);
                num_vertices += (uint32_t)(m.positions.size() / 3 - before);
            }
    }
    return m;
}

}  // namespace b200rt_host
