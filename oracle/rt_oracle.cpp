// rt_oracle.cpp — CPU ORACLE for the frame hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the checker, never the product: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  libb200rt.so
// does not link, include or call anything in oracle/.
//
// PARITY STATUS: shading arithmetic PINNED to the reference's own compiled shaders; traversal pinned by float64
// brute force (the reference holds nothing for it).  The reference (Rust + Vulkan KHR ray tracing) cannot be built or
// run here (no rustc/cargo, no Vulkan loader/ICD, no RT-capable device), it ships no golden images, and its single
// unit test (shaders/ray-tracing/src/pbr.rs:52-69) only asserts finiteness.  What it does ship is its arithmetic as
// SPIR-V: shaders/*.spv, all seven stages.  tests/spirv_interp.py EXECUTES those modules on the CPU
// (tests/spirv_pipeline.py plays the driver: intersections, texture sampling and buffer loads are callbacks) and
// tests/golden/make_spirv_golden.py stores their outputs for ~6 000 pixels of five scenes, seeded any-hit candidates
// and heat-map deltas (tests/golden/spirv_*.npz).  tests/test_spirv_pin.py holds this oracle to those vectors at
// <= 1e-5 relative (payload colour, sRGB store, hit IDs, trace-call counts, any-hit decisions) and the CUDA path to the
// same vectors at the north star's 1e-3.  BVH build, traversal, ray/triangle intersection and texture filtering live
// inside the Vulkan driver, not under /root/reference: for those the oracle restates the VK_KHR_ray_tracing_pipeline
// semantics and is pinned by independent numpy float64 brute-force frames (tests/test_oracle.py), BVH-vs-brute-force
// equality, the KAT above, the constants of the shipped .spv files and the struct layouts.
//
// What follows what (reference paths relative to /root/reference):
//   ray_generation          shaders/ray-tracing/src/lib.rs:94-191
//   linear_to_srgb          shaders/ray-tracing/src/lib.rs:85-92
//   primary_ray_miss        shaders/ray-tracing/src/lib.rs:40-51
//   shadow_ray_miss         shaders/ray-tracing/src/lib.rs:33-36
//   closest_hit_portal      shaders/ray-tracing/src/lib.rs:300-312
//   heatmap_temperature     shaders/ray-tracing/src/heatmap.rs:5-54 (+ lib.rs:120-124, 174-186)
//   closest_hit_textured    shaders/closest_hit_textured.glsl:13-226
//   hit_shader_common       shaders/hit_shader_common.glsl:75-165
//   brdf & friends          shaders/pbr.glsl:25-103, 174-211
//   closest_hit_mirror      shaders/closest_hit_mirror.glsl:11-29
//   any_hit_alpha_clip      shaders/any_hit_alpha_clip.glsl:11-28
//   hit-group selection     src/main.rs:289-305, 360-384
//   texture table/samplers  src/util_structs.rs:1296-1377, src/util_functions.rs:216-265
//   instance record         src/gpu_structs.rs:20-58
//
// Arithmetic contract (DESIGN.md "Arithmetic contract"): Vulkan leaves fp32
// evaluation order and fusing implementation-defined.  To make hit IDs
// comparable bit-for-bit with the CUDA path, the *geometric* functions below
// fix one evaluation order (explicit fused multiply-adds where written, every
// other operation rounded individually; build with -ffp-contract=off).  The
// CUDA kernels were written separately against the same written contract.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -mfma -fopenmp -shared).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../include/b200rt.h"
#include <atomic>
#include <thread>

namespace {

// ------------------------------------------------------------------ contract math
struct V3 { float x, y, z; };
struct V2 { float x, y; };

inline float fma_(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 sub3(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 add3(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 scale3(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
// dot(a,b) := fma(az,bz, fma(ay,by, ax*bx))
inline float dot3(V3 a, V3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
// cross(a,b).x := fma(ay,bz, -(az*by)) ...
inline V3 cross3(V3 a, V3 b) {
    return v3(fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x)));
}
// normalize(v) := v / sqrt(dot(v,v))  (IEEE sqrt and divide)
inline V3 normalize3(V3 v) {
    float l = std::sqrt(dot3(v, v));
    return v3(v.x / l, v.y / l, v.z / l);
}
// rows of a 3x4: r_i = fma(m[i][2],z, fma(m[i][1],y, fma(m[i][0],x, m[i][3])))
inline V3 xform_point(const float* m, V3 p) {
    return v3(fma_(m[2], p.z, fma_(m[1], p.y, fma_(m[0], p.x, m[3]))),
              fma_(m[6], p.z, fma_(m[5], p.y, fma_(m[4], p.x, m[7]))),
              fma_(m[10], p.z, fma_(m[9], p.y, fma_(m[8], p.x, m[11]))));
}
inline V3 xform_vec(const float* m, V3 v) {
    return v3(fma_(m[2], v.z, fma_(m[1], v.y, m[0] * v.x)),
              fma_(m[6], v.z, fma_(m[5], v.y, m[4] * v.x)),
              fma_(m[10], v.z, fma_(m[9], v.y, m[8] * v.x)));
}
// mat3(gl_WorldToObject3x4EXT) * n  ==  transpose(inverse 3x3) * n
inline V3 xform_normal(const float* inv, V3 n) {
    return v3(fma_(inv[8], n.z, fma_(inv[4], n.y, inv[0] * n.x)),
              fma_(inv[9], n.z, fma_(inv[5], n.y, inv[1] * n.x)),
              fma_(inv[10], n.z, fma_(inv[6], n.y, inv[2] * n.x)));
}
// a*bx + b*by + c*bz := fma(c,bz, fma(b,by, a*bx))       hit_shader_common.glsl:79-81
inline float interp1(float a, float b, float c, V3 w) { return fma_(c, w.z, fma_(b, w.y, a * w.x)); }
inline V3 interp3(V3 a, V3 b, V3 c, V3 w) {
    return v3(interp1(a.x, b.x, c.x, w), interp1(a.y, b.y, c.y, w), interp1(a.z, b.z, c.z, w));
}
inline V2 interp2(V2 a, V2 b, V2 c, V3 w) { return V2{interp1(a.x, b.x, c.x, w), interp1(a.y, b.y, c.y, w)}; }

// Inverse of the instance's 3x4 (object->world) => world->object 3x4.  Unfused,
// left to right.  inv3 = adj / det;  inv_t = -(inv3 * t).
void invert_3x4(const float* m, float* o) {
    float a00 = m[0], a01 = m[1], a02 = m[2], tx = m[3];
    float a10 = m[4], a11 = m[5], a12 = m[6], ty = m[7];
    float a20 = m[8], a21 = m[9], a22 = m[10], tz = m[11];
    float c00 = a11 * a22 - a12 * a21;
    float c01 = a12 * a20 - a10 * a22;
    float c02 = a10 * a21 - a11 * a20;
    float det = (a00 * c00 + a01 * c01) + a02 * c02;
    float id = 1.0f / det;
    o[0] = c00 * id;
    o[1] = (a02 * a21 - a01 * a22) * id;
    o[2] = (a01 * a12 - a02 * a11) * id;
    o[4] = c01 * id;
    o[5] = (a00 * a22 - a02 * a20) * id;
    o[6] = (a02 * a10 - a00 * a12) * id;
    o[8] = c02 * id;
    o[9] = (a01 * a20 - a00 * a21) * id;
    o[10] = (a00 * a11 - a01 * a10) * id;
    o[3] = -((o[0] * tx + o[1] * ty) + o[2] * tz);
    o[7] = -((o[4] * tx + o[5] * ty) + o[6] * tz);
    o[11] = -((o[8] * tx + o[9] * ty) + o[10] * tz);
}

// Ray/triangle candidate test in object space (Moller-Trumbore, no culling).
// The triangle is stored as v0, e1 = v1 - v0, e2 = v2 - v0.  Returns true and
// (t,u,v) when the ray line crosses the triangle; the caller applies the
// exclusive interval tmin < t < tmax (VK_KHR_ray_tracing_pipeline).
inline bool tri_candidate(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float& t, float& u, float& v) {
    V3 p = cross3(d, e2);
    float det = dot3(e1, p);
    if (!(det != 0.0f)) return false;  // also rejects NaN
    float inv = 1.0f / det;
    V3 tv = sub3(o, v0);
    u = dot3(tv, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    V3 q = cross3(tv, e1);
    v = dot3(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot3(e2, q) * inv;
    return true;
}

// ------------------------------------------------------------------ scene data
struct Texture {
    uint32_t w = 0, h = 0, format = 0;
    bool linear = false;
    std::vector<uint8_t> rgba8;
    std::vector<float> rgba32f;
};

struct Tri {
    V3 v0, e1, e2;
    uint32_t geom, prim;
    bool opaque;
};

struct BvhNode {  // BVH2: count>0 => leaf [a, a+count), else children a, a+1
    float lo[3], hi[3];
    uint32_t a, count;
};

struct Geometry {
    std::vector<uint32_t> indices;
    bool opaque;
    RtGeometryImages images;
};

struct Model {
    std::vector<V3> positions, normals;
    std::vector<V2> uvs;
    std::vector<Geometry> geoms;
    std::vector<Tri> tris;  // BVH leaf order
    std::vector<BvhNode> nodes;
};

struct Inst {
    RtInstance rec;
    float inv[12];
    float lo[3], hi[3];
    int32_t blas;  // model index or -1
};

float g_srgb_lut[256];
bool g_lut_ready = false;
void init_lut() {
    if (g_lut_ready) return;
    for (int i = 0; i < 256; i++) {
        float c = (float)i / 255.0f;
        g_srgb_lut[i] = c <= 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f);
    }
    g_lut_ready = true;
}

}  // namespace

struct OrcContext {
    std::vector<Texture> textures;
    std::vector<Model> models;
    std::vector<Inst> insts;
    std::vector<BvhNode> tlas;
    std::vector<uint32_t> tlas_order;
    bool brute_force = false;
    int threads = 0;
    std::string err;
};

namespace {

// ------------------------------------------------------------------ BVH2 build (binned SAH)
struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
    }
    void grow(const float* l, const float* h) {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); }
    }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0 && dy >= 0 && dz >= 0)) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct PrimBox { float lo[3], hi[3], c[3]; };

void build_bvh2(const std::vector<PrimBox>& prims, uint32_t max_leaf, std::vector<BvhNode>& nodes,
                std::vector<uint32_t>& order) {
    uint32_t n = (uint32_t)prims.size();
    order.resize(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    nodes.clear();
    nodes.reserve(2 * (size_t)n + 1);
    nodes.push_back(BvhNode{});
    if (n == 0) {
        BvhNode& r = nodes[0];
        for (int k = 0; k < 3; k++) { r.lo[k] = INFINITY; r.hi[k] = -INFINITY; }
        r.a = 0; r.count = 0;
        // empty leaf: mark as leaf with zero prims via count=0 and a=0xFFFFFFFF
        r.a = 0xFFFFFFFFu;
        return;
    }
    struct Task { uint32_t node, first, count; };
    std::vector<Task> stack;
    stack.push_back({0, 0, n});
    const int NB = 16;
    while (!stack.empty()) {
        Task t = stack.back();
        stack.pop_back();
        Box b, cb;
        b.reset(); cb.reset();
        for (uint32_t i = t.first; i < t.first + t.count; i++) {
            const PrimBox& p = prims[order[i]];
            b.grow(p.lo, p.hi);
            cb.grow(p.c, p.c);
        }
        // copy (nodes may reallocate later)
        {
            BvhNode& nd = nodes[t.node];
            for (int k = 0; k < 3; k++) { nd.lo[k] = b.lo[k]; nd.hi[k] = b.hi[k]; }
        }
        auto make_leaf = [&]() { nodes[t.node].a = t.first; nodes[t.node].count = t.count; };
        if (t.count <= max_leaf) { make_leaf(); continue; }
        int best_axis = -1, best_split = -1;
        float best_cost = INFINITY;
        for (int ax = 0; ax < 3; ax++) {
            float ext = cb.hi[ax] - cb.lo[ax];
            if (!(ext > 0)) continue;
            Box bins[NB];
            uint32_t cnt[NB] = {0};
            for (int i = 0; i < NB; i++) bins[i].reset();
            float k1 = NB / ext;
            for (uint32_t i = t.first; i < t.first + t.count; i++) {
                const PrimBox& p = prims[order[i]];
                int bi = std::min(NB - 1, std::max(0, (int)((p.c[ax] - cb.lo[ax]) * k1)));
                bins[bi].grow(p.lo, p.hi);
                cnt[bi]++;
            }
            float right_area[NB];
            uint32_t right_cnt[NB];
            Box acc; acc.reset();
            uint32_t c = 0;
            for (int i = NB - 1; i > 0; i--) {
                if (cnt[i]) acc.grow(bins[i].lo, bins[i].hi);
                c += cnt[i];
                right_area[i] = acc.area();
                right_cnt[i] = c;
            }
            acc.reset(); c = 0;
            for (int i = 0; i < NB - 1; i++) {
                if (cnt[i]) acc.grow(bins[i].lo, bins[i].hi);
                c += cnt[i];
                if (c == 0 || right_cnt[i + 1] == 0) continue;
                float cost = acc.area() * c + right_area[i + 1] * right_cnt[i + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = ax; best_split = i; }
            }
        }
        uint32_t mid;
        if (best_axis < 0) {
            mid = t.first + t.count / 2;  // all centroids coincide: split by index
        } else {
            float ext = cb.hi[best_axis] - cb.lo[best_axis];
            float k1 = NB / ext;
            auto it = std::partition(order.begin() + t.first, order.begin() + t.first + t.count, [&](uint32_t id) {
                int bi = std::min(NB - 1, std::max(0, (int)((prims[id].c[best_axis] - cb.lo[best_axis]) * k1)));
                return bi <= best_split;
            });
            mid = (uint32_t)(it - order.begin());
            if (mid == t.first || mid == t.first + t.count) mid = t.first + t.count / 2;
        }
        uint32_t left = (uint32_t)nodes.size();
        nodes.push_back(BvhNode{});
        nodes.push_back(BvhNode{});
        nodes[t.node].a = left;
        nodes[t.node].count = 0;
        stack.push_back({left, t.first, mid - t.first});
        stack.push_back({left + 1, mid, t.first + t.count - mid});
    }
}

// Conservative slab test.  Boxes were padded at build time and the far
// distance is scaled up, so a ray that the triangle test accepts is never
// culled by rounding in the box test.
inline bool slab(const BvhNode& nd, V3 o, V3 id, float tmin, float tmax, float& tnear) {
    float t0x = (nd.lo[0] - o.x) * id.x, t1x = (nd.hi[0] - o.x) * id.x;
    float t0y = (nd.lo[1] - o.y) * id.y, t1y = (nd.hi[1] - o.y) * id.y;
    float t0z = (nd.lo[2] - o.z) * id.z, t1z = (nd.hi[2] - o.z) * id.z;
    float tn = std::fmax(std::fmax(std::fmin(t0x, t1x), std::fmin(t0y, t1y)), std::fmax(std::fmin(t0z, t1z), tmin));
    float tf = std::fmin(std::fmin(std::fmax(t0x, t1x), std::fmax(t0y, t1y)), std::fmin(std::fmax(t0z, t1z), tmax));
    tf = tf * 1.00001f + 1e-30f;
    tn = tn * 0.99999f;
    tnear = tn;
    return tn <= tf;
}

inline V3 safe_inv_dir(V3 d) {
    auto inv = [](float x) {
        float ax = std::fabs(x);
        if (!(ax >= 1e-30f)) x = std::signbit(x) ? -1e-30f : 1e-30f;
        return 1.0f / x;
    };
    return v3(inv(d.x), inv(d.y), inv(d.z));
}

void pad_box(float* lo, float* hi) {
    float m = 0.f;
    for (int k = 0; k < 3; k++) m = std::max(m, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
    float pad = m * 1e-6f + 1e-30f;
    for (int k = 0; k < 3; k++) { lo[k] -= pad; hi[k] += pad; }
}

// ------------------------------------------------------------------ textures
// Vulkan sampling rules at LOD 0 (explicit-lod only in the shipped .spv files):
// REPEAT addressing, normalised coords, texel centres at +0.5, sRGB decode of
// RGB before filtering, alpha linear.   src/util_structs.rs:1306-1320
struct F4 { float r, g, b, a; };

inline int wrap(int i, int n) {
    int m = i % n;
    return m < 0 ? m + n : m;
}

inline F4 texel(const Texture& t, int x, int y) {
    if (t.format == RT_FORMAT_RGBA32_SFLOAT) {
        const float* p = &t.rgba32f[((size_t)y * t.w + x) * 4];
        return F4{p[0], p[1], p[2], p[3]};
    }
    const uint8_t* p = &t.rgba8[((size_t)y * t.w + x) * 4];
    if (t.format == RT_FORMAT_RGBA8_SRGB)
        return F4{g_srgb_lut[p[0]], g_srgb_lut[p[1]], g_srgb_lut[p[2]], (float)p[3] / 255.0f};
    return F4{(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f};
}

F4 sample_texture(const OrcContext& c, uint32_t index, float u, float v) {
    if (index >= c.textures.size()) return F4{0, 0, 0, 0};  // robustness2 null descriptor (src/main.rs:183-184)
    const Texture& t = c.textures[index];
    if (!t.linear) {
        int x = wrap((int)std::floor(u * (float)t.w), (int)t.w);
        int y = wrap((int)std::floor(v * (float)t.h), (int)t.h);
        return texel(t, x, y);
    }
    float fx = u * (float)t.w - 0.5f, fy = v * (float)t.h - 0.5f;
    float flx = std::floor(fx), fly = std::floor(fy);
    float ax = fx - flx, ay = fy - fly;
    int x0 = wrap((int)flx, (int)t.w), y0 = wrap((int)fly, (int)t.h);
    int x1 = wrap(x0 + 1, (int)t.w), y1 = wrap(y0 + 1, (int)t.h);
    F4 t00 = texel(t, x0, y0), t10 = texel(t, x1, y0), t01 = texel(t, x0, y1), t11 = texel(t, x1, y1);
    float w00 = (1.0f - ax) * (1.0f - ay), w10 = ax * (1.0f - ay), w01 = (1.0f - ax) * ay, w11 = ax * ay;
    return F4{t00.r * w00 + t10.r * w10 + t01.r * w01 + t11.r * w11,
              t00.g * w00 + t10.g * w10 + t01.g * w01 + t11.g * w11,
              t00.b * w00 + t10.b * w10 + t01.b * w01 + t11.b * w11,
              t00.a * w00 + t10.a * w10 + t01.a * w01 + t11.a * w11};
}

// ------------------------------------------------------------------ trace
struct Hit {
    float t, u, v;
    uint32_t inst, tri;  // tri = index into the BLAS model's tris[]
    bool valid;
};

struct Ray { V3 o, d; float tmin, tmax; };

// gl_InstanceCustomIndexEXT -> ModelInfo; any_hit_alpha_clip.glsl:11-28
bool anyhit_accepts(const OrcContext& c, const Inst& in, const Tri& tr, float u, float v) {
    uint32_t custom = in.rec.instance_custom_index_and_mask & 0xFFFFFFu;
    if (custom >= c.models.size()) return true;
    const Model& m = c.models[custom];
    if (tr.geom >= m.geoms.size()) return true;
    const Geometry& g = m.geoms[tr.geom];
    if ((size_t)tr.prim * 3 + 2 >= g.indices.size()) return true;
    uint32_t ia = g.indices[tr.prim * 3], ib = g.indices[tr.prim * 3 + 1], ic = g.indices[tr.prim * 3 + 2];
    V3 w = v3((1.0f - u) - v, u, v);
    V2 uv = interp2(m.uvs[ia], m.uvs[ib], m.uvs[ic], w);
    float alpha = sample_texture(c, g.images.diffuse_image_index, uv.x, uv.y).a;
    return !(alpha < 0.5f);
}

// Candidate ordering rule: smaller t wins; exact tie -> lowest
// (instance, geometry, primitive).
inline bool better(float t, uint32_t inst, const Tri& tr, const Hit& h, const OrcContext& c) {
    if (!h.valid) return true;
    if (t < h.t) return true;
    if (t > h.t) return false;
    if (inst != h.inst) return inst < h.inst;
    const Tri& ht = c.models[c.insts[h.inst].blas].tris[h.tri];
    if (tr.geom != ht.geom) return tr.geom < ht.geom;
    return tr.prim < ht.prim;
}

// Tests one triangle.  `any` = shadow-ray mode (TerminateOnFirstHit).  Returns true if traversal can stop.
inline bool test_tri(const OrcContext& c, uint32_t inst_id, const Inst& in, const Model& m, uint32_t tri_id,
                     V3 oo, V3 od, const Ray& r, Hit& h, bool any) {
    const Tri& tr = m.tris[tri_id];
    float t, u, v;
    if (!tri_candidate(oo, od, tr.v0, tr.e1, tr.e2, t, u, v)) return false;
    if (!(t > r.tmin && t < r.tmax)) return false;
    if (!any && !better(t, inst_id, tr, h, c)) return false;
    if (!tr.opaque && !anyhit_accepts(c, in, tr, u, v)) return false;
    h.valid = true; h.t = t; h.u = u; h.v = v; h.inst = inst_id; h.tri = tri_id;
    return any;
}

bool trace_instance(const OrcContext& c, uint32_t inst_id, const Ray& r, Hit& h, bool any) {
    const Inst& in = c.insts[inst_id];
    if (in.blas < 0) return false;
    if ((in.rec.instance_custom_index_and_mask >> 24) == 0) return false;  // mask & 0xFF cull mask
    const Model& m = c.models[in.blas];
    V3 oo = xform_point(in.inv, r.o), od = xform_vec(in.inv, r.d);
    if (c.brute_force) {
        for (uint32_t i = 0; i < m.tris.size(); i++)
            if (test_tri(c, inst_id, in, m, i, oo, od, r, h, any)) return true;
        return false;
    }
    if (m.nodes.empty() || m.tris.empty()) return false;
    V3 id = safe_inv_dir(od);
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const BvhNode& nd = m.nodes[stack[--sp]];
        float tn;
        float tmax = h.valid ? h.t : r.tmax;
        if (!slab(nd, oo, id, r.tmin, tmax, tn)) continue;
        if (nd.count) {
            for (uint32_t i = nd.a; i < nd.a + nd.count; i++)
                if (test_tri(c, inst_id, in, m, i, oo, od, r, h, any)) return true;
        } else if (sp + 2 <= 128) {
            stack[sp++] = nd.a;
            stack[sp++] = nd.a + 1;
        }
    }
    return false;
}

Hit trace(const OrcContext& c, const Ray& r, bool any) {
    Hit h;
    h.valid = false; h.t = r.tmax; h.u = h.v = 0; h.inst = h.tri = 0xFFFFFFFFu;
    if (c.insts.empty()) return h;
    if (c.brute_force) {
        for (uint32_t i = 0; i < c.insts.size(); i++)
            if (trace_instance(c, i, r, h, any)) return h;
        return h;
    }
    V3 id = safe_inv_dir(r.d);
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const BvhNode& nd = c.tlas[stack[--sp]];
        float tn;
        float tmax = h.valid ? h.t : r.tmax;
        if (!slab(nd, r.o, id, r.tmin, tmax, tn)) continue;
        if (nd.count) {
            for (uint32_t i = nd.a; i < nd.a + nd.count; i++)
                if (trace_instance(c, c.tlas_order[i], r, h, any)) return h;
        } else if (nd.a != 0xFFFFFFFFu && sp + 2 <= 128) {
            stack[sp++] = nd.a;
            stack[sp++] = nd.a + 1;
        }
    }
    return h;
}

// ------------------------------------------------------------------ shading
const float PI = 3.141592653589793f;

struct TriAttr {
    V3 pa, pb, pc, na, nb, nc;
    V2 ta, tb, tc;
};

// pbr.glsl:25-103
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
struct DotParams { float NoH, NoV, NoL, LoH, roughness; };

inline float D_GGX(const DotParams& p) {
    float a = p.NoH * p.roughness;
    float k = p.roughness / (1.0f - p.NoH * p.NoH + a * a);
    return k * k * (1.0f / PI);
}
inline float V_SmithGGXCorrelated(const DotParams& p) {
    float a2 = p.roughness * p.roughness;
    float GGXV = p.NoL * std::sqrt(p.NoV * p.NoV * (1.0f - a2) + a2);
    float GGXL = p.NoV * std::sqrt(p.NoL * p.NoL * (1.0f - a2) + a2);
    return 0.5f / (GGXV + GGXL);
}
inline float F_Schlick1(float u, float f0, float f90) { return f0 + (f90 - f0) * std::pow(1.0f - u, 5.0f); }
inline float compute_f90(const DotParams& p) { return 0.5f + 2.0f * p.roughness * p.LoH * p.LoH; }
inline float Fd_Burley(const DotParams& p) {
    float f90 = compute_f90(p);
    float ls = F_Schlick1(p.NoL, 1.0f, f90);
    float vs = F_Schlick1(p.NoV, 1.0f, f90);
    return ls * vs * (1.0f / PI);
}

struct BrdfIn {
    V3 normal, view, light, base;
    float perceptual_roughness, metallic, reflectance;
    V3 light_intensity;
};

inline float plain_dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// pbr.glsl:174-211
V3 brdf(const BrdfIn& in) {
    V3 hsum = add3(in.view, in.light);
    float hl = std::sqrt(plain_dot(hsum, hsum));
    V3 h = v3(hsum.x / hl, hsum.y / hl, hsum.z / hl);
    DotParams p;
    p.roughness = in.perceptual_roughness * in.perceptual_roughness;
    p.NoV = clampf(plain_dot(in.normal, in.view), 10.0e-10f, 1.0f);
    p.NoH = clampf(plain_dot(in.normal, h), 0.0f, 1.0f);
    p.NoL = clampf(plain_dot(in.normal, in.light), 0.0f, 1.0f);
    p.LoH = clampf(plain_dot(in.light, h), 0.0f, 1.0f);
    float D = D_GGX(p);
    float dielectric_f0 = 0.16f * in.reflectance * in.reflectance;
    V3 f0 = v3(dielectric_f0 * (1.0f - in.metallic) + in.base.x * in.metallic,
               dielectric_f0 * (1.0f - in.metallic) + in.base.y * in.metallic,
               dielectric_f0 * (1.0f - in.metallic) + in.base.z * in.metallic);
    float f90 = compute_f90(p);
    float fw = std::pow(1.0f - p.LoH, 5.0f);
    V3 F = v3(f0.x + (f90 - f0.x) * fw, f0.y + (f90 - f0.y) * fw, f0.z + (f90 - f0.z) * fw);
    float G = V_SmithGGXCorrelated(p);
    float DG = D * G;
    float fd = Fd_Burley(p);
    V3 comb = v3(in.base.x * fd + DG * F.x, in.base.y * fd + DG * F.y, in.base.z * fd + DG * F.z);
    return v3(in.light_intensity.x * p.NoL * comb.x, in.light_intensity.y * p.NoL * comb.y,
              in.light_intensity.z * p.NoL * comb.z);
}

// closest_hit_textured.glsl:13-19
inline V3 project_onto_tangent_plane(V3 point, V3 vpos, V3 vnormal) {
    V3 vtp = sub3(point, vpos);
    float dp = std::fmin(0.0f, dot3(vtp, vnormal));
    return v3(fma_(-dp, vnormal.x, vtp.x), fma_(-dp, vnormal.y, vtp.y), fma_(-dp, vnormal.z, vtp.z));
}

// closest_hit_textured.glsl:25-39 (object->world applied with the instance 3x4)
inline V3 terminator_origin(const TriAttr& a, V3 p, V3 w, const float* o2w) {
    V3 oa = project_onto_tangent_plane(p, a.pa, a.na);
    V3 ob = project_onto_tangent_plane(p, a.pb, a.nb);
    V3 oc = project_onto_tangent_plane(p, a.pc, a.nc);
    V3 off = interp3(oa, ob, oc, w);
    return xform_point(o2w, add3(p, off));
}

// closest_hit_textured.glsl:99-120.  Nearest + REPEAT on exact k/64 coordinates
// is an integer modular index; `.r` of an RGBA8_UNORM texel.
inline V2 blue_noise_xi(const OrcContext& c, uint32_t tex_index, uint32_t px, uint32_t py, uint32_t iteration,
                        uint32_t frame_index) {
    uint32_t ox1 = iteration * 2u * 13u, oy1 = iteration * 2u * 41u;
    uint32_t ox2 = (iteration * 2u + 1u) * 13u, oy2 = (iteration * 2u + 1u) * 41u;
    float a = sample_texture(c, tex_index, (float)(px + ox1) / 64.0f, (float)(py + oy1) / 64.0f).r;
    float b = sample_texture(c, tex_index, (float)(px + ox2) / 64.0f, (float)(py + oy2) / 64.0f).r;
    float k = (float)(frame_index % 32u) * 0.618033988749f;
    float sa = a + k, sb = b + k;
    return V2{sa - std::floor(sa), sb - std::floor(sb)};
}

// closest_hit_textured.glsl:77-94
inline V3 sample_directional_light(V2 rng, V3 center, float radius) {
    float r = std::sqrt(rng.x);
    float angle = rng.y * 2.0f * PI;
    float px = r * std::cos(angle) * radius, py = r * std::sin(angle) * radius;
    V3 up = v3(0.f, 1.f, 0.f);
    V3 tangent = normalize3(cross3(center, up));
    V3 bitangent = normalize3(cross3(tangent, center));
    return normalize3(v3(center.x + px * tangent.x + py * bitangent.x, center.y + px * tangent.y + py * bitangent.y,
                         center.z + px * tangent.z + py * bitangent.z));
}

// lib.rs:85-92
inline float linear_to_srgb1(float c) {
    return c <= 0.0031308f ? c * 12.92f : 1.055f * std::pow(c, 1.0f / 2.4f) - 0.055f;
}
inline uint8_t unorm8(float c) {
    if (!(c > 0.0f)) return 0;  // NaN and negatives -> 0
    if (c >= 1.0f) return 255;
    return (uint8_t)std::floor(c * 255.0f + 0.5f);
}

struct Payload { V3 colour, new_origin, new_dir; };

bool load_tri_attr(const OrcContext& c, const Inst& in, const Tri& tr, TriAttr& a, const Geometry** geom) {
    uint32_t custom = in.rec.instance_custom_index_and_mask & 0xFFFFFFu;
    if (custom >= c.models.size()) return false;
    const Model& m = c.models[custom];
    if (tr.geom >= m.geoms.size()) return false;
    const Geometry& g = m.geoms[tr.geom];
    if ((size_t)tr.prim * 3 + 2 >= g.indices.size()) return false;
    uint32_t ia = g.indices[tr.prim * 3], ib = g.indices[tr.prim * 3 + 1], ic = g.indices[tr.prim * 3 + 2];
    if (ia >= m.positions.size() || ib >= m.positions.size() || ic >= m.positions.size()) return false;
    a.pa = m.positions[ia]; a.pb = m.positions[ib]; a.pc = m.positions[ic];
    a.na = m.normals[ia]; a.nb = m.normals[ib]; a.nc = m.normals[ic];
    a.ta = m.uvs[ia]; a.tb = m.uvs[ib]; a.tc = m.uvs[ic];
    *geom = &g;
    return true;
}

// closest_hit_textured.glsl:174-226
void closest_hit_textured(const OrcContext& c, const RtUniforms& u, uint32_t shadow_rays, uint32_t px, uint32_t py,
                          const Ray& ray, const Hit& h, Payload& pl, uint64_t& shadow_count) {
    const Inst& in = c.insts[h.inst];
    const Tri& tr = c.models[in.blas].tris[h.tri];
    TriAttr a;
    const Geometry* g;
    if (!load_tri_attr(c, in, tr, a, &g)) return;
    V3 w = v3((1.0f - h.u) - h.v, h.u, h.v);
    V3 ipos = interp3(a.pa, a.pb, a.pc, w);
    V3 inrm = interp3(a.na, a.nb, a.nc, w);
    V2 iuv = interp2(a.ta, a.tb, a.tc, w);

    V3 shadow_origin = terminator_origin(a, ipos, w, in.rec.transform);
    V3 sun = v3(u.sun_dir[0], u.sun_dir[1], u.sun_dir[2]);
    float sum = 0.0f;
    for (uint32_t i = 0; i < shadow_rays; i++) {
        V2 xi = blue_noise_xi(c, u.blue_noise_texture_index, px, py, i, u.frame_index);
        V3 dir = sample_directional_light(xi, sun, u.sun_radius);
        // cast_shadow_ray, closest_hit_textured.glsl:159-172
        Ray sr{shadow_origin, dir, 0.001f, 10000.0f};
        Hit sh = trace(c, sr, true);
        shadow_count++;
        sum += sh.valid ? 0.0f : 1.0f;
    }
    float sun_factor = sum / (float)shadow_rays;

    // read_material_from_textures, :47-63
    F4 dcol = sample_texture(c, g->images.diffuse_image_index, iuv.x, iuv.y);
    F4 mr = sample_texture(c, g->images.metallic_roughness_image_index, iuv.x, iuv.y);
    float metallic = mr.b, roughness = mr.g;

    // calculate_normal, :141-157
    V3 normal;
    if (g->images.normal_map_image_index < 0) {
        normal = normalize3(xform_normal(in.inv, inrm));
    } else {
        F4 nm = sample_texture(c, (uint32_t)g->images.normal_map_image_index, iuv.x, iuv.y);
        V3 mn = v3(nm.r * 2.0f - 1.0f, nm.g * 2.0f - 1.0f, nm.b * 2.0f - 1.0f);
        // compute_cotangent_frame, :123-139
        V3 dp1 = sub3(a.pb, a.pa), dp2 = sub3(a.pc, a.pa);
        V2 duv1 = V2{a.tb.x - a.ta.x, a.tb.y - a.ta.y}, duv2 = V2{a.tc.x - a.ta.x, a.tc.y - a.ta.y};
        V3 dp2perp = cross3(dp2, inrm), dp1perp = cross3(inrm, dp1);
        V3 T = add3(scale3(dp2perp, duv1.x), scale3(dp1perp, duv2.x));
        V3 B = add3(scale3(dp2perp, duv1.y), scale3(dp1perp, duv2.y));
        float invmax = 1.0f / std::sqrt(std::fmax(dot3(T, T), dot3(B, B)));
        V3 mnn = normalize3(mn);
        V3 Ts = scale3(T, invmax), Bs = scale3(B, invmax);
        V3 local = v3(Ts.x * mnn.x + Bs.x * mnn.y + inrm.x * mnn.z, Ts.y * mnn.x + Bs.y * mnn.y + inrm.y * mnn.z,
                      Ts.z * mnn.x + Bs.z * mnn.y + inrm.z * mnn.z);
        normal = normalize3(xform_normal(in.inv, local));
    }

    BrdfIn bi;
    bi.normal = normal;
    bi.view = v3(-ray.d.x, -ray.d.y, -ray.d.z);
    bi.light = sun;
    bi.base = v3(dcol.r, dcol.g, dcol.b);
    bi.metallic = metallic;
    bi.perceptual_roughness = roughness;
    bi.reflectance = 0.5f;
    bi.light_intensity = v3(sun_factor, sun_factor, sun_factor);
    V3 lo = brdf(bi);
    pl.colour = v3(lo.x + 0.1f * dcol.r, lo.y + 0.1f * dcol.g, lo.z + 0.1f * dcol.b);
}

// closest_hit_mirror.glsl:11-29
void closest_hit_mirror(const OrcContext& c, const Ray& ray, const Hit& h, Payload& pl) {
    const Inst& in = c.insts[h.inst];
    const Tri& tr = c.models[in.blas].tris[h.tri];
    TriAttr a;
    const Geometry* g;
    if (!load_tri_attr(c, in, tr, a, &g)) return;
    V3 w = v3((1.0f - h.u) - h.v, h.u, h.v);
    V3 n = normalize3(xform_normal(in.inv, interp3(a.na, a.nb, a.nc, w)));
    float k = 2.0f * dot3(n, ray.d);  // reflect(I,N) = I - 2 dot(N,I) N
    pl.new_dir = v3(fma_(-k, n.x, ray.d.x), fma_(-k, n.y, ray.d.y), fma_(-k, n.z, ray.d.z));
    pl.new_origin = v3(fma_(ray.d.x, h.t, ray.o.x), fma_(ray.d.y, h.t, ray.o.y), fma_(ray.d.z, h.t, ray.o.z));
}

// lib.rs:300-312
void closest_hit_portal(const Ray& ray, const Hit& h, Payload& pl) {
    pl.new_dir = ray.d;
    pl.new_origin = v3(fma_(ray.d.x, h.t, ray.o.x), fma_(ray.d.y, h.t, ray.o.y) + 5.0f, fma_(ray.d.z, h.t, ray.o.z));
}

// lib.rs:40-51
void primary_ray_miss(const RtUniforms& u, const Ray& ray, Payload& pl) {
    V3 sun = v3(u.sun_dir[0], u.sun_dir[1], u.sun_dir[2]);
    if (dot3(ray.d, sun) > std::cos(u.sun_radius)) pl.colour = v3(1.f, 1.f, 1.f);
    else pl.colour = v3(0.0f, 0.0f, 0.05f);
}

// lib.rs:126-142
Ray generate_primary_ray(const RtUniforms& u, uint32_t x, uint32_t y, uint32_t W, uint32_t H) {
    float pcx = (float)x + 0.5f, pcy = (float)y + 0.5f;
    float ndx = (pcx / (float)W) * 2.0f - 1.0f, ndy = (pcy / (float)H) * 2.0f - 1.0f;
    const float* V = u.view_inverse;
    const float* P = u.proj_inverse;
    V3 origin = v3(V[12], V[13], V[14]);
    // proj_inverse * (ndx, ndy, 1, 1): r_i = fma(P[12+i],1, fma(P[8+i],1, fma(P[4+i],y, P[i]*x)))
    V3 target = v3(fma_(P[12], 1.0f, fma_(P[8], 1.0f, fma_(P[4], ndy, P[0] * ndx))),
                   fma_(P[13], 1.0f, fma_(P[9], 1.0f, fma_(P[5], ndy, P[1] * ndx))),
                   fma_(P[14], 1.0f, fma_(P[10], 1.0f, fma_(P[6], ndy, P[2] * ndx))));
    V3 ld = normalize3(target);
    V3 dir = v3(fma_(V[8], ld.z, fma_(V[4], ld.y, V[0] * ld.x)), fma_(V[9], ld.z, fma_(V[5], ld.y, V[1] * ld.x)),
                fma_(V[10], ld.z, fma_(V[6], ld.y, V[2] * ld.x)));
    return Ray{origin, dir, 0.01f, 10000.0f};
}

// heatmap.rs:43-54
inline float saturate1(float x) { return std::fmin(std::fmax(x, 0.0f), 1.0f); }
inline float smoothstep1(float e0, float e1, float x) {
    float t = saturate1((x - e0) / (e1 - e0));
    return (t * t) * (3.0f - 2.0f * t);
}
// heatmap.rs:5-41.  `colours[heat as i32]` reads one past the table at heat == 1.0 exactly (undefined in
// SPIR-V): `cur` is clamped to the last entry, the same rule as the CUDA path.
V3 heatmap_temperature(float heat) {
    static const float k[10][3] = {
        {0.0f / 255.0f, 2.0f / 255.0f, 91.0f / 255.0f},    {0.0f / 255.0f, 108.0f / 255.0f, 251.0f / 255.0f},
        {0.0f / 255.0f, 221.0f / 255.0f, 221.0f / 255.0f}, {51.0f / 255.0f, 221.0f / 255.0f, 0.0f / 255.0f},
        {255.0f / 255.0f, 252.0f / 255.0f, 0.0f / 255.0f}, {255.0f / 255.0f, 180.0f / 255.0f, 0.0f / 255.0f},
        {255.0f / 255.0f, 104.0f / 255.0f, 0.0f / 255.0f}, {226.0f / 255.0f, 22.0f / 255.0f, 0.0f / 255.0f},
        {191.0f / 255.0f, 0.0f / 255.0f, 83.0f / 255.0f},  {145.0f / 255.0f, 0.0f / 255.0f, 65.0f / 255.0f}};
    heat = saturate1(heat) * 10.0f;
    int idx = (int)heat;
    int cur = std::min(idx, 9), prv = std::max(idx - 1, 0), nxt = std::min(idx + 1, 9);
    float lo = std::floor(heat), hi = std::ceil(heat), blur = 0.8f;
    float s_lo = smoothstep1(lo - blur, lo + blur, heat);
    float s_hi = smoothstep1(hi - blur, hi + blur, heat);
    float wc = s_lo * (1.0f - s_hi), wp = 1.0f - s_lo, wn = s_hi;
    V3 r = add3(add3(scale3(v3(k[cur][0], k[cur][1], k[cur][2]), wc), scale3(v3(k[prv][0], k[prv][1], k[prv][2]), wp)),
                scale3(v3(k[nxt][0], k[nxt][1], k[nxt][2]), wn));
    return v3(saturate1(r.x), saturate1(r.y), saturate1(r.z));
}
// lib.rs:174-186
V3 heatmap_pixel(uint32_t cycles, float scale, V3 colour) {
    return add3(heatmap_temperature((float)cycles / scale), scale3(colour, 0.000001f));
}
// The oracle has no shader clock.  Its show_heatmap frames use this stand-in for `end_time - start_time`:
// a fixed number of ticks per trace call the pixel issued (ray-gen segments + shadow rays).  The CUDA path's real
// clock values are never compared with it — tests feed the GPU's own cost_cycles through heatmap_pixel().
constexpr uint32_t kOracleTicksPerTrace = 20000u;

struct PixelOut {
    V3 colour;
    uint32_t ids[9];
};

void render_pixel(const OrcContext& c, const RtUniforms& u, const RtRenderParams& p, uint32_t x, uint32_t y,
                  PixelOut& out, uint64_t& n_primary, uint64_t& n_shadow) {
    for (int i = 0; i < 9; i++) out.ids[i] = 0xFFFFFFFFu;
    Ray ray = generate_primary_ray(u, x, y, p.width, p.height);
    Payload pl;
    pl.colour = v3(0, 0, 0);
    uint32_t segs = p.max_segments;
    for (uint32_t s = 0; s < segs; s++) {
        pl.colour = v3(0, 0, 0); pl.new_origin = v3(0, 0, 0); pl.new_dir = v3(0, 0, 0);
        Hit h = trace(c, ray, false);
        n_primary++;
        if (!h.valid) {
            primary_ray_miss(u, ray, pl);
        } else {
            const Inst& in = c.insts[h.inst];
            const Tri& tr = c.models[in.blas].tris[h.tri];
            if (s < 3) { out.ids[s * 3] = h.inst; out.ids[s * 3 + 1] = tr.geom; out.ids[s * 3 + 2] = tr.prim; }
            // hit group = instance.sbt_offset + 0 (src/main.rs:289-305)
            uint32_t kind = in.rec.sbt_record_offset_and_flags & 0xFFFFFFu;
            if (kind == RT_HIT_TEXTURED) closest_hit_textured(c, u, p.shadow_rays, x, y, ray, h, pl, n_shadow);
            else if (kind == RT_HIT_MIRROR) closest_hit_mirror(c, ray, h, pl);
            else if (kind == RT_HIT_PORTAL) closest_hit_portal(ray, h, pl);
            // out-of-table hit group: payload untouched (colour 0, no continuation)
        }
        if (pl.new_dir.x == 0.0f && pl.new_dir.y == 0.0f && pl.new_dir.z == 0.0f) break;
        ray.o = pl.new_origin;
        ray.d = pl.new_dir;
    }
    out.colour = pl.colour;
}

void finish_instance(OrcContext& c, Inst& in) {
    invert_3x4(in.rec.transform, in.inv);
    uint64_t handle = in.rec.acceleration_structure_device_address;
    in.blas = (handle >= 1 && handle <= c.models.size()) ? (int32_t)(handle - 1) : -1;
    for (int k = 0; k < 3; k++) { in.lo[k] = INFINITY; in.hi[k] = -INFINITY; }
    if (in.blas < 0) return;
    const Model& m = c.models[in.blas];
    if (m.nodes.empty() || m.tris.empty()) return;
    const BvhNode& root = m.nodes[0];
    for (int corner = 0; corner < 8; corner++) {
        V3 p = v3(corner & 1 ? root.hi[0] : root.lo[0], corner & 2 ? root.hi[1] : root.lo[1],
                  corner & 4 ? root.hi[2] : root.lo[2]);
        V3 wpt = xform_point(in.rec.transform, p);
        float a[3] = {wpt.x, wpt.y, wpt.z};
        for (int k = 0; k < 3; k++) { in.lo[k] = std::min(in.lo[k], a[k]); in.hi[k] = std::max(in.hi[k], a[k]); }
    }
    pad_box(in.lo, in.hi);
}

void rebuild_tlas(OrcContext& c) {
    std::vector<PrimBox> pb(c.insts.size());
    for (size_t i = 0; i < c.insts.size(); i++) {
        Inst& in = c.insts[i];
        finish_instance(c, in);
        for (int k = 0; k < 3; k++) {
            bool ok = in.lo[k] <= in.hi[k];
            pb[i].lo[k] = ok ? in.lo[k] : 0.f;
            pb[i].hi[k] = ok ? in.hi[k] : 0.f;
            pb[i].c[k] = 0.5f * (pb[i].lo[k] + pb[i].hi[k]);
            if (!std::isfinite(pb[i].c[k])) { pb[i].lo[k] = pb[i].hi[k] = pb[i].c[k] = 0.f; }
        }
        if (!(in.lo[0] <= in.hi[0])) {  // empty / degenerate instance: unreachable box
            for (int k = 0; k < 3; k++) { pb[i].lo[k] = INFINITY; pb[i].hi[k] = -INFINITY; pb[i].c[k] = 0.f; }
        }
    }
    build_bvh2(pb, 1, c.tlas, c.tlas_order);
}

}  // namespace

// ====================================================================== C API
extern "C" {

int orc_create(OrcContext** out) {
    init_lut();
    *out = new OrcContext();
    return 0;
}
void orc_destroy(OrcContext* c) { delete c; }
const char* orc_last_error(const OrcContext* c) { return c ? c->err.c_str() : ""; }
void orc_set_brute_force(OrcContext* c, int on) { c->brute_force = on != 0; }
void orc_set_threads(OrcContext* c, int n) { c->threads = n; }
int orc_get_threads(const OrcContext* c) {
    if (c->threads > 0) return c->threads;
    unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

int orc_push_image(OrcContext* c, const void* texels, uint32_t w, uint32_t h, uint32_t format, int linear,
                   uint32_t* out_index) {
    if (!texels || !w || !h || format > 2) { c->err = "orc_push_image: bad argument"; return RT_ERR_INVALID_ARGUMENT; }
    if (c->textures.size() >= RT_MAX_BOUND_IMAGES) { c->err = "image table full"; return RT_ERR_OUT_OF_RANGE; }
    Texture t;
    t.w = w; t.h = h; t.format = format; t.linear = linear != 0;
    size_t n = (size_t)w * h * 4;
    if (format == RT_FORMAT_RGBA32_SFLOAT) t.rgba32f.assign((const float*)texels, (const float*)texels + n);
    else t.rgba8.assign((const uint8_t*)texels, (const uint8_t*)texels + n);
    if (out_index) *out_index = (uint32_t)c->textures.size();
    c->textures.push_back(std::move(t));
    return 0;
}

int orc_create_model(OrcContext* c, const RtModelDesc* d, uint32_t* out_id, uint64_t* out_handle) {
    if (!d || (d->num_vertices && (!d->positions || !d->normals || !d->uvs))) {
        c->err = "orc_create_model: bad argument";
        return RT_ERR_INVALID_ARGUMENT;
    }
    Model m;
    m.positions.resize(d->num_vertices); m.normals.resize(d->num_vertices); m.uvs.resize(d->num_vertices);
    if (d->num_vertices) {
        memcpy(m.positions.data(), d->positions, sizeof(float) * 3 * d->num_vertices);
        memcpy(m.normals.data(), d->normals, sizeof(float) * 3 * d->num_vertices);
        memcpy(m.uvs.data(), d->uvs, sizeof(float) * 2 * d->num_vertices);
    }
    std::vector<Tri> tris;
    std::vector<PrimBox> pb;
    for (uint32_t gi = 0; gi < d->num_geometries; gi++) {
        const RtGeometryDesc& gd = d->geometries[gi];
        Geometry g;
        g.indices.assign(gd.indices, gd.indices + gd.num_indices);
        g.opaque = gd.opaque != 0;
        g.images = gd.images;
        for (uint32_t i = 0; i < gd.num_indices; i++)
            if (gd.indices[i] >= d->num_vertices) { c->err = "index out of range"; return RT_ERR_OUT_OF_RANGE; }
        for (uint32_t p = 0; p < gd.num_indices / 3; p++) {
            V3 a = m.positions[g.indices[p * 3]], b = m.positions[g.indices[p * 3 + 1]], cc = m.positions[g.indices[p * 3 + 2]];
            Tri t;
            t.v0 = a; t.e1 = sub3(b, a); t.e2 = sub3(cc, a);
            t.geom = gi; t.prim = p; t.opaque = g.opaque;
            PrimBox q;
            float xs[3][3] = {{a.x, a.y, a.z}, {b.x, b.y, b.z}, {cc.x, cc.y, cc.z}};
            for (int k = 0; k < 3; k++) {
                q.lo[k] = std::min(xs[0][k], std::min(xs[1][k], xs[2][k]));
                q.hi[k] = std::max(xs[0][k], std::max(xs[1][k], xs[2][k]));
            }
            bool finite = true;
            for (int k = 0; k < 3; k++) finite = finite && std::isfinite(q.lo[k]) && std::isfinite(q.hi[k]);
            if (!finite) continue;  // inactive primitive (NaN position): never hit
            pad_box(q.lo, q.hi);
            for (int k = 0; k < 3; k++) q.c[k] = 0.5f * (q.lo[k] + q.hi[k]);
            tris.push_back(t);
            pb.push_back(q);
        }
        m.geoms.push_back(std::move(g));
    }
    std::vector<uint32_t> order;
    build_bvh2(pb, 4, m.nodes, order);
    m.tris.resize(tris.size());
    for (size_t i = 0; i < order.size(); i++) m.tris[i] = tris[order[i]];
    if (out_id) *out_id = (uint32_t)c->models.size();
    if (out_handle) *out_handle = (uint64_t)c->models.size() + 1;
    c->models.push_back(std::move(m));
    return 0;
}

int orc_build_tlas(OrcContext* c, const RtInstance* inst, uint32_t n) {
    c->insts.resize(n);
    for (uint32_t i = 0; i < n; i++) c->insts[i].rec = inst[i];
    rebuild_tlas(*c);
    return 0;
}

int orc_update_instances(OrcContext* c, uint32_t first, uint32_t count, const RtInstance* inst) {
    if ((uint64_t)first + count > c->insts.size()) { c->err = "instance range out of bounds"; return RT_ERR_OUT_OF_RANGE; }
    for (uint32_t i = 0; i < count; i++) c->insts[first + i].rec = inst[i];
    return 0;
}

int orc_update_tlas(OrcContext* c, uint32_t /*mode*/) {
    rebuild_tlas(*c);
    return 0;
}

static int rows_of(const RtRenderParams& p, std::vector<uint32_t>& ys, uint32_t& x0, uint32_t& tw, std::string& err) {
    uint32_t ty0 = p.tile_y0, th = p.tile_h;
    x0 = p.tile_x0; tw = p.tile_w;
    if (tw == 0) { x0 = 0; ty0 = 0; tw = p.width; th = p.height; }
    if (x0 + tw > p.width || ty0 + th > p.height) { err = "tile outside image"; return RT_ERR_OUT_OF_RANGE; }
    for (uint32_t r = 0; r < th; r++) {
        if (p.strip_height && p.strip_count > 1) {
            uint32_t s = r / p.strip_height;
            if (s % p.strip_count != p.strip_index) continue;
        }
        ys.push_back(ty0 + r);
    }
    return 0;
}

int orc_render(OrcContext* c, const RtUniforms* u, const RtRenderParams* p, const RtFrameOutputs* out) {
    if (!u || !p || !p->width || !p->height || !p->shadow_rays) { c->err = "orc_render: bad argument"; return RT_ERR_INVALID_ARGUMENT; }
    std::vector<uint32_t> ys;
    uint32_t x0, tw;
    int rc = rows_of(*p, ys, x0, tw, c->err);
    if (rc) return rc;
    uint64_t total_primary = 0, total_shadow = 0;
    uint32_t segs_out = std::min(p->max_segments, 3u);
    int nrows = (int)ys.size();
    std::atomic<int> next_row{0};
    std::atomic<uint64_t> acc_primary{0}, acc_shadow{0};
    auto worker = [&]() {
        uint64_t np = 0, ns = 0;
        for (;;) {
            int r = next_row.fetch_add(1);
            if (r >= nrows) break;
            for (uint32_t i = 0; i < tw; i++) {
                PixelOut po;
                const uint64_t traces_before = np + ns;
                render_pixel(*c, *u, *p, x0 + i, ys[r], po, np, ns);
                size_t pix = (size_t)r * tw + i;
                if (u->show_heatmap) {
                    uint64_t ticks = (np + ns - traces_before) * kOracleTicksPerTrace;
                    uint32_t cycles = ticks > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)ticks;
                    if (out && out->cost_cycles) out->cost_cycles[pix] = cycles;
                    po.colour = heatmap_pixel(cycles, p->heatmap_scale > 0.0f ? p->heatmap_scale : 1000000.0f, po.colour);
                }
                if (out && out->radiance) { out->radiance[pix * 3] = po.colour.x; out->radiance[pix * 3 + 1] = po.colour.y; out->radiance[pix * 3 + 2] = po.colour.z; }
                if (out && out->rgba8) {
                    out->rgba8[pix * 4] = unorm8(linear_to_srgb1(po.colour.x));
                    out->rgba8[pix * 4 + 1] = unorm8(linear_to_srgb1(po.colour.y));
                    out->rgba8[pix * 4 + 2] = unorm8(linear_to_srgb1(po.colour.z));
                    out->rgba8[pix * 4 + 3] = 255;
                }
                if (out && out->hit_ids)
                    for (uint32_t s = 0; s < p->max_segments; s++)
                        for (int k = 0; k < 3; k++)
                            out->hit_ids[(pix * p->max_segments + s) * 3 + k] = s < segs_out ? po.ids[s * 3 + k] : 0xFFFFFFFFu;
            }
        }
        acc_primary += np;
        acc_shadow += ns;
    };
    int nt = std::max(1, std::min(orc_get_threads(c), nrows));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    total_primary = acc_primary;
    total_shadow = acc_shadow;
    if (out && out->ray_counts) { out->ray_counts[0] = total_primary; out->ray_counts[1] = total_shadow; }
    return 0;
}

// ---- unit-level exports for known-answer tests -------------------------------
void orc_brdf(const float* normal, const float* view, const float* light, const float* base, float rough, float metallic,
              float sun_factor, float* out3) {
    BrdfIn bi;
    bi.normal = v3(normal[0], normal[1], normal[2]); bi.view = v3(view[0], view[1], view[2]);
    bi.light = v3(light[0], light[1], light[2]); bi.base = v3(base[0], base[1], base[2]);
    bi.perceptual_roughness = rough; bi.metallic = metallic; bi.reflectance = 0.5f;
    bi.light_intensity = v3(sun_factor, sun_factor, sun_factor);
    V3 r = brdf(bi);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
// pbr.rs:52-69 restated for the GLSL that ships (eps 10.0e-10).
float orc_v_smith_ggx(const float* normal, const float* view, const float* light, float roughness) {
    V3 n = v3(normal[0], normal[1], normal[2]), vv = v3(view[0], view[1], view[2]), l = v3(light[0], light[1], light[2]);
    DotParams p;
    p.roughness = roughness;
    p.NoV = clampf(plain_dot(n, vv), 10.0e-10f, 1.0f);
    p.NoL = clampf(plain_dot(n, l), 0.0f, 1.0f);
    p.NoH = p.LoH = 0.f;
    return V_SmithGGXCorrelated(p);
}
void orc_heatmap_temperature(float heat, float* out3) {
    V3 r = heatmap_temperature(heat);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void orc_heatmap_pixel(uint32_t cycles, float scale, const float* colour3, float* out3) {
    V3 r = heatmap_pixel(cycles, scale > 0.0f ? scale : 1000000.0f, v3(colour3[0], colour3[1], colour3[2]));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
float orc_linear_to_srgb(float c) { return linear_to_srgb1(c); }
uint8_t orc_unorm8(float c) { return unorm8(c); }
void orc_blue_noise_xi(OrcContext* c, uint32_t tex, uint32_t px, uint32_t py, uint32_t iteration, uint32_t frame, float* out2) {
    V2 r = blue_noise_xi(*c, tex, px, py, iteration, frame);
    out2[0] = r.x; out2[1] = r.y;
}
void orc_sample_directional_light(const float* xi, const float* center, float radius, float* out3) {
    V3 r = sample_directional_light(V2{xi[0], xi[1]}, v3(center[0], center[1], center[2]), radius);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void orc_sample_texture(OrcContext* c, uint32_t index, float u, float v, float* out4) {
    F4 r = sample_texture(*c, index, u, v);
    out4[0] = r.r; out4[1] = r.g; out4[2] = r.b; out4[3] = r.a;
}
int orc_intersect_triangle(const float* o, const float* d, const float* a, const float* b, const float* cc, float* tuv) {
    V3 v0 = v3(a[0], a[1], a[2]);
    V3 e1 = sub3(v3(b[0], b[1], b[2]), v0), e2 = sub3(v3(cc[0], cc[1], cc[2]), v0);
    float t, u, v;
    if (!tri_candidate(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), v0, e1, e2, t, u, v)) return 0;
    tuv[0] = t; tuv[1] = u; tuv[2] = v;
    return 1;
}
void orc_invert_3x4(const float* m, float* out12) { invert_3x4(m, out12); }
void orc_primary_ray(const RtUniforms* u, uint32_t x, uint32_t y, uint32_t W, uint32_t H, float* o3, float* d3) {
    Ray r = generate_primary_ray(*u, x, y, W, H);
    o3[0] = r.o.x; o3[1] = r.o.y; o3[2] = r.o.z; d3[0] = r.d.x; d3[1] = r.d.y; d3[2] = r.d.z;
}
void orc_terminator_origin(const float* pos9, const float* nrm9, const float* bary3, const float* o2w12, float* out3) {
    TriAttr a;
    a.pa = v3(pos9[0], pos9[1], pos9[2]); a.pb = v3(pos9[3], pos9[4], pos9[5]); a.pc = v3(pos9[6], pos9[7], pos9[8]);
    a.na = v3(nrm9[0], nrm9[1], nrm9[2]); a.nb = v3(nrm9[3], nrm9[4], nrm9[5]); a.nc = v3(nrm9[6], nrm9[7], nrm9[8]);
    V3 w = v3(bary3[0], bary3[1], bary3[2]);
    V3 p = interp3(a.pa, a.pb, a.pc, w);
    V3 r = terminator_origin(a, p, w, o2w12);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
// Trace one world-space ray; returns 1 on hit and fills (inst, geom, prim) and (t,u,v).
int orc_trace(OrcContext* c, const float* o, const float* d, float tmin, float tmax, int any, uint32_t* ids3, float* tuv3) {
    Ray r{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), tmin, tmax};
    Hit h = trace(*c, r, any != 0);
    if (!h.valid) return 0;
    const Tri& tr = c->models[c->insts[h.inst].blas].tris[h.tri];
    ids3[0] = h.inst; ids3[1] = tr.geom; ids3[2] = tr.prim;
    tuv3[0] = h.t; tuv3[1] = h.u; tuv3[2] = h.v;
    return 1;
}

// any_hit_alpha_clip.glsl:11-28 on one candidate (instance, geometry, primitive, barycentrics): 1 = the candidate is kept,
// 0 = ignoreIntersectionEXT.  -1: no such instance / geometry / primitive.
int orc_anyhit_accepts(OrcContext* c, uint32_t instance_id, uint32_t geom, uint32_t prim, float u, float v) {
    if (instance_id >= c->insts.size()) return -1;
    const Inst& in = c->insts[instance_id];
    Tri tr{};
    tr.geom = geom; tr.prim = prim;
    uint32_t custom = in.rec.instance_custom_index_and_mask & 0xFFFFFFu;
    if (custom >= c->models.size() || geom >= c->models[custom].geoms.size()) return -1;
    if ((size_t)prim * 3 + 2 >= c->models[custom].geoms[geom].indices.size()) return -1;
    return anyhit_accepts(*c, in, tr, u, v) ? 1 : 0;
}

}  // extern "C"
