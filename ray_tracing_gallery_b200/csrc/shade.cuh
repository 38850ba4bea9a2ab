// shade.cuh — the reference's shader stages as device functions.
//
//   sample_texture            Vulkan sampler rules at LOD 0 (src/util_structs.rs:1306-1320;
//                             only OpImageSampleExplicitLod in the shipped .spv files)
//   anyhit_accepts            shaders/any_hit_alpha_clip.glsl:11-28
//   shade_textured_*          shaders/closest_hit_textured.glsl:13-226 + shaders/pbr.glsl:25-211
//   shade_mirror              shaders/closest_hit_mirror.glsl:11-29
//   shade_portal              shaders/ray-tracing/src/lib.rs:300-312
//   miss_colour               shaders/ray-tracing/src/lib.rs:40-51
//   linear_to_srgb / unorm8   shaders/ray-tracing/src/lib.rs:85-92, :188-190
//   primary_ray               shaders/ray-tracing/src/lib.rs:126-142
//
// Geometry that decides which primitive a later ray hits (ray generation, barycentric
// interpolation, normal transform, reflect, the shadow-terminator origin) follows the arithmetic
// contract of contract.cuh; BRDF and texture filtering are plain fp32.
#pragma once
#include "contract.cuh"
#include "rt_types.h"

namespace b200rt {

#define RT_PI 3.141592653589793f

struct F4 { float r, g, b, a; };

__device__ __forceinline__ int wrap_i(int i, int n) {
    if ((n & (n - 1)) == 0) return i & (n - 1);  // power-of-two images: two's-complement mask == REPEAT
    int m = i % n;
    return m < 0 ? m + n : m;
}

// One texel of image `k`.  `k` MUST be warp-uniform (see sample_texture).
__device__ __forceinline__ F4 fetch_texel_uniform(const SceneDev& S, uint32_t k, int x, int y) {
    const TexEntry* t = S.textures + k;
    cudaTextureObject_t obj = t->obj;
    uint32_t format = t->format;
    F4 r;
    if (format == RT_FORMAT_RGBA32_SFLOAT) {
        float4 c = tex2D<float4>(obj, (float)x + 0.5f, (float)y + 0.5f);
        r.r = c.x; r.g = c.y; r.b = c.z; r.a = c.w;
        return r;
    }
    uchar4 c = tex2D<uchar4>(obj, (float)x + 0.5f, (float)y + 0.5f);
    // exact per 8-bit code: [0,256) sRGB EOTF, [256,512) code / 255 (correctly rounded), see rt_create
    const float* lut = S.srgb_lut + (format == RT_FORMAT_RGBA8_SRGB ? 0 : 256);
    r.r = __ldg(lut + c.x); r.g = __ldg(lut + c.y); r.b = __ldg(lut + c.z);
    r.a = __ldg(S.srgb_lut + 256 + c.w);
    return r;
}

// Sampling of image `k` (warp-uniform) at per-lane coordinates.  REPEAT addressing, normalised
// coordinates, texel centres at +0.5, sRGB decode before filtering (Vulkan rules at LOD 0).
__device__ __forceinline__ F4 sample_image_uniform(const SceneDev& S, uint32_t k, float u, float v) {
    const TexEntry* t = S.textures + k;
    int w = (int)t->w, h = (int)t->h;
    if (!t->linear) {
        int x = wrap_i((int)floorf(u * (float)w), w);
        int y = wrap_i((int)floorf(v * (float)h), h);
        return fetch_texel_uniform(S, k, x, y);
    }
    float fx = u * (float)w - 0.5f, fy = v * (float)h - 0.5f;
    float flx = floorf(fx), fly = floorf(fy);
    float ax = fx - flx, ay = fy - fly;
    int x0 = wrap_i((int)flx, w), y0 = wrap_i((int)fly, h);
    int x1 = wrap_i(x0 + 1, w), y1 = wrap_i(y0 + 1, h);
    F4 t00 = fetch_texel_uniform(S, k, x0, y0), t10 = fetch_texel_uniform(S, k, x1, y0);
    F4 t01 = fetch_texel_uniform(S, k, x0, y1), t11 = fetch_texel_uniform(S, k, x1, y1);
    float w00 = (1.0f - ax) * (1.0f - ay), w10 = ax * (1.0f - ay), w01 = (1.0f - ax) * ay, w11 = ax * ay;
    F4 r;
    r.r = t00.r * w00 + t10.r * w10 + t01.r * w01 + t11.r * w11;
    r.g = t00.g * w00 + t10.g * w10 + t01.g * w01 + t11.g * w11;
    r.b = t00.b * w00 + t10.b * w10 + t01.b * w01 + t11.b * w11;
    r.a = t00.a * w00 + t10.a * w10 + t01.a * w01 + t11.a * w11;
    return r;
}

// Bindless fetch with a per-lane image index (`nonuniformEXT`, closest_hit_textured.glsl:50-52).
// A TEX instruction takes its texture header from a uniform register, so lanes holding different images have to take
// turns.  The turns are made explicit here and cost O(distinct images in the warp), not O(images in the table): the
// lanes that reached this call together OR their indices into one presence word per 32 table entries (redux.sync), and
// the loop walks only the set bits; each lane samples in the iteration that names its image.  1x1 images are constants
// in the table and need no TEX.  (The first version of the wavefront shading kernel returned texels of the wrong image
// when lanes mixed lain and fence hits, and round 1 therefore walked the whole table with a uniform counter; the
// stand-alone check of the pattern, tests/cuda/tex_divergent_handles.cu, shows every variant — including this one and
// the plain per-lane handle — correct on B200 with nvcc 12.9, and the all-table walk 1.5-1.8x slower at 96-128 images:
// profiles/r02c_tex_repro.txt.  tests/test_gpu_parity.py::test_many_images_in_one_warp holds this path to the oracle
// with 64+ real images mixed inside single warps.)
__device__ __forceinline__ F4 sample_texture(const SceneDev& S, uint32_t index, float u, float v) {
    F4 r;
    r.r = r.g = r.b = r.a = 0.0f;
    bool need = false;
    if (index < S.num_textures) {  // else: robustness2 null descriptor, src/main.rs:183-184
        const TexEntry* t = S.textures + index;
        if (t->obj == 0) {
            r.r = t->constant[0]; r.g = t->constant[1]; r.b = t->constant[2]; r.a = t->constant[3];
        } else {
            need = true;
        }
    }
    const uint32_t together = __activemask();
    for (uint32_t word = 0; word * 32u < S.num_textures; word++) {
        uint32_t present = __reduce_or_sync(together, (need && (index >> 5) == word) ? 1u << (index & 31u) : 0u);
        while (present) {
            const uint32_t k = word * 32u + (uint32_t)__ffs(present) - 1u;
            present &= present - 1u;
            if (need && k == index) r = sample_image_uniform(S, k, u, v);
        }
    }
    return r;
}
// Same, for an index that is uniform by construction (the blue-noise image named by the Uniforms).
__device__ __forceinline__ F4 sample_texture_uniform(const SceneDev& S, uint32_t index, float u, float v) {
    F4 r;
    r.r = r.g = r.b = r.a = 0.0f;
    if (index >= S.num_textures) return r;
    const TexEntry* t = S.textures + index;
    if (t->obj == 0) {
        r.r = t->constant[0]; r.g = t->constant[1]; r.b = t->constant[2]; r.a = t->constant[3];
        return r;
    }
    return sample_image_uniform(S, index, u, v);
}

// ---- bindless vertex fetch through the reference-layout ModelInfo / GeometryInfo tables
struct TriAttr {
    V3 pa, pb, pc, na, nb, nc;
    V2 ta, tb, tc;
};

__device__ __forceinline__ V3 ld_v3(const float* p, uint32_t i) { return v3(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)); }
__device__ __forceinline__ V2 ld_v2(const float* p, uint32_t i) { V2 r; r.x = __ldg(p + 2 * i); r.y = __ldg(p + 2 * i + 1); return r; }

__device__ __forceinline__ bool load_geometry(const SceneDev& S, uint32_t custom_index, uint32_t geom, RtModelInfo& mi, RtGeometryInfo& gi) {
    if (custom_index >= S.num_models) return false;
    mi = S.model_info[custom_index];
    gi = reinterpret_cast<const RtGeometryInfo*>(mi.geometry_info_address)[geom];
    return true;
}
__device__ __forceinline__ void load_indices(const RtGeometryInfo& gi, uint32_t prim, uint32_t& ia, uint32_t& ib, uint32_t& ic) {
    const uint32_t* idx = reinterpret_cast<const uint32_t*>(gi.index_buffer_address);
    ia = __ldg(idx + 3 * prim); ib = __ldg(idx + 3 * prim + 1); ic = __ldg(idx + 3 * prim + 2);
}

// any_hit_alpha_clip.glsl:11-28 — true = keep the candidate
__device__ __noinline__ bool anyhit_accepts(const SceneDev& S, uint32_t custom_index, uint32_t geom, uint32_t prim, float u, float v) {
    RtModelInfo mi; RtGeometryInfo gi;
    if (!load_geometry(S, custom_index, geom, mi, gi)) return true;
    uint32_t ia, ib, ic;
    load_indices(gi, prim, ia, ib, ic);
    const float* uvs = reinterpret_cast<const float*>(mi.uv_buffer_address);
    V2 uv = interp2(ld_v2(uvs, ia), ld_v2(uvs, ib), ld_v2(uvs, ic), bary_weights(u, v));
    float alpha = sample_texture(S, gi.images.diffuse_image_index, uv.x, uv.y).a;
    return !(alpha < 0.5f);
}

// ---- pbr.glsl
struct DotParams { float NoH, NoV, NoL, LoH, roughness; };
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float pdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float compute_f90(const DotParams& p) { return 0.5f + 2.0f * p.roughness * p.LoH * p.LoH; }
// pow(x, 5.0) of pbr.glsl:70 for x in [0,1]: three multiplications (within 2 ulp of powf; BRDF terms carry the 1e-3 tolerance)
__device__ __forceinline__ float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }
__device__ __forceinline__ float schlick1(float u, float f0, float f90) { return f0 + (f90 - f0) * pow5(1.0f - u); }

// pbr.glsl:174-211.  The result is light_intensity * NoL * (diffuse + specular) (:200-210); the
// light-independent part is returned so that the wavefront can finish the product after its shadow
// rays have been counted (shade_finish), with exactly the operations of the one-pass path.
struct BrdfTerms {
    float NoL;
    V3 comb;   // base * Fd + D * V * F
};
__device__ __forceinline__ BrdfTerms brdf_terms(V3 normal, V3 view, V3 light, V3 base, float perceptual_roughness, float metallic) {
    V3 hs = v3(view.x + light.x, view.y + light.y, view.z + light.z);
    float hl = sqrtf(pdot(hs, hs));
    V3 h = v3(hs.x / hl, hs.y / hl, hs.z / hl);
    DotParams p;
    p.roughness = perceptual_roughness * perceptual_roughness;
    p.NoV = clampf(pdot(normal, view), 10.0e-10f, 1.0f);
    p.NoH = clampf(pdot(normal, h), 0.0f, 1.0f);
    p.NoL = clampf(pdot(normal, light), 0.0f, 1.0f);
    p.LoH = clampf(pdot(light, h), 0.0f, 1.0f);
    // D_GGX, :44-51
    float a = p.NoH * p.roughness;
    float k = p.roughness / (1.0f - p.NoH * p.NoH + a * a);
    float D = k * k * (1.0f / RT_PI);
    // f0, :190-193 (perceptual_dielectric_reflectance = 0.5, closest_hit_textured.glsl:217)
    float dielectric_f0 = 0.16f * 0.5f * 0.5f;
    float f0x = dielectric_f0 * (1.0f - metallic) + base.x * metallic;
    float f0y = dielectric_f0 * (1.0f - metallic) + base.y * metallic;
    float f0z = dielectric_f0 * (1.0f - metallic) + base.z * metallic;
    float f90 = compute_f90(p);
    float fw = pow5(1.0f - p.LoH);
    float Fx = f0x + (f90 - f0x) * fw, Fy = f0y + (f90 - f0y) * fw, Fz = f0z + (f90 - f0z) * fw;
    // V_SmithGGXCorrelated, :56-65
    float a2 = p.roughness * p.roughness;
    float GGXV = p.NoL * sqrtf(p.NoV * p.NoV * (1.0f - a2) + a2);
    float GGXL = p.NoV * sqrtf(p.NoL * p.NoL * (1.0f - a2) + a2);
    float G = 0.5f / (GGXV + GGXL);
    float DG = D * G;
    // Fd_Burley, :95-103
    float fd = schlick1(p.NoL, 1.0f, f90) * schlick1(p.NoV, 1.0f, f90) * (1.0f / RT_PI);
    BrdfTerms t;
    t.NoL = p.NoL;
    t.comb = v3(base.x * fd + DG * Fx, base.y * fd + DG * Fy, base.z * fd + DG * Fz);
    return t;
}
// primary_payload.colour = brdf(params) + 0.1 * base (closest_hit_textured.glsl:222-225), unfused
__device__ __forceinline__ V3 shade_finish(float sun_factor, float NoL, V3 comb, V3 base) {
    float k = mul_(sun_factor, NoL);
    return v3(add_(mul_(k, comb.x), mul_(0.1f, base.x)), add_(mul_(k, comb.y), mul_(0.1f, base.y)), add_(mul_(k, comb.z), mul_(0.1f, base.z)));
}

// ---- closest_hit_textured.glsl
// :13-19
__device__ __forceinline__ V3 project_onto_tangent_plane(V3 point, V3 vpos, V3 vnormal) {
    V3 vtp = sub3(point, vpos);
    float dp = fminf(0.0f, dot3(vtp, vnormal));
    return v3(fmaf(-dp, vnormal.x, vtp.x), fmaf(-dp, vnormal.y, vtp.y), fmaf(-dp, vnormal.z, vtp.z));
}
// :25-39
__device__ __forceinline__ V3 terminator_origin(const TriAttr& a, V3 p, V3 w, const float* o2w) {
    V3 oa = project_onto_tangent_plane(p, a.pa, a.na);
    V3 ob = project_onto_tangent_plane(p, a.pb, a.nb);
    V3 oc = project_onto_tangent_plane(p, a.pc, a.nc);
    V3 off = interp3(oa, ob, oc, w);
    return xform_point(o2w, add3(p, off));
}
// :99-120
__device__ __forceinline__ V2 blue_noise_xi(const SceneDev& S, uint32_t tex, uint32_t px, uint32_t py, uint32_t iteration, uint32_t frame_index) {
    uint32_t ox1 = iteration * 2u * 13u, oy1 = iteration * 2u * 41u;
    uint32_t ox2 = (iteration * 2u + 1u) * 13u, oy2 = (iteration * 2u + 1u) * 41u;
    float a, b;
    const TexEntry* te = S.textures + tex;
    if (tex < S.num_textures && te->obj != 0 && te->w == 64u && te->h == 64u && !te->linear) {
        // the shipped 64x64 nearest-filtered image: coordinate k/64 selects texel k mod 64 exactly
        a = fetch_texel_uniform(S, tex, (int)((px + ox1) & 63u), (int)((py + oy1) & 63u)).r;
        b = fetch_texel_uniform(S, tex, (int)((px + ox2) & 63u), (int)((py + oy2) & 63u)).r;
    } else {
        a = sample_texture_uniform(S, tex, __fdiv_rn((float)(px + ox1), 64.0f), __fdiv_rn((float)(py + oy1), 64.0f)).r;
        b = sample_texture_uniform(S, tex, __fdiv_rn((float)(px + ox2), 64.0f), __fdiv_rn((float)(py + oy2), 64.0f)).r;
    }
    float k = mul_((float)(frame_index % 32u), 0.618033988749f);
    float sa = add_(a, k), sb = add_(b, k);
    V2 r; r.x = sub_(sa, floorf(sa)); r.y = sub_(sb, floorf(sb));
    return r;
}
// :77-94.  The tangent frame depends only on the light direction, so it is built once per thread
// (same operations, same order as the shader, which rebuilds it per sample).
struct SunFrame { V3 center, tangent, bitangent; };
__device__ __forceinline__ SunFrame make_sun_frame(const RtUniforms& U) {
    SunFrame f;
    f.center = v3(U.sun_dir[0], U.sun_dir[1], U.sun_dir[2]);
    f.tangent = normalize3(cross3(f.center, v3(0.f, 1.f, 0.f)));
    f.bitangent = normalize3(cross3(f.tangent, f.center));
    return f;
}
__device__ __forceinline__ V3 sample_directional_light(V2 rng, const SunFrame& f, float radius) {
    float r = __fsqrt_rn(rng.x);
    float angle = mul_(mul_(rng.y, 2.0f), RT_PI);
    float px = mul_(mul_(r, cosf(angle)), radius), py = mul_(mul_(r, sinf(angle)), radius);
    return normalize3(v3(add_(add_(f.center.x, mul_(px, f.tangent.x)), mul_(py, f.bitangent.x)),
                         add_(add_(f.center.y, mul_(px, f.tangent.y)), mul_(py, f.bitangent.y)),
                         add_(add_(f.center.z, mul_(px, f.tangent.z)), mul_(py, f.bitangent.z))));
}

struct TexturedHit {       // what the textured closest-hit needs after traversal
    uint32_t px, py;       // gl_LaunchIDEXT.xy
    uint32_t inst_pos;     // TLAS leaf position (-> InstRT)
    uint32_t custom_index; // gl_InstanceCustomIndexEXT and gl_InstanceID, carried from the traversal so that the
    uint32_t instance_id;  //   ModelInfo fetch does not have to wait for an InstRT read
    uint32_t geom, prim;
    float u, v;
    V3 dir;                // gl_WorldRayDirectionEXT
};

struct ShadeCtx {          // state kept between the shadow-ray phase and the BRDF phase
    TriAttr tri;
    RtGeometryInfo gi;
    V3 w;
    V2 uv;
    V3 inrm;
    uint32_t instance_id;
};

// closest_hit_textured main(), part 1: ModelInfo -> GeometryInfo -> load_triangle -> interpolate (:175-188)
__device__ __forceinline__ bool shade_textured_load(const SceneDev& S, const TexturedHit& h, ShadeCtx& c) {
    c.instance_id = h.instance_id;
    RtModelInfo mi;
    if (!load_geometry(S, h.custom_index, h.geom, mi, c.gi)) return false;
    uint32_t ia, ib, ic;
    load_indices(c.gi, h.prim, ia, ib, ic);
    const float* pos = reinterpret_cast<const float*>(mi.position_buffer_address);
    const float* nrm = reinterpret_cast<const float*>(mi.normal_buffer_address);
    const float* uvs = reinterpret_cast<const float*>(mi.uv_buffer_address);
    c.tri.pa = ld_v3(pos, ia); c.tri.pb = ld_v3(pos, ib); c.tri.pc = ld_v3(pos, ic);
    c.tri.na = ld_v3(nrm, ia); c.tri.nb = ld_v3(nrm, ib); c.tri.nc = ld_v3(nrm, ic);
    c.tri.ta = ld_v2(uvs, ia); c.tri.tb = ld_v2(uvs, ib); c.tri.tc = ld_v2(uvs, ic);
    c.w = bary_weights(h.u, h.v);
    c.inrm = interp3(c.tri.na, c.tri.nb, c.tri.nc, c.w);
    c.uv = interp2(c.tri.ta, c.tri.tb, c.tri.tc, c.w);
    return true;
}
// part 2: get_shadow_terminator_fix_shadow_origin (:193)
__device__ __forceinline__ V3 shade_textured_shadow_origin(const SceneDev& S, const ShadeCtx& c) {
    V3 ipos = interp3(c.tri.pa, c.tri.pb, c.tri.pc, c.w);
    float o2w[12];
    const float* tr = S.instances[c.instance_id].transform;
#pragma unroll
    for (int i = 0; i < 12; i++) o2w[i] = __ldg(tr + i);
    return terminator_origin(c.tri, ipos, c.w, o2w);
}
__device__ __forceinline__ bool shade_textured_begin(const SceneDev& S, const TexturedHit& h, ShadeCtx& c, V3& shadow_origin) {
    if (!shade_textured_load(S, h, c)) return false;
    shadow_origin = shade_textured_shadow_origin(S, c);
    return true;
}

// part 3: material, normal, BRDF terms (:205-221).  `base` receives the diffuse texel.
__device__ __forceinline__ BrdfTerms shade_textured_terms(const SceneDev& S, const RtUniforms& U, const TexturedHit& h, const ShadeCtx& c, V3& base) {
    F4 dcol = sample_texture(S, c.gi.images.diffuse_image_index, c.uv.x, c.uv.y);
    F4 mr = sample_texture(S, c.gi.images.metallic_roughness_image_index, c.uv.x, c.uv.y);
    float metallic = mr.b, roughness = mr.g;  // `.bg` swizzle, closest_hit_textured.glsl:54-60
    float inv[12];
    const float* ip = S.inst_rt[h.inst_pos].inv;
#pragma unroll
    for (int i = 0; i < 12; i++) inv[i] = __ldg(ip + i);
    V3 normal;
    if (c.gi.images.normal_map_image_index < 0) {
        normal = normalize3(xform_normal(inv, c.inrm));
    } else {
        F4 nm = sample_texture(S, (uint32_t)c.gi.images.normal_map_image_index, c.uv.x, c.uv.y);
        V3 mn = v3(nm.r * 2.0f - 1.0f, nm.g * 2.0f - 1.0f, nm.b * 2.0f - 1.0f);
        // compute_cotangent_frame, :123-139
        V3 dp1 = sub3(c.tri.pb, c.tri.pa), dp2 = sub3(c.tri.pc, c.tri.pa);
        float du1x = c.tri.tb.x - c.tri.ta.x, du1y = c.tri.tb.y - c.tri.ta.y;
        float du2x = c.tri.tc.x - c.tri.ta.x, du2y = c.tri.tc.y - c.tri.ta.y;
        V3 dp2perp = cross3(dp2, c.inrm), dp1perp = cross3(c.inrm, dp1);
        V3 T = add3(scale3(dp2perp, du1x), scale3(dp1perp, du2x));
        V3 B = add3(scale3(dp2perp, du1y), scale3(dp1perp, du2y));
        float invmax = 1.0f / sqrtf(fmaxf(dot3(T, T), dot3(B, B)));
        V3 mnn = normalize3(mn);
        V3 Ts = scale3(T, invmax), Bs = scale3(B, invmax);
        V3 local = v3(Ts.x * mnn.x + Bs.x * mnn.y + c.inrm.x * mnn.z, Ts.y * mnn.x + Bs.y * mnn.y + c.inrm.y * mnn.z,
                      Ts.z * mnn.x + Bs.z * mnn.y + c.inrm.z * mnn.z);
        normal = normalize3(xform_normal(inv, local));
    }
    V3 sun = v3(U.sun_dir[0], U.sun_dir[1], U.sun_dir[2]);
    base = v3(dcol.r, dcol.g, dcol.b);
    return brdf_terms(normal, v3(-h.dir.x, -h.dir.y, -h.dir.z), sun, base, roughness, metallic);
}

// closest_hit_mirror.glsl:11-29.  Returns false if the model tables cannot be resolved.
__device__ __forceinline__ bool shade_mirror(const SceneDev& S, uint32_t inst_pos, uint32_t geom, uint32_t prim, float u, float v,
                                             float t, V3 o, V3 d, V3& new_o, V3& new_d) {
    const InstRT* ir = S.inst_rt + inst_pos;
    uint32_t custom = __ldg(&ir->custom_sbt) & 0xFFFFFFu;
    RtModelInfo mi; RtGeometryInfo gi;
    if (!load_geometry(S, custom, geom, mi, gi)) return false;
    uint32_t ia, ib, ic;
    load_indices(gi, prim, ia, ib, ic);
    const float* nrm = reinterpret_cast<const float*>(mi.normal_buffer_address);
    V3 n = interp3(ld_v3(nrm, ia), ld_v3(nrm, ib), ld_v3(nrm, ic), bary_weights(u, v));
    float inv[12];
#pragma unroll
    for (int i = 0; i < 12; i++) inv[i] = __ldg(ir->inv + i);
    n = normalize3(xform_normal(inv, n));
    float k = mul_(2.0f, dot3(n, d));  // reflect(I,N) = I - 2 dot(N,I) N
    new_d = v3(fmaf(-k, n.x, d.x), fmaf(-k, n.y, d.y), fmaf(-k, n.z, d.z));
    new_o = v3(fmaf(d.x, t, o.x), fmaf(d.y, t, o.y), fmaf(d.z, t, o.z));
    return true;
}

// lib.rs:300-312
__device__ __forceinline__ void shade_portal(float t, V3 o, V3 d, V3& new_o, V3& new_d) {
    new_d = d;
    new_o = v3(fmaf(d.x, t, o.x), add_(fmaf(d.y, t, o.y), 5.0f), fmaf(d.z, t, o.z));
}

// lib.rs:40-51 (cos(sun_radius) is evaluated once on the host)
__device__ __forceinline__ V3 miss_colour(const RtUniforms& U, float cos_sun_radius, V3 d) {
    V3 sun = v3(U.sun_dir[0], U.sun_dir[1], U.sun_dir[2]);
    if (dot3(d, sun) > cos_sun_radius) return v3(1.0f, 1.0f, 1.0f);
    return v3(0.0f, 0.0f, 0.05f);
}

// lib.rs:85-92
__device__ __forceinline__ float linear_to_srgb1(float c) {
    return c <= 0.0031308f ? c * 12.92f : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}
// UNORM8 image store: clamp, round to nearest, NaN -> 0
__device__ __forceinline__ uint32_t unorm8(float c) {
    if (!(c > 0.0f)) return 0u;
    if (c >= 1.0f) return 255u;
    return (uint32_t)floorf(c * 255.0f + 0.5f);
}

// heatmap.rs:5-54 — `heatmap_temperature`, every operation rounded once (contract arithmetic, no fusing).
// The reference indexes `colours[heat as i32]`, which is one past the table at heat == 1.0 exactly
// (undefined in SPIR-V); `cur` is clamped to the last entry here and in the oracle.
__device__ __forceinline__ float saturate1(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
__device__ __forceinline__ float smoothstep1(float e0, float e1, float x) {
    float t = saturate1(div_(sub_(x, e0), sub_(e1, e0)));
    return mul_(mul_(t, t), sub_(3.0f, mul_(2.0f, t)));
}
__device__ __forceinline__ V3 heatmap_temperature(float heat) {
    const float k[10][3] = {
        {0.0f / 255.0f, 2.0f / 255.0f, 91.0f / 255.0f},    {0.0f / 255.0f, 108.0f / 255.0f, 251.0f / 255.0f},
        {0.0f / 255.0f, 221.0f / 255.0f, 221.0f / 255.0f}, {51.0f / 255.0f, 221.0f / 255.0f, 0.0f / 255.0f},
        {255.0f / 255.0f, 252.0f / 255.0f, 0.0f / 255.0f}, {255.0f / 255.0f, 180.0f / 255.0f, 0.0f / 255.0f},
        {255.0f / 255.0f, 104.0f / 255.0f, 0.0f / 255.0f}, {226.0f / 255.0f, 22.0f / 255.0f, 0.0f / 255.0f},
        {191.0f / 255.0f, 0.0f / 255.0f, 83.0f / 255.0f},  {145.0f / 255.0f, 0.0f / 255.0f, 65.0f / 255.0f}};
    heat = mul_(saturate1(heat), 10.0f);
    const int idx = (int)heat;
    const int cur = min(idx, 9), prv = max(idx - 1, 0), nxt = min(idx + 1, 9);
    const float lo = floorf(heat), hi = ceilf(heat), blur = 0.8f;
    const float s_lo = smoothstep1(sub_(lo, blur), add_(lo, blur), heat);
    const float s_hi = smoothstep1(sub_(hi, blur), add_(hi, blur), heat);
    const float wc = mul_(s_lo, sub_(1.0f, s_hi)), wp = sub_(1.0f, s_lo), wn = s_hi;
    V3 r = add3(add3(scale3(v3(k[cur][0], k[cur][1], k[cur][2]), wc), scale3(v3(k[prv][0], k[prv][1], k[prv][2]), wp)),
                scale3(v3(k[nxt][0], k[nxt][1], k[nxt][2]), wn));
    return v3(saturate1(r.x), saturate1(r.y), saturate1(r.z));
}
// lib.rs:174-186: `heatmap_temperature(delta_time as f32 / heatmap_scale) + payload.colour * 0.000001`
__device__ __forceinline__ V3 heatmap_pixel(uint32_t cycles, float scale, V3 colour) {
    return add3(heatmap_temperature(div_((float)cycles, scale)), scale3(colour, 0.000001f));
}

// lib.rs:126-142
__device__ __forceinline__ void primary_ray(const RtUniforms& U, uint32_t x, uint32_t y, uint32_t W, uint32_t H, V3& o, V3& d) {
    float pcx = add_((float)x, 0.5f), pcy = add_((float)y, 0.5f);
    float ndx = sub_(mul_(div_(pcx, (float)W), 2.0f), 1.0f), ndy = sub_(mul_(div_(pcy, (float)H), 2.0f), 1.0f);
    const float* V = U.view_inverse;
    const float* P = U.proj_inverse;
    o = v3(V[12], V[13], V[14]);
    V3 target = v3(fmaf(P[12], 1.0f, fmaf(P[8], 1.0f, fmaf(P[4], ndy, mul_(P[0], ndx)))),
                   fmaf(P[13], 1.0f, fmaf(P[9], 1.0f, fmaf(P[5], ndy, mul_(P[1], ndx)))),
                   fmaf(P[14], 1.0f, fmaf(P[10], 1.0f, fmaf(P[6], ndy, mul_(P[2], ndx)))));
    V3 ld = normalize3(target);
    d = v3(fmaf(V[8], ld.z, fmaf(V[4], ld.y, mul_(V[0], ld.x))), fmaf(V[9], ld.z, fmaf(V[5], ld.y, mul_(V[1], ld.x))),
           fmaf(V[10], ld.z, fmaf(V[6], ld.y, mul_(V[2], ld.x))));
}

}  // namespace b200rt
