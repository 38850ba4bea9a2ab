// render.cu — the frame: what one vkCmdTraceRaysKHR(width, height, 1) does in the reference
// (src/command_buffer_recording.rs:116-126 -> ray_generation, shaders/ray-tracing/src/lib.rs:94-191).
//
// Wavefront pipeline (default).  One frame = counter memset + five launches:
//   k_sun_dirs   the frame's 4096 x N shadow-ray directions (blue noise repeats every 64 pixels).
//   k_trace0     ray generation + closest-hit traversal; miss -> sky colour to the framebuffer;
//                mirror/portal -> next ray pushed to the compacted ray queue; textured hit ->
//                HitRec pushed to the compacted hit queue (warp ballot + one atomic per warp).
//   k_prep       one thread per queued textured hit: triangle fetch, shadow-terminator origin,
//                textures, normal, BRDF terms; the HitRec is rewritten in place.
//   k_shadow     one thread per shadow ray (hit x sample): first-hit traversal; unshadowed rays
//                are counted into HitRec.lit.
//   k_tail       cooperative: segment 0's resolve phase (sun_factor, final colour, sRGB encode,
//                framebuffer store), then ray-gen segments >= 1 (bounce rays) as trace / prep /
//                shadow / resolve phases separated by grid barriers, then the ray-count export.
//                k_trace_n / k_resolve are the same phases as separate launches (RT_RENDER_SPLIT_TAIL).
// Small kernels on purpose: the one-pass shade kernel of the first version was instruction-fetch
// bound (profiles/r01a_shade_details.txt: 31 % of issue stalls "no instruction", 128 registers).
// Traversal kernels are persistent: warps pull 32-item batches from a device-side cursor, so a launch
// never needs a queue length on the host; stages chain with programmatic dependent launches.
// Megakernel (A/B baseline): one thread per pixel runs the whole segment loop.
#include <cooperative_groups.h>

#include "launch_count.h"
#include "render.h"
#include "trace.cuh"

namespace cg = cooperative_groups;

namespace b200rt {
namespace {

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialisation attribute may
// be scheduled while its predecessor in the stream is still draining; pdl_wait() blocks until the predecessor
// has completed and its writes are visible, pdl_release() lets the successor start being scheduled.  Both are
// no-ops for plain launches.  Used so that launch latency and block scheduling of the next stage hide behind
// the tail of the current one.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;"); }

// compact local row -> global y (strips dealt round-robin to ranks)
__device__ __forceinline__ uint32_t global_y(const FrameDev& F, uint32_t ly) {
    if (F.strip_count > 1 && F.strip_height > 0) {
        uint32_t j = ly / F.strip_height, r = ly - j * F.strip_height;
        return F.y0 + (j * F.strip_count + F.strip_index) * F.strip_height + r;
    }
    return F.y0 + ly;
}

// work item (8x4 pixel tile per warp) -> local pixel
__device__ __forceinline__ bool item_to_pixel(const FrameDev& F, uint32_t item, uint32_t& lx, uint32_t& ly) {
    uint32_t tiles_x = (F.tw + 7u) >> 3;
    uint32_t tile = item >> 5, lane = item & 31u;
    uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
    lx = tx * 8u + (lane & 7u);
    ly = ty * 4u + (lane >> 3);
    return lx < F.tw && ly < F.rows;
}

__device__ __forceinline__ uint32_t encode_rgba8(V3 c) {
    uint32_t r = unorm8(linear_to_srgb1(c.x)), g = unorm8(linear_to_srgb1(c.y)), b = unorm8(linear_to_srgb1(c.z));
    return r | (g << 8) | (b << 16) | 0xFF000000u;
}
// Output index of compact pixel `pixel`.  Default: the compact index itself (outputs hold the rendered rows only).
// RT_RENDER_OUTPUT_IMAGE_ROWS: outputs are whole tile images and this call fills its own rows in place, so that
// several ranks can store into ONE frame (e.g. rank 0's, through NVLink peer memory) with no gather afterwards.
__device__ __forceinline__ size_t out_index(const FrameDev& F, uint32_t pixel) {
    if (!F.image_rows) return pixel;
    uint32_t ly = pixel / F.tw, lx = pixel - ly * F.tw;
    return (size_t)(global_y(F, ly) - F.y0) * F.tw + lx;
}

__device__ __forceinline__ void write_pixel(const FrameDev& F, uint32_t cpixel, V3 c) {
    const size_t pixel = out_index(F, cpixel);
    if (F.radiance) {
        F.radiance[3 * pixel] = c.x;
        F.radiance[3 * pixel + 1] = c.y;
        F.radiance[3 * pixel + 2] = c.z;
    }
    if (F.rgba8) {
        reinterpret_cast<uint32_t*>(F.rgba8)[pixel] = encode_rgba8(c);
    }
}

// same store with the sRGB encode already done (the two miss colours are constants of the frame)
__device__ __forceinline__ void write_pixel_encoded(const FrameDev& F, uint32_t cpixel, V3 c, uint32_t rgba) {
    const size_t pixel = out_index(F, cpixel);
    if (F.radiance) {
        F.radiance[3 * pixel] = c.x;
        F.radiance[3 * pixel + 1] = c.y;
        F.radiance[3 * pixel + 2] = c.z;
    }
    if (F.rgba8) reinterpret_cast<uint32_t*>(F.rgba8)[pixel] = rgba;
}

__device__ __forceinline__ void write_hit_ids(const FrameDev& F, uint32_t pixel, uint32_t seg, const Hit& h) {
    if (!F.hit_ids) return;
    uint32_t* p = F.hit_ids + ((size_t)pixel * F.max_segments + seg) * 3;
    p[0] = h.instance_id; p[1] = h.geom; p[2] = h.prim;
}

// Queue push: one atomicAdd per warp (ballot + popc prefix); every lane of the warp must call both halves.
// Two halves, so that independent work can run while the atomic is in flight:
// reserve() issues it (leader lane), slot() reads the result.
struct WarpPush {
    uint32_t mask, base;
    __device__ __forceinline__ void reserve(unsigned int* counter, bool pred) {
        mask = __ballot_sync(0xFFFFFFFFu, pred);
        base = 0;
        if (mask && lane_id() == (uint32_t)(__ffs(mask) - 1)) base = atomicAdd(counter, __popc(mask));
    }
    __device__ __forceinline__ uint32_t slot() const {
        if (mask == 0) return 0;
        return __shfl_sync(0xFFFFFFFFu, base, __ffs(mask) - 1) + __popc(mask & ((1u << lane_id()) - 1u));
    }
};

__device__ __forceinline__ void warp_add(unsigned long long* counter, uint32_t v) {
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if (lane_id() == 0 && v) atomicAdd(counter, (unsigned long long)v);
}

// `phase`: 0 = closest-hit rays, 1 = shadow rays
template <bool COUNT>
__device__ __forceinline__ void flush_trace_counters(const FrameDev& F, int phase, const TraceCounters& tc) {
    if (COUNT) {
        warp_add(&F.counters->nodes_visited[phase], tc.nodes);
        warp_add(&F.counters->instances_entered[phase], tc.instances);
        warp_add(&F.counters->triangles_tested[phase], tc.tris);
        warp_add(&F.counters->anyhit_calls[phase], tc.anyhits);
    }
    uint32_t ov = __reduce_add_sync(0xFFFFFFFFu, tc.overflow);
    if (lane_id() == 0 && ov) {
        atomicAdd(&F.counters->stack_overflow, ov);
        if (F.overflow_flag) *F.overflow_flag = 1u;  // a subtree was skipped: the API must not return this frame as good
    }
}

__device__ __forceinline__ void flush_ray_counters(const FrameDev& F, uint32_t n_primary, uint32_t n_shadow, uint32_t n_textured) {
    warp_add(&F.counters->primary_rays, n_primary);
    warp_add(&F.counters->shadow_rays, n_shadow);
    warp_add(&F.counters->textured_hits, n_textured);
}

// cast_shadow_ray for sample `i` of a textured hit (closest_hit_textured.glsl:159-172, :195-201):
// TerminateOnFirstHit | SkipClosestHit, tmin 0.001, tmax 10000.  Returns 1 - shadowed.
template <bool COUNT>
__device__ __forceinline__ bool shadow_sample_lit(const SceneDev& S, const FrameDev& F, const SunFrame& sun, uint32_t px, uint32_t py,
                                                  uint32_t i, V3 origin, TraceCounters& tc) {
    V3 dir;
    if (F.sun_dirs) {
        float4 t = __ldg(F.sun_dirs + ((size_t)(((py & 63u) << 6) | (px & 63u)) * F.shadow_rays + i));
        dir = v3(t.x, t.y, t.z);
    } else {
        V2 xi = blue_noise_xi(S, F.uniforms.blue_noise_texture_index, px, py, i, F.uniforms.frame_index);
        dir = sample_directional_light(xi, sun, F.uniforms.sun_radius);
    }
    Hit sh;
    return !trace_ray<true, COUNT>(S, origin, dir, 0.001f, 10000.0f, sh, tc);
}

// The textured closest-hit in one pass (megakernel).  closest_hit_textured.glsl:174-226
template <bool COUNT>
__device__ __forceinline__ V3 shade_textured(const SceneDev& S, const FrameDev& F, const TexturedHit& th, uint32_t& n_shadow, TraceCounters& tc /* shadow-ray counters */) {
    ShadeCtx ctx;
    if (!shade_textured_load(S, th, ctx)) return v3(0.f, 0.f, 0.f);
    V3 shadow_origin = shade_textured_shadow_origin(S, ctx);
    SunFrame sun = make_sun_frame(F.uniforms);
    uint32_t lit = 0;
    for (uint32_t i = 0; i < F.shadow_rays; i++) lit += shadow_sample_lit<COUNT>(S, F, sun, th.px, th.py, i, shadow_origin, tc) ? 1u : 0u;
    n_shadow += F.shadow_rays;
    float sun_factor = div_((float)lit, (float)F.shadow_rays);
    V3 base;
    BrdfTerms bt = shade_textured_terms(S, F.uniforms, th, ctx, base);
    return shade_finish(sun_factor, bt.NoL, bt.comb, base);
}

// ------------------------------------------------------------------------------------ wavefront
enum { K_TRACE = 0, K_PREP = 1, K_SHADOW = 2, K_RESOLVE = 3, K_MEGA = 4, K_TAIL = 5 };

#ifndef RT_TRACE_MINB
#define RT_TRACE_MINB 6   // resident blocks per SM asked of k_trace0: 5 / 6 / 7 measured after the bf16 box test, profiles/r02y_ab.txt
#endif
#ifndef RT_TRACE_N_MINB
#define RT_TRACE_N_MINB 6   // same for the bounce-segment launches (k_trace_n): C4 -3 %
#endif
#ifndef RT_SHADOW_MINB
#define RT_SHADOW_MINB 7
#endif
#define RT_CHUNK 32u   // work items a warp takes from a queue cursor per atomic (128 measured slower: coarser tail)

__device__ __forceinline__ SegCounters* seg_counters(const FrameDev& F, uint32_t seg) { return &F.counters->seg[seg & (RT_SEG_SLOTS - 1u)]; }

// Warp-wide work distribution: a warp takes RT_CHUNK items at a time from the device-side cursor with
// one atomic.  The atomic for the NEXT batch is issued before the current batch is traced and its
// result is only read (shuffled) when that batch starts, so the round trip to L2 (about 20 % of
// k_trace0's stall samples in profiles/r01g) hides behind the traversal.  Dealing most batches statically and
// leaving only the tail to the cursor was measured 9 % slower (profiles/r02a_ab.txt, variant ss6): the cursor is
// what balances rays of very different length.
struct WarpChunk {
    uint32_t pending;  // lane 0: base of the next batch, possibly still in flight
    __device__ __forceinline__ void init(unsigned int* cursor, uint32_t /*total*/) { pending = lane_id() == 0 ? atomicAdd(cursor, RT_CHUNK) : 0u; }
    // base of the next batch into `base`; false when the queue has run dry
    __device__ __forceinline__ bool next(unsigned int* cursor, uint32_t total, uint32_t& base) {
        base = __shfl_sync(0xFFFFFFFFu, pending, 0);
        if (base >= total) return false;
        if (lane_id() == 0) pending = atomicAdd(cursor, RT_CHUNK);
        return true;
    }
};

// Ray generation (segment 0) or ray-queue read, closest-hit traversal, miss / mirror / portal shaders,
// textured hits -> hit queue.
template <bool SEG0, bool COUNT>
__device__ __forceinline__ void trace_phase(const SceneDev& S, const FrameDev& F, uint32_t seg, uint32_t total) {
    TraceCounters tc = {0, 0, 0, 0, 0};
    uint32_t n_primary = 0;
    const RayRec* __restrict__ in_q = F.ray_q[(seg + 1u) & 1u];
    RayRec* __restrict__ out_q = F.ray_q[seg & 1u];
    SegCounters* sc = seg_counters(F, seg);
    const uint32_t rgba_sky = encode_rgba8(v3(0.0f, 0.0f, 0.05f)), rgba_sun = encode_rgba8(v3(1.0f, 1.0f, 1.0f));
    WarpChunk wc;
    wc.init(&sc->work_next[K_TRACE], total);
    uint32_t batch;
    while (wc.next(&sc->work_next[K_TRACE], total, batch)) {
        uint32_t item = batch + lane_id();
        bool active = false;
        uint32_t pixel = 0;
        V3 o = v3(0, 0, 0), d = v3(0, 0, 1);
        if (SEG0) {
            uint32_t lx, ly;
            active = item < total && item_to_pixel(F, item, lx, ly);
            if (active) {
                pixel = ly * F.tw + lx;
                primary_ray(F.uniforms, F.x0 + lx, global_y(F, ly), F.width, F.height, o, d);
            }
        } else {
            active = item < total;
            if (active) {
                const float4* rp = reinterpret_cast<const float4*>(in_q + item);
                float4 a = __ldcg(rp), b = __ldcg(rp + 1);
                o = v3(a.x, a.y, a.z); pixel = __float_as_uint(a.w);
                d = v3(b.x, b.y, b.z);
            }
        }
        Hit h;
        bool got = false;
        if (active) {
            got = trace_ray<false, COUNT, SEG0>(S, o, d, 0.01f, 10000.0f, h, tc);  // lib.rs:151-163
            n_primary++;
        }
        uint32_t kind = got ? (h.custom_sbt >> 24) : 0xFFFFFFFFu;  // hit group = instance sbt offset (main.rs:289-305)
        bool textured = active && got && kind == RT_HIT_TEXTURED;
        bool bounce = active && got && (kind == RT_HIT_MIRROR || kind == RT_HIT_PORTAL);
        V3 no = o, ndir = d;
        if (active && got) write_hit_ids(F, pixel, seg, h);
        if (bounce) {
            if (kind == RT_HIT_MIRROR) {
                if (!shade_mirror(S, h.inst_pos, h.geom, h.prim, h.u, h.v, h.t, o, d, no, ndir)) bounce = false;
            } else {
                shade_portal(h.t, o, d, no, ndir);
            }
            // `if payload.new_ray_direction == 0 break` (lib.rs:165-171)
            if (bounce && ndir.x == 0.0f && ndir.y == 0.0f && ndir.z == 0.0f) bounce = false;
            if (bounce && seg + 1u >= F.max_segments) bounce = false;  // loop bound reached: colour stays 0
        }
        WarpPush hp, rp;
        hp.reserve(&sc->hit_count, textured);   // the two queue reservations are in flight ...
        rp.reserve(&sc->ray_count, bounce);
        if (active && !textured && !bounce) {    // ... while the finished pixels are encoded and stored
            if (got) {
                write_pixel_encoded(F, pixel, v3(0.f, 0.f, 0.f), 0xFF000000u);  // segment budget used up: colour stays 0
            } else {
                V3 c = miss_colour(F.uniforms, F.cos_sun_radius, d);  // primary_ray_miss: sun disc or SKY_COLOUR (lib.rs:38-51)
                write_pixel_encoded(F, pixel, c, c.x == 1.0f ? rgba_sun : rgba_sky);
            }
        }
        uint32_t hslot = hp.slot();
        if (textured) {
            HitRec* hr = F.hit_q + hslot;
            reinterpret_cast<uint4*>(hr)[0] = make_uint4(pixel, h.inst_pos, h.geom, h.prim);
            reinterpret_cast<float4*>(hr)[1] = make_float4(h.u, h.v, __uint_as_float(h.custom_sbt & 0xFFFFFFu), 0.f);
            reinterpret_cast<float4*>(hr)[2] = make_float4(d.x, d.y, d.z, __uint_as_float(h.instance_id));
        }
        uint32_t rslot = rp.slot();
        if (bounce) {
            float4* rq = reinterpret_cast<float4*>(out_q + rslot);
            rq[0] = make_float4(no.x, no.y, no.z, __uint_as_float(pixel));
            rq[1] = make_float4(ndir.x, ndir.y, ndir.z, 0.f);
        }
    }
    flush_ray_counters(F, n_primary, 0, 0);
    flush_trace_counters<COUNT>(F, 0, tc);
}

// closest_hit_textured.glsl:174-221 for every queued hit, minus the shadow rays: triangle fetch,
// shadow-terminator origin, material textures, normal, BRDF terms.  Rewrites the HitRec in place.
// Uniform cost per item: plain grid-stride loop, no cursor.
__device__ __forceinline__ void prep_phase(const SceneDev& S, const FrameDev& F, uint32_t seg) {
    const uint32_t total = *((volatile unsigned int*)&seg_counters(F, seg)->hit_count);
    uint32_t n_textured = 0;
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
        HitRec* hr = F.hit_q + item;
        uint4 a = __ldcg(reinterpret_cast<const uint4*>(hr));
        float4 b = __ldcg(reinterpret_cast<const float4*>(hr) + 1);
        float4 c = __ldcg(reinterpret_cast<const float4*>(hr) + 2);
        TexturedHit th;
        uint32_t ly = a.x / F.tw, lx = a.x - ly * F.tw;
        th.px = F.x0 + lx; th.py = global_y(F, ly);
        th.inst_pos = a.y; th.geom = a.z; th.prim = a.w;
        th.u = b.x; th.v = b.y;
        th.custom_index = __float_as_uint(b.z); th.instance_id = __float_as_uint(c.w);
        th.dir = v3(c.x, c.y, c.z);
        ShadeCtx ctx;
        float4 r0 = make_float4(__uint_as_float(a.x), 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, 0.f), r2 = r1, r3 = r1;
        if (shade_textured_load(S, th, ctx)) {
            V3 so = shade_textured_shadow_origin(S, ctx);
            V3 base;
            BrdfTerms bt = shade_textured_terms(S, F.uniforms, th, ctx, base);
            r0.y = bt.NoL;
            r1 = make_float4(bt.comb.x, bt.comb.y, bt.comb.z, 0.f);  // .w: lit = 0
            r2 = make_float4(base.x, base.y, base.z, 0.f);
            // .w: valid bit | gl_LaunchIDEXT.x << 1 | gl_LaunchIDEXT.y << 17 (k_shadow's blue-noise cell without a division)
            r3 = make_float4(so.x, so.y, so.z, __uint_as_float(1u | (th.px << 1) | (th.py << 17)));
        }
        reinterpret_cast<float4*>(hr)[0] = r0;
        reinterpret_cast<float4*>(hr)[1] = r1;
        reinterpret_cast<float4*>(hr)[2] = r2;
        reinterpret_cast<float4*>(hr)[3] = r3;
        n_textured++;
    }
    // every lane of the warp must reach the warp-wide flush
    flush_ray_counters(F, 0, 0, n_textured);
}

// One thread per shadow ray: item = hit * shadow_rays + sample, so the samples of one hit sit in
// adjacent lanes (same origin, directions inside the sun's cone: coherent traversal).  A warp runs
// its 32 rays to completion before taking the next 32: replacing finished rays lane by lane
// (persistent threads with ray replacement) was measured slower on C2 — profiles/r01_notes.md.
template <bool COUNT>
__device__ __forceinline__ void shadow_phase(const SceneDev& S, const FrameDev& F, uint32_t seg) {
    TraceCounters tc = {0, 0, 0, 0, 0};
    SegCounters* sc = seg_counters(F, seg);
    const uint32_t n = F.shadow_rays;
    const uint32_t total = *((volatile unsigned int*)&sc->hit_count) * n;
    const SunFrame sun = make_sun_frame(F.uniforms);
    const bool pow2 = (n & (n - 1u)) == 0u;
    const uint32_t shift = 31u - __clz(n);
    uint32_t n_shadow = 0;
    WarpChunk wc;
    wc.init(&sc->work_next[K_SHADOW], total);
    uint32_t batch;
    while (wc.next(&sc->work_next[K_SHADOW], total, batch)) {
        uint32_t item = batch + lane_id();
        if (item < total) {
            uint32_t hi = pow2 ? item >> shift : item / n, i = item - hi * n;
            HitRec* hr = F.hit_q + hi;
            float4 so = __ldg(reinterpret_cast<const float4*>(hr) + 3);
            const uint32_t w = __float_as_uint(so.w);
            if (w & 1u) {
                bool lit = shadow_sample_lit<COUNT>(S, F, sun, (w >> 1) & 0xFFFFu, w >> 17, i, v3(so.x, so.y, so.z), tc);
                if (lit) atomicAdd(&hr->lit, 1u);  // shadow_ray_miss: shadowed = false (lib.rs:33-36)
                n_shadow++;
            }
        }
    }
    flush_ray_counters(F, 0, n_shadow, 0);
    flush_trace_counters<COUNT>(F, 1, tc);
}

// sun_factor = lit / N, colour = sun_factor * NoL * comb + 0.1 * base (closest_hit_textured.glsl:203, :222-225),
// then the ray-gen tail: linear_to_srgb + image store (lib.rs:188-190).
__device__ __forceinline__ void resolve_phase(const FrameDev& F, uint32_t seg) {
    const uint32_t total = *((volatile unsigned int*)&seg_counters(F, seg)->hit_count);
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
        const HitRec* hr = F.hit_q + item;
        float4 r0 = __ldcg(reinterpret_cast<const float4*>(hr));
        float4 r1 = __ldcg(reinterpret_cast<const float4*>(hr) + 1);
        float4 r2 = __ldcg(reinterpret_cast<const float4*>(hr) + 2);
        uint32_t valid = __ldcg(&hr->shadow_valid) & 1u;
        V3 col = v3(0.f, 0.f, 0.f);
        if (valid) {
            float sun_factor = div_((float)__float_as_uint(r1.w), (float)F.shadow_rays);
            if (seg == 0 && F.sun_factor) sun_factor = __ldcg(F.sun_factor + __float_as_uint(r0.x));  // what the denoise hook left
            col = shade_finish(sun_factor, r0.y, v3(r1.x, r1.y, r1.z), v3(r2.x, r2.y, r2.z));
        }
        write_pixel(F, __float_as_uint(r0.x), col);
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(128, RT_TRACE_MINB) k_trace0(SceneDev S, FrameDev F, uint32_t total) {
    pdl_release();
    pdl_wait();
    trace_phase<true, COUNT>(S, F, 0, total);
}
// bounce segment as its own launch (RT_RENDER_SPLIT_TAIL): the ray count is read on the device
template <bool COUNT>
__global__ void __launch_bounds__(128, RT_TRACE_N_MINB) k_trace_n(SceneDev S, FrameDev F, uint32_t seg) {
    trace_phase<false, COUNT>(S, F, seg, *((volatile unsigned int*)&seg_counters(F, seg - 1u)->ray_count));
}
__global__ void __launch_bounds__(128) k_prep(SceneDev S, FrameDev F, uint32_t seg) {
    pdl_release();
    pdl_wait();
    prep_phase(S, F, seg);
}
template <bool COUNT>
__global__ void __launch_bounds__(128, RT_SHADOW_MINB) k_shadow(SceneDev S, FrameDev F, uint32_t seg) {
    pdl_release();
    pdl_wait();
    shadow_phase<COUNT>(S, F, seg);
}
// How many bounce rays segment 0 queued: the host's hint for the tail policy of the frames that follow (host-mapped
// word: no copy, no synchronisation).  Called by the first kernel after segment 0's shadow rays, before any counter slot
// can be reused.
__device__ __forceinline__ void export_bounce_hint(const FrameDev& F) {
    if (F.bounce_hint && blockIdx.x == 0 && threadIdx.x == 0) *F.bounce_hint = *((volatile unsigned int*)&seg_counters(F, 0)->ray_count);
}
__global__ void __launch_bounds__(128) k_resolve(FrameDev F, uint32_t seg) {
    if (seg == 0) export_bounce_hint(F);
    resolve_phase(F, seg);
}

// Denoise hook, segment 0: the sun factor of every textured hit (closest_hit_textured.glsl:203, the same division the resolve
// phase does) and its guide values in image space, compact pixel order; pixels without a hit keep the 1.0 / 0 fill.
__global__ void __launch_bounds__(256) k_fill_sun_factor(float* sun_factor, float4* position_nol, uint32_t pixels) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += gridDim.x * blockDim.x) {
        sun_factor[i] = 1.0f;
        position_nol[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
__global__ void __launch_bounds__(128) k_export_sun_factor(FrameDev F) {
    const uint32_t total = *((volatile unsigned int*)&seg_counters(F, 0)->hit_count);
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
        const HitRec* hr = F.hit_q + item;
        const float4 r0 = __ldcg(reinterpret_cast<const float4*>(hr));
        const float4 r1 = __ldcg(reinterpret_cast<const float4*>(hr) + 1);
        const float4 r3 = __ldcg(reinterpret_cast<const float4*>(hr) + 3);
        if (!(__float_as_uint(r3.w) & 1u)) continue;
        const uint32_t pixel = __float_as_uint(r0.x);
        F.sun_factor[pixel] = div_((float)__float_as_uint(r1.w), (float)F.shadow_rays);
        F.position_nol[pixel] = make_float4(r3.x, r3.y, r3.z, r0.y);
    }
}
// rt_denoise_bilateral: 5x5 cross-bilateral average of the sun factor
__global__ void __launch_bounds__(256) k_denoise_bilateral(const float* __restrict__ in, float* __restrict__ out, const float4* __restrict__ guide, uint32_t w,
                                                           uint32_t h, float inv_sigma2) {
    const uint32_t x = blockIdx.x * 16u + (threadIdx.x & 15u), y = blockIdx.y * 16u + (threadIdx.x >> 4);
    if (x >= w || y >= h) return;
    const float4 g0 = guide[(size_t)y * w + x];
    float acc = 0.f, wsum = 0.f;
    const bool hit0 = g0.x != 0.f || g0.y != 0.f || g0.z != 0.f || g0.w != 0.f;
    if (!hit0) { out[(size_t)y * w + x] = in[(size_t)y * w + x]; return; }
    for (int dy = -2; dy <= 2; dy++)
        for (int dx = -2; dx <= 2; dx++) {
            const int xx = (int)x + dx, yy = (int)y + dy;
            if (xx < 0 || yy < 0 || xx >= (int)w || yy >= (int)h) continue;
            const float4 g = guide[(size_t)yy * w + xx];
            if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;   // no textured hit there
            if ((g.w > 0.f) != (g0.w > 0.f)) continue;                           // lit side / unlit side do not mix
            const float ddx = g.x - g0.x, ddy = g.y - g0.y, ddz = g.z - g0.z;
            const float wt = __expf(-(ddx * ddx + ddy * ddy + ddz * ddz) * inv_sigma2);
            acc += wt * in[(size_t)yy * w + xx];
            wsum += wt;
        }
    out[(size_t)y * w + x] = wsum > 0.f ? acc / wsum : in[(size_t)y * w + x];
}

// segments >= RT_SEG_SLOTS reuse a counter slot (split-tail path)
__global__ void k_reset_segment(FrameCounters* c, uint32_t seg) {
    SegCounters z = {};
    c->seg[seg & (RT_SEG_SLOTS - 1u)] = z;
}

// The rest of the frame in ONE cooperative launch: segment 0's resolve phase, then segments
// 1 .. max_segments-1 (mirror / portal bounces) with the four phases separated by grid-wide barriers,
// then the ray-count export.  Most frames have few or no bounce rays; the kernel leaves as soon as a
// segment's ray queue is empty, so such frames pay one launch instead of four per segment plus two.
template <bool COUNT>
__global__ void __launch_bounds__(128, 5) k_tail(SceneDev S, FrameDev F, uint64_t* ray_counts_out) {
    cg::grid_group grid = cg::this_grid();
    export_bounce_hint(F);
    resolve_phase(F, 0);
    bool traced = false;
    for (uint32_t seg = 1; seg < F.max_segments; seg++) {
        const uint32_t rays = *((volatile unsigned int*)&seg_counters(F, seg - 1u)->ray_count);
        if (rays == 0) break;  // grid-uniform: written before the last barrier (or by the previous kernel)
        if (seg >= RT_SEG_SLOTS && grid.thread_rank() == 0) {  // counter slots are reused round-robin
            SegCounters z = {};
            *seg_counters(F, seg) = z;
        }
        grid.sync();  // the previous resolve phase has read the hit queue this segment will overwrite
        trace_phase<false, COUNT>(S, F, seg, rays);
        traced = true;
        grid.sync();
        const uint32_t hits = *((volatile unsigned int*)&seg_counters(F, seg)->hit_count);
        if (hits) {
            prep_phase(S, F, seg);
            grid.sync();
            shadow_phase<COUNT>(S, F, seg);
            grid.sync();
            resolve_phase(F, seg);
        }
    }
    if (ray_counts_out) {
        if (traced) grid.sync();  // this kernel added to the counters: wait for every warp's flush
        if (grid.thread_rank() == 0) {
            ray_counts_out[0] = *((volatile unsigned long long*)&F.counters->primary_rays);
            ray_counts_out[1] = *((volatile unsigned long long*)&F.counters->shadow_rays);
        }
    }
}

// ------------------------------------------------------------------------------------ megakernel
// HEAT = Uniforms.show_heatmap (lib.rs:120-124, 174-186): the pixel shows how many clock ticks its ray-gen
// invocation took instead of its colour.  The reference's definition is per invocation (read_clock_khr before
// and after the segment loop), so heatmap frames always run this one-thread-per-pixel kernel.
template <bool COUNT, bool HEAT>
__global__ void __launch_bounds__(128) k_mega(SceneDev S, FrameDev F, uint32_t total) {
    TraceCounters tc = {0, 0, 0, 0, 0}, tcs = {0, 0, 0, 0, 0};
    uint32_t n_primary = 0, n_shadow = 0, n_textured = 0;
    uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lx, ly;
    bool active = item < total && item_to_pixel(F, item, lx, ly);
    if (active) {
        const long long start_time = HEAT ? clock64() : 0ll;
        uint32_t pixel = ly * F.tw + lx;
        uint32_t px = F.x0 + lx, py = global_y(F, ly);
        V3 o, d;
        primary_ray(F.uniforms, px, py, F.width, F.height, o, d);
        V3 colour = v3(0.f, 0.f, 0.f);
        for (uint32_t seg = 0; seg < F.max_segments; seg++) {
            colour = v3(0.f, 0.f, 0.f);
            Hit h;
            bool got = trace_ray<false, COUNT>(S, o, d, 0.01f, 10000.0f, h, tc);
            n_primary++;
            if (!got) { colour = miss_colour(F.uniforms, F.cos_sun_radius, d); break; }
            write_hit_ids(F, pixel, seg, h);
            uint32_t kind = h.custom_sbt >> 24;
            if (kind == RT_HIT_TEXTURED) {
                TexturedHit th;
                th.px = px; th.py = py; th.inst_pos = h.inst_pos; th.geom = h.geom; th.prim = h.prim;
                th.custom_index = h.custom_sbt & 0xFFFFFFu; th.instance_id = h.instance_id;
                th.u = h.u; th.v = h.v; th.dir = d;
                colour = shade_textured<COUNT>(S, F, th, n_shadow, tcs);
                n_textured++;
                break;
            }
            V3 no, nd;
            if (kind == RT_HIT_MIRROR) {
                if (!shade_mirror(S, h.inst_pos, h.geom, h.prim, h.u, h.v, h.t, o, d, no, nd)) break;
            } else if (kind == RT_HIT_PORTAL) {
                shade_portal(h.t, o, d, no, nd);
            } else {
                break;
            }
            if (nd.x == 0.0f && nd.y == 0.0f && nd.z == 0.0f) break;
            o = no; d = nd;
        }
        if (HEAT) {
            const unsigned long long delta_time = (unsigned long long)(clock64() - start_time);
            const uint32_t cycles = delta_time > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)delta_time;
            if (F.cost) F.cost[out_index(F, pixel)] = cycles;
            colour = heatmap_pixel(cycles, F.heatmap_scale, colour);
        }
        write_pixel(F, pixel, colour);
    }
    flush_ray_counters(F, n_primary, n_shadow, n_textured);
    flush_trace_counters<COUNT>(F, 0, tc);
    flush_trace_counters<COUNT>(F, 1, tcs);
}

// sample_directional_light(animated_blue_noise(..)) for every (pixel mod 64, sample) of the frame
__global__ void __launch_bounds__(128) k_sun_dirs(SceneDev S, FrameDev F, float4* out) {
    pdl_release();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = F.shadow_rays;
    if (t >= 4096u * n) return;
    uint32_t cell = t / n, i = t - cell * n;
    const SunFrame sun = make_sun_frame(F.uniforms);
    V2 xi = blue_noise_xi(S, F.uniforms.blue_noise_texture_index, cell & 63u, cell >> 6, i, F.uniforms.frame_index);
    V3 d = sample_directional_light(xi, sun, F.uniforms.sun_radius);
    out[t] = make_float4(d.x, d.y, d.z, 0.f);
}

__global__ void k_export_counts(const FrameCounters* c, uint64_t* out) {
    out[0] = c->primary_rays;
    out[1] = c->shadow_rays;
}

// Box-test audit (rt_debug_box_test): the traversal's node_hit_mask for every (ray, node) pair, first-hit and closest-hit form.
__global__ void __launch_bounds__(128) k_debug_box_test(const Node8* __restrict__ nodes, uint32_t num_nodes, const float4* __restrict__ rays, uint32_t num_rays,
                                                        uint8_t* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_rays * num_nodes) return;
    const uint32_t r = t / num_nodes, nidx = t - r * num_nodes;
    const float4 ro = rays[2 * r], rd = rays[2 * r + 1];
    const V3 co = v3(ro.x, ro.y, ro.z), cd = v3(rd.x, rd.y, rd.z);
    const float idx = safe_rcp(cd.x), idy = safe_rcp(cd.y), idz = safe_rcp(cd.z);
    const uint32_t sx = cd.x < 0.0f ? 1u : 0u, sy = cd.y < 0.0f ? 1u : 0u, sz = cd.z < 0.0f ? 1u : 0u;
    const uint4* np = reinterpret_cast<const uint4*>(nodes + nidx);
    const uint4 n0 = __ldg(np);
    const uint4 nx = __ldg(np + 2 + sx), fx = __ldg(np + 3 - sx), ny = __ldg(np + 4 + sy), fy = __ldg(np + 5 - sy), nz = __ldg(np + 6 + sz), fz = __ldg(np + 7 - sz);
    out[2 * (size_t)t] = (uint8_t)node_hit_mask<true>(n0, nx, fx, ny, fy, nz, fz, co, idx, idy, idz, ro.w, rd.w);
    out[2 * (size_t)t + 1] = (uint8_t)node_hit_mask<false>(n0, nx, fx, ny, fy, nz, fz, co, idx, idy, idz, ro.w, rd.w);
}

}  // namespace

// Launch with the programmatic-stream-serialisation attribute (see pdl_wait); plain launch when `pdl` is false.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, cudaStream_t stream, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename K>
static int persistent_grid(K kernel, int sms) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

cudaError_t init_launch_geometry(LaunchGeometry& g, int sms) {
    if (g.ready) return cudaSuccess;
    g.trace0[0] = persistent_grid(k_trace0<false>, sms); g.trace0[1] = persistent_grid(k_trace0<true>, sms);
    g.trace_n[0] = persistent_grid(k_trace_n<false>, sms); g.trace_n[1] = persistent_grid(k_trace_n<true>, sms);
    g.shadow[0] = persistent_grid(k_shadow<false>, sms); g.shadow[1] = persistent_grid(k_shadow<true>, sms);
    g.tail[0] = persistent_grid(k_tail<false>, sms);     g.tail[1] = persistent_grid(k_tail<true>, sms);
    g.prep = persistent_grid(k_prep, sms);
    g.resolve = persistent_grid(k_resolve, sms);
    g.ready = true;
    return cudaGetLastError();
}

cudaError_t launch_frame(const SceneDev& S, const FrameDev& F, uint32_t pipeline, bool count, bool split_tail, bool no_pdl,
                         const LaunchGeometry& geom, uint64_t* d_ray_counts,
                         FrameTiming* timing, cudaStream_t stream, const DenoiseHook* hook, int* hook_status) {
    if (hook_status) *hook_status = 0;
    cudaMemsetAsync(F.counters, 0, sizeof(FrameCounters), stream);
    if (F.hit_ids) cudaMemsetAsync(F.hit_ids, 0xFF, (size_t)F.rows * F.tw * F.max_segments * 3 * sizeof(uint32_t), stream);
    uint32_t tiles = ((F.tw + 7u) / 8u) * ((F.rows + 3u) / 4u);
    uint32_t total = tiles * 32u;
    if (timing) {
        timing->n = 0;
        cudaEventRecord(timing->ev[0], stream);
    }
    auto mark = [&](int kind) {
        note_launch();
        if (timing && timing->n < FrameTiming::MAX_INTERVALS) {
            timing->kind[timing->n] = kind;
            cudaEventRecord(timing->ev[timing->n + 1], stream);
            timing->n++;
        }
    };
    if (total == 0 || F.max_segments == 0) {
        if (d_ray_counts) { k_export_counts<<<1, 1, 0, stream>>>(F.counters, d_ray_counts); note_launch(); }
        return cudaGetLastError();
    }
    if (F.sun_dirs) {
        k_sun_dirs<<<(4096u * F.shadow_rays + 127u) / 128u, 128, 0, stream>>>(S, F, const_cast<float4*>(F.sun_dirs));
        note_launch();
    }
    const bool heat = F.uniforms.show_heatmap != 0;
    if (pipeline == RT_PIPELINE_MEGAKERNEL || heat) {
        const unsigned g = (total + 127) / 128;
        if (heat) {
            if (count) k_mega<true, true><<<g, 128, 0, stream>>>(S, F, total);
            else k_mega<false, true><<<g, 128, 0, stream>>>(S, F, total);
        } else {
            if (count) k_mega<true, false><<<g, 128, 0, stream>>>(S, F, total);
            else k_mega<false, false><<<g, 128, 0, stream>>>(S, F, total);
        }
        mark(K_MEGA);
    } else {
        const int ci = count ? 1 : 0;
        const int* g_trace0 = geom.trace0; const int* g_shadow = geom.shadow; const int* g_tail = geom.tail;
        const int g_prep = geom.prep, g_resolve = geom.resolve;
        int cap = (int)((total + 127) / 128);  // no more blocks than there could be work
        auto fit = [&](int g, uint32_t per_item) { long long c = (long long)cap * per_item; return (int)(c < g ? c : g); };
        // stage-to-stage launches overlap the predecessor's tail (programmatic dependent launch); not when
        // per-kernel timing events sit between them
        const bool pdl = timing == nullptr && !no_pdl;
        cudaError_t le;
        if (count) le = launch_pdl(k_trace0<true>, fit(g_trace0[ci], 1), stream, pdl && F.sun_dirs != nullptr, S, F, total);
        else le = launch_pdl(k_trace0<false>, fit(g_trace0[ci], 1), stream, pdl && F.sun_dirs != nullptr, S, F, total);
        if (le != cudaSuccess) return le;
        mark(K_TRACE);
        le = launch_pdl(k_prep, fit(g_prep, 1), stream, pdl, S, F, 0u);
        if (le != cudaSuccess) return le;
        mark(K_PREP);
        if (count) le = launch_pdl(k_shadow<true>, fit(g_shadow[ci], F.shadow_rays), stream, pdl, S, F, 0u);
        else le = launch_pdl(k_shadow<false>, fit(g_shadow[ci], F.shadow_rays), stream, pdl, S, F, 0u);
        if (le != cudaSuccess) return le;
        mark(K_SHADOW);
        const bool hooked = hook && hook->fn && F.sun_factor && F.position_nol;
        if (hooked) {
            // the sun factor in image space -> the host's filter (enqueues on this stream) -> read back by k_resolve(seg 0)
            const uint32_t pixels = F.rows * F.tw;
            k_fill_sun_factor<<<fit(geom.resolve, 1), 256, 0, stream>>>(F.sun_factor, F.position_nol, pixels);
            k_export_sun_factor<<<fit(geom.resolve, 1), 128, 0, stream>>>(F);
            note_launch(2);
            RtDenoiseBuffers db;
            db.width = F.tw; db.rows = F.rows; db.sun_factor = F.sun_factor; db.position_nol = reinterpret_cast<const float*>(F.position_nol);
            db.shadow_rays = F.shadow_rays; db.frame_index = F.uniforms.frame_index;
            const int hs = hook->fn(hook->user, (void*)stream, &db);
            if (hs != 0) {
                if (hook_status) *hook_status = hs;
                return cudaErrorUnknown;
            }
        }
        if (F.max_segments == 1 || split_tail || hooked) {
            k_resolve<<<fit(g_resolve, 1), 128, 0, stream>>>(F, 0);
            mark(K_RESOLVE);
            for (uint32_t seg = 1; seg < F.max_segments; seg++) {
                if (seg >= RT_SEG_SLOTS) { k_reset_segment<<<1, 1, 0, stream>>>(F.counters, seg); note_launch(); }
                if (count) k_trace_n<true><<<fit(geom.trace_n[ci], 1), 128, 0, stream>>>(S, F, seg);
                else k_trace_n<false><<<fit(geom.trace_n[ci], 1), 128, 0, stream>>>(S, F, seg);
                mark(K_TRACE);
                k_prep<<<fit(g_prep, 1), 128, 0, stream>>>(S, F, seg);
                mark(K_PREP);
                if (count) k_shadow<true><<<fit(g_shadow[ci], F.shadow_rays), 128, 0, stream>>>(S, F, seg);
                else k_shadow<false><<<fit(g_shadow[ci], F.shadow_rays), 128, 0, stream>>>(S, F, seg);
                mark(K_SHADOW);
                k_resolve<<<fit(g_resolve, 1), 128, 0, stream>>>(F, seg);
                mark(K_RESOLVE);
            }
        } else {
            SceneDev s_arg = S;
            FrameDev f_arg = F;
            uint64_t* rc_arg = d_ray_counts;
            void* args[] = {&s_arg, &f_arg, &rc_arg};
            cudaError_t ce = cudaLaunchCooperativeKernel(count ? (void*)k_tail<true> : (void*)k_tail<false>, dim3(fit(g_tail[ci], 1)), dim3(128), args, 0, stream);
            if (ce != cudaSuccess) return ce;
            mark(K_TAIL);
            return cudaGetLastError();  // k_tail exported the ray counts
        }
    }
    if (d_ray_counts) { k_export_counts<<<1, 1, 0, stream>>>(F.counters, d_ray_counts); note_launch(); }
    return cudaGetLastError();
}

}  // namespace b200rt

namespace b200rt {
cudaError_t launch_debug_box_test(const Node8* nodes, uint32_t num_nodes, const float4* rays, uint32_t num_rays, uint8_t* out, cudaStream_t stream) {
    const uint64_t total = (uint64_t)num_nodes * num_rays;
    if (total) { k_debug_box_test<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(nodes, num_nodes, rays, num_rays, out); note_launch(); }
    return cudaGetLastError();
}
}  // namespace b200rt

#ifdef RT_LANE_HIST
extern "C" int rt_debug_lane_hist(unsigned long long* out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, b200rt::g_lane_hist, sizeof(unsigned long long) * 66);
    if (reset) { unsigned long long z[66] = {}; cudaMemcpyToSymbol(b200rt::g_lane_hist, z, sizeof(z)); }
    return (int)cudaGetLastError();
}
#endif

extern "C" int rt_denoise_bilateral(void* user, void* cuda_stream, const RtDenoiseBuffers* b) {
    if (!b || !b->sun_factor || !b->position_nol) return RT_ERR_INVALID_ARGUMENT;
    const size_t pixels = (size_t)b->width * b->rows;
    if (!pixels) return RT_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    float* tmp = nullptr;
    if (cudaMallocAsync(&tmp, pixels * sizeof(float), st) != cudaSuccess) return RT_ERR_CUDA;
    const float sigma = user ? *static_cast<const float*>(user) : 0.25f;
    cudaMemcpyAsync(tmp, b->sun_factor, pixels * sizeof(float), cudaMemcpyDeviceToDevice, st);
    dim3 grid((b->width + 15u) / 16u, (b->rows + 15u) / 16u);
    b200rt::k_denoise_bilateral<<<grid, 256, 0, st>>>(tmp, b->sun_factor, reinterpret_cast<const float4*>(b->position_nol), b->width, b->rows,
                                                      1.0f / (sigma * sigma));
    b200rt::note_launch();
    cudaFreeAsync(tmp, st);
    return cudaGetLastError() == cudaSuccess ? RT_OK : RT_ERR_CUDA;
}
