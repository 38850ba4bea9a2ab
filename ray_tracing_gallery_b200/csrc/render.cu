// render.cu — the frame: what one vkCmdTraceRaysKHR(width, height, 1) does in the reference
// (src/command_buffer_recording.rs:116-126 -> ray_generation, shaders/ray-tracing/src/lib.rs:94-191).
//
// Wavefront pipeline (default), per ray-gen segment s = 0 .. max_segments-1:
//   k_trace<s>   ray generation (s = 0) or ray-queue read (s > 0) + closest-hit traversal;
//                miss -> sky colour to the framebuffer; mirror/portal -> next ray pushed to the
//                compacted ray queue; textured hit -> HitRec pushed to the compacted hit queue
//                (warp ballot + one atomic per warp).
//   k_shade      one thread per queued textured hit (all lanes busy): shadow-terminator origin,
//                N blue-noise shadow rays (first-hit traversal), textures, BRDF, framebuffer write.
// Both are persistent kernels: warps pull 32-item batches from a device-side cursor, so a launch
// never needs the queue length on the host.
// Megakernel (A/B baseline): one thread per pixel runs the whole segment loop.
#include "launch_count.h"
#include "render.h"
#include "trace.cuh"

namespace b200rt {
namespace {

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

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

__device__ __forceinline__ void write_pixel(const FrameDev& F, uint32_t pixel, V3 c) {
    if (F.radiance) {
        F.radiance[3 * (size_t)pixel] = c.x;
        F.radiance[3 * (size_t)pixel + 1] = c.y;
        F.radiance[3 * (size_t)pixel + 2] = c.z;
    }
    if (F.rgba8) {
        uint32_t r = unorm8(linear_to_srgb1(c.x)), g = unorm8(linear_to_srgb1(c.y)), b = unorm8(linear_to_srgb1(c.z));
        reinterpret_cast<uint32_t*>(F.rgba8)[pixel] = r | (g << 8) | (b << 16) | 0xFF000000u;
    }
}

__device__ __forceinline__ void write_hit_ids(const FrameDev& F, uint32_t pixel, uint32_t seg, const Hit& h) {
    if (!F.hit_ids) return;
    uint32_t* p = F.hit_ids + ((size_t)pixel * F.max_segments + seg) * 3;
    p[0] = h.instance_id; p[1] = h.geom; p[2] = h.prim;
}

// one atomicAdd per warp; every lane of the warp must call this
__device__ __forceinline__ uint32_t warp_push(unsigned int* counter, bool pred) {
    uint32_t mask = __ballot_sync(0xFFFFFFFFu, pred);
    if (mask == 0) return 0;
    uint32_t leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(mask & ((1u << lane_id()) - 1u));
}

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
    if (lane_id() == 0 && ov) atomicAdd(&F.counters->stack_overflow, ov);
}

__device__ __forceinline__ void flush_ray_counters(const FrameDev& F, uint32_t n_primary, uint32_t n_shadow, uint32_t n_textured) {
    warp_add(&F.counters->primary_rays, n_primary);
    warp_add(&F.counters->shadow_rays, n_shadow);
    warp_add(&F.counters->textured_hits, n_textured);
}

// The textured closest-hit: shadow rays + shading.  closest_hit_textured.glsl:174-226
template <bool COUNT>
__device__ __forceinline__ V3 shade_textured(const SceneDev& S, const FrameDev& F, const TexturedHit& th, uint32_t& n_shadow, TraceCounters& tc /* shadow-ray counters */) {
    ShadeCtx ctx;
    V3 shadow_origin;
    if (!shade_textured_begin(S, th, ctx, shadow_origin)) return v3(0.f, 0.f, 0.f);
    V3 sun = v3(F.uniforms.sun_dir[0], F.uniforms.sun_dir[1], F.uniforms.sun_dir[2]);
    float sum = 0.0f;
    for (uint32_t i = 0; i < F.shadow_rays; i++) {
        V2 xi = blue_noise_xi(S, F.uniforms.blue_noise_texture_index, th.px, th.py, i, F.uniforms.frame_index);
        V3 dir = sample_directional_light(xi, sun, F.uniforms.sun_radius);
        Hit sh;
        // cast_shadow_ray: TerminateOnFirstHit | SkipClosestHit, tmin 0.001, tmax 10000 (:159-172)
        bool shadowed = trace_ray<true, COUNT>(S, shadow_origin, dir, 0.001f, 10000.0f, sh, tc);
        sum += shadowed ? 0.0f : 1.0f;
    }
    n_shadow += F.shadow_rays;
    float sun_factor = sum / (float)F.shadow_rays;
    return shade_textured_end(S, F.uniforms, th, ctx, sun_factor);
}

// ------------------------------------------------------------------------------------ wavefront
template <bool SEG0, bool COUNT>
__global__ void __launch_bounds__(128) k_trace(SceneDev S, FrameDev F, uint32_t seg, uint32_t total_seg0) {
    TraceCounters tc = {0, 0, 0, 0, 0};
    uint32_t n_primary = 0;
    const RayRec* __restrict__ in_q = F.ray_q[(seg + 1u) & 1u];
    RayRec* __restrict__ out_q = F.ray_q[seg & 1u];
    unsigned int* cursor = &F.counters->work_next[0];
    const uint32_t total = SEG0 ? total_seg0 : *((volatile unsigned int*)&F.counters->ray_count[(seg + 1u) & 1u]);
    for (;;) {
        uint32_t base = 0;
        if (lane_id() == 0) base = atomicAdd(cursor, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= total) break;
        uint32_t item = base + lane_id();
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
                float4 a = __ldg(rp), b = __ldg(rp + 1);
                o = v3(a.x, a.y, a.z); pixel = __float_as_uint(a.w);
                d = v3(b.x, b.y, b.z);
            }
        }
        Hit h;
        bool got = false;
        if (active) {
            got = trace_ray<false, COUNT>(S, o, d, 0.01f, 10000.0f, h, tc);  // lib.rs:151-163
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
        if (active && !textured && !bounce) {
            V3 c = got ? v3(0.f, 0.f, 0.f) : miss_colour(F.uniforms, F.cos_sun_radius, d);
            write_pixel(F, pixel, c);
        }
        uint32_t hslot = warp_push(&F.counters->hit_count, textured);
        if (textured) {
            HitRec* hr = F.hit_q + hslot;
            reinterpret_cast<uint4*>(hr)[0] = make_uint4(pixel, h.inst_pos, h.geom, h.prim);
            reinterpret_cast<float4*>(hr)[1] = make_float4(h.u, h.v, h.t, 0.f);
            reinterpret_cast<float4*>(hr)[2] = make_float4(d.x, d.y, d.z, 0.f);
        }
        uint32_t rslot = warp_push(&F.counters->ray_count[seg & 1u], bounce);
        if (bounce) {
            float4* rp = reinterpret_cast<float4*>(out_q + rslot);
            rp[0] = make_float4(no.x, no.y, no.z, __uint_as_float(pixel));
            rp[1] = make_float4(ndir.x, ndir.y, ndir.z, 0.f);
        }
    }
    flush_ray_counters(F, n_primary, 0, 0);
    flush_trace_counters<COUNT>(F, 0, tc);
}

template <bool COUNT>
__global__ void __launch_bounds__(128) k_shade(SceneDev S, FrameDev F) {
    TraceCounters tc = {0, 0, 0, 0, 0};
    uint32_t n_shadow = 0, n_textured = 0;
    unsigned int* cursor = &F.counters->work_next[1];
    const uint32_t total = *((volatile unsigned int*)&F.counters->hit_count);
    for (;;) {
        uint32_t base = 0;
        if (lane_id() == 0) base = atomicAdd(cursor, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= total) break;
        uint32_t item = base + lane_id();
        if (item < total) {
            const HitRec* hr = F.hit_q + item;
            uint4 a = __ldg(reinterpret_cast<const uint4*>(hr));
            float4 b = __ldg(reinterpret_cast<const float4*>(hr) + 1);
            float4 c = __ldg(reinterpret_cast<const float4*>(hr) + 2);
            TexturedHit th;
            uint32_t ly = a.x / F.tw, lx = a.x - ly * F.tw;
            th.px = F.x0 + lx; th.py = global_y(F, ly);
            th.inst_pos = a.y; th.geom = a.z; th.prim = a.w;
            th.u = b.x; th.v = b.y;
            th.dir = v3(c.x, c.y, c.z);
            V3 col = shade_textured<COUNT>(S, F, th, n_shadow, tc);
            n_textured++;
            write_pixel(F, a.x, col);
        }
    }
    flush_ray_counters(F, 0, n_shadow, n_textured);
    flush_trace_counters<COUNT>(F, 1, tc);
}

// between segments: the hit queue and the next output ray queue start empty, cursors rewind
__global__ void k_next_segment(FrameCounters* c, uint32_t next_seg) {
    c->hit_count = 0;
    c->ray_count[next_seg & 1u] = 0;
    c->work_next[0] = 0;
    c->work_next[1] = 0;
}

// ------------------------------------------------------------------------------------ megakernel
template <bool COUNT>
__global__ void __launch_bounds__(128) k_mega(SceneDev S, FrameDev F, uint32_t total) {
    TraceCounters tc = {0, 0, 0, 0, 0}, tcs = {0, 0, 0, 0, 0};
    uint32_t n_primary = 0, n_shadow = 0, n_textured = 0;
    uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lx, ly;
    bool active = item < total && item_to_pixel(F, item, lx, ly);
    if (active) {
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
        write_pixel(F, pixel, colour);
    }
    flush_ray_counters(F, n_primary, n_shadow, n_textured);
    flush_trace_counters<COUNT>(F, 0, tc);
    flush_trace_counters<COUNT>(F, 1, tcs);
}

__global__ void k_export_counts(const FrameCounters* c, uint64_t* out) {
    out[0] = c->primary_rays;
    out[1] = c->shadow_rays;
}

}  // namespace

template <typename K>
static int persistent_grid(K kernel, int sms) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

cudaError_t launch_frame(const SceneDev& S, const FrameDev& F, uint32_t pipeline, bool count, int sms, uint64_t* d_ray_counts,
                         FrameTiming* timing, cudaStream_t stream) {
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
    if (pipeline == RT_PIPELINE_MEGAKERNEL) {
        if (count) k_mega<true><<<(total + 127) / 128, 128, 0, stream>>>(S, F, total);
        else k_mega<false><<<(total + 127) / 128, 128, 0, stream>>>(S, F, total);
        mark(2);
    } else {
        static int g_trace0[2] = {0, 0}, g_trace[2] = {0, 0}, g_shade[2] = {0, 0};
        int ci = count ? 1 : 0;
        if (!g_trace0[ci]) {
            g_trace0[ci] = count ? persistent_grid(k_trace<true, true>, sms) : persistent_grid(k_trace<true, false>, sms);
            g_trace[ci] = count ? persistent_grid(k_trace<false, true>, sms) : persistent_grid(k_trace<false, false>, sms);
            g_shade[ci] = count ? persistent_grid(k_shade<true>, sms) : persistent_grid(k_shade<false>, sms);
        }
        for (uint32_t seg = 0; seg < F.max_segments; seg++) {
            if (seg == 0) {
                int grid = (int)((total + 127) / 128) < g_trace0[ci] ? (int)((total + 127) / 128) : g_trace0[ci];
                if (count) k_trace<true, true><<<grid, 128, 0, stream>>>(S, F, 0, total);
                else k_trace<true, false><<<grid, 128, 0, stream>>>(S, F, 0, total);
            } else {
                k_next_segment<<<1, 1, 0, stream>>>(F.counters, seg);
                note_launch();
                if (count) k_trace<false, true><<<g_trace[ci], 128, 0, stream>>>(S, F, seg, 0);
                else k_trace<false, false><<<g_trace[ci], 128, 0, stream>>>(S, F, seg, 0);
            }
            mark(0);
            if (count) k_shade<true><<<g_shade[ci], 128, 0, stream>>>(S, F);
            else k_shade<false><<<g_shade[ci], 128, 0, stream>>>(S, F);
            mark(1);
        }
    }
    if (d_ray_counts) { k_export_counts<<<1, 1, 0, stream>>>(F.counters, d_ray_counts); note_launch(); }
    return cudaGetLastError();
}

}  // namespace b200rt
