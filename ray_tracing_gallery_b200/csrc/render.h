// render.h — frame launch (render.cu) and scene-preparation kernels (scene_kernels.cu).
#pragma once
#include "rt_types.h"

namespace b200rt {

// Optional per-kernel timing of one frame: ev[0] is recorded before the first kernel, ev[i+1]
// after the i-th timed kernel, kind[i] = 0 k_trace / 1 k_prep / 2 k_shadow / 3 k_resolve / 4 k_mega / 5 k_tail.
struct FrameTiming {
    enum { MAX_INTERVALS = 40 };
    cudaEvent_t ev[MAX_INTERVALS + 1];
    int kind[MAX_INTERVALS];
    int n;
};

// Grid sizes of the persistent / cooperative frame kernels on ONE device: SMs x resident blocks from the occupancy API.
// Owned by the context (one per GPU) and filled on its first frame, so contexts on different GPUs, or rendering from
// different host threads, never share them.  [0] = plain kernels, [1] = the RT_RENDER_COUNTERS instantiations.
struct LaunchGeometry {
    bool ready = false;
    int trace0[2] = {0, 0}, trace_n[2] = {0, 0}, shadow[2] = {0, 0}, tail[2] = {0, 0}, prep = 0, resolve = 0;
};
cudaError_t init_launch_geometry(LaunchGeometry& g, int sms);

// Denoise hook of the context (rt_set_denoise_hook): called on the host while the frame is enqueued, between the shadow
// rays of segment 0 and its resolve; works on F.sun_factor / F.position_nol with work enqueued on the frame's stream.
struct DenoiseHook {
    RtDenoiseFn fn = nullptr;
    void* user = nullptr;
};

// Enqueue one frame on `stream`.  d_ray_counts (device, optional) receives {ray-gen segments, shadow rays}.
cudaError_t launch_frame(const SceneDev& S, const FrameDev& F, uint32_t pipeline, bool count, bool split_tail, bool no_pdl,
                         const LaunchGeometry& geom, uint64_t* d_ray_counts,
                         FrameTiming* timing, cudaStream_t stream, const DenoiseHook* hook = nullptr, int* hook_status = nullptr);

// rt_debug_box_test: node_hit_mask of every (ray, node) pair -> out[ray][node][2] (first-hit form, closest-hit form with tlimit = tmax)
cudaError_t launch_debug_box_test(const Node8* nodes, uint32_t num_nodes, const float4* rays, uint32_t num_rays, uint8_t* out, cudaStream_t stream);

// BLAS input: per flattened triangle (geometry-major) the padded AABB.
struct ModelGeomDev {
    const float* positions;         // float3 per vertex
    const uint32_t* const* indices; // per geometry
    const uint32_t* geom_start;     // [num_geoms + 1] flattened triangle offsets
    const uint8_t* geom_opaque;     // [num_geoms]
    uint32_t num_geoms, num_tris, num_vertices;
};
cudaError_t launch_triangle_boxes(const ModelGeomDev& M, Aabb* boxes, cudaStream_t stream);
cudaError_t launch_gather_triangles(const ModelGeomDev& M, const uint32_t* leaf_order, TriRec* out, cudaStream_t stream);

// TLAS input: per instance the inverse transform, traversal record and padded world AABB.
// leaf_order == nullptr: output slot i = instance i;  otherwise slot i = instance leaf_order[i].
cudaError_t launch_prepare_instances(const RtInstance* instances, uint32_t n, const BlasInfo* blas, uint32_t num_models,
                                     const uint32_t* leaf_order, InstRT* out_rt, Aabb* out_boxes, cudaStream_t stream);
cudaError_t launch_gather_instances(const InstRT* in, const uint32_t* leaf_order, uint32_t n, InstRT* out, cudaStream_t stream);
// Copy the topology of one TLAS set to the other (wide nodes in use, leaf order, node count) before a refit of the copy.
cudaError_t launch_copy_tlas(const Node8* src_nodes, Node8* dst_nodes, const uint32_t* src_count, uint32_t* dst_count,
                             const uint32_t* src_order, uint32_t* dst_order, uint32_t n, uint32_t node_cap, int sms, cudaStream_t stream);

}  // namespace b200rt
