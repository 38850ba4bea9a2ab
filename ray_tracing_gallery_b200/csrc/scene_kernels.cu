// scene_kernels.cu — builder inputs: triangle boxes / leaf-ordered triangle records for a BLAS
// (AccelerationStructure::build_blas, src/util_structs.rs:140-224) and per-instance traversal
// records + world boxes for the TLAS (build_tlas / update_tlas, src/util_functions.rs:453-510,
// src/util_structs.rs:285-357).
#include <math_constants.h>

#include "bvh_build.h"
#include "contract.cuh"
#include "launch_count.h"
#include "render.h"

namespace b200rt {
namespace {

__device__ __forceinline__ void pad_box(Aabb& b) {
    float m = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) m = fmaxf(m, fmaxf(fabsf(b.lo[k]), fabsf(b.hi[k])));
    float pad = m * 1e-6f + 1e-30f;
#pragma unroll
    for (int k = 0; k < 3; k++) { b.lo[k] -= pad; b.hi[k] += pad; }
}

__device__ __forceinline__ Aabb invalid_box() {
    Aabb b;
    b.lo[0] = b.lo[1] = b.lo[2] = CUDART_INF_F;
    b.hi[0] = b.hi[1] = b.hi[2] = -CUDART_INF_F;
    return b;
}

__device__ __forceinline__ void locate(const ModelGeomDev& M, uint32_t flat, uint32_t& geom, uint32_t& prim) {
    uint32_t g = 0;
    while (g + 1 < M.num_geoms && flat >= M.geom_start[g + 1]) g++;
    geom = g;
    prim = flat - M.geom_start[g];
}

__global__ void k_triangle_boxes(ModelGeomDev M, Aabb* boxes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M.num_tris) return;
    uint32_t geom, prim;
    locate(M, i, geom, prim);
    const uint32_t* idx = M.indices[geom] + 3 * (size_t)prim;
    Aabb b = invalid_box();
    bool ok = true;
#pragma unroll
    for (int v = 0; v < 3; v++) {
        uint32_t vi = idx[v];
        if (vi >= M.num_vertices) { ok = false; break; }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float p = M.positions[3 * (size_t)vi + k];
            if (!isfinite(p)) ok = false;
            b.lo[k] = fminf(b.lo[k], p);
            b.hi[k] = fmaxf(b.hi[k], p);
        }
    }
    if (ok) pad_box(b);
    else b = invalid_box();  // inactive primitive: never hit
    boxes[i] = b;
}

__global__ void k_gather_triangles(ModelGeomDev M, const uint32_t* __restrict__ leaf_order, TriRec* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M.num_tris) return;
    uint32_t geom, prim;
    locate(M, leaf_order[i], geom, prim);
    const uint32_t* idx = M.indices[geom] + 3 * (size_t)prim;
    uint32_t ia = idx[0], ib = idx[1], ic = idx[2];
    TriRec r;
    bool ok = ia < M.num_vertices && ib < M.num_vertices && ic < M.num_vertices;
    if (ok) {
        V3 a = v3(M.positions[3 * (size_t)ia], M.positions[3 * (size_t)ia + 1], M.positions[3 * (size_t)ia + 2]);
        V3 b = v3(M.positions[3 * (size_t)ib], M.positions[3 * (size_t)ib + 1], M.positions[3 * (size_t)ib + 2]);
        V3 c = v3(M.positions[3 * (size_t)ic], M.positions[3 * (size_t)ic + 1], M.positions[3 * (size_t)ic + 2]);
        V3 e1 = sub3(b, a), e2 = sub3(c, a);
        r.v0[0] = a.x; r.v0[1] = a.y; r.v0[2] = a.z;
        r.e1[0] = e1.x; r.e1[1] = e1.y; r.e1[2] = e1.z;
        r.e2[0] = e2.x; r.e2[1] = e2.y; r.e2[2] = e2.z;
    } else {
        for (int k = 0; k < 3; k++) r.v0[k] = r.e1[k] = r.e2[k] = CUDART_NAN_F;
    }
    r.prim = prim;
    r.geom_flags = geom | (M.geom_opaque[geom] ? 0u : RT_TRI_NON_OPAQUE);
    r._pad = 0;
    float4* d = reinterpret_cast<float4*>(out + i);
    const float4* s = reinterpret_cast<const float4*>(&r);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
}

__global__ void k_prepare_instances(const RtInstance* __restrict__ instances, uint32_t n, const BlasInfo* __restrict__ blas,
                                    uint32_t num_models, const uint32_t* __restrict__ leaf_order, InstRT* out_rt, Aabb* out_boxes) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    uint32_t id = leaf_order ? leaf_order[slot] : slot;
    const uint4* rp = reinterpret_cast<const uint4*>(instances + id);
    uint4 q0 = rp[0], q1 = rp[1], q2 = rp[2], q3 = rp[3];
    float m[12] = {__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z), __uint_as_float(q0.w),
                   __uint_as_float(q1.x), __uint_as_float(q1.y), __uint_as_float(q1.z), __uint_as_float(q1.w),
                   __uint_as_float(q2.x), __uint_as_float(q2.y), __uint_as_float(q2.z), __uint_as_float(q2.w)};
    uint32_t custom_mask = q3.x, sbt_flags = q3.y;
    uint64_t handle = ((uint64_t)q3.w << 32) | q3.z;
    InstRT r;
    invert_3x4(m, r.inv);
    r.instance_id = id;
    r.custom_sbt = (custom_mask & 0xFFFFFFu) | ((sbt_flags & 0xFFu) << 24);
    r.mask = custom_mask >> 24;
    r.blas_root = 0xFFFFFFFFu;
    Aabb wb = invalid_box();
    uint32_t model = (uint32_t)(handle & 0xFFFFFFFFu) - 1u;
    if ((handle >> 48) == 0xB200u && model < num_models) {
        BlasInfo bi = blas[model];
        if (bi.num_tris > 0 && bi.lo[0] <= bi.hi[0]) {
            r.blas_root = bi.root;
            if (bi.num_tris <= RT_TINY_BLAS_TRIS) {
                r.blas_root = bi.tri_first;
                r.mask |= bi.num_tris << 8;
            }
            bool finite = true;
            for (int c = 0; c < 8; c++) {
                float px = c & 1 ? bi.hi[0] : bi.lo[0], py = c & 2 ? bi.hi[1] : bi.lo[1], pz = c & 4 ? bi.hi[2] : bi.lo[2];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    float w = m[4 * k] * px + m[4 * k + 1] * py + m[4 * k + 2] * pz + m[4 * k + 3];
                    if (!isfinite(w)) finite = false;
                    wb.lo[k] = fminf(wb.lo[k], w);
                    wb.hi[k] = fmaxf(wb.hi[k], w);
                }
            }
            // a singular transform has no inverse: the instance cannot be entered
            for (int k = 0; k < 12; k++)
                if (!isfinite(r.inv[k])) finite = false;
            if (finite) pad_box(wb);
            else { wb = invalid_box(); r.blas_root = 0xFFFFFFFFu; }
        }
    }
    if ((sbt_flags & 0xFFFFFFu) > 0xFFu) r.custom_sbt = (r.custom_sbt & 0xFFFFFFu) | (0xFFu << 24);  // out-of-table group
    float4* d = reinterpret_cast<float4*>(out_rt + slot);
    const float4* s = reinterpret_cast<const float4*>(&r);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
    out_boxes[slot] = wb;
}

// Exact world bounds for the instances of small models: one warp per instance walks the model's vertices through the
// instance transform.  The box of the eight transformed corners of the object-space bounds (k_prepare_instances, what a
// Vulkan driver's TLAS build uses) is up to 1.41x too wide per axis for a rotated model — every instance box of C4 / C5 is a
// torus pair turned about y — and a wider leaf box is entered by rays that have no business inside the BLAS.  Triangles lie in
// the convex hull of their vertices, so the result (padded like the corner box and intersected with it) stays conservative.
// The point list (rt_create_model) holds the finite vertices that some triangle references, as float4.
__device__ __forceinline__ int ordered_int(float f) {
    int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7FFFFFFF);
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7FFFFFFF)); }

__global__ void __launch_bounds__(256) k_tighten_instance_boxes(const RtInstance* __restrict__ instances, uint32_t n,
                                                                const BlasInfo* __restrict__ blas, uint32_t num_models,
                                                                const uint32_t* __restrict__ leaf_order, Aabb* boxes) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (slot >= n) return;
    const uint32_t id = leaf_order ? leaf_order[slot] : slot;
    const uint4* rp = reinterpret_cast<const uint4*>(instances + id);
    const uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3);
    const uint64_t handle = ((uint64_t)q3.w << 32) | q3.z;
    const uint32_t model = (uint32_t)(handle & 0xFFFFFFFFu) - 1u;
    if ((handle >> 48) != 0xB200u || model >= num_models) return;
    const uint32_t nv = blas[model].num_verts;
    const float4* __restrict__ verts = blas[model].verts;
    if (nv == 0 || blas[model].num_tris == 0) return;
    const float m00 = __uint_as_float(q0.x), m01 = __uint_as_float(q0.y), m02 = __uint_as_float(q0.z), m03 = __uint_as_float(q0.w);
    const float m10 = __uint_as_float(q1.x), m11 = __uint_as_float(q1.y), m12 = __uint_as_float(q1.z), m13 = __uint_as_float(q1.w);
    const float m20 = __uint_as_float(q2.x), m21 = __uint_as_float(q2.y), m22 = __uint_as_float(q2.z), m23 = __uint_as_float(q2.w);
    float lx = CUDART_INF_F, ly = CUDART_INF_F, lz = CUDART_INF_F, hx = -CUDART_INF_F, hy = -CUDART_INF_F, hz = -CUDART_INF_F;
#pragma unroll 4
    for (uint32_t v = lane; v < nv; v += 32) {
        const float4 p = __ldg(verts + v);
        const float wx = fmaf(m00, p.x, fmaf(m01, p.y, fmaf(m02, p.z, m03)));
        const float wy = fmaf(m10, p.x, fmaf(m11, p.y, fmaf(m12, p.z, m13)));
        const float wz = fmaf(m20, p.x, fmaf(m21, p.y, fmaf(m22, p.z, m23)));
        lx = fminf(lx, wx); hx = fmaxf(hx, wx);
        ly = fminf(ly, wy); hy = fmaxf(hy, wy);
        lz = fminf(lz, wz); hz = fmaxf(hz, wz);
    }
    Aabb t;
    t.lo[0] = ordered_float(__reduce_min_sync(0xFFFFFFFFu, ordered_int(lx)));
    t.lo[1] = ordered_float(__reduce_min_sync(0xFFFFFFFFu, ordered_int(ly)));
    t.lo[2] = ordered_float(__reduce_min_sync(0xFFFFFFFFu, ordered_int(lz)));
    t.hi[0] = ordered_float(__reduce_max_sync(0xFFFFFFFFu, ordered_int(hx)));
    t.hi[1] = ordered_float(__reduce_max_sync(0xFFFFFFFFu, ordered_int(hy)));
    t.hi[2] = ordered_float(__reduce_max_sync(0xFFFFFFFFu, ordered_int(hz)));
    if (lane != 0) return;
    Aabb c = boxes[slot];
    if (!(c.lo[0] <= c.hi[0])) return;  // instance without geometry / refused by k_prepare_instances
    // (the points are finite; a transform that overflows on one of them gives an infinite or empty bound here: keep the corner box)
    if (!(isfinite(t.lo[0]) && isfinite(t.lo[1]) && isfinite(t.lo[2]) && isfinite(t.hi[0]) && isfinite(t.hi[1]) && isfinite(t.hi[2]))) return;
    pad_box(t);
#pragma unroll
    for (int k = 0; k < 3; k++) { c.lo[k] = fmaxf(c.lo[k], t.lo[k]); c.hi[k] = fminf(c.hi[k], t.hi[k]); }
    boxes[slot] = c;
}

__global__ void k_gather_instances(const InstRT* __restrict__ in, const uint32_t* __restrict__ leaf_order, uint32_t n, InstRT* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* s = reinterpret_cast<const float4*>(in + leaf_order[i]);
    float4* d = reinterpret_cast<float4*>(out + i);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

// Topology hand-over between the two TLAS sets (refit of the set that is not being rendered from): the wide nodes in
// use, the leaf order and the node count, all sized by the device-side count (no host read-back on the update path).
__global__ void __launch_bounds__(256) k_copy_tlas(const Node8* __restrict__ src_nodes, Node8* __restrict__ dst_nodes,
                                                   const uint32_t* __restrict__ src_count, uint32_t* __restrict__ dst_count,
                                                   const uint32_t* __restrict__ src_order, uint32_t* __restrict__ dst_order, uint32_t n,
                                                   uint32_t node_cap) {
    uint32_t count = *src_count;
    if (count > node_cap) count = node_cap;
    const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const uint4* s = reinterpret_cast<const uint4*>(src_nodes);
    uint4* d = reinterpret_cast<uint4*>(dst_nodes);
    const size_t quads = (size_t)count * (sizeof(Node8) / sizeof(uint4));
    for (size_t i = t0; i < quads; i += stride) d[i] = __ldg(s + i);
    for (uint32_t i = t0; i < n; i += stride) dst_order[i] = __ldg(src_order + i);
    if (t0 == 0) *dst_count = count;
}

}  // namespace

cudaError_t launch_copy_tlas(const Node8* src_nodes, Node8* dst_nodes, const uint32_t* src_count, uint32_t* dst_count,
                             const uint32_t* src_order, uint32_t* dst_order, uint32_t n, uint32_t node_cap, int sms, cudaStream_t stream) {
    // the node count is only known on the device: enough blocks for the largest TLAS, grid-stride inside
    uint32_t want = (max_wide_nodes(n) * 8u + 255u) / 256u;
    uint32_t cap = (uint32_t)(sms > 0 ? sms : 1) * 8u;
    uint32_t grid = want < cap ? (want ? want : 1u) : cap;
    k_copy_tlas<<<grid, 256, 0, stream>>>(src_nodes, dst_nodes, src_count, dst_count, src_order, dst_order, n, node_cap);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_triangle_boxes(const ModelGeomDev& M, Aabb* boxes, cudaStream_t stream) {
    if (M.num_tris) { k_triangle_boxes<<<(M.num_tris + 255) / 256, 256, 0, stream>>>(M, boxes); note_launch(); }
    return cudaGetLastError();
}
cudaError_t launch_gather_triangles(const ModelGeomDev& M, const uint32_t* leaf_order, TriRec* out, cudaStream_t stream) {
    if (M.num_tris) { k_gather_triangles<<<(M.num_tris + 255) / 256, 256, 0, stream>>>(M, leaf_order, out); note_launch(); }
    return cudaGetLastError();
}
cudaError_t launch_prepare_instances(const RtInstance* instances, uint32_t n, const BlasInfo* blas, uint32_t num_models,
                                     const uint32_t* leaf_order, InstRT* out_rt, Aabb* out_boxes, cudaStream_t stream) {
    if (n) {
        k_prepare_instances<<<(n + 127) / 128, 128, 0, stream>>>(instances, n, blas, num_models, leaf_order, out_rt, out_boxes);
        note_launch();
#ifndef RT_NO_TIGHT_BOXES
        k_tighten_instance_boxes<<<(uint32_t)(((size_t)n * 32 + 255) / 256), 256, 0, stream>>>(instances, n, blas, num_models, leaf_order, out_boxes);
        note_launch();
#endif
    }
    return cudaGetLastError();
}
cudaError_t launch_gather_instances(const InstRT* in, const uint32_t* leaf_order, uint32_t n, InstRT* out, cudaStream_t stream) {
    if (n) { k_gather_instances<<<(n + 127) / 128, 128, 0, stream>>>(in, leaf_order, n, out); note_launch(); }
    return cudaGetLastError();
}

}  // namespace b200rt
