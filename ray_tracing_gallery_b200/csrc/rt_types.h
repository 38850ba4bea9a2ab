// rt_types.h — device-side records of libb200rt (internal; the public layouts are in include/rt_abi.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200rt.h"

namespace b200rt {

// ---------------------------------------------------------------------------------------------
// Compressed 8-wide BVH node: two 128-byte lines, 256-byte aligned.  Traversal reads the FIRST line only (eight
// ld.global.nc.v4: header + the near and far planes of the ray's octant); the second line is build / refit state.
//
// Child boxes are quantised to 8 bits per plane on a power-of-two grid anchored at `origin`:
//     plane = origin[k] + q * 2^(exp[k] - 127),   q = 0..255
// The origin is snapped onto that grid and exp is clamped so that every plane is exactly representable in fp32
// (builder: quantise_node).  lo planes are rounded down, hi planes up.  The planes are STORED as bf16 numbers (the
// integer q, exact in bf16's 8 significant bits), two slots per 32-bit word, so that the box test runs on packed
// bf16x2 arithmetic with no unpacking: one HFMA2.BF16 per two planes (trace.cuh node_hit_mask).
// Slots are assigned by the octant of the child centre relative to the node centre, so that visiting hit slots in
// order of (slot XOR ray_octant) is front-to-back without sorting.  Internal children are stored contiguously from
// `child_base` in slot order (child index = child_base + popc(imask & ((1<<slot)-1))); leaf primitives are stored
// contiguously from `prim_base`: meta[slot] = offset (5 bits) | count << 5 (count 1..7) for a leaf slot, 0xFF for an
// internal slot, 0 for an empty slot (which also has q lo = 255, q hi = 0).
struct __align__(256) Node8 {
    float    origin[3];   //   0
    uint8_t  exp[3];      //  12  biased like an fp32 exponent field: cell size = 2^(exp - 127)
    uint8_t  imask;       //  15  bit s: slot s is an internal child
    uint32_t child_base;  //  16
    uint32_t prim_base;   //  20
    uint8_t  meta[8];     //  24
    uint16_t q[6][8];     //  32  bf16 bit patterns: [2 * axis] = lo planes, [2 * axis + 1] = hi planes; [slot]
    // ---- second line: not read by traversal
    float    lo[3];       // 128  exact bounds of this node
    float    hi[3];       // 140
    uint32_t parent;      // 152  wide-node index of the parent (0xFFFFFFFF for the root)
    uint32_t parent_slot; // 156
    uint32_t lmask;       // 160  bit s: slot s is a leaf child
    uint32_t _pad[23];
};
static_assert(sizeof(Node8) == 256, "Node8 is two 128-byte lines");
#define RT_NODE_QUADS (sizeof(Node8) / sizeof(uint4))
#define RT_NODE_BOUNDS_FLOAT 32   // Node8::lo as a float index

// Entries of a ray's traversal stack (trace.cuh): one per visited node that still has other hit children, one per TLAS leaf
// found but not entered yet.  Typical depth < 10; an overflow is reported by the API (api.cu: check_stack_overflow).
#ifndef RT_STACK_SIZE
#define RT_STACK_SIZE 48
#endif

// Triangle record in BVH leaf order: three float4.
struct __align__(16) TriRec {
    float v0[3];  uint32_t prim;        // gl_PrimitiveID
    float e1[3];  uint32_t geom_flags;  // gl_GeometryIndexEXT | (non-opaque ? 0x80000000 : 0)
    float e2[3];  uint32_t _pad;
};
static_assert(sizeof(TriRec) == 48, "TriRec is 48 bytes");
#define RT_TRI_NON_OPAQUE 0x80000000u
// A BLAS of at most this many triangles (ground planes, quads, billboards) is tested directly on instance entry:
// its single node would cost more than the triangles it culls.
#define RT_TINY_BLAS_TRIS 4u
// Triangles per BLAS leaf slot.  Measured on C2/C4 with 2/4/6/7: a triangle test costs about a third of a node
// visit and culls nothing, so small leaves win (profiles/r01_notes.md).
#ifndef RT_BLAS_LEAF_TRIS
#define RT_BLAS_LEAF_TRIS 2u
#endif

// Binary tree under the wide nodes (bvh_build.cu): top-down binned-SAH splits (step 4b, one cooperative launch) for every BLAS and every
// TLAS build or rebuild; 0 = the Morton radix tree (A/B runs).  Measured (profiles/r04cd_sah_builder_ab.txt): C5 55.8 -> 42.8 ms,
// C4 7.61 -> 6.31 ms with both, of which the TLAS is the larger part.
#ifndef RT_BLAS_SAH
#define RT_BLAS_SAH 1
#endif
#ifndef RT_BLAS_SAH_COLLAPSE
#define RT_BLAS_SAH_COLLAPSE 0
#endif
#ifndef RT_TLAS_SAH_COLLAPSE
#define RT_TLAS_SAH_COLLAPSE 1
#endif
#ifndef RT_TLAS_SAH
#define RT_TLAS_SAH 1
#endif

// Per-instance traversal record in TLAS leaf order: four float4.
struct __align__(16) InstRT {
    float    inv[12];      // world -> object, row-major 3x4
    uint32_t blas_root;    // index into the BLAS node pool, 0xFFFFFFFF = no geometry; first TriRec for a tiny BLAS
    uint32_t instance_id;  // gl_InstanceID (index of the 64-byte record)
    uint32_t custom_sbt;   // custom_index (24 low) | sbt_offset (8 high)
    uint32_t mask;         // visibility mask (bits 0..7) | tiny-BLAS triangle count (bits 8..15, 0 = traverse nodes)
};
static_assert(sizeof(InstRT) == 64, "InstRT is 64 bytes");

struct BlasInfo {
    uint32_t root;       // wide-node index of the BLAS root in the pool
    uint32_t num_tris;
    uint32_t tri_first;  // index of the model's first TriRec
    uint32_t num_verts;  // > 0: `verts` may be walked for exact instance bounds (k_tighten_instance_boxes), 0: corners of lo / hi only
    float    lo[3], hi[3];
    const float4* verts; // the finite vertices some triangle references (xyz, w unused)
};

// Instances of a BLAS with at most this many referenced vertices get the exact world bounds of their transformed vertices instead of
// the bounds of the eight transformed box corners (a model rotated by 45 degrees: up to 1.41x per axis).
#ifndef RT_TIGHT_BOX_MAX_VERTS
#define RT_TIGHT_BOX_MAX_VERTS 4096u
#endif

// Bindless image table entry (<= RT_MAX_BOUND_IMAGES).
struct __align__(16) TexEntry {
    cudaTextureObject_t obj;  // point-sampled, unnormalised coords; 0 for 1x1 constants
    uint32_t w, h;
    uint32_t format;          // RtFormat
    uint32_t linear;          // sampler choice
    float    constant[4];     // decoded texel of a 1x1 image
};

// Generic AABB used by the builder.
struct Aabb {
    float lo[3];
    float hi[3];
};

// ---------------------------------------------------------------------------------------------
// Wavefront queues.
struct __align__(16) RayRec {   // 32 B: a ray-gen segment waiting to be traced
    float ox, oy, oz; uint32_t pixel;
    float dx, dy, dz; uint32_t _pad;
};
// 64 B: a textured hit on its way through k_trace -> k_prep -> k_shadow -> k_resolve.
// k_trace writes the hit (words 0..11); k_prep replaces it in place by what the later stages need.
struct __align__(16) HitRec {
    uint32_t pixel; uint32_t a1, a2, a3;  // k_trace: inst_pos, geom, prim       k_prep: NoL, -, -
    float    b0, b1, b2; uint32_t lit;    // k_trace: u, v, custom index         k_prep: comb.xyz; lit = unshadowed rays (k_shadow, atomic)
    float    c0, c1, c2; uint32_t c3;     // k_trace: gl_WorldRayDirectionEXT, gl_InstanceID      k_prep: base colour
    float    sox, soy, soz; uint32_t shadow_valid;  // k_prep: shadow-ray origin; bit 0: triangle resolved, bits 1..16 px, 17..31 py
};
static_assert(sizeof(HitRec) == 64, "HitRec is 64 bytes");

// Per ray-gen segment: queue fills and the work cursors of the persistent kernels.
#define RT_SEG_SLOTS 8
struct SegCounters {
    unsigned int hit_count;      // HitRec queue fill (k_trace)
    unsigned int ray_count;      // RayRec queue fill: rays for the NEXT segment (k_trace)
    unsigned int work_next[4];   // cursors: trace, prep, shadow, resolve
    unsigned int _pad[2];
};
struct FrameCounters {
    unsigned long long primary_rays, shadow_rays, textured_hits;
    // [0] closest-hit rays (k_trace), [1] shadow rays (k_shadow)
    unsigned long long nodes_visited[2], instances_entered[2], triangles_tested[2], anyhit_calls[2];
    SegCounters seg[RT_SEG_SLOTS];  // slot = segment % RT_SEG_SLOTS
    unsigned int stack_overflow;    // traversal stack overflow events (must stay 0)
    unsigned int _pad;
};

// Everything a render kernel needs, passed by value (< 4 KB).
struct SceneDev {
    const Node8*      tlas_nodes;
    const InstRT*     inst_rt;       // TLAS leaf order
    const RtInstance* instances;     // original 64-byte records, by gl_InstanceID
    const Node8*      blas_nodes;
    const TriRec*     tris;
    const RtModelInfo* model_info;   // reference layout, device pointers inside
    uint32_t          num_models;
    uint32_t          num_instances;
    const TexEntry*   textures;
    uint32_t          num_textures;
    const float*      srgb_lut;      // 512 floats: sRGB EOTF per 8-bit code, then code / 255
};

struct FrameDev {
    RtUniforms uniforms;
    uint32_t width, height;         // launch size
    uint32_t max_segments, shadow_rays;
    uint32_t x0, y0, tw, rows;      // rendered rectangle: tw columns, `rows` compact rows
    uint32_t tile_h;
    uint32_t strip_height, strip_count, strip_index;
    float    cos_sun_radius;
    uint32_t image_rows;            // RT_RENDER_OUTPUT_IMAGE_ROWS: rgba8 / radiance are whole tile images, rows stored in place
    uint8_t*  rgba8;
    float*    radiance;
    uint32_t* hit_ids;
    uint32_t* cost;                 // show_heatmap frames: clock ticks per pixel (optional)
    float     heatmap_scale;        // ticks that map to heat 1.0
    FrameCounters* counters;
    unsigned int* bounce_hint;      // host-mapped word: bounce rays queued by segment 0 of the latest wavefront frame (optional)
    // denoise hook (segment 0 only): image-space sun factor written after the shadow rays, read by the resolve phase when set
    float*    sun_factor;
    float4*   position_nol;
    unsigned int* overflow_flag;    // host-mapped word, set when a traversal stack overflowed: the frame is incomplete and the
                                    // next synchronising call of the API reports RT_ERR_OUT_OF_RANGE (optional)
    RayRec*   ray_q[2];
    HitRec*   hit_q;
    // Shadow-ray directions of this frame, [64][64][shadow_rays] float4 indexed by (py & 63, px & 63, sample): the
    // blue-noise sample depends on the pixel only through (px, py) mod 64 (closest_hit_textured.glsl:99-111), so the
    // 4096 * N distinct directions are computed once per frame (k_sun_dirs).  nullptr: compute per ray.
    const float4* sun_dirs;
};

}  // namespace b200rt
