/* b200rt.h — C ABI of the B200-native frame hot path (libb200rt.so).
 *
 * The reference (expenses/ray-tracing-gallery) has no FFI of its own: the seam
 * is the set of Rust calls that reach the Vulkan driver.  Each export below
 * replaces one of those call sites (cited per function, paths relative to the
 * reference root).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (RtStatus);
 *     rt_last_error() gives the message.  Nothing throws or aborts across the
 *     ABI (the reference logs `anyhow` errors and keeps looping,
 *     src/main.rs:1030-1032).
 *   - the caller owns every input array; it is copied during the call (like
 *     the reference's staging buffers, src/util_functions.rs:289-294).
 *   - one host thread per context; calls are ordered on the context's CUDA
 *     stream and asynchronous unless stated otherwise.
 *   - there is NO CPU fallback: without a CUDA device rt_create fails.
 */
#ifndef B200RT_H
#define B200RT_H

#include "rt_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RtContext RtContext;

typedef enum RtStatus {
    RT_OK = 0,
    RT_ERR_INVALID_ARGUMENT = -1,
    RT_ERR_CUDA = -2,
    RT_ERR_OUT_OF_RANGE = -3,
    RT_ERR_NOT_BUILT = -4,      /* render/update before rt_build_tlas */
    RT_ERR_NO_DEVICE = -5
} RtStatus;

/* Texel formats of the bindless image table.
 * vk::Format::R8G8B8A8_UNORM / R8G8B8A8_SRGB: src/util_functions.rs:239-265,
 * R32G32B32A32_SFLOAT (1x1 constants): src/util_functions.rs:216-237. */
typedef enum RtFormat {
    RT_FORMAT_RGBA8_UNORM = 0,
    RT_FORMAT_RGBA8_SRGB = 1,
    RT_FORMAT_RGBA32_SFLOAT = 2
} RtFormat;

/* One geometry = one glTF material's triangles (src/util_structs.rs:911-916). */
typedef struct RtGeometryDesc {
    const uint32_t*  indices;       /* 3 per triangle, already rebased to the model's vertex arrays */
    uint32_t         num_indices;
    uint8_t          opaque;        /* alphaMode == OPAQUE (src/util_structs.rs:992); 0 => any-hit alpha clip */
    uint8_t          _pad[3];
    RtGeometryImages images;
} RtGeometryDesc;

/* `ModelArrays`, src/util_structs.rs:903-909. */
typedef struct RtModelDesc {
    const float*          positions;   /* num_vertices * 3 */
    const float*          normals;     /* num_vertices * 3 */
    const float*          uvs;         /* num_vertices * 2 */
    uint32_t              num_vertices;
    uint32_t              num_geometries;
    const RtGeometryDesc* geometries;
} RtModelDesc;

typedef enum RtUpdateMode {
    RT_UPDATE_AUTO = 0,     /* library picks: refit, and a full rebuild once the instance records written since the last
                               build add up to 4x the instance count (topology drift) */
    RT_UPDATE_REFIT = 1,    /* keep topology, refit boxes: VK mode UPDATE, src/util_structs.rs:309 */
    RT_UPDATE_REBUILD = 2,  /* full rebuild, stream-ordered: a new binned-SAH tree in one cooperative launch (0.75 ms for 10 k instances, 9 ms for 1 M) */
    RT_UPDATE_REBUILD_FAST = 3 /* full rebuild on the Morton radix tree (VK PREFER_FAST_BUILD): 0.33 ms / 2.1 ms, frames trace 20-30 % slower */
} RtUpdateMode;

typedef enum RtPipeline {
    RT_PIPELINE_WAVEFRONT = 0,  /* ray-gen/trace -> compacted hit + ray queues -> shade/shadow */
    RT_PIPELINE_MEGAKERNEL = 1  /* one thread per pixel runs the whole segment loop (A/B baseline) */
} RtPipeline;

enum {
    RT_RENDER_COUNTERS = 1u,    /* also count nodes/instances/triangles visited (slower; for the roofline audit) */
    RT_RENDER_TIMING = 2u,      /* CUDA events around every kernel of the frame -> RtStats.kernel_ms */
    RT_RENDER_SPLIT_TAIL = 4u,  /* bounce segments as four launches each instead of the one cooperative k_tail.  Without this
                                   flag (and without RT_RENDER_COOP_TAIL) the library picks per frame from the bounce-ray
                                   count of the latest finished frame; the frames are identical either way */
    RT_RENDER_NO_PDL = 8u,      /* plain stream-ordered launches instead of programmatic dependent launches (A/B) */
    RT_RENDER_COOP_TAIL = 32u,  /* always the one cooperative k_tail (A/B) */
    RT_RENDER_OUTPUT_IMAGE_ROWS = 16u /* rt_render_device: out->rgba8 / out->radiance are whole [tile_h][tile_w] images and the call
                                   stores its own rows in place (strip partition).  All ranks of a multi-GPU frame can then
                                   store into ONE frame — rank 0's, mapped through NVLink peer memory — instead of
                                   gathering compact slabs afterwards */
};

/* What `cmd_trace_rays(width, height, 1)` + the hard-coded shader constants
 * mean (src/command_buffer_recording.rs:116-126, lib.rs:144,
 * closest_hit_textured.glsl:195). */
typedef struct RtRenderParams {
    uint32_t width, height;     /* gl_LaunchSizeEXT: ALWAYS the full image, also when rendering a tile */
    uint32_t max_segments;      /* ray-gen loop bound; reference = 3 */
    uint32_t shadow_rays;       /* shadow rays per textured hit; reference = 2 */
    uint32_t tile_x0, tile_y0;  /* rectangle of global pixel coords to render; */
    uint32_t tile_w, tile_h;    /*   tile_w == 0 => full image */
    uint32_t strip_height;      /* >0: rows of the tile are dealt to `strip_count` ranks in strips */
    uint32_t strip_count;       /*   of this height, round-robin; this call renders strips with */
    uint32_t strip_index;       /*   (strip % strip_count) == strip_index (the last strip may be partial, */
                                /*   shares may differ by one strip).  0 => no interleave */
    uint32_t pipeline;          /* RtPipeline */
    uint32_t flags;             /* RT_RENDER_* */
    float    heatmap_scale;     /* Uniforms.show_heatmap: clock ticks that map to heat 1.0; 0 => 1 000 000, the
                                   reference's hard-coded `heatmap_scale` (lib.rs:179) */
    uint32_t _reserved[2];
} RtRenderParams;

/* Output arrays are compact over the rendered rows (ascending global y), row
 * pitch = tile_w pixels — or, with RT_RENDER_OUTPUT_IMAGE_ROWS (rt_render_device*), whole
 * [tile_h][tile_w] images of which the call stores its own rows.  Any pointer may be NULL. */
typedef struct RtFrameOutputs {
    uint8_t*  rgba8;        /* [rows][tile_w][4]  linear_to_srgb + UNORM8 store, alpha 255 (lib.rs:188-190) */
    float*    radiance;     /* [rows][tile_w][3]  payload.colour before the sRGB encode */
    uint32_t* hit_ids;      /* [rows][tile_w][max_segments][3] = (gl_InstanceID, gl_GeometryIndexEXT,
                               gl_PrimitiveID) per ray-gen segment; 0xFFFFFFFF x3 = miss / segment not traced */
    uint64_t* ray_counts;   /* [2] trace calls issued: {ray-gen segments, shadow rays} */
    uint32_t* cost_cycles;  /* [rows][tile_w]  SM clock ticks the pixel's ray-gen invocation took (saturating) — the
                               `delta_time` of lib.rs:174-177.  Only written by frames with Uniforms.show_heatmap set */
} RtFrameOutputs;

typedef struct RtStats {
    uint64_t primary_rays;      /* ray-gen trace calls of the last render */
    uint64_t shadow_rays;
    uint64_t textured_hits;
    /* the next four only with RT_RENDER_COUNTERS; [0] = closest-hit rays (ray-gen segments),
       [1] = shadow rays */
    uint64_t nodes_visited[2];
    uint64_t instances_entered[2];
    uint64_t triangles_tested[2];
    uint64_t anyhit_calls[2];
    float    last_render_ms;    /* CUDA-event time of the last rt_render*, valid after rt_sync */
    float    last_tlas_ms;      /* CUDA-event time of the last rt_build_tlas / rt_update_tlas */
    float    kernel_ms[6];      /* with RT_RENDER_TIMING: {trace kernels, k_prep, k_shadow, k_resolve (split-tail path only),
                                   k_mega, k_tail (resolve + bounce segments + ray-count export)} */
    uint32_t kernel_launches[6];/* launches behind kernel_ms */
    uint32_t tlas_nodes;        /* wide (8-child) nodes in the current TLAS */
    uint32_t blas_nodes;        /* over all models */
    uint32_t num_instances;
    uint32_t num_triangles;     /* over all models */
    uint32_t segment_rays[8];   /* wavefront: bounce rays queued BY ray-gen segment s (traced in segment s+1) */
    uint32_t segment_hits[8];   /* wavefront: textured hits queued by ray-gen segment s */
} RtStats;

#if defined(__cplusplus)
static_assert(sizeof(RtRenderParams) == 64, "RtRenderParams is 64 bytes");
static_assert(sizeof(RtFrameOutputs) == 5 * sizeof(void*), "RtFrameOutputs is five pointers");
#else
_Static_assert(sizeof(RtRenderParams) == 64, "RtRenderParams is 64 bytes");
_Static_assert(sizeof(RtFrameOutputs) == 5 * sizeof(void*), "RtFrameOutputs is five pointers");
#endif

/* Device/allocator creation: src/main.rs:157-204,337.  One context per GPU. */
int  rt_create(int cuda_device, RtContext** out);
/* Explicit teardown, like the reference's cleanup() chain, src/main.rs:997-1023. */
void rt_destroy(RtContext* ctx);
/* Message of the last failed call on this context (ctx may be NULL: last rt_create failure). */
const char* rt_last_error(const RtContext* ctx);
/* Run this context's work on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the
 * context's own non-blocking stream.  NULL names the legacy default stream, as in the CUDA runtime. */
int  rt_set_stream(RtContext* ctx, void* cuda_stream);

/* load_png_image_from_bytes / create_single_colour_image + ImageManager::push_image
 * (src/util_functions.rs:216-265, src/util_structs.rs:1330-1349).  Indices are
 * dense in push order; the reference pushes green, pink, blue-noise, GGX LUT
 * as 0..3 first (src/main.rs:416-460).  linear_filter: the sampler choice of
 * src/util_structs.rs:954-955.  texels: w*h*4 bytes (RGBA8) or w*h*16 (RGBA32F). */
int  rt_push_image(RtContext* ctx, const void* texels, uint32_t width, uint32_t height,
                   uint32_t format, int linear_filter, uint32_t* out_index);

/* Model::new + AccelerationStructure::build_blas + ModelInfo push
 * (src/util_structs.rs:1158-1236, 140-224, 1148-1153).  out_model_id is the
 * index into ModelInfo[] (= instance custom index); out_blas_handle is what
 * the host writes into instance bytes 56..64. */
int  rt_create_model(RtContext* ctx, const RtModelDesc* desc,
                     uint32_t* out_model_id, uint64_t* out_blas_handle);

/* build_tlas (src/util_functions.rs:453-510; caller src/main.rs:524-530).  The one-time, PREFER_FAST_TRACE build: binned-SAH
 * tree (0.75 ms for 10 k instances, 9 ms for 1 M; stream-ordered).  Per-frame changes go through rt_update_tlas. */
int  rt_build_tlas(RtContext* ctx, const RtInstance* instances, uint32_t count);

/* Buffer::write_mapped on the instance buffer (src/scene.rs:177-181) ... */
int  rt_update_instances(RtContext* ctx, uint32_t first, uint32_t count, const RtInstance* host_records);
/* ... same, from device memory (e.g. the landing buffer of an NCCL broadcast). */
int  rt_update_instances_device(RtContext* ctx, uint32_t first, uint32_t count, const void* device_records);
/* ... then AccelerationStructure::update_tlas + the AS-write -> RT-read barrier
 * (src/util_structs.rs:285-357, src/scene.rs:183-201).  Stream-ordered. */
int  rt_update_tlas(RtContext* ctx, uint32_t mode /* RtUpdateMode */);

/* Uniforms write + push constants + cmd_trace_rays
 * (src/main.rs:942-948, src/command_buffer_recording.rs:102-126).
 * rt_render: `out` holds HOST pointers (or is NULL); the frame is rendered into
 * the context's own device framebuffer, copied to the host arrays, and the call
 * returns when they are filled.
 * rt_render_device: `out` holds DEVICE pointers supplied by the caller; the call
 * only enqueues work on the context's stream. */
int  rt_render(RtContext* ctx, const RtUniforms* uniforms, const RtRenderParams* params,
               const RtFrameOutputs* out);
int  rt_render_device(RtContext* ctx, const RtUniforms* uniforms, const RtRenderParams* params,
                      const RtFrameOutputs* out);
/* Two frames in flight, like the reference's PerFrameResources + one fence per frame
 * (src/command_buffer_recording.rs:22-30, src/main.rs:917-928).  rt_render_async enqueues the frame into
 * the next of two frame slots and returns at once: the render runs on the context's stream, the
 * copy of the finished RGBA8 rows (and ray counts) to `out`'s HOST pointers on a second stream, so the
 * copy of frame i overlaps the rendering of frame i+1.  The host arrays (use pinned memory) are valid
 * after rt_wait_frame(slot).  If the slot is still busy with the frame from two calls ago, the call
 * waits for it first (the reference's wait_for_fences).  Only out->rgba8 and out->ray_counts are
 * supported here. */
int  rt_render_async(RtContext* ctx, const RtUniforms* uniforms, const RtRenderParams* params,
                     const RtFrameOutputs* out, uint32_t* out_slot);
int  rt_wait_frame(RtContext* ctx, uint32_t slot);
/* The device-output form of the same idea, for hosts that own the streams (e.g. one process per GPU with the frame
 * going to another rank): render into caller-owned DEVICE outputs on the caller's cudaStream_t using the private
 * queues of frame slot 0 or 1, so that two frames on two streams overlap.  Frames that use the same slot must be
 * ordered by the caller (same stream, or events).  Scene changes made through this context are ordered before and
 * after the frame by the library. */
int  rt_render_device_slot(RtContext* ctx, uint32_t slot, void* cuda_stream, const RtUniforms* uniforms,
                           const RtRenderParams* params, const RtFrameOutputs* out);
/* Storage-image copy / present (src/command_buffer_recording.rs:165-179): copy the
 * RGBA8 rows of the last rt_render(…, NULL) / rt_render frame to host memory (blocking). */
int  rt_readback(RtContext* ctx, void* host_rgba8, size_t capacity_bytes);
/* Fence wait (src/main.rs:919-923). */
int  rt_sync(RtContext* ctx);

/* Shadow-denoise hook.  The reference's readme lists "shadow denoising" as its next step (readme.md:17-20); this is the seam
 * for it in the frame: between the shadow rays of the first ray-gen segment and the colour resolve, the per-pixel sun
 * factor (`sun_factor = unshadowed / N`, closest_hit_textured.glsl:203) is laid out in image space and handed to the hook,
 * which may filter it in place with work enqueued on the given stream; the resolve then uses the filtered value.  Pixels
 * whose first segment has no textured hit hold 1.0 and are not read back.  Wavefront frames only (the megakernel and heat-
 * map paths shade inside one thread and ignore the hook); bounce segments keep their raw factor.  With an identity
 * hook the frame is bit-identical to a frame without hook. */
typedef struct RtDenoiseBuffers {
    uint32_t width, rows;        /* the rendered rectangle, compact like the frame outputs */
    float*       sun_factor;     /* [rows][width] in / out */
    const float* position_nol;   /* [rows][width][4] guide: world-space shadow-ray origin of the hit (xyz) and N.L (w); 0 where no hit */
    uint32_t     shadow_rays;    /* N of this frame */
    uint32_t     frame_index;
} RtDenoiseBuffers;
typedef int (*RtDenoiseFn)(void* user, void* cuda_stream, const RtDenoiseBuffers* buffers);  /* return 0, or the frame fails */
int  rt_set_denoise_hook(RtContext* ctx, RtDenoiseFn fn, void* user);   /* fn == NULL removes the hook */
/* A ready-made hook with the RtDenoiseFn signature: 5x5 cross-bilateral average of the sun factor, weights from the
 * world-space distance between the hits (`user` points to a float: the distance at which the weight has dropped to 1/e, in scene
 * units; NULL = 0.25) and from N.L sign agreement. */
int  rt_denoise_bilateral(void* user, void* cuda_stream, const RtDenoiseBuffers* buffers);

/* Page-locked host memory for frame read-back and instance uploads (what the reference gets from its host-visible
 * `Buffer`s, src/util_structs.rs:17-120): copies to/from it are asynchronous, so rt_render_async really overlaps. */
int  rt_host_alloc(RtContext* ctx, size_t bytes, void** out);
int  rt_host_free(RtContext* ctx, void* ptr);

int  rt_get_stats(RtContext* ctx, RtStats* out);
/* The three device addresses the reference pushes as push constants
 * (ModelInfo[] table, Uniforms copy, TLAS root) — exposed for layout parity checks. */
int  rt_get_push_constants(RtContext* ctx, RtPushConstantBufferAddresses* out);
/* Copy `count` ModelInfo / `count` GeometryInfo records of a model back (tests: the
 * device tables carry the reference's 32/24-byte layouts). */
int  rt_debug_read_model_info(RtContext* ctx, uint32_t model_id, RtModelInfo* out_info,
                              RtGeometryInfo* out_geoms, uint32_t max_geoms);
/* ---------------------------------------------------------------------------------------------------------------
 * One frame on several GPUs of one box.  The reference is single-GPU: its device creation
 * (src/main.rs:157-204) and its per-frame scene update (src/scene.rs:167-204) are the two seams a multi-GPU host
 * goes through, and they map to rt_group_create and rt_group_update_instances.  One process (or thread) per GPU, one
 * RtContext each, every rank holding the whole scene; a group joins the ranks:
 *   - image rows are dealt to the ranks in strips of RT_GROUP_STRIP_ROWS rows, round-robin (pixels are independent,
 *     blue noise is indexed by the global pixel: shaders/closest_hit_textured.glsl:100);
 *   - on a TLAS change the instance records go from the root rank to all ranks with ONE ncclBroadcast over NVLink that
 *     lands directly in the buffer the TLAS builder reads, followed by the refit / rebuild on every rank;
 *   - the frame reaches rank 0 either in DEVICE memory — every rank's render kernels store their rows straight into rank
 *     0's frame through NVLink peer memory (cudaIpc mapping) and raise a flag there, no collective, no copy kernel — or
 *     in HOST memory — every rank copies its own strips over its own PCIe link into one page-locked frame shared by
 *     the ranks (POSIX shared memory, registered with CUDA on every rank), which rank 0 then reads in place.
 * NCCL is loaded at run time (dlopen of libnccl.so.2, the copy already in the process if there is one); single-GPU
 * users of this library do not need it.  No torch, no MPI: the only thing the host has to move between the ranks
 * is the 128-byte id of rt_group_unique_id. */
typedef struct RtGroup RtGroup;
enum { RT_GROUP_ID_BYTES = 128, RT_GROUP_STRIP_ROWS = 8, RT_GROUP_MAX_RANKS = 16, RT_GROUP_FRAME_SLOTS = 4 };

/* Rank 0: make the id (ncclGetUniqueId); hand its bytes to every rank by any means. */
int  rt_group_unique_id(void* out_id_128_bytes);
/* All ranks, collectively.  `ctx` is this rank's context (its GPU, its scene).  Frames of `width` x `height` pixels. */
int  rt_group_create(RtContext* ctx, int n_ranks, int rank, const void* id_128_bytes, uint32_t width, uint32_t height,
                     RtGroup** out);
void rt_group_destroy(RtGroup* group);
const char* rt_group_last_error(const RtGroup* group);
/* Fill the strip fields of `params` for this rank (rows r with (r / 8) % n_ranks == rank); rows this rank renders. */
uint32_t rt_group_partition(const RtGroup* group, RtRenderParams* params);

/* All ranks, collectively: DefaultScene::write_resources for a multi-GPU box (src/scene.rs:167-204).  `host_records`
 * (count x 64 bytes, rank `root` only, may be NULL elsewhere) -> ncclBroadcast into the staging instance buffer at
 * `first` -> rt_update_tlas(mode) on every rank.  Stream-ordered on each rank's context stream. */
int  rt_group_update_instances(RtGroup* group, int root, uint32_t first, uint32_t count, const RtInstance* host_records,
                               uint32_t mode /* RtUpdateMode */);
/* Same, the root's records already in device memory of the root's GPU (the counterpart of rt_update_instances_device). */
int  rt_group_update_instances_device(RtGroup* group, int root, uint32_t first, uint32_t count, const void* device_records,
                                      uint32_t mode /* RtUpdateMode */);

/* All ranks, collectively: build_tlas for a multi-GPU box (src/util_functions.rs:453-510), SHARDED.  The root's
 * `count` records are broadcast; every rank computes the Morton keys, the ranks agree on n_ranks contiguous key ranges of
 * equal population, each rank builds the wide BVH of ITS range only (1/N of the SAH tree, fit, SAH collapse), the
 * treelets are exchanged over NVLink (NCCL broadcasts of exactly the nodes in use) and every rank puts the same top node
 * over them.  Frames do not depend on topology (closest hit + tie rule), so they equal those of rt_build_tlas bit for
 * bit; later rt_group_update_instances refits work on the assembled tree.  Falls back to the replicated build for more
 * than 8 ranks or fewer than 1024 instances per rank, unless RT_GROUP_BUILD_FORCE_SHARDED is set (tests: one rank). */
enum { RT_GROUP_BUILD_FORCE_SHARDED = 1u };
int  rt_group_build_tlas(RtGroup* group, int root, const RtInstance* host_records, uint32_t count, uint32_t flags);

/* All ranks: enqueue frame number `seq` (1, 2, 3 ... the same on every rank).  This rank renders its strips; `params`
 * carries width / height / max_segments / shadow_rays / pipeline / flags, the strip fields are filled by the group.
 *   rt_group_render_device: rows are stored into rank 0's DEVICE frame (slot seq % RT_GROUP_FRAME_SLOTS) over NVLink
 *     peer memory, then this rank's arrival flag is raised in rank 0's memory.
 *   rt_group_render_host: rows are rendered locally and copied into the shared page-locked HOST frame, this rank's
 *     ray counts next to them; arrival is published from a stream callback.
 * A slot is reused every RT_GROUP_FRAME_SLOTS frames: the calls wait until rank 0 has released the frame that used it. */
int  rt_group_render_device(RtGroup* group, uint64_t seq, const RtUniforms* uniforms, const RtRenderParams* params);
int  rt_group_render_host(RtGroup* group, uint64_t seq, const RtUniforms* uniforms, const RtRenderParams* params);
/* Rank 0: the finished frame.
 *   rt_group_acquire_device enqueues, on rank 0's context stream, a wait for every rank's arrival flag of frame `seq`;
 *     work enqueued on that stream afterwards sees the whole frame at *out_device_rgba8 ([height][width][4]).
 *   rt_group_acquire_host blocks until every rank's rows of frame `seq` are in host memory (bounded wait: an error
 *     after `timeout_ms`); *out_host_rgba8 points into the shared frame, ray_counts[2] are summed over the ranks.
 * rt_group_release(seq) gives the slot back (device path: stream-ordered after what rank 0 enqueued so far). */
int  rt_group_acquire_device(RtGroup* group, uint64_t seq, uint8_t** out_device_rgba8);
int  rt_group_acquire_host(RtGroup* group, uint64_t seq, uint32_t timeout_ms, const uint8_t** out_host_rgba8, uint64_t* ray_counts);
int  rt_group_release(RtGroup* group, uint64_t seq);
/* Rank 0, for hosts without a CUDA runtime of their own (the counterpart of rt_readback): wait for frame `seq` of the
 * device path and copy it to host memory (blocking).  Does not release the slot. */
int  rt_group_readback(RtGroup* group, uint64_t seq, void* host_rgba8, size_t capacity_bytes);
/* All ranks: ray counts {ray-gen segments, shadow rays} of this rank's share of the last device-path frame (device pointer,
 * valid in stream order after rt_group_render_device). */
int  rt_group_local_ray_counts(RtGroup* group, uint64_t** out_device_counts);
/* All ranks, collectively: every rank's stream work is finished (ncclAllReduce of one word + stream sync). */
int  rt_group_barrier(RtGroup* group);

/* Audit of the traversal's box test (tests/test_gpu_parity.py::test_box_test_is_conservative): runs it — the packed-bf16
 * node_hit_mask of trace.cuh, exactly as a node visit does — for every pair of `num_rays` rays and wide nodes
 * [first_node, first_node + num_nodes) of the BLAS pool (tlas == 0) or of the current TLAS (tlas != 0).
 * rays: num_rays x 8 floats {origin xyz, tmin, direction xyz, tmax}.  out_masks: [num_rays][num_nodes][2] bytes, bit s = child
 * slot s reported hit (before the empty-slot mask): [0] first-hit form (no far clamp), [1] closest-hit form with the far limit
 * at tmax.  out_node_lines (optional): the nodes' first 128-byte lines (header + bf16 planes), num_nodes x 128 bytes. */
int  rt_debug_box_test(RtContext* ctx, int tlas, uint32_t first_node, uint32_t num_nodes, const float* rays, uint32_t num_rays,
                       uint8_t* out_masks, void* out_node_lines);
/* Micro-benchmark behind bench.py's `l2_frac`: read a device buffer of `bytes` (cache-resident when well below the 126 MB L2)
 * `repeats` times with 16-byte loads from a full persistent grid and report read bytes / kernel time (CUDA events). */
int  rt_debug_l2_read_bandwidth(RtContext* ctx, size_t bytes, uint32_t repeats, float* out_gb_per_s);
/* Number of this library's own kernels launched so far in the process (all contexts). */
uint64_t rt_kernel_launches(void);
/* Library/ABI version: (major << 16) | minor. */
uint32_t rt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
