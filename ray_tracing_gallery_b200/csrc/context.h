// context.h — the state behind an RtContext (shared by api.cu and group.cu; internal, not part of the C ABI).
#pragma once
#include <string>
#include <vector>

#include "bvh_build.h"
#include "render.h"

namespace b200rt {

template <typename T>
struct DevVec {  // growable device array; indices stay valid across growth
    T* ptr = nullptr;
    size_t size = 0, cap = 0;
    cudaError_t reserve(size_t want, cudaStream_t stream) {
        if (want <= cap) return cudaSuccess;
        size_t ncap = cap ? cap : 1024;
        while (ncap < want) ncap += ncap / 2 + 1024;
        T* np = nullptr;
        cudaError_t e = cudaMalloc(&np, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (size) {
            e = cudaMemcpyAsync(np, ptr, size * sizeof(T), cudaMemcpyDeviceToDevice, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) { cudaFree(np); return e; }
        }
        if (ptr) cudaFree(ptr);
        ptr = np;
        cap = ncap;
        return cudaSuccess;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        size = cap = 0;
    }
};

struct ModelRes {
    float *positions = nullptr, *normals = nullptr, *uvs = nullptr;
    float4* bound_pts = nullptr;  // finite vertices some triangle references, when there are few enough (BlasInfo::verts)
    RtGeometryInfo* geom_info = nullptr;
    std::vector<uint32_t*> index_bufs;
    uint32_t num_geoms = 0;
    BlasInfo blas{};
};

struct TexRes {
    cudaArray_t array = nullptr;
    cudaTextureObject_t obj = 0;
};

}  // namespace b200rt

using namespace b200rt;  // internal header: RtContext itself has to live in the global namespace (opaque type of the C ABI)

// What one frame in flight owns: the stream its kernels run on and every buffer they write besides the outputs.
// The scene (models, BLASes, TLAS, images) is shared and read-only while frames render.
struct FrameResources {
    cudaStream_t stream = nullptr;
    FrameCounters* d_counters = nullptr;
    RayRec* d_ray_q[2] = {nullptr, nullptr};
    HitRec* d_hit_q = nullptr;
    size_t queue_cap = 0;
    float4* d_sun_dirs = nullptr;
    size_t sun_dirs_cap = 0;
    float* d_sun_factor = nullptr;       // denoise hook: image-space sun factor + guide of segment 0
    float4* d_position_nol = nullptr;
    size_t denoise_cap = 0;
    void release() {
        cudaFree(d_counters); cudaFree(d_ray_q[0]); cudaFree(d_ray_q[1]); cudaFree(d_hit_q); cudaFree(d_sun_dirs);
        cudaFree(d_sun_factor); cudaFree(d_position_nol);
        d_sun_factor = nullptr; d_position_nol = nullptr; denoise_cap = 0;
        d_counters = nullptr; d_ray_q[0] = d_ray_q[1] = nullptr; d_hit_q = nullptr; d_sun_dirs = nullptr;
        queue_cap = sun_dirs_cap = 0;
    }
};

struct RtContext {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // render start/stop, tlas start/stop
    bool render_timed = false, tlas_timed = false;
    FrameTiming timing;
    bool timing_ready = false, timing_valid = false;
    DenoiseHook denoise;             // rt_set_denoise_hook
    LaunchGeometry launch_geometry;  // grids of the persistent kernels on this context's device
    // Tail policy.  Bounce segments run either inside the one cooperative k_tail (one launch; best when there are few or no
    // bounce rays: C2 0.453 against 0.505 ms) or as separate launches at each kernel's own occupancy (better when a frame
    // bounces a lot: C3 0.927 -> 0.903 ms, C4 8.76 -> 8.55 ms, and consecutive frames overlap better: C3 e2e +10 %).  Frames
    // are coherent, so the choice follows the bounce-ray count of the latest finished frame, which the frame kernels leave
    // in a host-mapped word (no copy, no synchronisation).  B200RT_SPLIT_TAIL=0 / 1 pins the choice (A/B).
    int tail_policy = -1;                    // -1 adaptive, 0 always cooperative, 1 always split
    volatile unsigned int* h_bounce = nullptr;  // cudaHostAllocMapped, two words: [0] bounce hint, [1] traversal-stack overflow flag
    unsigned int* d_bounce = nullptr;           // its device alias
    std::string err;

    // images
    std::vector<TexRes> tex_res;
    std::vector<TexEntry> tex_host;
    TexEntry* d_textures = nullptr;
    float* d_srgb_lut = nullptr;
    float srgb_lut[512];

    // models
    std::vector<ModelRes> models;
    DevVec<RtModelInfo> d_model_info;
    DevVec<BlasInfo> d_blas_info;
    DevVec<Node8> blas_nodes;
    DevVec<TriRec> tris;          // BVH leaf order; a tiny BLAS is followed by one bounds record
    uint32_t num_triangles = 0;

    // instances / TLAS.  Two sets, like the reference's PerFrameResources (instance buffer + TLAS per frame in flight,
    // src/command_buffer_recording.rs:22-30): frames read set `cur`; instance writes and the TLAS update that follows go
    // to the other set and `cur` flips when the update is enqueued, so updating the scene for frame i+1 only waits for the
    // frames that still read the set being written (frame i-1), not for frame i.
    struct TlasSet {
        RtInstance* d_instances = nullptr;   // the caller's 64-byte records, by gl_InstanceID
        InstRT* d_inst_rt = nullptr;         // traversal records, TLAS leaf order
        uint32_t* d_leaf_order = nullptr;
        Node8* d_tlas_nodes = nullptr;
        uint32_t* d_node_count = nullptr;
    } sets[2];
    uint32_t cur = 0;
    bool staged = false;                     // set cur^1 holds the records of `cur` plus the writes since the last flip
    uint32_t num_instances = 0, inst_cap = 0;
    bool tlas_built = false;
    uint64_t writes_since_build = 0;         // instance records written since the last full build (RT_UPDATE_AUTO)
    InstRT* d_inst_unsorted = nullptr;       // builder inputs: only touched on the context's stream
    Aabb* d_inst_boxes = nullptr;
    uint32_t tlas_node_cap = 0;
    BvhBuilder builder;

    // frame
    RtUniforms* d_uniforms = nullptr;  // copy kept for the push-constant parity view
    FrameResources main;               // rt_render / rt_render_device (main.stream == stream)
    FrameResources* last_res = nullptr;  // resources of the most recent frame (rt_get_stats)
    uint8_t* d_fb_rgba8 = nullptr;
    float* d_fb_radiance = nullptr;
    uint32_t* d_fb_hit_ids = nullptr;
    uint32_t* d_fb_cost = nullptr;
    size_t fb_rgba8_cap = 0, fb_radiance_cap = 0, fb_hit_ids_cap = 0, fb_cost_cap = 0;
    size_t last_rows = 0, last_tw = 0;
    uint64_t* d_ray_counts = nullptr;

    // two frames in flight (rt_render_async): each slot renders on its own stream with its own queues, so consecutive
    // frames overlap on the GPU wherever one frame alone leaves it idle (kernel tails, stage boundaries)
    struct FrameSlot {
        FrameResources res;
        uint8_t* d_rgba8 = nullptr;
        size_t cap = 0;
        uint64_t* d_ray_counts = nullptr;
        cudaEvent_t scene_ready = nullptr, rendered = nullptr, copied = nullptr;
        bool pending = false;       // host has not waited for `copied` yet
        bool rendering = false;     // `rendered` may not have fired yet: scene changes must wait for it
        uint32_t reads_set = 0;     // the TLAS set the frame renders from
    } slots[2];
    cudaStream_t copy_stream = nullptr;
    uint32_t next_slot = 0;

    // rt_render_device_slot: the same per-frame resources for caller-owned streams and device outputs
    struct DeviceSlot {
        FrameResources res;
        cudaEvent_t scene_ready = nullptr, rendered = nullptr;
        bool rendering = false;
        uint32_t reads_set = 0;
    } dev_slots[2];
};

