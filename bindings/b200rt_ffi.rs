//! Rust binding of libb200rt (include/b200rt.h).  NOT compiled in this repository: the image has no Rust
//! toolchain.  It is the file a maintainer of the reference drops into `src/` next to `gpu_structs.rs`
//!; `tests/test_abi.py` keeps it in step with the header (every export declared).
//! The reference's own `#[repr(C)]` types pass through unchanged:
//!   shared_structs::{Uniforms, ModelInfo, GeometryInfo, GeometryImages, PushConstantBufferAddresses}
//!   gpu_structs::AccelerationStructureInstance
#![allow(non_camel_case_types, dead_code)]

use crate::gpu_structs::AccelerationStructureInstance;
use shared_structs::{GeometryImages, GeometryInfo, ModelInfo, PushConstantBufferAddresses, Uniforms};
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct RtContext {
    _private: [u8; 0],
}

/// Several GPUs of one box rendering one frame (`rt_group_*`): one process per GPU, one `RtContext` each.
#[repr(C)]
pub struct RtGroup {
    _private: [u8; 0],
}

pub const RT_GROUP_ID_BYTES: usize = 128;
pub const RT_GROUP_STRIP_ROWS: u32 = 8;
pub const RT_GROUP_MAX_RANKS: usize = 16;
pub const RT_GROUP_FRAME_SLOTS: u64 = 4;
pub const RT_GROUP_BUILD_FORCE_SHARDED: u32 = 1;

pub const RT_OK: c_int = 0;
pub const RT_ERR_INVALID_ARGUMENT: c_int = -1;
pub const RT_ERR_CUDA: c_int = -2;
pub const RT_ERR_OUT_OF_RANGE: c_int = -3;
pub const RT_ERR_NOT_BUILT: c_int = -4;
pub const RT_ERR_NO_DEVICE: c_int = -5;

pub const RT_FORMAT_RGBA8_UNORM: u32 = 0;
pub const RT_FORMAT_RGBA8_SRGB: u32 = 1;
pub const RT_FORMAT_RGBA32_SFLOAT: u32 = 2;

pub const RT_UPDATE_AUTO: u32 = 0;
pub const RT_UPDATE_REFIT: u32 = 1; // vk::BuildAccelerationStructureModeKHR::UPDATE, src/util_structs.rs:309
pub const RT_UPDATE_REBUILD: u32 = 2;
pub const RT_UPDATE_REBUILD_FAST: u32 = 3; // Morton radix tree instead of the SAH tree (vk PREFER_FAST_BUILD)

pub const RT_PIPELINE_WAVEFRONT: u32 = 0;
pub const RT_PIPELINE_MEGAKERNEL: u32 = 1;

pub const RT_RENDER_COUNTERS: u32 = 1;
pub const RT_RENDER_TIMING: u32 = 2;
pub const RT_RENDER_SPLIT_TAIL: u32 = 4;
pub const RT_RENDER_NO_PDL: u32 = 8;
pub const RT_RENDER_OUTPUT_IMAGE_ROWS: u32 = 16;
pub const RT_RENDER_COOP_TAIL: u32 = 32;

#[repr(C)]
pub struct RtGeometryDesc {
    pub indices: *const u32,
    pub num_indices: u32,
    pub opaque: u8,
    pub _pad: [u8; 3],
    pub images: GeometryImages, // shared-structs/src/lib.rs:42-47
}

/// `ModelArrays`, src/util_structs.rs:903-909
#[repr(C)]
pub struct RtModelDesc {
    pub positions: *const f32,
    pub normals: *const f32,
    pub uvs: *const f32,
    pub num_vertices: u32,
    pub num_geometries: u32,
    pub geometries: *const RtGeometryDesc,
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct RtRenderParams {
    pub width: u32,
    pub height: u32,
    pub max_segments: u32, // 3 = the loop bound of ray_generation, lib.rs:144
    pub shadow_rays: u32,  // 2 = closest_hit_textured.glsl:195
    pub tile_x0: u32,
    pub tile_y0: u32,
    pub tile_w: u32,
    pub tile_h: u32,
    pub strip_height: u32,
    pub strip_count: u32,
    pub strip_index: u32,
    pub pipeline: u32,
    pub flags: u32,
    pub heatmap_scale: f32, // 0 = 1_000_000.0, the `heatmap_scale` of lib.rs:179
    pub _reserved: [u32; 2],
}

#[repr(C)]
pub struct RtFrameOutputs {
    pub rgba8: *mut u8,
    pub radiance: *mut f32,
    pub hit_ids: *mut u32,
    pub ray_counts: *mut u64,
    pub cost_cycles: *mut u32, // show_heatmap frames: clock ticks per pixel (lib.rs:174-177)
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct RtStats {
    pub primary_rays: u64,
    pub shadow_rays: u64,
    pub textured_hits: u64,
    pub nodes_visited: [u64; 2],
    pub instances_entered: [u64; 2],
    pub triangles_tested: [u64; 2],
    pub anyhit_calls: [u64; 2],
    pub last_render_ms: f32,
    pub last_tlas_ms: f32,
    pub kernel_ms: [f32; 6],
    pub kernel_launches: [u32; 6],
    pub tlas_nodes: u32,
    pub blas_nodes: u32,
    pub num_instances: u32,
    pub num_triangles: u32,
    pub segment_rays: [u32; 8],
    pub segment_hits: [u32; 8],
}

/// Image-space buffers handed to the shadow-denoise hook (readme.md:17-20 names shadow denoising as the next step).
#[repr(C)]
pub struct RtDenoiseBuffers {
    pub width: u32,
    pub rows: u32,
    pub sun_factor: *mut f32,
    pub position_nol: *const f32,
    pub shadow_rays: u32,
    pub frame_index: u32,
}
pub type RtDenoiseFn = Option<unsafe extern "C" fn(user: *mut c_void, cuda_stream: *mut c_void, buffers: *const RtDenoiseBuffers) -> c_int>;

extern "C" {
    pub fn rt_create(cuda_device: c_int, out: *mut *mut RtContext) -> c_int; // src/main.rs:157-204,337
    pub fn rt_destroy(ctx: *mut RtContext); // src/main.rs:997-1023
    pub fn rt_last_error(ctx: *const RtContext) -> *const c_char;
    pub fn rt_set_stream(ctx: *mut RtContext, cuda_stream: *mut c_void) -> c_int;
    pub fn rt_push_image(ctx: *mut RtContext, texels: *const c_void, width: u32, height: u32, format: u32, linear_filter: c_int,
                         out_index: *mut u32) -> c_int; // src/util_structs.rs:1330-1349
    pub fn rt_create_model(ctx: *mut RtContext, desc: *const RtModelDesc, out_model_id: *mut u32, out_blas_handle: *mut u64) -> c_int; // src/util_structs.rs:1158-1236, 140-224
    pub fn rt_build_tlas(ctx: *mut RtContext, instances: *const AccelerationStructureInstance, count: u32) -> c_int; // src/util_functions.rs:453-510
    pub fn rt_update_instances(ctx: *mut RtContext, first: u32, count: u32, host_records: *const AccelerationStructureInstance) -> c_int; // src/scene.rs:177-181
    pub fn rt_update_instances_device(ctx: *mut RtContext, first: u32, count: u32, device_records: *const c_void) -> c_int;
    pub fn rt_update_tlas(ctx: *mut RtContext, mode: u32) -> c_int; // src/util_structs.rs:285-357
    pub fn rt_render(ctx: *mut RtContext, uniforms: *const Uniforms, params: *const RtRenderParams, out: *const RtFrameOutputs) -> c_int; // src/command_buffer_recording.rs:102-126
    pub fn rt_render_device(ctx: *mut RtContext, uniforms: *const Uniforms, params: *const RtRenderParams, out: *const RtFrameOutputs) -> c_int;
    pub fn rt_render_async(ctx: *mut RtContext, uniforms: *const Uniforms, params: *const RtRenderParams, out: *const RtFrameOutputs,
                           out_slot: *mut u32) -> c_int; // two PerFrameResources, src/main.rs:917-928
    pub fn rt_wait_frame(ctx: *mut RtContext, slot: u32) -> c_int; // wait_for_fences, src/main.rs:919-923
    pub fn rt_render_device_slot(ctx: *mut RtContext, slot: u32, cuda_stream: *mut c_void, uniforms: *const Uniforms,
                                 params: *const RtRenderParams, out: *const RtFrameOutputs) -> c_int;
    pub fn rt_readback(ctx: *mut RtContext, host_rgba8: *mut c_void, capacity_bytes: usize) -> c_int; // src/command_buffer_recording.rs:165-179
    pub fn rt_sync(ctx: *mut RtContext) -> c_int;
    pub fn rt_set_denoise_hook(ctx: *mut RtContext, hook: RtDenoiseFn, user: *mut c_void) -> c_int;
    pub fn rt_denoise_bilateral(user: *mut c_void, cuda_stream: *mut c_void, buffers: *const RtDenoiseBuffers) -> c_int;
    pub fn rt_host_alloc(ctx: *mut RtContext, bytes: usize, out: *mut *mut c_void) -> c_int; // host-visible Buffer, src/util_structs.rs:17-120
    pub fn rt_host_free(ctx: *mut RtContext, ptr: *mut c_void) -> c_int;
    pub fn rt_get_stats(ctx: *mut RtContext, out: *mut RtStats) -> c_int;
    pub fn rt_get_push_constants(ctx: *mut RtContext, out: *mut PushConstantBufferAddresses) -> c_int;
    pub fn rt_debug_read_model_info(ctx: *mut RtContext, model_id: u32, out_info: *mut ModelInfo, out_geoms: *mut GeometryInfo, max_geoms: u32) -> c_int;
    // ---- multi-GPU: device creation (src/main.rs:157-204) and the per-frame scene update (src/scene.rs:167-204) for a group of ranks
    pub fn rt_group_unique_id(out_id_128_bytes: *mut c_void) -> c_int;
    pub fn rt_group_create(ctx: *mut RtContext, n_ranks: c_int, rank: c_int, id_128_bytes: *const c_void, width: u32, height: u32,
                           out: *mut *mut RtGroup) -> c_int;
    pub fn rt_group_destroy(group: *mut RtGroup);
    pub fn rt_group_last_error(group: *const RtGroup) -> *const c_char;
    pub fn rt_group_partition(group: *const RtGroup, params: *mut RtRenderParams) -> u32;
    pub fn rt_group_update_instances(group: *mut RtGroup, root: c_int, first: u32, count: u32, host_records: *const AccelerationStructureInstance,
                                     mode: u32) -> c_int; // ncclBroadcast into the TLAS builder's input, then update_tlas on every rank
    pub fn rt_group_update_instances_device(group: *mut RtGroup, root: c_int, first: u32, count: u32, device_records: *const c_void, mode: u32) -> c_int;
    pub fn rt_group_build_tlas(group: *mut RtGroup, root: c_int, host_records: *const AccelerationStructureInstance, count: u32, flags: u32) -> c_int; // sharded build_tlas
    pub fn rt_group_render_device(group: *mut RtGroup, seq: u64, uniforms: *const Uniforms, params: *const RtRenderParams) -> c_int;
    pub fn rt_group_render_host(group: *mut RtGroup, seq: u64, uniforms: *const Uniforms, params: *const RtRenderParams) -> c_int;
    pub fn rt_group_acquire_device(group: *mut RtGroup, seq: u64, out_device_rgba8: *mut *mut u8) -> c_int;
    pub fn rt_group_acquire_host(group: *mut RtGroup, seq: u64, timeout_ms: u32, out_host_rgba8: *mut *const u8, ray_counts: *mut u64) -> c_int;
    pub fn rt_group_release(group: *mut RtGroup, seq: u64) -> c_int;
    pub fn rt_group_readback(group: *mut RtGroup, seq: u64, host_rgba8: *mut c_void, capacity_bytes: usize) -> c_int;
    pub fn rt_group_local_ray_counts(group: *mut RtGroup, out_device_counts: *mut *mut u64) -> c_int;
    pub fn rt_group_barrier(group: *mut RtGroup) -> c_int;
    pub fn rt_debug_box_test(ctx: *mut RtContext, tlas: c_int, first_node: u32, num_nodes: u32, rays: *const f32, num_rays: u32, out_masks: *mut u8,
                             out_node_lines: *mut c_void) -> c_int;
    pub fn rt_debug_l2_read_bandwidth(ctx: *mut RtContext, bytes: usize, repeats: u32, out_gb_per_s: *mut f32) -> c_int;
    pub fn rt_kernel_launches() -> u64;
    pub fn rt_version() -> u32;
}

/// `anyhow`-style error mapping: every call returns 0 or a negative `RtStatus`; the message is `rt_last_error`.
pub fn check(ctx: *mut RtContext, rc: c_int) -> anyhow::Result<()> {
    if rc == RT_OK {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(rt_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(anyhow::anyhow!("b200rt error {}: {}", rc, msg))
}
