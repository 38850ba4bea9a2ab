"""ctypes mirrors of include/rt_abi.h and include/b200rt.h.

Layouts follow the reference's `#[repr(C)]` structs byte for byte:
  Uniforms / ModelInfo / GeometryInfo / GeometryImages / PushConstantBufferAddresses
      shared-structs/src/lib.rs:10-56
  AccelerationStructureInstance          src/gpu_structs.rs:20-25
"""
import ctypes as C

import numpy as np


class RtUniforms(C.Structure):
    _fields_ = [
        ("view_inverse", C.c_float * 16),
        ("proj_inverse", C.c_float * 16),
        ("sun_dir", C.c_float * 3),
        ("_padding", C.c_uint32),
        ("sun_radius", C.c_float),
        ("blue_noise_texture_index", C.c_uint32),
        ("ggx_lut_texture_index", C.c_uint32),
        ("frame_index", C.c_uint32),
        ("show_heatmap", C.c_uint8),
        ("_tail_padding", C.c_uint8 * 15),
    ]


class RtModelInfo(C.Structure):
    _fields_ = [
        ("position_buffer_address", C.c_uint64),
        ("normal_buffer_address", C.c_uint64),
        ("uv_buffer_address", C.c_uint64),
        ("geometry_info_address", C.c_uint64),
    ]


class RtGeometryImages(C.Structure):
    _fields_ = [
        ("diffuse_image_index", C.c_uint32),
        ("metallic_roughness_image_index", C.c_uint32),
        ("normal_map_image_index", C.c_int32),
        ("_padding", C.c_uint32),
    ]


class RtGeometryInfo(C.Structure):
    _fields_ = [("index_buffer_address", C.c_uint64), ("images", RtGeometryImages)]


class RtPushConstantBufferAddresses(C.Structure):
    _fields_ = [("model_info", C.c_uint64), ("uniforms", C.c_uint64), ("acceleration_structure", C.c_uint64)]


class RtInstance(C.Structure):
    _fields_ = [
        ("transform", C.c_float * 12),
        ("instance_custom_index_and_mask", C.c_uint32),
        ("sbt_record_offset_and_flags", C.c_uint32),
        ("acceleration_structure_device_address", C.c_uint64),
    ]


# numpy view of the same 64 bytes, for bulk instance generation
INSTANCE_DTYPE = np.dtype(
    [("transform", "<f4", (12,)), ("custom_index_and_mask", "<u4"), ("sbt_offset_and_flags", "<u4"), ("blas", "<u8")]
)
assert INSTANCE_DTYPE.itemsize == 64


class RtGeometryDesc(C.Structure):
    _fields_ = [
        ("indices", C.POINTER(C.c_uint32)),
        ("num_indices", C.c_uint32),
        ("opaque", C.c_uint8),
        ("_pad", C.c_uint8 * 3),
        ("images", RtGeometryImages),
    ]


class RtModelDesc(C.Structure):
    _fields_ = [
        ("positions", C.POINTER(C.c_float)),
        ("normals", C.POINTER(C.c_float)),
        ("uvs", C.POINTER(C.c_float)),
        ("num_vertices", C.c_uint32),
        ("num_geometries", C.c_uint32),
        ("geometries", C.POINTER(RtGeometryDesc)),
    ]


class RtRenderParams(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("max_segments", C.c_uint32),
        ("shadow_rays", C.c_uint32),
        ("tile_x0", C.c_uint32),
        ("tile_y0", C.c_uint32),
        ("tile_w", C.c_uint32),
        ("tile_h", C.c_uint32),
        ("strip_height", C.c_uint32),
        ("strip_count", C.c_uint32),
        ("strip_index", C.c_uint32),
        ("pipeline", C.c_uint32),
        ("flags", C.c_uint32),
        ("heatmap_scale", C.c_float),
        ("_reserved", C.c_uint32 * 2),
    ]


class RtFrameOutputs(C.Structure):
    _fields_ = [
        ("rgba8", C.c_void_p),
        ("radiance", C.c_void_p),
        ("hit_ids", C.c_void_p),
        ("ray_counts", C.c_void_p),
        ("cost_cycles", C.c_void_p),
    ]


class RtDenoiseBuffers(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("rows", C.c_uint32),
        ("sun_factor", C.c_void_p),
        ("position_nol", C.c_void_p),
        ("shadow_rays", C.c_uint32),
        ("frame_index", C.c_uint32),
    ]


RT_DENOISE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(RtDenoiseBuffers))


class RtStats(C.Structure):
    _fields_ = [
        ("primary_rays", C.c_uint64),
        ("shadow_rays", C.c_uint64),
        ("textured_hits", C.c_uint64),
        ("nodes_visited", C.c_uint64 * 2),
        ("instances_entered", C.c_uint64 * 2),
        ("triangles_tested", C.c_uint64 * 2),
        ("anyhit_calls", C.c_uint64 * 2),
        ("last_render_ms", C.c_float),
        ("last_tlas_ms", C.c_float),
        ("kernel_ms", C.c_float * 6),
        ("kernel_launches", C.c_uint32 * 6),
        ("tlas_nodes", C.c_uint32),
        ("blas_nodes", C.c_uint32),
        ("num_instances", C.c_uint32),
        ("num_triangles", C.c_uint32),
        ("segment_rays", C.c_uint32 * 8),
        ("segment_hits", C.c_uint32 * 8),
    ]


RT_FORMAT_RGBA8_UNORM, RT_FORMAT_RGBA8_SRGB, RT_FORMAT_RGBA32_SFLOAT = 0, 1, 2
RT_HIT_TEXTURED, RT_HIT_MIRROR, RT_HIT_PORTAL = 0, 1, 2
RT_UPDATE_AUTO, RT_UPDATE_REFIT, RT_UPDATE_REBUILD, RT_UPDATE_REBUILD_FAST = 0, 1, 2, 3
RT_PIPELINE_WAVEFRONT, RT_PIPELINE_MEGAKERNEL = 0, 1
RT_RENDER_COUNTERS = 1
RT_RENDER_TIMING = 2
RT_RENDER_SPLIT_TAIL = 4
RT_RENDER_NO_PDL = 8
RT_RENDER_OUTPUT_IMAGE_ROWS = 16
RT_RENDER_COOP_TAIL = 32
RT_INSTANCE_TRIANGLE_FACING_CULL_DISABLE = 1
MISS_ID = 0xFFFFFFFF

assert C.sizeof(RtUniforms) == 176
assert C.sizeof(RtModelInfo) == 32
assert C.sizeof(RtGeometryInfo) == 24
assert C.sizeof(RtPushConstantBufferAddresses) == 24
assert C.sizeof(RtInstance) == 64
assert C.sizeof(RtRenderParams) == 64


def declare_api(lib, prefix):
    """Attach argtypes/restype for the shared entry points (`rt_*` in the product,
    `orc_*` in the oracle, which mirrors the same signatures minus the device id)."""
    p = C.c_void_p
    u32 = C.c_uint32

    def f(name, argtypes, restype=C.c_int):
        fn = getattr(lib, prefix + name)
        fn.argtypes = argtypes
        fn.restype = restype
        return fn

    f("push_image", [p, C.c_void_p, u32, u32, u32, C.c_int, C.POINTER(u32)])
    f("create_model", [p, C.POINTER(RtModelDesc), C.POINTER(u32), C.POINTER(C.c_uint64)])
    f("build_tlas", [p, C.c_void_p, u32])
    f("update_instances", [p, u32, u32, C.c_void_p])
    f("update_tlas", [p, u32])
    f("render", [p, C.POINTER(RtUniforms), C.POINTER(RtRenderParams), C.POINTER(RtFrameOutputs)])
    f("last_error", [p], C.c_char_p)
    f("destroy", [p], None)
