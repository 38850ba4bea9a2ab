/* rt_abi.h — byte-exact host/device record layouts of the frame hot path.
 *
 * Every struct here mirrors one `#[repr(C)]` struct of the reference so that a
 * host written against the reference can hand its bytes to this library
 * unchanged.  Offsets are asserted below (C11 / C++11).
 *
 *   RtUniforms                  <- shared-structs/src/lib.rs:10-19
 *                                  (GLSL mirror shaders/hit_shader_common.glsl:3-13)
 *   RtModelInfo                 <- shared-structs/src/lib.rs:24-29
 *   RtGeometryImages            <- shared-structs/src/lib.rs:42-47
 *   RtGeometryInfo              <- shared-structs/src/lib.rs:34-37
 *   RtPushConstantBufferAddresses <- shared-structs/src/lib.rs:52-56
 *   RtInstance                  <- src/gpu_structs.rs:20-25
 *                                  (= VkAccelerationStructureInstanceKHR)
 */
#ifndef B200RT_RT_ABI_H
#define B200RT_RT_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 176 bytes, align 16.  Matrices are column-major (glam Mat4 / GLSL mat4). */
typedef struct RtUniforms {
    float    view_inverse[16];          /*   0 */
    float    proj_inverse[16];          /*  64 */
    float    sun_dir[3];                /* 128  glam Vec3A: 12 B payload + 4 B pad */
    uint32_t _padding;                  /* 140 */
    float    sun_radius;                /* 144 */
    uint32_t blue_noise_texture_index;  /* 148  (=2, src/main.rs:595) */
    uint32_t ggx_lut_texture_index;     /* 152  (=3, src/main.rs:596; bound, never sampled) */
    uint32_t frame_index;               /* 156  first rendered frame is 1 (src/main.rs:944) */
    uint8_t  show_heatmap;              /* 160  per-pixel clock heatmap instead of the colour (lib.rs:120-124, 174-186) */
    uint8_t  _tail_padding[15];         /* 161..175 */
} RtUniforms;

/* 32 bytes.  The u64 fields hold CUDA device pointers in this library. */
typedef struct RtModelInfo {
    uint64_t position_buffer_address;   /* float3[], stride 12 */
    uint64_t normal_buffer_address;     /* float3[], stride 12 */
    uint64_t uv_buffer_address;         /* float2[], stride 8  */
    uint64_t geometry_info_address;     /* RtGeometryInfo[]    */
} RtModelInfo;

typedef struct RtGeometryImages {
    uint32_t diffuse_image_index;
    uint32_t metallic_roughness_image_index;
    int32_t  normal_map_image_index;    /* -1 = none */
    uint32_t _padding;
} RtGeometryImages;

/* 24 bytes. */
typedef struct RtGeometryInfo {
    uint64_t         index_buffer_address;  /* uint32[3] per triangle */
    RtGeometryImages images;
} RtGeometryInfo;

/* 24 bytes: what the reference pushes before vkCmdTraceRaysKHR
 * (src/command_buffer_recording.rs:102-114). */
typedef struct RtPushConstantBufferAddresses {
    uint64_t model_info;
    uint64_t uniforms;
    uint64_t acceleration_structure;
} RtPushConstantBufferAddresses;

/* 64 bytes.  transform = rows 0..2 of the object->world matrix (row-major 3x4). */
typedef struct RtInstance {
    float    transform[12];                  /*  0 */
    uint32_t instance_custom_index_and_mask; /* 48  custom_index: low 24 bits, mask: high 8 */
    uint32_t sbt_record_offset_and_flags;    /* 52  sbt offset: low 24 bits, flags: high 8 */
    uint64_t acceleration_structure_device_address; /* 56  BLAS handle from rt_create_model */
} RtInstance;

/* Hit-shader selection: `HitShader` enum, src/main.rs:46-50. */
enum { RT_HIT_TEXTURED = 0, RT_HIT_MIRROR = 1, RT_HIT_PORTAL = 2 };

/* VK_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE_BIT_KHR — carried, no-op
 * (no ray of the path sets a cull flag, src/gpu_structs.rs:35-39). */
enum { RT_INSTANCE_TRIANGLE_FACING_CULL_DISABLE = 1 };

#define RT_MAX_BOUND_IMAGES 128u         /* src/main.rs:44 */

#if defined(__cplusplus)
#define RT_SA(c, m) static_assert(c, m)
#else
#define RT_SA(c, m) _Static_assert(c, m)
#endif
RT_SA(sizeof(RtUniforms) == 176, "Uniforms is 176 bytes");
RT_SA(offsetof(RtUniforms, proj_inverse) == 64, "proj_inverse @64");
RT_SA(offsetof(RtUniforms, sun_dir) == 128, "sun_dir @128");
RT_SA(offsetof(RtUniforms, sun_radius) == 144, "sun_radius @144");
RT_SA(offsetof(RtUniforms, blue_noise_texture_index) == 148, "blue_noise @148");
RT_SA(offsetof(RtUniforms, ggx_lut_texture_index) == 152, "ggx_lut @152");
RT_SA(offsetof(RtUniforms, frame_index) == 156, "frame_index @156");
RT_SA(offsetof(RtUniforms, show_heatmap) == 160, "show_heatmap @160");
RT_SA(sizeof(RtModelInfo) == 32, "ModelInfo is 32 bytes");
RT_SA(sizeof(RtGeometryImages) == 16, "GeometryImages is 16 bytes");
RT_SA(sizeof(RtGeometryInfo) == 24, "GeometryInfo is 24 bytes");
RT_SA(offsetof(RtGeometryInfo, images) == 8, "images @8");
RT_SA(sizeof(RtPushConstantBufferAddresses) == 24, "push constants are 24 bytes");
RT_SA(sizeof(RtInstance) == 64, "instance record is 64 bytes");
RT_SA(offsetof(RtInstance, instance_custom_index_and_mask) == 48, "custom index @48");
RT_SA(offsetof(RtInstance, sbt_record_offset_and_flags) == 52, "sbt offset @52");
RT_SA(offsetof(RtInstance, acceleration_structure_device_address) == 56, "blas @56");
#undef RT_SA

#ifdef __cplusplus
}
#endif
#endif /* B200RT_RT_ABI_H */
