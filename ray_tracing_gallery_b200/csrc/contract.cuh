// contract.cuh — the fp32 "arithmetic contract" of the geometric part of the path.
//
// Vulkan leaves evaluation order and fusing of shader arithmetic implementation-defined, and the
// reference never sees the traversal arithmetic at all (it is inside the driver).  This library
// fixes one order so that results are reproducible bit for bit: every function here is written with
// explicit round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn are never contracted by
// nvcc; fmaf is a single fused op; __fdiv_rn/__fsqrt_rn are IEEE regardless of -use_fast_math).
#pragma once
#include <cuda_runtime.h>

namespace b200rt {

struct V3 { float x, y, z; };
struct V2 { float x, y; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return v3(sub_(a.x, b.x), sub_(a.y, b.y), sub_(a.z, b.z)); }
__device__ __forceinline__ V3 add3(V3 a, V3 b) { return v3(add_(a.x, b.x), add_(a.y, b.y), add_(a.z, b.z)); }
__device__ __forceinline__ V3 scale3(V3 a, float s) { return v3(mul_(a.x, s), mul_(a.y, s), mul_(a.z, s)); }

// dot(a,b) := fma(az,bz, fma(ay,by, ax*bx))
__device__ __forceinline__ float dot3(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, mul_(a.x, b.x))); }
// cross(a,b).x := fma(ay,bz, -(az*by)), cyclic
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -mul_(a.z, b.y)), fmaf(a.z, b.x, -mul_(a.x, b.z)), fmaf(a.x, b.y, -mul_(a.y, b.x)));
}
// normalize(v) := v / sqrt(dot(v,v))
__device__ __forceinline__ V3 normalize3(V3 v) {
    float l = __fsqrt_rn(dot3(v, v));
    return v3(div_(v.x, l), div_(v.y, l), div_(v.z, l));
}
// row-major 3x4 times point / vector
__device__ __forceinline__ V3 xform_point(const float* m, V3 p) {
    return v3(fmaf(m[2], p.z, fmaf(m[1], p.y, fmaf(m[0], p.x, m[3]))),
              fmaf(m[6], p.z, fmaf(m[5], p.y, fmaf(m[4], p.x, m[7]))),
              fmaf(m[10], p.z, fmaf(m[9], p.y, fmaf(m[8], p.x, m[11]))));
}
__device__ __forceinline__ V3 xform_vec(const float* m, V3 v) {
    return v3(fmaf(m[2], v.z, fmaf(m[1], v.y, mul_(m[0], v.x))),
              fmaf(m[6], v.z, fmaf(m[5], v.y, mul_(m[4], v.x))),
              fmaf(m[10], v.z, fmaf(m[9], v.y, mul_(m[8], v.x))));
}
// mat3(gl_WorldToObject3x4EXT) * n: transpose of the inverse 3x3 (closest_hit_textured.glsl:69-72)
__device__ __forceinline__ V3 xform_normal(const float* inv, V3 n) {
    return v3(fmaf(inv[8], n.z, fmaf(inv[4], n.y, mul_(inv[0], n.x))),
              fmaf(inv[9], n.z, fmaf(inv[5], n.y, mul_(inv[1], n.x))),
              fmaf(inv[10], n.z, fmaf(inv[6], n.y, mul_(inv[2], n.x))));
}
// a*w.x + b*w.y + c*w.z := fma(c,wz, fma(b,wy, a*wx))   (hit_shader_common.glsl:79-81)
__device__ __forceinline__ float interp1(float a, float b, float c, V3 w) { return fmaf(c, w.z, fmaf(b, w.y, mul_(a, w.x))); }
__device__ __forceinline__ V3 interp3(V3 a, V3 b, V3 c, V3 w) {
    return v3(interp1(a.x, b.x, c.x, w), interp1(a.y, b.y, c.y, w), interp1(a.z, b.z, c.z, w));
}
__device__ __forceinline__ V2 interp2(V2 a, V2 b, V2 c, V3 w) {
    V2 r; r.x = interp1(a.x, b.x, c.x, w); r.y = interp1(a.y, b.y, c.y, w); return r;
}
// barycentric weights of (a,b,c) from the hit attributes (hit_shader_common.glsl:75-77)
__device__ __forceinline__ V3 bary_weights(float u, float v) { return v3(sub_(sub_(1.0f, u), v), u, v); }

// world->object 3x4 from the instance's object->world 3x4: adj/det, unfused, left to right.
__device__ __forceinline__ void invert_3x4(const float* m, float* o) {
    float a00 = m[0], a01 = m[1], a02 = m[2], tx = m[3];
    float a10 = m[4], a11 = m[5], a12 = m[6], ty = m[7];
    float a20 = m[8], a21 = m[9], a22 = m[10], tz = m[11];
    float c00 = sub_(mul_(a11, a22), mul_(a12, a21));
    float c01 = sub_(mul_(a12, a20), mul_(a10, a22));
    float c02 = sub_(mul_(a10, a21), mul_(a11, a20));
    float det = add_(add_(mul_(a00, c00), mul_(a01, c01)), mul_(a02, c02));
    float id = div_(1.0f, det);
    o[0] = mul_(c00, id);
    o[1] = mul_(sub_(mul_(a02, a21), mul_(a01, a22)), id);
    o[2] = mul_(sub_(mul_(a01, a12), mul_(a02, a11)), id);
    o[4] = mul_(c01, id);
    o[5] = mul_(sub_(mul_(a00, a22), mul_(a02, a20)), id);
    o[6] = mul_(sub_(mul_(a02, a10), mul_(a00, a12)), id);
    o[8] = mul_(c02, id);
    o[9] = mul_(sub_(mul_(a01, a20), mul_(a00, a21)), id);
    o[10] = mul_(sub_(mul_(a00, a11), mul_(a01, a10)), id);
    o[3] = -add_(add_(mul_(o[0], tx), mul_(o[1], ty)), mul_(o[2], tz));
    o[7] = -add_(add_(mul_(o[4], tx), mul_(o[5], ty)), mul_(o[6], tz));
    o[11] = -add_(add_(mul_(o[8], tx), mul_(o[9], ty)), mul_(o[10], tz));
}

// Moller-Trumbore candidate test against (v0, e1, e2); no culling.  The caller applies the
// exclusive interval tmin < t < tmax.
__device__ __forceinline__ bool tri_candidate(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float& t, float& u, float& v) {
    V3 p = cross3(d, e2);
    float det = dot3(e1, p);
    if (!(det != 0.0f)) return false;
    float inv = div_(1.0f, det);
    V3 tv = sub3(o, v0);
    u = mul_(dot3(tv, p), inv);
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    V3 q = cross3(tv, e1);
    v = mul_(dot3(d, q), inv);
    if (!(v >= 0.0f && add_(u, v) <= 1.0f)) return false;
    t = mul_(dot3(e2, q), inv);
    return true;
}

}  // namespace b200rt
