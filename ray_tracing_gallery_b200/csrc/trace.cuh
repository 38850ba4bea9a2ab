// trace.cuh — two-level traversal of the compressed 8-wide BVH: what OpTraceRayKHR does for the
// reference (shaders/ray-tracing/src/lib.rs:67-82, shaders/closest_hit_textured.glsl:164-169).
//
// Semantics (VK_KHR_ray_tracing_pipeline):
//   * the ray is taken to object space with the instance's inverse 3x4; t is preserved;
//   * a triangle candidate is valid iff tmin < t < tmax (exclusive), no face culling;
//   * closest-hit rays commit the smallest t; an exact tie goes to the lowest
//     (gl_InstanceID, gl_GeometryIndexEXT, gl_PrimitiveID) so the result does not depend on
//     BVH topology or traversal order;
//   * candidates on non-opaque geometry run the alpha-clip any-hit (for both ray types);
//   * shadow rays (ANY = true) stop at the first accepted candidate.
//
// Mechanics: one thread per ray.  A node visit is five ld.global.nc.v4 (80 of the node's 128
// bytes).  The traversal stack holds "node groups" — (child_base, pending-hit mask, imask) — so a
// node costs one stack slot however many of its children were hit; slots were assigned at build
// time by child octant, so visiting pending bits in order of (slot XOR ray_octant) is
// front-to-back with no distance sort.  TLAS leaves (instances) that are hit but not entered yet
// are stacked as single entries.
#pragma once
#include "shade.cuh"

namespace b200rt {

#define RT_STACK_SIZE 48

struct Hit {
    float t, u, v;
    uint32_t inst_pos;     // TLAS leaf position of the instance (0xFFFFFFFF = miss)
    uint32_t instance_id;  // gl_InstanceID
    uint32_t geom, prim;
    uint32_t custom_sbt;
};

struct TraceCounters {
    uint32_t nodes, instances, tris, anyhits, overflow;
};

__device__ __forceinline__ float safe_rcp(float x) {
    float ax = fabsf(x);
    if (!(ax >= 1e-12f)) x = copysignf(1e-12f, x);
    return __frcp_rn(x);
}

// byte `sel` of `w` as float (exact): place the byte in the mantissa of 2^23 and subtract
__device__ __forceinline__ float byte_f(uint32_t w, uint32_t sel) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | sel)) - 8388608.0f;
}

// 8-bit mask of non-zero bytes of (lo, hi)
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t lo, uint32_t hi) {
    uint32_t a = ((__vcmpne4(lo, 0u) & 0x08040201u) * 0x01010101u) >> 24;
    uint32_t b = ((__vcmpne4(hi, 0u) & 0x08040201u) * 0x01010101u) >> 24;
    return a | (b << 4);
}

// reorder the 8 slot bits so that bit i holds slot (i ^ oct)
__device__ __forceinline__ uint32_t permute_by_octant(uint32_t m, uint32_t oct) {
    if (oct & 1u) m = ((m & 0xAAu) >> 1) | ((m & 0x55u) << 1);
    if (oct & 2u) m = ((m & 0xCCu) >> 2) | ((m & 0x33u) << 2);
    if (oct & 4u) m = ((m & 0xF0u) >> 4) | ((m & 0x0Fu) << 4);
    return m;
}

template <bool ANY, bool COUNT>
__device__ __forceinline__ bool trace_ray(const SceneDev& S, V3 o, V3 d, float tmin, float tmax, Hit& hit, TraceCounters& tc) {
    uint2 stack[RT_STACK_SIZE];
    int sp = 0;

    hit.t = tmax;
    hit.inst_pos = 0xFFFFFFFFu;
    hit.instance_id = 0xFFFFFFFFu;
    hit.geom = hit.prim = 0xFFFFFFFFu;
    hit.u = hit.v = 0.0f;
    hit.custom_sbt = 0;

    const Node8* __restrict__ nodes = S.tlas_nodes;
    V3 co = o, cd = d;
    float idx = safe_rcp(cd.x), idy = safe_rcp(cd.y), idz = safe_rcp(cd.z);
    uint32_t oct = (cd.x < 0.0f ? 1u : 0u) | (cd.y < 0.0f ? 2u : 0u) | (cd.z < 0.0f ? 4u : 0u);

    bool in_blas = false;
    int inst_sp = 0;
    uint32_t cur_inst_pos = 0, cur_instance_id = 0, cur_custom_sbt = 0;

    // current node group: root of the TLAS as slot 0 of a virtual parent
    uint32_t ng_base = 0, ng_bits = (1u << oct) | (1u << 8);
    uint32_t enter_inst = 0xFFFFFFFFu;  // TLAS leaf to enter next

    for (;;) {
        if (ng_bits & 0xFFu) {
            // ---- visit the nearest pending child of the current group
            uint32_t i = __ffs(ng_bits & 0xFFu) - 1;
            ng_bits &= ~(1u << i);
            uint32_t slot = i ^ oct;
            uint32_t child = ng_base + __popc((ng_bits >> 8) & 0xFFu & ((1u << slot) - 1u));
            if (ng_bits & 0xFFu) {
                if (sp < RT_STACK_SIZE) stack[sp++] = make_uint2(ng_base, ng_bits);
                else tc.overflow++;
            }
            const uint4* np = reinterpret_cast<const uint4*>(nodes + child);
            uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) tc.nodes++;

            float sx = __uint_as_float((uint32_t)((int)(int8_t)(n0.w & 0xFFu) + 127) << 23);
            float sy = __uint_as_float((uint32_t)((int)(int8_t)((n0.w >> 8) & 0xFFu) + 127) << 23);
            float sz = __uint_as_float((uint32_t)((int)(int8_t)((n0.w >> 16) & 0xFFu) + 127) << 23);
            float ax = sx * idx, ay = sy * idy, az = sz * idz;
            float bx = (__uint_as_float(n0.x) - co.x) * idx;
            float by = (__uint_as_float(n0.y) - co.y) * idy;
            float bz = (__uint_as_float(n0.z) - co.z) * idz;
            // near / far plane words per axis, chosen by ray direction
            uint32_t nx0, nx1, fx0, fx1, ny0, ny1, fy0, fy1, nz0, nz1, fz0, fz1;
            if (oct & 1u) { nx0 = n3.z; nx1 = n3.w; fx0 = n2.x; fx1 = n2.y; } else { nx0 = n2.x; nx1 = n2.y; fx0 = n3.z; fx1 = n3.w; }
            if (oct & 2u) { ny0 = n4.x; ny1 = n4.y; fy0 = n2.z; fy1 = n2.w; } else { ny0 = n2.z; ny1 = n2.w; fy0 = n4.x; fy1 = n4.y; }
            if (oct & 4u) { nz0 = n4.z; nz1 = n4.w; fz0 = n3.x; fz1 = n3.y; } else { nz0 = n3.x; nz1 = n3.y; fz0 = n4.z; fz1 = n4.w; }
            float tlimit = hit.t;
            uint32_t h = 0;
#pragma unroll
            for (int s = 0; s < 8; s++) {
                uint32_t sel = s & 3;
                float tnx = fmaf(byte_f(s < 4 ? nx0 : nx1, sel), ax, bx);
                float tny = fmaf(byte_f(s < 4 ? ny0 : ny1, sel), ay, by);
                float tnz = fmaf(byte_f(s < 4 ? nz0 : nz1, sel), az, bz);
                float tfx = fmaf(byte_f(s < 4 ? fx0 : fx1, sel), ax, bx);
                float tfy = fmaf(byte_f(s < 4 ? fy0 : fy1, sel), ay, by);
                float tfz = fmaf(byte_f(s < 4 ? fz0 : fz1, sel), az, bz);
                float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
                float tf = fminf(fminf(tfx, tfy), fminf(tfz, tlimit));
                if (tn <= tf * 1.000002f) h |= 1u << s;
            }
            uint32_t imask = n0.w >> 24;
            h &= nonzero_bytes(n1.z, n1.w);
            uint32_t hl = h & ~imask;
            ng_base = n1.x;
            ng_bits = permute_by_octant(h & imask, oct) | (imask << 8);

            // ---- leaves of this node
            uint64_t meta = ((uint64_t)n1.w << 32) | n1.z;
            if (in_blas) {
                while (hl) {
                    uint32_t s = __ffs(hl) - 1;
                    hl &= hl - 1;
                    uint32_t m = (uint32_t)(meta >> (8 * s)) & 0xFFu;
                    uint32_t first = n1.y + (m & 31u), cnt = m >> 5;
                    for (uint32_t k = 0; k < cnt; k++) {
                        const float4* tp = reinterpret_cast<const float4*>(S.tris + first + k);
                        float4 ta = __ldg(tp), tb = __ldg(tp + 1), tcv = __ldg(tp + 2);
                        if (COUNT) tc.tris++;
                        float t, u, v;
                        if (!tri_candidate(co, cd, v3(ta.x, ta.y, ta.z), v3(tb.x, tb.y, tb.z), v3(tcv.x, tcv.y, tcv.z), t, u, v)) continue;
                        if (!(t > tmin && t < tmax)) continue;
                        uint32_t prim = __float_as_uint(ta.w), gf = __float_as_uint(tb.w);
                        uint32_t geom = gf & ~RT_TRI_NON_OPAQUE;
                        if (!ANY) {
                            if (t > hit.t) continue;
                            if (t == hit.t && hit.inst_pos != 0xFFFFFFFFu) {
                                // exact tie: lowest (instance, geometry, primitive) wins
                                bool lower = cur_instance_id != hit.instance_id ? cur_instance_id < hit.instance_id
                                             : geom != hit.geom                ? geom < hit.geom
                                                                               : prim < hit.prim;
                                if (!lower) continue;
                            }
                        }
                        if (gf & RT_TRI_NON_OPAQUE) {
                            if (COUNT) tc.anyhits++;
                            if (!anyhit_accepts(S, cur_custom_sbt & 0xFFFFFFu, geom, prim, u, v)) continue;
                        }
                        hit.t = t; hit.u = u; hit.v = v;
                        hit.inst_pos = cur_inst_pos; hit.instance_id = cur_instance_id;
                        hit.geom = geom; hit.prim = prim; hit.custom_sbt = cur_custom_sbt;
                        if (ANY) return true;
                    }
                }
            } else if (hl) {
                // TLAS leaves: enter the first now, stack the others
                uint32_t s = __ffs(hl) - 1;
                hl &= hl - 1;
                enter_inst = n1.y + ((uint32_t)(meta >> (8 * s)) & 31u);
                while (hl) {
                    s = __ffs(hl) - 1;
                    hl &= hl - 1;
                    uint32_t pos = n1.y + ((uint32_t)(meta >> (8 * s)) & 31u);
                    if (sp < RT_STACK_SIZE) stack[sp++] = make_uint2(pos, 0x80000000u);
                    else tc.overflow++;
                }
            }
        }

        if (enter_inst != 0xFFFFFFFFu) {
            // ---- enter an instance: world ray -> object ray
            const float4* ip = reinterpret_cast<const float4*>(S.inst_rt + enter_inst);
            float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
            uint32_t pos = enter_inst;
            enter_inst = 0xFFFFFFFFu;
            uint32_t root = __float_as_uint(r3.x);
            if (root != 0xFFFFFFFFu && (__float_as_uint(r3.w) & 0xFFu)) {
                if (COUNT) tc.instances++;
                if (ng_bits & 0xFFu) {
                    if (sp < RT_STACK_SIZE) stack[sp++] = make_uint2(ng_base, ng_bits);
                    else tc.overflow++;
                }
                float inv[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
                co = xform_point(inv, o);
                cd = xform_vec(inv, d);
                idx = safe_rcp(cd.x); idy = safe_rcp(cd.y); idz = safe_rcp(cd.z);
                oct = (cd.x < 0.0f ? 1u : 0u) | (cd.y < 0.0f ? 2u : 0u) | (cd.z < 0.0f ? 4u : 0u);
                cur_inst_pos = pos;
                cur_instance_id = __float_as_uint(r3.y);
                cur_custom_sbt = __float_as_uint(r3.z);
                nodes = S.blas_nodes;
                ng_base = root;
                ng_bits = (1u << oct) | (1u << 8);
                in_blas = true;
                inst_sp = sp;
            }
            continue;
        }

        if ((ng_bits & 0xFFu) == 0u) {
            // ---- current group exhausted: pop
            if (in_blas && sp == inst_sp) {
                in_blas = false;
                nodes = S.tlas_nodes;
                co = o; cd = d;
                idx = safe_rcp(cd.x); idy = safe_rcp(cd.y); idz = safe_rcp(cd.z);
                oct = (cd.x < 0.0f ? 1u : 0u) | (cd.y < 0.0f ? 2u : 0u) | (cd.z < 0.0f ? 4u : 0u);
            }
            if (sp == 0) break;
            uint2 e = stack[--sp];
            if (e.y & 0x80000000u) enter_inst = e.x;
            else { ng_base = e.x; ng_bits = e.y; }
        }
    }
    return hit.inst_pos != 0xFFFFFFFFu;
}

}  // namespace b200rt
