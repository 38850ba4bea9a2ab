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
// Mechanics: one thread per ray, written as a resumable state machine (`Traverser::step` = one
// node visit, instance entry or stack pop).  A node visit is eight ld.global.nc.v4: the first 128-byte
// line of the node (header + bf16 planes; the near / far rows of the ray's octant are chosen in the address).
// The stack holds "node groups" — (child_base, pending-hit mask, imask) — so a node costs one
// slot however many of its children were hit; slots were assigned at build time by child octant,
// so visiting pending bits in order of (slot XOR ray_octant) is front-to-back with no distance
// sort.  TLAS leaves (instances) that are hit but not entered yet are stacked as single entries.
//
// Box test: packed bf16, two children per instruction — see node_hit_mask below.
#pragma once
#include "shade.cuh"

namespace b200rt {

#define RT_NONE 0xFFFFFFFFu

struct Hit {
    float t, u, v;
    uint32_t inst_pos;     // TLAS leaf position of the instance (RT_NONE = miss)
    uint32_t instance_id;  // gl_InstanceID
    uint32_t geom, prim;
    uint32_t custom_sbt;   // custom index (24 low) | hit-shader kind (8 high)
};

struct TraceCounters {
    uint32_t nodes, instances, tris, anyhits, overflow;
};

// 1/x for the slab tests only (never for hit distances): one MUFU.RCP.  Its relative error (2^-23) moves every
// plane parameter by |t|*2^-23 <= (|b| + 255|a|)*2^-23, well inside the outward padding p of the box test.
__device__ __forceinline__ float safe_rcp(float x) {
    float ax = fabsf(x);
    if (!(ax >= 1e-12f)) x = copysignf(1e-12f, x);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
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

// ---- the box test: eight quantised child boxes against one ray, on packed bf16 arithmetic
//
// Per axis the ray parameter of grid plane q is t(q) = b + q*a with a = cell/d, b = (origin - o)/d (fp32).  The planes are
// stored as bf16 integers, two slots per word, so one HFMA2.BF16 evaluates a plane of two children.  bf16 has 8 significant
// bits; to keep that enough, everything is measured from a reference close to where the ray is inside the node:
//     t_ref = max(entry of the ray into the node's grid box [0, 255 cells]^3, tmin)      (fp32, guarded downwards)
//     u(q)  = t(q) - t_ref = c + q*a,  c = b - t_ref
// At t_ref the ray sits on (or inside) the grid box, i.e. at a cell position q_o in [0, 255] on every axis, and
// u(q) = (q - q_o)*a: |c| <= 255|a| and |u| <= 255|a| for every plane of the node.  Rounding errors (half an ulp = 2^-9):
//     a -> bf16:  q * |a| * 2^-9 <= 0.498 |a|          c -/+ pad -> bf16:  (255 + PAD)|a| * 2^-9 <= 0.502 |a|
//     the fma's own rounding:  |u| * 2^-9 <= 0.502 |a|
// together at most 1.503 |a| = 1.503 cells of that axis; near planes are therefore evaluated with c - pad, far planes with
// c + pad, pad = 1.5625 |a| plus the fp32 guard (|b| + |t_ref| + 256|a|) * 2^-21 that the fp32 version of this test carried:
// the test stays conservative (a child is never missed), its boxes are about 1.6 cells = 0.6 % of the node's extent fatter.
// A ray that misses the grid box altogether has no such bounds, but then it misses every child: whatever the arithmetic
// yields is at worst a wasted visit.  min / max are exact; tf - tn >= 0 exactly when tf >= tn (round to nearest, x - x = +0).
//
// Cost: 24 HFMA2.BF16 + 8 VHMNMX.BF16 (3-input) + 4..8 HMNMX2.BF16 + 4 HADD2 for the 48 planes, against 48 PRMT + 48 FFMA + 24 FMNMX(3) + 16 mask
// instructions of the fp32 version; the alu pipe (PRMT, FMNMX, LOP3, SHF: one warp instruction per two cycles per scheduler)
// was what bounded a node visit (profiles/r02p_summary.md: alu 71 %, fma 29 %).
__device__ __forceinline__ uint32_t bf2_fma(uint32_t q, uint32_t a, uint32_t c) {
    uint32_t r;
    asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(q), "r"(a), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t bf2_max(uint32_t x, uint32_t y) {
    uint32_t r;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    return r;
}
__device__ __forceinline__ uint32_t bf2_min(uint32_t x, uint32_t y) {
    uint32_t r;
    asm("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    return r;
}
__device__ __forceinline__ uint32_t bf2_sub(uint32_t x, uint32_t y) {
    uint32_t r;
    asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    return r;
}
// both halves = x rounded to nearest
__device__ __forceinline__ uint32_t bf2_splat(float x) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x), "f"(x));
    return r;
}

#ifndef RT_BOX_PAD_CELLS
#define RT_BOX_PAD_CELLS 1.5625f
#endif

// Hit mask (bit s = slot s) of the eight child boxes of a node against the ray (origin co, 1 / direction = id*) on
// [tmin, tlimit].  n0 = the node's first 16 bytes; nx..fz = its near / far planes per axis for this ray (Node8::q rows chosen
// by the sign of the direction: the caller loads them from octant-dependent addresses).  Empty slots are stored as inverted boxes and fail on
// their own.  FIRST_HIT: no far clamp.
template <bool FIRST_HIT>
__device__ __forceinline__ uint32_t node_hit_mask(const uint4& n0, const uint4& nx, const uint4& fx, const uint4& ny, const uint4& fy, const uint4& nz,
                                                  const uint4& fz, V3 co, float idx, float idy, float idz, float tmin, float tlimit) {
    // (the exponent bytes are stored biased: shifted into place they ARE the cell sizes 2^e)
    const float ax = __uint_as_float((n0.w << 23) & 0x7F800000u) * idx;
    const float ay = __uint_as_float((n0.w << 15) & 0x7F800000u) * idy;
    const float az = __uint_as_float((n0.w << 7) & 0x7F800000u) * idz;
    const float bx = (__uint_as_float(n0.x) - co.x) * idx;
    const float by = (__uint_as_float(n0.y) - co.y) * idy;
    const float bz = (__uint_as_float(n0.z) - co.z) * idz;
    // fp32 guards per axis, then the entry into the grid box
    const float gx = fmaf(fabsf(ax), 256.0f, fabsf(bx)) * 4.76837158e-7f;
    const float gy = fmaf(fabsf(ay), 256.0f, fabsf(by)) * 4.76837158e-7f;
    const float gz = fmaf(fabsf(az), 256.0f, fabsf(bz)) * 4.76837158e-7f;
    const float ex = fminf(bx, fmaf(255.0f, ax, bx)) - gx;
    const float ey = fminf(by, fmaf(255.0f, ay, by)) - gy;
    const float ez = fminf(bz, fmaf(255.0f, az, bz)) - gz;
    const float t_ref = fmaxf(fmaxf(ex, ey), fmaxf(ez, tmin));
    const float gr = fabsf(t_ref) * 4.76837158e-7f;
    const float px = fmaf(fabsf(ax), RT_BOX_PAD_CELLS, gx + gr);
    const float py = fmaf(fabsf(ay), RT_BOX_PAD_CELLS, gy + gr);
    const float pz = fmaf(fabsf(az), RT_BOX_PAD_CELLS, gz + gr);
    const float cx = bx - t_ref, cy = by - t_ref, cz = bz - t_ref;
    const uint32_t Ax = bf2_splat(ax), Ay = bf2_splat(ay), Az = bf2_splat(az);
    const uint32_t Nx = bf2_splat(cx - px), Ny = bf2_splat(cy - py), Nz = bf2_splat(cz - pz);
    const uint32_t Fx = bf2_splat(cx + px), Fy = bf2_splat(cy + py), Fz = bf2_splat(cz + pz);
    // closest-hit rays cull children beyond the committed hit (rounded up); first-hit rays keep tmax for the triangles only
    uint32_t TL = 0;
    if (!FIRST_HIT) {
        const float tl = tlimit - t_ref;
        TL = bf2_splat(fmaf(fabsf(tl), 0.0078125f, tl) + gr);
    }
    // two children per word: .x = slots 0,1  .y = 2,3  .z = 4,5  .w = 6,7 (even slot in the low half)
#define RT_BOX2(W)                                                                                         \
    ({                                                                                                     \
        uint32_t tn = bf2_max(bf2_max(bf2_fma(nx.W, Ax, Nx), bf2_fma(ny.W, Ay, Ny)), bf2_max(bf2_fma(nz.W, Az, Nz), 0u)); \
        uint32_t tf = bf2_min(bf2_min(bf2_fma(fx.W, Ax, Fx), bf2_fma(fy.W, Ay, Fy)), bf2_fma(fz.W, Az, Fz));               \
        if (!FIRST_HIT) tf = bf2_min(tf, TL);                                                              \
        bf2_sub(tf, tn);                                                                                   \
    })
    const uint32_t d0 = RT_BOX2(x), d1 = RT_BOX2(y), d2 = RT_BOX2(z), d3 = RT_BOX2(w);
#undef RT_BOX2
    // sign bytes of the eight differences (bytes 1 and 3 of each word) -> one bit per slot; sign set = miss
    const uint32_t s03 = __byte_perm(d0, d1, 0x7531), s47 = __byte_perm(d2, d3, 0x7531);
    const uint32_t m03 = (((s03 >> 7) & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t m47 = (((s47 >> 7) & 0x01010101u) * 0x01020408u) >> 24;
    return ~(m03 | (m47 << 4)) & 0xFFu;
}

// The traversal stack of one ray.  RT_SMEM_STACK = N > 0 (A/B): the first N entries live in shared memory (entry e of thread t at
// [e * 128 + t]: consecutive lanes, consecutive 8-byte words, no bank conflicts), deeper entries in local memory; 0: all of
// it in local memory.  The top of the stack proper — the node group being worked on — is always in registers (ng_base, ng_bits).
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 0
#endif
#define RT_SMEM_ENTRIES (RT_SMEM_STACK < RT_STACK_SIZE ? RT_SMEM_STACK : RT_STACK_SIZE)
#if RT_SMEM_STACK > 0
__device__ __forceinline__ uint2* block_stack_base() {
    __shared__ uint2 s_stack[RT_SMEM_ENTRIES * 128];  // every kernel of the path runs 128-thread blocks
    return s_stack;
}
#endif
struct RayStack {
#if RT_SMEM_STACK > 0
    uint2* sm;
    uint2 lm[RT_STACK_SIZE > RT_SMEM_ENTRIES ? RT_STACK_SIZE - RT_SMEM_ENTRIES : 1];
    __device__ __forceinline__ void init() { sm = block_stack_base() + threadIdx.x; }
    __device__ __forceinline__ void store(int i, uint2 v) {
        if (i < RT_SMEM_ENTRIES) sm[i * 128] = v;
        else lm[i - RT_SMEM_ENTRIES] = v;
    }
    __device__ __forceinline__ uint2 load(int i) const { return i < RT_SMEM_ENTRIES ? sm[i * 128] : lm[i - RT_SMEM_ENTRIES]; }
#else
    uint2 lm[RT_STACK_SIZE];
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void store(int i, uint2 v) { lm[i] = v; }
    __device__ __forceinline__ uint2 load(int i) const { return lm[i]; }
#endif
};

// Shadow rays stop at the first accepted hit, whatever its distance: their children need no front-to-back order, so
// the octant permutation of the pending mask is skipped for them (measured: profiles/r01n_ab.txt).
#define RT_UNORDERED(ANY) (ANY)

template <bool ANY, bool COUNT>
struct Traverser {
    V3 o, d;    // world-space ray
    V3 co, cd;  // ray in the current space (world in the TLAS, object inside an instance)
    float idx, idy, idz;
    float tmin, tmax;
    uint32_t oct;
    uint32_t ng_base, ng_bits;  // current node group: child_base | pending (bits 0..7, priority order) + imask (bits 8..15)
    uint32_t enter_inst;        // TLAS leaf to enter next, or RT_NONE
    int sp, inst_sp;            // inst_sp >= 0: inside an instance entered at this stack height
    uint32_t cur_inst_pos, cur_instance_id, cur_custom_sbt;
    Hit hit;

    __device__ __forceinline__ void set_space(V3 no, V3 nd) {
        co = no; cd = nd;
        idx = safe_rcp(cd.x); idy = safe_rcp(cd.y); idz = safe_rcp(cd.z);
        oct = (cd.x < 0.0f ? 1u : 0u) | (cd.y < 0.0f ? 2u : 0u) | (cd.z < 0.0f ? 4u : 0u);
    }

    __device__ __forceinline__ void begin(V3 ro, V3 rd, float t_min, float t_max) {
        o = ro; d = rd; tmin = t_min; tmax = t_max;
        set_space(ro, rd);
        hit.t = t_max; hit.u = hit.v = 0.0f;
        hit.inst_pos = RT_NONE; hit.instance_id = RT_NONE; hit.geom = hit.prim = RT_NONE; hit.custom_sbt = 0;
        sp = 0; inst_sp = -1;
        cur_inst_pos = cur_instance_id = cur_custom_sbt = 0;
        enter_inst = RT_NONE;
        ng_base = 0;                       // TLAS root = slot 0 of a virtual parent
        ng_bits = (RT_UNORDERED(ANY) ? 1u : (1u << oct)) | (1u << 8);
    }

    __device__ __forceinline__ bool found() const { return hit.inst_pos != RT_NONE; }

    // Call right after begin(): false when the ray cannot touch the scene at all (slab test against the exact
    // bounds of the TLAS root, padded for the approximate reciprocals).  Saves the 8-child test of the root
    // for sky rays; only worth its ~20 instructions for rays that start outside the scene (primary rays).
    __device__ __forceinline__ bool touches_scene(const SceneDev& S) const {
        const float4* rb = reinterpret_cast<const float4*>(S.tlas_nodes) + RT_NODE_BOUNDS_FLOAT / 4;  // Node8::lo[3], hi[3], parent, slot
        float4 a = __ldg(rb), b = __ldg(rb + 1);
        float x0 = (a.x - o.x) * idx, x1 = (a.w - o.x) * idx;
        float y0 = (a.y - o.y) * idy, y1 = (b.x - o.y) * idy;
        float z0 = (a.z - o.z) * idz, z1 = (b.y - o.z) * idz;
        float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
        float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
        return !(tn > tf + 4e-6f * (fabsf(tn) + fabsf(tf)) + 1e-30f);  // NaN (empty scene bounds) -> traverse
    }

    __device__ __forceinline__ void push(RayStack& stack, uint32_t x, uint32_t y, TraceCounters& tc) {
        if (sp < RT_STACK_SIZE) stack.store(sp++, make_uint2(x, y));
        else tc.overflow++;
    }

    // Candidate test of TriRec `ti` against the object-space ray (ro, rd) of instance (ipos, iid, isbt):
    // exclusive interval, closest-t + tie rule (closest-hit rays), any-hit on non-opaque geometry.
    // Returns true when the candidate was accepted and committed to `hit`.
    __device__ __forceinline__ bool test_triangle(const SceneDev& S, uint32_t ti, V3 ro, V3 rd, uint32_t ipos, uint32_t iid, uint32_t isbt,
                                                  TraceCounters& tc) {
        const float4* tp = reinterpret_cast<const float4*>(S.tris + ti);
        float4 ta = __ldg(tp), tb = __ldg(tp + 1), tcv = __ldg(tp + 2);
        if (COUNT) tc.tris++;
        float t, u, v;
        if (!tri_candidate(ro, rd, v3(ta.x, ta.y, ta.z), v3(tb.x, tb.y, tb.z), v3(tcv.x, tcv.y, tcv.z), t, u, v)) return false;
        if (!(t > tmin && t < tmax)) return false;
        uint32_t prim = __float_as_uint(ta.w), gf = __float_as_uint(tb.w);
        uint32_t geom = gf & ~RT_TRI_NON_OPAQUE;
        if (!ANY) {
            if (t > hit.t) return false;
            if (t == hit.t && hit.inst_pos != RT_NONE) {
                // exact tie: lowest (instance, geometry, primitive) wins
                bool lower = iid != hit.instance_id ? iid < hit.instance_id : geom != hit.geom ? geom < hit.geom : prim < hit.prim;
                if (!lower) return false;
            }
        }
        if (gf & RT_TRI_NON_OPAQUE) {
            if (COUNT) tc.anyhits++;
            if (!anyhit_accepts(S, isbt & 0xFFFFFFu, geom, prim, u, v)) return false;
        }
        hit.t = t; hit.u = u; hit.v = v;
        hit.inst_pos = ipos; hit.instance_id = iid;
        hit.geom = geom; hit.prim = prim; hit.custom_sbt = isbt;
        return true;
    }

    // One traversal step: pop (when the current group is exhausted), node visit, instance entry — in this order, so
    // that a popped group is visited and a popped or freshly found instance is entered in the same step, and the warp
    // runs as few passes over the three blocks as its longest ray needs.  Returns true when the ray is finished (then
    // found() tells hit or miss).
    __device__ __forceinline__ bool step(const SceneDev& S, RayStack& stack, TraceCounters& tc) {
        if ((ng_bits & 0xFFu) == 0u) {
            // ---- current group exhausted: pop
            if (inst_sp >= 0 && sp == inst_sp) {
                inst_sp = -1;
                set_space(o, d);
            }
            if (sp == 0) return true;
            uint2 e = stack.load(--sp);
            if (e.y & 0x80000000u) enter_inst = e.x;
            else { ng_base = e.x; ng_bits = e.y; }
        }
        if (ng_bits & 0xFFu) {
            // ---- visit the nearest pending child of the current group
            uint32_t i = __ffs(ng_bits & 0xFFu) - 1;
            ng_bits &= ~(1u << i);
            uint32_t slot = RT_UNORDERED(ANY) ? i : (i ^ oct);
            uint32_t child = ng_base + __popc((ng_bits >> 8) & ((1u << slot) - 1u));
            if (ng_bits & 0xFFu) push(stack, ng_base, ng_bits, tc);
            const Node8* nodes = inst_sp >= 0 ? S.blas_nodes : S.tlas_nodes;
            const uint4* np = reinterpret_cast<const uint4*>(nodes + child);
            // header, then per axis the near and the far planes of this ray's octant (Node8::q rows 2k and 2k + 1, swapped
            // for a negative direction): the choice is made in the address, not with selects afterwards
            const uint32_t sx = oct & 1u, sy = (oct >> 1) & 1u, sz = oct >> 2;
            const uint4 n0 = __ldg(np), n1 = __ldg(np + 1);
            const uint4 nx = __ldg(np + 2 + sx), fx = __ldg(np + 3 - sx);
            const uint4 ny = __ldg(np + 4 + sy), fy = __ldg(np + 5 - sy);
            const uint4 nz = __ldg(np + 6 + sz), fz = __ldg(np + 7 - sz);
            if (COUNT) tc.nodes++;

            uint32_t h = node_hit_mask<ANY>(n0, nx, fx, ny, fy, nz, fz, co, idx, idy, idz, tmin, hit.t);
            // Empty slots need no mask: their boxes are stored inverted (q lo = 255, q hi = 0), 255 cells against a padding of
            // 1.6 — they fail the test on their own.  Only non-finite arithmetic (a ray that misses the node altogether, at the
            // edge of the number range) could report one: an empty slot is never an internal child (imask bit clear), its meta
            // byte is 0 = no triangles in a BLAS, and the TLAS leaf loop below skips meta 0.
            // First-hit rays leave it at that (-4 % on k_shadow, profiles/r03hij_empty_slot_mask_ab.txt); closest-hit rays keep the explicit
            // mask from the meta bytes: without it k_trace0 compiles to a 10 % slower kernel at identical counters.
            uint32_t imask = n0.w >> 24;
            if (!ANY) h &= nonzero_bytes(n1.z, n1.w);
            uint32_t hl = h & ~imask;
            ng_base = n1.x;
            ng_bits = (RT_UNORDERED(ANY) ? (h & imask) : permute_by_octant(h & imask, oct)) | (imask << 8);

            // ---- leaves of this node
            uint64_t meta = ((uint64_t)n1.w << 32) | n1.z;
            if (inst_sp >= 0) {
                while (hl) {
                    uint32_t s = __ffs(hl) - 1;
                    hl &= hl - 1;
                    uint32_t m = (uint32_t)(meta >> (8 * s)) & 0xFFu;
                    uint32_t first = n1.y + (m & 31u), cnt = m >> 5;
                    for (uint32_t k = 0; k < cnt; k++)
                        if (test_triangle(S, first + k, co, cd, cur_inst_pos, cur_instance_id, cur_custom_sbt, tc) && ANY) return true;
                }
            } else if (hl) {
                // TLAS leaves: enter the first now, stack the others (meta 0: an empty slot reported by non-finite arithmetic)
                uint32_t s = __ffs(hl) - 1;
                hl &= hl - 1;
                uint32_t m = (uint32_t)(meta >> (8 * s)) & 0xFFu;
                enter_inst = m ? n1.y + (m & 31u) : RT_NONE;
                while (hl) {
                    s = __ffs(hl) - 1;
                    hl &= hl - 1;
                    m = (uint32_t)(meta >> (8 * s)) & 0xFFu;
                    if (m) push(stack, n1.y + (m & 31u), 0x80000000u, tc);
                }
            }
        }

        if (enter_inst != RT_NONE) {
            // ---- enter an instance: world ray -> object ray
            const float4* ip = reinterpret_cast<const float4*>(S.inst_rt + enter_inst);
            float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
            uint32_t pos = enter_inst;
            enter_inst = RT_NONE;
            uint32_t root = __float_as_uint(r3.x);
            const uint32_t mw = __float_as_uint(r3.w);
            if (root != RT_NONE && (mw & 0xFFu)) {
                if (COUNT) tc.instances++;
                float inv[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
                const uint32_t tiny = (mw >> 8) & 0xFFu;
                if (tiny) {
                    // tiny BLAS: its triangles are tested right here, traversal stays in the TLAS.  First the exact
                    // object-space bounds (the record after the triangles): the TLAS slot box is quantised to the
                    // root's grid, far too coarse to reject e.g. a shadow ray leaving the ground plane it starts on.
                    V3 oo = xform_point(inv, o), od = xform_vec(inv, d);
                    const float4* bp = reinterpret_cast<const float4*>(S.tris + root + tiny);
                    float4 blo = __ldg(bp), bhi = __ldg(bp + 1);
                    float ix = safe_rcp(od.x), iy = safe_rcp(od.y), iz = safe_rcp(od.z);
                    float x0 = (blo.x - oo.x) * ix, x1 = (bhi.x - oo.x) * ix;
                    float y0 = (blo.y - oo.y) * iy, y1 = (bhi.y - oo.y) * iy;
                    float z0 = (blo.z - oo.z) * iz, z1 = (bhi.z - oo.z) * iz;
                    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
                    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), hit.t));
                    if (!(tn > tf + 4e-6f * (fabsf(tn) + fabsf(tf)) + 1e-30f)) {
                        for (uint32_t k = 0; k < tiny; k++)
                            if (test_triangle(S, root + k, oo, od, pos, __float_as_uint(r3.y), __float_as_uint(r3.z), tc) && ANY) return true;
                    }
                } else {
                    if (ng_bits & 0xFFu) push(stack, ng_base, ng_bits, tc);
                    set_space(xform_point(inv, o), xform_vec(inv, d));
                    cur_inst_pos = pos;
                    cur_instance_id = __float_as_uint(r3.y);
                    cur_custom_sbt = __float_as_uint(r3.z);
                    ng_base = root;
                    ng_bits = (RT_UNORDERED(ANY) ? 1u : (1u << oct)) | (1u << 8);
                    inst_sp = sp;
                }
            }
        }

        // nothing pending and nothing stacked: finished (saves the pass that would only find the stack empty)
        return (ng_bits & 0xFFu) == 0u && sp == 0;
    }
};

#ifdef RT_LANE_HIST
__device__ unsigned long long g_lane_hist[2][33];
#endif

// Run one ray to completion (megakernel path, tests).
template <bool ANY, bool COUNT, bool OUTSIDE_START = false>
__device__ __forceinline__ bool trace_ray(const SceneDev& S, V3 o, V3 d, float tmin, float tmax, Hit& hit, TraceCounters& tc) {
    RayStack stack;
    stack.init();
    Traverser<ANY, COUNT> T;
    T.begin(o, d, tmin, tmax);
    if (OUTSIDE_START && !T.touches_scene(S)) { hit = T.hit; return false; }
#ifdef RT_LANE_HIST
    // scratch instrumentation (tools/gpu_lane_hist.py): passes of the batch by the number of lanes still traversing
    while (!T.step(S, stack, tc)) {
        const uint32_t m = __activemask();
        if ((threadIdx.x & 31u) == (uint32_t)(__ffs(m) - 1)) atomicAdd(&g_lane_hist[ANY ? 1 : 0][__popc(m)], 1ull);
    }
#else
    while (!T.step(S, stack, tc)) {}
#endif
    hit = T.hit;
    return T.found();
}

}  // namespace b200rt
