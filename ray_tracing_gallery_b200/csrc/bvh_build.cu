// bvh_build.cu — binary tree (binned SAH or Morton radix tree) -> SAH-guided collapse -> compressed 8-wide BVH, entirely on the GPU.
//
//   1. centroid bounds            k_centroid_bounds    (warp shuffles + ordered-int atomics)
//   2. 63-bit Morton keys         k_morton
//   3. radix sort                 cub::DeviceRadixSort::SortPairs (64-bit key, 32-bit payload)
//   4. binary radix tree          k_hierarchy          (Karras 2012, index tie-break for duplicates)
//   4b. INSTEAD of 1-4: top-down binned SAH in one cooperative launch, k_sah_tree — level by level, one block per node (the
//      whole grid for large nodes), 32 bins per axis; 1-4 remain as the path without cooperative launches
//   5. bottom-up AABB fit         k_fit                (one atomic flag per internal node)
//      + in the same pass: SAH-optimal collapse table per binary node (dynamic programme of
//        Ylitie et al. 2017, "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide
//        BVHs", sec. 3.1): cost[n][i] = cheapest way to hand subtree n to a parent that can spare
//        i slots, i = 1..7, with the decisions that achieve it
//   6. collapse to 8-wide nodes   k_collapse_coop      (breadth-first, one cooperative launch:
//                                                      task index == wide-node index, so the BFS
//                                                      queue IS the node array; grid.sync per level)
//      - children of a wide node: the 8-slot cut of the binary tree the table says is cheapest
//      - slots assigned by child-centre octant => front-to-back order is (slot XOR ray octant);
//        for BLASes the assignment is refined by pairwise exchanges (RT_SLOT_OPT)
//      - boxes quantised to 8 bits on an exactly representable power-of-two grid
//   7. refit (TLAS update)        k_refit              (bottom-up over wide nodes)
//
// What the Vulkan driver does for the reference at src/util_structs.rs:269-274 / :345-354.
#include <cooperative_groups.h>
#include <math_constants.h>

#include <cub/device/device_radix_sort.cuh>

#include "bvh_build.h"
#include "launch_count.h"

namespace cg = cooperative_groups;

namespace b200rt {
namespace {

// Slot assignment refined by pairwise exchanges (collapse_task): 0 = off, 1 = BLASes only (shipped: C2's closest-hit rays -14 %, the
// TLAS of C5 loses 2 % with it), 2 = every tree.  profiles/r04ef_slot_assignment_ab.txt
#ifndef RT_SLOT_OPT
#define RT_SLOT_OPT 1
#endif

enum { ST_LEVEL_BEGIN = 0, ST_LEVEL_END = 1, ST_WIDE_COUNT = 2, ST_PRIM_CURSOR = 3 };

__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__device__ __forceinline__ bool box_valid(const Aabb& b) {
    return b.lo[0] <= b.hi[0] && b.lo[1] <= b.hi[1] && b.lo[2] <= b.hi[2] && isfinite(b.lo[0]) && isfinite(b.lo[1]) &&
           isfinite(b.lo[2]) && isfinite(b.hi[0]) && isfinite(b.hi[1]) && isfinite(b.hi[2]);
}
__device__ __forceinline__ Aabb box_empty() {
    Aabb b;
    b.lo[0] = b.lo[1] = b.lo[2] = CUDART_INF_F;
    b.hi[0] = b.hi[1] = b.hi[2] = -CUDART_INF_F;
    return b;
}
__device__ __forceinline__ void box_grow(Aabb& a, const Aabb& b) {
    if (!box_valid(b)) return;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a.lo[k] = fminf(a.lo[k], b.lo[k]);
        a.hi[k] = fmaxf(a.hi[k], b.hi[k]);
    }
}
__device__ __forceinline__ float box_area(const Aabb& b) {
    if (!box_valid(b)) return 0.0f;
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ Aabb ldcg_box(const Aabb* p) {
    Aabb b;
    const float* f = reinterpret_cast<const float*>(p);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        b.lo[k] = __ldcg(f + k);
        b.hi[k] = __ldcg(f + 3 + k);
    }
    return b;
}

// ------------------------------------------------------------------------------------ 1, 2
__global__ void k_init_state(int* bounds, uint32_t* state, int* task_node, uint32_t* task_parent, int root_ref) {
    int t = threadIdx.x;
    if (t == 0) {
        task_node[0] = root_ref;  // binary root (0) or the only leaf (~0)
        task_parent[0] = 0;
    }
    if (t < 3) bounds[t] = f2ord(CUDART_INF_F);
    else if (t < 6) bounds[t] = f2ord(-CUDART_INF_F);
    if (t == 0) {
        for (int i = 0; i < 8; i++) state[i] = 0;
        state[ST_LEVEL_BEGIN] = 0;
        state[ST_LEVEL_END] = 1;
        state[ST_WIDE_COUNT] = 1;
        state[ST_PRIM_CURSOR] = 0;
    }
}

__global__ void k_centroid_bounds(const Aabb* __restrict__ boxes, uint32_t n, int* bounds) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
    float hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    if (i < n) {
        Aabb b = boxes[i];
        if (box_valid(b)) {
#pragma unroll
            for (int k = 0; k < 3; k++) lo[k] = hi[k] = 0.5f * b.lo[k] + 0.5f * b.hi[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(&bounds[k], f2ord(lo[k]));
            atomicMax(&bounds[3 + k], f2ord(hi[k]));
        }
    }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x) {
    x &= 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// `shift` drops low key bits so that the radix sort needs fewer passes (key_bits_for)
__global__ void k_morton(const Aabb* __restrict__ boxes, uint32_t n, const int* __restrict__ bounds, uint32_t shift, uint64_t* keys,
                         uint32_t* vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Aabb b = boxes[i];
    uint64_t key = 0x7FFFFFFFFFFFFFFFull >> shift;  // invalid boxes sort last
    if (box_valid(b)) {
        uint64_t q[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float lo = ord2f(bounds[k]), hi = ord2f(bounds[3 + k]);
            float c = 0.5f * b.lo[k] + 0.5f * b.hi[k];
            float ext = hi - lo;
            float f = ext > 0.0f ? (c - lo) / ext : 0.0f;
            f = fminf(fmaxf(f, 0.0f), 1.0f);
            q[k] = (uint64_t)(f * 2097151.0f);
        }
        key = ((spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2])) >> shift;
    }
    keys[i] = key;
    vals[i] = i;
}

// ------------------------------------------------------------------------------------ 4
struct Tree2 {
    const uint64_t* keys;   // sorted
    const uint32_t* vals;   // sorted position -> input primitive
    int* left;              // internal node -> child ref (>=0 internal, <0: ~leaf position)
    int* right;
    int* parent_int;        // parent of an internal node (-1 for root)
    int* parent_leaf;       // parent of a leaf position
    uint32_t* first;        // primitive range of an internal node (sorted positions)
    uint32_t* last;
    Aabb* ibox;             // internal node boxes
    uint32_t* flags;
    const Aabb* boxes;      // input boxes (by input primitive)
    int n;
    // SAH collapse table, per internal node
    float* dp_cost;         // [n][7]: cost with i+1 slots
    uint8_t* dp_dec;        // [n][8]: [0] 0 = leaf / 1 = internal node when given ONE slot; [i] (i = 1..6) left child's share of i+1
                            //         slots, 0 = no better than i slots; [7] left child's share of this node's own 8 slots
    uint32_t max_leaf;
    float c_node, c_prim;   // cost of visiting a wide node / of processing one leaf primitive, per unit of box area
};

__device__ __forceinline__ int delta(const uint64_t* __restrict__ k, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = k[i], b = k[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_hierarchy(Tree2 T) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = T.n;
    if (i >= n - 1) return;
    const uint64_t* k = T.keys;
    int d = (delta(k, n, i, i + 1) - delta(k, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(k, n, i, i - d);
    int lmax = 2;
    while (delta(k, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(k, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(k, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(k, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lref, rref;
    if (lo == gamma) { lref = ~gamma; T.parent_leaf[gamma] = i; } else { lref = gamma; T.parent_int[gamma] = i; }
    if (hi == gamma + 1) { rref = ~(gamma + 1); T.parent_leaf[gamma + 1] = i; } else { rref = gamma + 1; T.parent_int[gamma + 1] = i; }
    T.left[i] = lref;
    T.right[i] = rref;
    T.first[i] = (uint32_t)lo;
    T.last[i] = (uint32_t)hi;
    if (i == 0) T.parent_int[0] = -1;
}


// ------------------------------------------------------------------------------------ 4b: top-down binned SAH
// The binary tree under the wide nodes (the reference asks its driver for PREFER_FAST_TRACE, src/util_structs.rs:236): grown top-down,
// every split the cheapest of 3 x 31 binned surface-area-heuristic candidates (32 bins per axis over the centroid bounds; Wald 2007).
// Per node: centroid bounds, binning with shared-memory atomics, plane search by warp scans, stable partition of the node's range of
// `order` (through `tmp`), two children.  Ranges stay contiguous, single primitives become leaves, internal nodes are numbered from an
// atomic counter (root = 0): the same Tree2 conventions as the radix tree, so that k_fit and the collapse run unchanged.
// Measured against the radix tree: sum of internal-node areas -22 % (tori) / -31 % (lain); node visits per ray -22 % (C5) / -31 % (C4).
#define SAH_BINS 32  // = the warp width: one warp scans the bins of an axis.  16 -> 32: C5 -2.8 %, C4 -1.9 %; 64: within noise (profiles/r04cd_sah_builder_ab.txt, r04ef)
struct SahArgs {
    const Aabb* boxes;
    uint32_t *order, *tmp;
    uint32_t* counters;  // [0] next internal node id, [2] set when the two passes over a node disagree (cannot happen; the node is then halved)
    int *left, *right, *parent_int, *parent_leaf;
    uint32_t *first, *last;
};

__device__ __forceinline__ float centroid_k(const Aabb& b, int k) { return __fadd_rn(__fmul_rn(0.5f, b.lo[k]), __fmul_rn(0.5f, b.hi[k])); }
__device__ __forceinline__ int sah_bin(float c, float cmn, float scale) {
    int b = (int)__fmul_rn(__fsub_rn(c, cmn), scale);
    return b < 0 ? 0 : b > SAH_BINS - 1 ? SAH_BINS - 1 : b;
}

// Shared memory of a block that splits a node: centroid bounds, the bins of the three axes, the outcome of the plane search.
struct SahBins {
    int cb[6];                                       // centroid bounds (ordered ints): lo xyz, hi xyz
    int lo[3][SAH_BINS][3], hi[3][SAH_BINS][3];      // [axis][bin][xyz], ordered ints
    uint32_t cnt[3][SAH_BINS];
    float cost[3];
    int best_split[3];
    int axis, split;                                 // axis < 0: no usable plane
};

__device__ __forceinline__ void sah_bins_clear(SahBins& B, uint32_t tid, uint32_t nthreads) {
    if (tid < 3) B.cb[tid] = f2ord(CUDART_INF_F);
    else if (tid < 6) B.cb[tid] = f2ord(-CUDART_INF_F);
    for (uint32_t i = tid; i < 3 * SAH_BINS * 3; i += nthreads) { (&B.lo[0][0][0])[i] = f2ord(CUDART_INF_F); (&B.hi[0][0][0])[i] = f2ord(-CUDART_INF_F); }
    for (uint32_t i = tid; i < 3 * SAH_BINS; i += nthreads) (&B.cnt[0][0])[i] = 0;
}

// bin = (centroid - cmn) * scale, from the centroid bounds of the node (every pass over a node's primitives derives it the same way)
__device__ __forceinline__ void sah_bin_scale(const int* cb, float* cmn, float* scale) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        cmn[k] = ord2f(cb[k]);
        const float ext = ord2f(cb[3 + k]) - cmn[k];
        scale[k] = (ext > 0.0f && isfinite(ext)) ? (float)SAH_BINS * 0.999999f / ext : 0.0f;  // 0: every primitive in bin 0, no split on this axis
        if (!isfinite(scale[k])) scale[k] = 0.0f;
        if (!isfinite(cmn[k])) cmn[k] = 0.0f;  // (no valid primitive at all)
    }
}

// The 3 x 31 candidate planes over filled bins (all threads of a block of >= 3 warps; B.axis / B.split on return): warp k scans the
// 32 bins of axis k (lane = bin) — an inclusive prefix of boxes and counts from the left, one from the right (shuffles) — lane sp
// prices the plane after bin sp, the warp keeps its cheapest (lowest sp wins ties), thread 0 the cheapest axis (lowest wins ties).
__device__ __forceinline__ void sah_pick_plane(SahBins& B) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (warp < 3) {
        const int k = (int)warp;
        float llo[3], lhi[3], rlo[3], rhi[3];
#pragma unroll
        for (int j = 0; j < 3; j++) { llo[j] = rlo[j] = ord2f(B.lo[k][lane][j]); lhi[j] = rhi[j] = ord2f(B.hi[k][lane][j]); }
        uint32_t nl = B.cnt[k][lane], nr = nl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const float a0 = __shfl_up_sync(0xFFFFFFFFu, llo[j], o), a1 = __shfl_up_sync(0xFFFFFFFFu, lhi[j], o);
                const float b0 = __shfl_down_sync(0xFFFFFFFFu, rlo[j], o), b1 = __shfl_down_sync(0xFFFFFFFFu, rhi[j], o);
                if ((int)lane >= o) { llo[j] = fminf(llo[j], a0); lhi[j] = fmaxf(lhi[j], a1); }
                if ((int)lane + o < 32) { rlo[j] = fminf(rlo[j], b0); rhi[j] = fmaxf(rhi[j], b1); }
            }
            const uint32_t cl = __shfl_up_sync(0xFFFFFFFFu, nl, o), cr = __shfl_down_sync(0xFFFFFFFFu, nr, o);
            if ((int)lane >= o) nl += cl;
            if ((int)lane + o < 32) nr += cr;
        }
        // (bins without primitives keep +inf / -inf planes: an empty side has area 0 and count 0)
        auto area = [](const float* lo, const float* hi) -> float {
            const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
            return (dx >= 0.0f && dy >= 0.0f && dz >= 0.0f) ? dx * dy + dy * dz + dz * dx : 0.0f;
        };
        const float right_cost = area(rlo, rhi) * (float)nr;   // bins lane .. 31
        const float rc = __shfl_down_sync(0xFFFFFFFFu, right_cost, 1);
        const uint32_t rn = __shfl_down_sync(0xFFFFFFFFu, nr, 1);
        float cost = (lane == 31u || nl == 0u || rn == 0u) ? CUDART_INF_F : area(llo, lhi) * (float)nl + rc;
        if (!(cost == cost)) cost = CUDART_INF_F;
        uint32_t sp = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float oc = __shfl_xor_sync(0xFFFFFFFFu, cost, o);
            const uint32_t os = __shfl_xor_sync(0xFFFFFFFFu, sp, o);
            if (oc < cost || (oc == cost && os < sp)) { cost = oc; sp = os; }
        }
        if (lane == 0) { B.cost[k] = cost; B.best_split[k] = (int)sp; }
    }
    __syncthreads();
    if (tid == 0) {
        int bk = -1;
        float best = CUDART_INF_F;
        for (int k = 0; k < 3; k++)
            if (B.cost[k] < best) { best = B.cost[k]; bk = k; }
        B.axis = bk;
        B.split = bk < 0 ? 0 : B.best_split[bk];
    }
    __syncthreads();
}

// Where the nodes of the next level are queued: nodes of at least big_min primitives go to a queue of their own (k_sah_tree splits
// them with the whole grid); big_min = 0xFFFFFFFF: one queue.
struct SahRoute {
    uint4* q_small; uint32_t* len_small;
    uint4* q_big;   uint32_t* len_big;
    uint32_t big_min;
};

// The two children of `node` = order[f .. f + c), cut after n_left primitives (one thread)
__device__ __forceinline__ void sah_emit_children(const SahArgs& A, uint32_t node, uint32_t f, uint32_t c, uint32_t n_left, const SahRoute& R) {
    A.first[node] = f;
    A.last[node] = f + c - 1u;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const uint32_t cf = side ? f + n_left : f, cc = side ? c - n_left : n_left;
        int ref;
        if (cc == 1u) { ref = ~(int)cf; A.parent_leaf[cf] = (int)node; }
        else {
            const uint32_t id = atomicAdd(&A.counters[0], 1u);
            ref = (int)id;
            A.parent_int[id] = (int)node;
            if (cc >= R.big_min) R.q_big[atomicAdd(R.len_big, 1u)] = make_uint4(id, cf, cc, 0u);
            else R.q_small[atomicAdd(R.len_small, 1u)] = make_uint4(id, cf, cc, 0u);
        }
        if (side) A.right[node] = ref; else A.left[node] = ref;
    }
}

// One node of the level: called by all TB threads of a block.  R: where nodes of the next level are queued.
// halve: no SAH, the range is cut in the middle (depth limit of the single-launch build).
template <int TB>
__device__ __forceinline__ void sah_split_node(const SahArgs& A, const uint4 item, const SahRoute& R, const bool halve) {
    const uint32_t node = item.x, f = item.y, c = item.z;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    __shared__ SahBins B;
    __shared__ uint32_t s_wl[TB / 32];
    if (c == 2u) {  // two primitives: two leaves (a third of all nodes)
        if (tid == 0) {
            A.first[node] = f; A.last[node] = f + 1u;
            A.left[node] = ~(int)f; A.right[node] = ~(int)(f + 1u);
            A.parent_leaf[f] = (int)node; A.parent_leaf[f + 1u] = (int)node;
        }
        return;
    }

    int axis = -1, split = 0;
    float cmn[3] = {0.0f, 0.0f, 0.0f}, scale[3] = {0.0f, 0.0f, 0.0f};
    if (!halve) {
    sah_bins_clear(B, tid, TB);
    __syncthreads();
    // ---- centroid bounds of the node's valid primitives
    {
        float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        for (uint32_t i = tid; i < c; i += TB) {
            const Aabb b = A.boxes[A.order[f + i]];
            if (box_valid(b)) {
#pragma unroll
                for (int k = 0; k < 3; k++) { float ck = centroid_k(b, k); lo[k] = fminf(lo[k], ck); hi[k] = fmaxf(hi[k], ck); }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
            }
            if (lane == 0) { atomicMin(&B.cb[k], f2ord(lo[k])); atomicMax(&B.cb[3 + k], f2ord(hi[k])); }
        }
    }
    __syncthreads();
    sah_bin_scale(B.cb, cmn, scale);
    // ---- binning (primitives without a valid box count in bin 0 and contribute no area)
    for (uint32_t i = tid; i < c; i += TB) {
        const Aabb b = A.boxes[A.order[f + i]];
        const bool ok = box_valid(b);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int bin = ok ? sah_bin(centroid_k(b, k), cmn[k], scale[k]) : 0;
            atomicAdd(&B.cnt[k][bin], 1u);
            if (ok) {
#pragma unroll
                for (int j = 0; j < 3; j++) { atomicMin(&B.lo[k][bin][j], f2ord(b.lo[j])); atomicMax(&B.hi[k][bin][j], f2ord(b.hi[j])); }
            }
        }
    }
    __syncthreads();
    sah_pick_plane(B);
    axis = B.axis; split = B.split;
    }
    uint32_t n_left = c / 2;  // no usable plane (coincident centroids): halve the range as it stands
    if (axis >= 0) {
        // ---- stable partition of order[f .. f + c) through tmp
        uint32_t total_left = 0;
#pragma unroll
        for (int bI = 0; bI < SAH_BINS; bI++) total_left += bI <= split ? B.cnt[axis][bI] : 0u;
        n_left = total_left;
        uint32_t done_l = 0, done_r = 0;
        for (uint32_t base = 0; base < c; base += TB) {
            const uint32_t i = base + tid;
            bool in = i < c, goes_left = false;
            uint32_t prim = 0;
            if (in) {
                prim = A.order[f + i];
                const Aabb b = A.boxes[prim];
                const int bin = box_valid(b) ? sah_bin(centroid_k(b, axis), cmn[axis], scale[axis]) : 0;
                goes_left = bin <= split;
            }
            const uint32_t ml = __ballot_sync(0xFFFFFFFFu, in && goes_left), ma = __ballot_sync(0xFFFFFFFFu, in);
            __syncthreads();  // (s_wl of the previous chunk has been read)
            if (lane == 0) s_wl[warp] = (uint32_t)__popc(ml) | ((uint32_t)__popc(ma) << 16);
            __syncthreads();
            uint32_t before_l = 0, before_a = 0, all_l = 0, all_a = 0;
#pragma unroll
            for (uint32_t w = 0; w < TB / 32; w++) {
                const uint32_t v = s_wl[w];
                if (w < warp) { before_l += v & 0xFFFFu; before_a += v >> 16; }
                all_l += v & 0xFFFFu; all_a += v >> 16;
            }
            if (in) {
                const uint32_t below = (1u << lane) - 1u;
                const uint32_t rank_l = before_l + (uint32_t)__popc(ml & below);
                const uint32_t rank_r = (before_a - before_l) + (uint32_t)__popc((ma & ~ml) & below);
                A.tmp[f + (goes_left ? done_l + rank_l : n_left + done_r + rank_r)] = prim;
            }
            done_l += all_l; done_r += all_a - all_l;
        }
        __syncthreads();
        if (done_l != n_left) {  // the two passes disagree: cannot happen with the explicitly rounded bin arithmetic
            if (tid == 0) atomicExch(&A.counters[2], 1u);
            n_left = c / 2;
        } else {
            for (uint32_t i = tid; i < c; i += TB) A.order[f + i] = A.tmp[f + i];
        }
    }
    if (tid == 0) sah_emit_children(A, node, f, c, n_left, R);
}

// Record of a node that is split by the whole grid (global memory): what SahBins is to a block, plus the partition cursors
struct SahBig {
    int cb[6];
    int lo[3][SAH_BINS][3], hi[3][SAH_BINS][3];
    uint32_t cnt[3][SAH_BINS];
    uint32_t cur[2];   // primitives written to the left / right part so far
    int axis, split;
    uint32_t n_left, _pad;
};

// ---- The whole tree in one cooperative launch (k_sah_tree): nothing for the host to follow, so static builds and per-frame rebuilds
//      alike are stream-ordered.  Level by level; the blocks share the nodes of a level (two alternating queues with alternating length
//      words, grid.sync() between levels); below level 48 ranges are halved, which bounds the depth at 48 + log2(n).  Nodes of at
//      least SAH_TREE_BIG_MIN primitives are queued apart and split by the whole grid, all large nodes of a level together, phase
//      by phase with grid.sync() between the phases — records cleared / centroid bounds / bins (per chunk in shared memory, flushed with
//      global atomics) / plane search + children / partition (blocks claim output ranges from the node's two cursors) / copy back —
//      each over the flattened list of (node, 4 096-primitive chunk) pairs, so the grid is busy however the primitives are spread.
#define SAH_TB 128
#define SAH_BLOCKS_PER_SM 4
#define SAH_CHUNK 4096u
#ifndef SAH_TREE_BIG_MIN
#define SAH_TREE_BIG_MIN 8192u
#endif
#define SAH_BIG_WORDS 704u   // one SahBig record, padded
static_assert(sizeof(SahBig) <= SAH_BIG_WORDS * sizeof(uint32_t), "SahBig record size");

struct SahTreeArgs {
    SahArgs A;
    uint4* qs[2];        // queues of small nodes: level L reads qs[(L & 1) ^ 1], fills qs[L & 1]
    uint4* qb[2];        // queues of large nodes, likewise
    uint32_t* words;     // [p] / [2 + p]: length of qs[p] / qb[p]
    uint32_t* big;       // SAH_BIG_WORDS words per large node of a level
};

__global__ void k_sah_tree_init(SahTreeArgs P, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) P.A.order[i] = i;
    if (i == 0) {
        P.A.counters[0] = 1; P.A.counters[1] = 0; P.A.counters[2] = 0; P.A.counters[3] = 0;
        P.A.parent_int[0] = -1;
        P.words[0] = P.words[1] = P.words[2] = P.words[3] = 0;
        if (n >= SAH_TREE_BIG_MIN) { P.qb[0][0] = make_uint4(0u, 0u, n, 0u); P.words[2] = 1; }
        else { P.qs[0][0] = make_uint4(0u, 0u, n, 0u); P.words[0] = 1; }
    }
}

// chunk `t` of the level's large nodes -> node j, primitives [begin, end) of its range; false: t is past the last chunk
__device__ __forceinline__ bool sah_locate_chunk(const uint4* qb, uint32_t n_big, uint32_t t, uint32_t& j, uint32_t& begin, uint32_t& end) {
    uint32_t first = 0;
    for (j = 0; j < n_big; j++) {
        const uint32_t c = qb[j].z, chunks = (c + SAH_CHUNK - 1u) / SAH_CHUNK;
        if (t < first + chunks) {
            begin = (t - first) * SAH_CHUNK;
            end = begin + SAH_CHUNK < c ? begin + SAH_CHUNK : c;
            return true;
        }
        first += chunks;
    }
    return false;
}

__global__ void __launch_bounds__(SAH_TB) k_sah_tree(SahTreeArgs P) {
    cg::grid_group grid = cg::this_grid();
    const SahArgs& A = P.A;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    __shared__ SahBins B;
    __shared__ uint32_t s_wl[SAH_TB / 32], s_base[2];
    uint32_t n_small = *((volatile uint32_t*)&P.words[0]), n_big = *((volatile uint32_t*)&P.words[2]);
    grid.sync();  // (everybody has read the lengths of level 1 before block 0 clears them below)
    for (uint32_t level = 1; n_small + n_big > 0; level++) {
        const uint32_t w = level & 1u;
        const uint4* qs_in = P.qs[w ^ 1u];
        const uint4* qb_in = P.qb[w ^ 1u];
        const SahRoute R = {P.qs[w], &P.words[w], P.qb[w], &P.words[2u + w], SAH_TREE_BIG_MIN};
        const bool halve = level > 48u;
        if (blockIdx.x == 0 && tid == 0) { P.words[w ^ 1u] = 0; P.words[2u + (w ^ 1u)] = 0; }  // the words level + 1 will fill (their old values, this level's lengths, have been read)
        if (n_big > 0) {
            // ---- 0: records
            for (uint32_t j = blockIdx.x; j < n_big; j += gridDim.x) {
                SahBig* G = reinterpret_cast<SahBig*>(P.big + (size_t)j * SAH_BIG_WORDS);
                if (tid < 3) G->cb[tid] = f2ord(CUDART_INF_F);
                else if (tid < 6) G->cb[tid] = f2ord(-CUDART_INF_F);
                for (uint32_t i = tid; i < 3 * SAH_BINS * 3; i += SAH_TB) { (&G->lo[0][0][0])[i] = f2ord(CUDART_INF_F); (&G->hi[0][0][0])[i] = f2ord(-CUDART_INF_F); }
                for (uint32_t i = tid; i < 3 * SAH_BINS; i += SAH_TB) (&G->cnt[0][0])[i] = 0;
                if (tid == 0) { G->cur[0] = G->cur[1] = 0; G->axis = -1; G->split = 0; G->n_left = 0; }
            }
            grid.sync();
            uint32_t j, begin, end;
            if (!halve) {
                // ---- 1: centroid bounds
                for (uint32_t t = blockIdx.x; sah_locate_chunk(qb_in, n_big, t, j, begin, end); t += gridDim.x) {
                    SahBig* G = reinterpret_cast<SahBig*>(P.big + (size_t)j * SAH_BIG_WORDS);
                    const uint32_t f = qb_in[j].y;
                    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
                    for (uint32_t i = begin + tid; i < end; i += SAH_TB) {
                        const Aabb b = A.boxes[A.order[f + i]];
                        if (box_valid(b)) {
#pragma unroll
                            for (int k = 0; k < 3; k++) { float ck = centroid_k(b, k); lo[k] = fminf(lo[k], ck); hi[k] = fmaxf(hi[k], ck); }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 3; k++) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
                            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
                        }
                        if (lane == 0) { atomicMin(&G->cb[k], f2ord(lo[k])); atomicMax(&G->cb[3 + k], f2ord(hi[k])); }
                    }
                }
                grid.sync();
                // ---- 2: bins (per chunk in shared memory, then into the node's record)
                for (uint32_t t = blockIdx.x; sah_locate_chunk(qb_in, n_big, t, j, begin, end); t += gridDim.x) {
                    SahBig* G = reinterpret_cast<SahBig*>(P.big + (size_t)j * SAH_BIG_WORDS);
                    const uint32_t f = qb_in[j].y;
                    sah_bins_clear(B, tid, SAH_TB);
                    __syncthreads();
                    float cmn[3], scale[3];
                    sah_bin_scale(G->cb, cmn, scale);
                    for (uint32_t i = begin + tid; i < end; i += SAH_TB) {
                        const Aabb b = A.boxes[A.order[f + i]];
                        const bool ok = box_valid(b);
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const int bin = ok ? sah_bin(centroid_k(b, k), cmn[k], scale[k]) : 0;
                            atomicAdd(&B.cnt[k][bin], 1u);
                            if (ok) {
#pragma unroll
                                for (int jj = 0; jj < 3; jj++) { atomicMin(&B.lo[k][bin][jj], f2ord(b.lo[jj])); atomicMax(&B.hi[k][bin][jj], f2ord(b.hi[jj])); }
                            }
                        }
                    }
                    __syncthreads();
                    for (uint32_t i = tid; i < 3 * SAH_BINS; i += SAH_TB) {
                        const uint32_t cnt = (&B.cnt[0][0])[i];
                        if (cnt == 0) continue;
                        atomicAdd(&(&G->cnt[0][0])[i], cnt);
#pragma unroll
                        for (int jj = 0; jj < 3; jj++) {
                            atomicMin(&(&G->lo[0][0][0])[3 * i + jj], (&B.lo[0][0][0])[3 * i + jj]);
                            atomicMax(&(&G->hi[0][0][0])[3 * i + jj], (&B.hi[0][0][0])[3 * i + jj]);
                        }
                    }
                    __syncthreads();  // B is cleared again for the block's next chunk
                }
                grid.sync();
            }
            // ---- 3: plane search and children, one block per node
            for (uint32_t jn = blockIdx.x; jn < n_big; jn += gridDim.x) {
                SahBig* G = reinterpret_cast<SahBig*>(P.big + (size_t)jn * SAH_BIG_WORDS);
                const uint4 item = qb_in[jn];
                if (!halve) {
                    for (uint32_t i = tid; i < 3 * SAH_BINS * 3; i += SAH_TB) { (&B.lo[0][0][0])[i] = (&G->lo[0][0][0])[i]; (&B.hi[0][0][0])[i] = (&G->hi[0][0][0])[i]; }
                    for (uint32_t i = tid; i < 3 * SAH_BINS; i += SAH_TB) (&B.cnt[0][0])[i] = (&G->cnt[0][0])[i];
                    __syncthreads();
                    sah_pick_plane(B);
                }
                if (tid == 0) {
                    uint32_t n_left = item.z / 2;
                    int axis = -1, split = 0;
                    if (!halve && B.axis >= 0) {
                        axis = B.axis; split = B.split;
                        n_left = 0;
                        for (int bI = 0; bI <= split; bI++) n_left += B.cnt[axis][bI];
                    }
                    G->axis = axis; G->split = split; G->n_left = n_left;
                    sah_emit_children(A, item.x, item.y, item.z, n_left, R);
                }
                __syncthreads();
            }
            if (!halve) {
                grid.sync();
                // ---- 4: partition: a block claims the output ranges of its 128 primitives from the node's two cursors
                for (uint32_t t = blockIdx.x; sah_locate_chunk(qb_in, n_big, t, j, begin, end); t += gridDim.x) {
                    SahBig* G = reinterpret_cast<SahBig*>(P.big + (size_t)j * SAH_BIG_WORDS);
                    const int axis = G->axis, split = G->split;
                    if (axis < 0) continue;  // halved as it stands (block-uniform)
                    const uint32_t f = qb_in[j].y, c = qb_in[j].z, n_left = G->n_left;
                    float cmn[3], scale[3];
                    sah_bin_scale(G->cb, cmn, scale);
                    for (uint32_t base = begin; base < end; base += SAH_TB) {
                        const uint32_t i = base + tid;
                        bool in = i < end, goes_left = false;
                        uint32_t prim = 0;
                        if (in) {
                            prim = A.order[f + i];
                            const Aabb b = A.boxes[prim];
                            const int bin = box_valid(b) ? sah_bin(centroid_k(b, axis), cmn[axis], scale[axis]) : 0;
                            goes_left = bin <= split;
                        }
                        const uint32_t ml = __ballot_sync(0xFFFFFFFFu, in && goes_left), ma = __ballot_sync(0xFFFFFFFFu, in);
                        __syncthreads();  // (s_wl / s_base of the previous round have been read)
                        if (lane == 0) s_wl[warp] = (uint32_t)__popc(ml) | ((uint32_t)__popc(ma) << 16);
                        __syncthreads();
                        uint32_t before_l = 0, before_a = 0, all_l = 0, all_a = 0;
#pragma unroll
                        for (uint32_t ww = 0; ww < SAH_TB / 32; ww++) {
                            const uint32_t v = s_wl[ww];
                            if (ww < warp) { before_l += v & 0xFFFFu; before_a += v >> 16; }
                            all_l += v & 0xFFFFu; all_a += v >> 16;
                        }
                        if (tid == 0) { s_base[0] = atomicAdd(&G->cur[0], all_l); s_base[1] = atomicAdd(&G->cur[1], all_a - all_l); }
                        __syncthreads();
                        if (in) {
                            const uint32_t below = (1u << lane) - 1u;
                            const uint32_t rank_l = before_l + (uint32_t)__popc(ml & below);
                            const uint32_t rank_r = (before_a - before_l) + (uint32_t)__popc((ma & ~ml) & below);
                            const uint32_t dst = goes_left ? s_base[0] + rank_l : n_left + s_base[1] + rank_r;
                            if (dst < c) A.tmp[f + dst] = prim;
                        }
                    }
                }
                grid.sync();
                // ---- 5: copy back (a node whose two passes disagree — cannot happen with the explicitly rounded bin arithmetic — keeps its order)
                for (uint32_t t = blockIdx.x; sah_locate_chunk(qb_in, n_big, t, j, begin, end); t += gridDim.x) {
                    const SahBig* G = reinterpret_cast<const SahBig*>(P.big + (size_t)j * SAH_BIG_WORDS);
                    const uint32_t f = qb_in[j].y, c = qb_in[j].z;
                    if (G->axis < 0) continue;
                    if (G->cur[0] != G->n_left || G->cur[0] + G->cur[1] != c) { if (tid == 0) atomicExch(&A.counters[2], 1u); continue; }
                    for (uint32_t i = begin + tid; i < end; i += SAH_TB) A.order[f + i] = A.tmp[f + i];
                }
            }
        }
        // ---- small nodes: one block each
        for (uint32_t q = blockIdx.x; q < n_small; q += gridDim.x) {
            sah_split_node<SAH_TB>(A, qs_in[q], R, halve);
            __syncthreads();  // shared bins are reused by the block's next node
        }
        grid.sync();
        n_small = *((volatile uint32_t*)&P.words[w]);
        n_big = *((volatile uint32_t*)&P.words[2u + w]);
        grid.sync();  // nobody may queue nodes of the next level (or clear this level's lengths) before everybody has read them
    }
}

// ------------------------------------------------------------------------------------ 5
__device__ __forceinline__ Aabb ref_box_cg(const Tree2& T, int ref) {
    if (ref < 0) return T.boxes[T.vals[~ref]];
    return ldcg_box(&T.ibox[ref]);
}

// cost[1..7] of handing child `ref` i slots (a single primitive costs the same whatever it is given)
__device__ __forceinline__ void child_costs(const Tree2& T, int ref, float* c /* [8], index 1..7 */) {
    if (ref < 0) {
        float a = box_area(T.boxes[T.vals[~ref]]) * T.c_prim;
#pragma unroll
        for (int i = 1; i <= 7; i++) c[i] = a;
    } else {
#pragma unroll
        for (int i = 1; i <= 7; i++) c[i] = __ldcg(T.dp_cost + (size_t)ref * 7 + (i - 1));
    }
}

__global__ void k_fit(Tree2 T) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= T.n) return;
    int cur = T.parent_leaf[p];
    while (cur >= 0) {
        __threadfence();
        unsigned old = atomicAdd(&T.flags[cur], 1u);
        if (old == 0) return;
        __threadfence();
        Aabb b = box_empty();
        int L = T.left[cur], R = T.right[cur];
        box_grow(b, ref_box_cg(T, L));
        box_grow(b, ref_box_cg(T, R));
        float* f = reinterpret_cast<float*>(&T.ibox[cur]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            __stcg(f + k, b.lo[k]);
            __stcg(f + 3 + k, b.hi[k]);
        }
        if (T.dp_cost == nullptr) { cur = T.parent_int[cur]; continue; }
        // ---- collapse table
        float cl[8], cr[8], cost[8];
        child_costs(T, L, cl);
        child_costs(T, R, cr);
        float area = box_area(b);
        float best = CUDART_INF_F;
        int bk = 1;
#pragma unroll
        for (int k = 1; k <= 7; k++) {
            float c = cl[k] + cr[8 - k];
            if (c < best) { best = c; bk = k; }
        }
        uint8_t dec[8];
        dec[7] = (uint8_t)bk;
        cost[1] = area * T.c_node + best;
        dec[0] = 1;
        uint32_t count = T.last[cur] - T.first[cur] + 1u;
        if (count <= T.max_leaf) {
            float lc = area * (float)count * T.c_prim;
            if (lc <= cost[1]) { cost[1] = lc; dec[0] = 0; }
        }
#pragma unroll
        for (int i = 2; i <= 7; i++) {
            float bi = cost[i - 1];
            int d = 0;
            for (int k = 1; k < i; k++) {
                float c = cl[k] + cr[i - k];
                if (c < bi) { bi = c; d = k; }
            }
            cost[i] = bi;
            dec[i - 1] = (uint8_t)d;
        }
#pragma unroll
        for (int i = 1; i <= 7; i++) __stcg(T.dp_cost + (size_t)cur * 7 + (i - 1), cost[i]);
        uint2 dw;
        dw.x = dec[0] | (dec[1] << 8) | (dec[2] << 16) | ((uint32_t)dec[3] << 24);
        dw.y = dec[4] | (dec[5] << 8) | (dec[6] << 16) | ((uint32_t)dec[7] << 24);
        __stcg(reinterpret_cast<uint2*>(T.dp_dec + (size_t)cur * 8), dw);
        cur = T.parent_int[cur];
    }
}

// ------------------------------------------------------------------------------------ quantisation
// Exact power-of-two helpers: the builder scales by 2^e with one multiplication (exact for the clamped
// exponents used here) instead of ldexpf/frexpf, which dominated the per-node cost of collapse and refit.
__device__ __forceinline__ float pow2i(int e) { return __int_as_float((e + 127) << 23); }  // -126 <= e <= 127
__device__ __forceinline__ int frexp_exponent(float x) {  // x > 0: x = m * 2^e with m in [0.5, 1)
    int b = (__float_as_int(x) >> 23) & 0xFF;
    return b == 0 ? -126 : b - 126;  // denormals count as 2^-126 (every caller clamps the result to >= -100)
}

// Choose the grid (origin, exponent) of one axis so that lo..hi spans <= 254 cells (one cell of
// slack for the outward rounding fix-up) and every plane origin + q*2^e is exactly representable.
__device__ __forceinline__ void axis_grid(float lo, float hi, float& origin, int& e) {
    float s = (hi - lo) / 253.0f;
    int es = s > 0.0f ? frexp_exponent(s) : -126;  // 2^es > s
    float mx = fmaxf(fabsf(lo), fabsf(hi));
    int em = mx > 0.0f ? frexp_exponent(mx) : -126;  // mx < 2^em
    e = max(max(es, em - 22), -100);
    for (;;) {
        origin = floorf(lo * pow2i(-e)) * pow2i(e);
        float top = ceilf((hi - origin) * pow2i(-e));
        if (top <= 254.0f || e >= 120) break;
        e++;
    }
}

// Fill geometry fields of `nd` (origin, exp, qlo, qhi, lo, hi) from the per-slot child boxes.
__device__ void quantise_node(Node8& nd, const Aabb* cb, uint32_t present) {
    Aabb nb = box_empty();
    for (int s = 0; s < 8; s++)
        if (present >> s & 1) box_grow(nb, cb[s]);
    bool any = box_valid(nb);
    float org[3] = {0.f, 0.f, 0.f};
    int ex[3] = {0, 0, 0};
    if (any) {
        for (int k = 0; k < 3; k++) axis_grid(nb.lo[k], nb.hi[k], org[k], ex[k]);
    }
    float cell[3], inv[3];
    for (int k = 0; k < 3; k++) {
        nd.origin[k] = org[k];
        nd.exp[k] = (uint8_t)(ex[k] + 127);  // -100 <= ex <= 120: the byte is the fp32 exponent field of the cell size
        nd.lo[k] = nb.lo[k];
        nd.hi[k] = nb.hi[k];
        cell[k] = pow2i(ex[k]);
        inv[k] = pow2i(-ex[k]);
    }
    for (int s = 0; s < 8; s++) {
        bool ok = (present >> s & 1) && box_valid(cb[s]);
        for (int k = 0; k < 3; k++) {
            uint8_t ql = 255, qh = 0;
            if (ok) {
                float a = floorf((cb[s].lo[k] - org[k]) * inv[k]);
                a = fminf(fmaxf(a, 0.0f), 255.0f);
                while (a > 0.0f && fmaf(a, cell[k], org[k]) > cb[s].lo[k]) a -= 1.0f;
                float b = ceilf((cb[s].hi[k] - org[k]) * inv[k]);
                b = fminf(fmaxf(b, 0.0f), 255.0f);
                while (b < 255.0f && fmaf(b, cell[k], org[k]) < cb[s].hi[k]) b += 1.0f;
                ql = (uint8_t)a;
                qh = (uint8_t)b;
            }
            // the integers as bf16 bit patterns (exact: q < 256 has at most 8 significant bits)
            nd.q[2 * k][s] = (uint16_t)(__float_as_uint((float)ql) >> 16);
            nd.q[2 * k + 1][s] = (uint16_t)(__float_as_uint((float)qh) >> 16);
        }
    }
}

__device__ __forceinline__ void store_node(Node8* dst, const Node8& nd) {
    const uint4* s = reinterpret_cast<const uint4*>(&nd);
    uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < (int)RT_NODE_QUADS; i++) d[i] = s[i];
}

// ------------------------------------------------------------------------------------ 6
struct CollapseArgs {
    Tree2 T;
    int* task_node;        // wide index -> binary subtree ref
    uint32_t* task_parent; // wide index -> parent wide index | slot << 29
    uint32_t* state;
    Node8* nodes;          // pool base
    uint32_t node_offset, prim_offset, max_leaf;
    uint32_t* leaf_order;
    uint32_t* node_count_out;  // optional: receives the number of wide nodes written
    bool greedy;               // children by greedy largest-area expansion instead of the SAH table
};

__device__ __forceinline__ uint32_t ref_count(const Tree2& T, int r) { return r < 0 ? 1u : T.last[r] - T.first[r] + 1u; }
__device__ __forceinline__ uint32_t ref_first(const Tree2& T, int r) { return r < 0 ? (uint32_t)~r : T.first[r]; }
__device__ __forceinline__ Aabb ref_box(const Tree2& T, int r) { return r < 0 ? T.boxes[T.vals[~r]] : T.ibox[r]; }

__device__ void collapse_task(const CollapseArgs& A, uint32_t w) {
    const Tree2& T = A.T;
    int ch[8];            // children of this wide node: binary subtree refs
    bool internal[8];     //   true: becomes a wide node of its own; false: leaf slot holding the subtree's primitives
    int cnt = 0;
    int r = A.task_node[w];
    if (r < 0) { ch[0] = r; internal[0] = false; cnt = 1; }
    else if (A.greedy) {
        // greedy expansion by surface area: first only subtrees too big to be a leaf, then any subtree
        ch[0] = T.left[r]; ch[1] = T.right[r]; cnt = 2;
        for (int pass = 0; pass < 2; pass++) {
            while (cnt < 8) {
                int best = -1;
                float best_area = -1.0f;
                for (int j = 0; j < cnt; j++) {
                    int c = ch[j];
                    if (c < 0) continue;
                    if (pass == 0 && ref_count(T, c) <= A.max_leaf) continue;
                    float a = box_area(T.ibox[c]);
                    if (a > best_area) { best_area = a; best = j; }
                }
                if (best < 0) break;
                int c = ch[best];
                ch[best] = T.left[c];
                ch[cnt++] = T.right[c];
            }
        }
        for (int j = 0; j < cnt; j++) internal[j] = ch[j] >= 0 && ref_count(T, ch[j]) > A.max_leaf;
    } else {
        // walk the cheapest 8-slot cut recorded by k_fit
        int st_node[8], st_slots[8], sp = 0;
        int k = T.dp_dec[(size_t)r * 8 + 7];
        st_node[sp] = T.right[r]; st_slots[sp++] = 8 - k;
        st_node[sp] = T.left[r]; st_slots[sp++] = k;
        while (sp > 0) {
            int n = st_node[--sp], i = st_slots[sp];
            if (n < 0) { ch[cnt] = n; internal[cnt++] = false; continue; }
            const uint8_t* dec = T.dp_dec + (size_t)n * 8;
            int d = 0;
            while (i >= 2 && (d = dec[i - 1]) == 0) i--;  // i slots are no better than i-1
            if (i == 1) { ch[cnt] = n; internal[cnt++] = dec[0] != 0; continue; }
            st_node[sp] = T.right[n]; st_slots[sp++] = i - d;
            st_node[sp] = T.left[n]; st_slots[sp++] = d;
        }
    }
    Aabb cb[8];
    Aabb nb = box_empty();
    for (int j = 0; j < cnt; j++) {
        cb[j] = ref_box(T, ch[j]);
        box_grow(nb, cb[j]);
    }
    // slot assignment by octant of the child centre
    int child_at[8];
    for (int s = 0; s < 8; s++) child_at[s] = -1;
    float cn[3];
    for (int k = 0; k < 3; k++) cn[k] = 0.5f * nb.lo[k] + 0.5f * nb.hi[k];
#ifdef RT_SLOT_MEAN
    {   // A/B: octants around the mean of the child centres instead of the centre of the node's box
        float sum[3] = {0, 0, 0}; int nvld = 0;
        for (int j = 0; j < cnt; j++) if (box_valid(cb[j])) { nvld++; for (int k = 0; k < 3; k++) sum[k] += 0.5f * cb[j].lo[k] + 0.5f * cb[j].hi[k]; }
        if (nvld) for (int k = 0; k < 3; k++) cn[k] = sum[k] / (float)nvld;
    }
#endif
    for (int j = 0; j < cnt; j++) {
        int pref = 0;
        if (box_valid(cb[j])) {
            for (int k = 0; k < 3; k++)
                if (0.5f * cb[j].lo[k] + 0.5f * cb[j].hi[k] > cn[k]) pref |= 1 << k;
        }
        int best = -1, best_cost = 99;
        for (int s = 0; s < 8; s++) {
            if (child_at[s] >= 0) continue;
            int cost = __popc(pref ^ s) * 8 + (pref ^ s);
            if (cost < best_cost) { best_cost = cost; best = s; }
        }
        child_at[best] = j;
    }
#if RT_SLOT_OPT
    if (RT_SLOT_OPT == 2 || A.max_leaf > 1)
    {   // pairwise-exchange refinement of the slot assignment.  Score of child j in slot s = dot(centre_j - cn, sign vector of s)
        // (Ylitie et al. 2017, sec. 3.2: the assignment that maximises the sum orders the children best for all eight ray octants).
        float rel[8][3];
        for (int j = 0; j < cnt; j++)
            for (int k = 0; k < 3; k++) rel[j][k] = box_valid(cb[j]) ? 0.5f * cb[j].lo[k] + 0.5f * cb[j].hi[k] - cn[k] : 0.0f;
        auto score = [&](int j, int sl) -> float {
            if (j < 0) return 0.0f;
            return ((sl & 1) ? rel[j][0] : -rel[j][0]) + ((sl & 2) ? rel[j][1] : -rel[j][1]) + ((sl & 4) ? rel[j][2] : -rel[j][2]);
        };
        for (int sweep = 0; sweep < 6; sweep++) {
            bool changed = false;
            for (int a = 0; a < 8; a++)
                for (int b = a + 1; b < 8; b++) {
                    const int ja = child_at[a], jb = child_at[b];
                    if (ja < 0 && jb < 0) continue;
                    if (score(ja, b) + score(jb, a) > score(ja, a) + score(jb, b) + 1e-12f) { child_at[a] = jb; child_at[b] = ja; changed = true; }
                }
            if (!changed) break;
        }
    }
#endif
    uint32_t k_int = 0, total_prims = 0;
    for (int j = 0; j < cnt; j++) {
        if (internal[j]) k_int++;
        else total_prims += ref_count(T, ch[j]);
    }
    uint32_t base = k_int ? atomicAdd(&A.state[ST_WIDE_COUNT], k_int) : 0u;
    uint32_t pbase = total_prims ? atomicAdd(&A.state[ST_PRIM_CURSOR], total_prims) : 0u;

    Node8 nd;
    memset(&nd, 0, sizeof(nd));
    Aabb sb[8];
    uint32_t present = 0, imask = 0, lmask = 0, rank = 0, off = 0;
    for (int s = 0; s < 8; s++) {
        int j = child_at[s];
        if (j < 0) continue;
        present |= 1u << s;
        sb[s] = cb[j];
        uint32_t c = ref_count(T, ch[j]);
        if (internal[j]) {
            imask |= 1u << s;
            nd.meta[s] = 0xFF;
            A.task_node[base + rank] = ch[j];
            A.task_parent[base + rank] = w | ((uint32_t)s << 29);
            rank++;
        } else {
            lmask |= 1u << s;
            uint32_t f = ref_first(T, ch[j]);
            for (uint32_t k = 0; k < c; k++) A.leaf_order[pbase + off + k] = T.vals[f + k];
            nd.meta[s] = (uint8_t)(off | (c << 5));
            off += c;
        }
    }
    quantise_node(nd, sb, present);
    nd.imask = (uint8_t)imask;
    nd.lmask = lmask;
    nd.child_base = A.node_offset + base;
    nd.prim_base = A.prim_offset + pbase;
    if (w == 0) { nd.parent = 0xFFFFFFFFu; nd.parent_slot = 0; }
    else {
        uint32_t tp = A.task_parent[w];
        nd.parent = A.node_offset + (tp & 0x1FFFFFFFu);
        nd.parent_slot = tp >> 29;
    }
    store_node(&A.nodes[A.node_offset + w], nd);
}

__global__ void __launch_bounds__(128) k_collapse_coop(CollapseArgs A) {
    cg::grid_group grid = cg::this_grid();
    uint32_t begin = 0, end = 1;
    uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    while (begin < end) {
        for (uint32_t w = begin + gtid; w < end; w += gsize) collapse_task(A, w);
        grid.sync();
        begin = end;
        end = *((volatile uint32_t*)&A.state[ST_WIDE_COUNT]);
        grid.sync();  // nobody may bump WIDE_COUNT before everyone has read it
    }
    if (gtid == 0 && A.node_count_out) *A.node_count_out = end;
}

// Fallback: one launch per level, the level range lives in A.state.
__global__ void k_collapse_level(CollapseArgs A, uint32_t begin, uint32_t end) {
    uint32_t w = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (w < end) collapse_task(A, w);
}

__global__ void k_empty_root(Node8* nodes, uint32_t node_offset, uint32_t prim_offset, uint32_t* state) {
    Node8 nd;
    memset(&nd, 0, sizeof(nd));
    Aabb sb[8];
    quantise_node(nd, sb, 0);
    nd.child_base = node_offset;
    nd.prim_base = prim_offset;
    nd.parent = 0xFFFFFFFFu;
    store_node(&nodes[node_offset], nd);
    state[ST_WIDE_COUNT] = 1;
}

__global__ void k_copy_u32(const uint32_t* src, uint32_t* dst) { *dst = *src; }

// ------------------------------------------------------------------------------------ 7
struct RefitArgs {
    Node8* nodes;
    uint32_t node_offset, prim_offset;
    const uint32_t* node_count;
    const Aabb* boxes;  // by leaf position
    uint32_t* counters;
};

__global__ void k_refit(RefitArgs A) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= *A.node_count) return;
    Node8* nd = &A.nodes[A.node_offset + w];
    if (nd->imask != 0) return;  // only bottom nodes start
    for (;;) {
        Node8 cur;
        {
            const uint4* s = reinterpret_cast<const uint4*>(nd);
            uint4* d = reinterpret_cast<uint4*>(&cur);
#pragma unroll
            for (int i = 0; i < (int)RT_NODE_QUADS; i++) d[i] = __ldcg(s + i);
        }
        Aabb sb[8];
        uint32_t present = cur.imask | cur.lmask, rank = 0;
        for (int s = 0; s < 8; s++) {
            if (cur.imask >> s & 1) {
                const Node8* c = &A.nodes[cur.child_base + rank++];
                const float* f = reinterpret_cast<const float*>(c) + RT_NODE_BOUNDS_FLOAT;  // Node8::lo, ::hi
                for (int k = 0; k < 3; k++) { sb[s].lo[k] = __ldcg(f + k); sb[s].hi[k] = __ldcg(f + 3 + k); }
            } else if (cur.lmask >> s & 1) {
                uint32_t off = cur.meta[s] & 31u, c = cur.meta[s] >> 5;
                Aabb b = box_empty();
                for (uint32_t k = 0; k < c; k++) box_grow(b, A.boxes[cur.prim_base - A.prim_offset + off + k]);
                sb[s] = b;
            }
        }
        quantise_node(cur, sb, present);
        {
            const uint4* s = reinterpret_cast<const uint4*>(&cur);
            uint4* d = reinterpret_cast<uint4*>(nd);
#pragma unroll
            for (int i = 0; i < (int)RT_NODE_QUADS; i++) __stcg(d + i, s[i]);
        }
        if (cur.parent == 0xFFFFFFFFu) return;
        __threadfence();
        uint32_t pl = cur.parent - A.node_offset;
        Node8* pn = &A.nodes[cur.parent];
        uint32_t need = __popc((uint32_t)__ldcg(reinterpret_cast<const unsigned char*>(pn) + 15));
        uint32_t old = atomicAdd(&A.counters[pl], 1u);
        if (old + 1 != need) return;
        nd = pn;
    }
}

// ------------------------------------------------------------------------------------ sharded TLAS builds
__device__ __forceinline__ uint32_t shard_bin(uint64_t key) { return (uint32_t)(key >> (63u - RT_SHARD_BITS)); }

__global__ void k_shard_hist(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* hist) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&hist[shard_bin(keys[i])], 1u);
}
// bin boundaries b[0..n_shards]: shard r owns bins [b[r], b[r+1]); the boundary goes where the running population first
// reaches r * n / n_shards.  One thread: 32 k bins, once per build.
__global__ void k_shard_split(const uint32_t* __restrict__ hist, uint32_t n, uint32_t n_shards, uint32_t* bounds, uint32_t* counts) {
    uint32_t r = 1, run = 0, last = 0;
    bounds[0] = 0;
    for (uint32_t b = 0; b < (1u << RT_SHARD_BITS); b++) {
        while (r < n_shards && run >= (uint32_t)(((unsigned long long)n * r) / n_shards)) {
            bounds[r] = b;
            counts[r - 1] = run - last;
            last = run;
            r++;
        }
        run += hist[b];
    }
    while (r < n_shards) { bounds[r] = 1u << RT_SHARD_BITS; counts[r - 1] = run - last; last = run; r++; }
    bounds[n_shards] = 1u << RT_SHARD_BITS;
    counts[n_shards - 1] = run - last;
    bounds[33] = 0;  // select cursor
}
__global__ void k_shard_select(const Aabb* __restrict__ boxes, const uint64_t* __restrict__ keys, uint32_t n, uint32_t* bounds, uint32_t shard,
                               Aabb* sel_boxes, uint32_t* sel_index) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t bin = shard_bin(keys[i]);
    if (bin < bounds[shard] || bin >= bounds[shard + 1]) return;
    const uint32_t pos = atomicAdd(&bounds[33], 1u);
    sel_boxes[pos] = boxes[i];
    sel_index[pos] = i;
}
__global__ void k_map_order(const uint32_t* __restrict__ treelet_order, const uint32_t* __restrict__ sel_index, uint32_t count, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = sel_index[treelet_order[i]];
}
__global__ void k_treelet_rebase(Node8* nodes, uint32_t root_at, uint32_t rest_at, uint32_t count, uint32_t slot) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;  // treelet-local node number
    if (j >= count) return;
    Node8* nd = &nodes[j == 0 ? root_at : rest_at + (j - 1)];
    const uint32_t delta = rest_at - 1u;  // local index l >= 1 lives at rest_at + l - 1
    if (nd->imask) nd->child_base += delta;
    if (j == 0) { nd->parent = 0u; nd->parent_slot = slot; }
    else nd->parent = nd->parent == 0u ? root_at : nd->parent + delta;
}
__global__ void k_tlas_top(Node8* nodes, uint32_t k, uint32_t total_nodes, uint32_t* node_count) {
    Node8 nd;
    memset(&nd, 0, sizeof(nd));
    Aabb sb[8];
    uint32_t present = 0;
    for (uint32_t s = 0; s < k && s < 8; s++) {
        const Node8& c = nodes[1 + s];
        for (int a = 0; a < 3; a++) { sb[s].lo[a] = c.lo[a]; sb[s].hi[a] = c.hi[a]; }
        present |= 1u << s;
        nd.meta[s] = 0xFF;
    }
    quantise_node(nd, sb, present);
    nd.imask = (uint8_t)present;
    nd.lmask = 0;
    nd.child_base = 1;
    nd.prim_base = 0;
    nd.parent = 0xFFFFFFFFu;
    nd.parent_slot = 0;
    store_node(&nodes[0], nd);
    if (node_count) *node_count = total_nodes;
}

template <typename T>
T* carve(char*& p, size_t count) {
    uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255);
    T* r = reinterpret_cast<T*>(a);
    p = reinterpret_cast<char*>(a + count * sizeof(T));
    return r;
}

struct Scratch {
    int* bounds;
    uint32_t* state;
    uint64_t *keys_in, *keys_out;
    uint32_t *vals_in, *vals_out;
    int *left, *right, *parent_int, *parent_leaf;
    uint32_t *first, *last, *flags;
    Aabb* ibox;
    float* dp_cost;
    uint8_t* dp_dec;
    int* task_node;
    uint32_t* task_parent;
    uint32_t* refit_counters;
    uint32_t* sah_tree;     // k_sah_tree: 16 length words, two queues of large nodes, one SahBig record per large node of a level
    uint32_t* shard_hist;   // (1 << RT_SHARD_BITS) bins + 64 words: shard bin boundaries [0..32], select cursor [33]
    void* cub_temp;
};

size_t layout(char* base, uint32_t n, size_t cub_bytes, Scratch& s) {
    char* p = base;
    uint32_t maxw = max_wide_nodes(n);
    s.bounds = carve<int>(p, 8);
    s.state = carve<uint32_t>(p, 8);
    s.keys_in = carve<uint64_t>(p, n);
    s.keys_out = carve<uint64_t>(p, n);
    s.vals_in = carve<uint32_t>(p, n);
    s.vals_out = carve<uint32_t>(p, n);
    s.left = carve<int>(p, n);
    s.right = carve<int>(p, n);
    s.parent_int = carve<int>(p, n);
    s.parent_leaf = carve<int>(p, n);
    s.first = carve<uint32_t>(p, n);
    s.last = carve<uint32_t>(p, n);
    s.flags = carve<uint32_t>(p, n);
    s.ibox = carve<Aabb>(p, n);
    s.dp_cost = carve<float>(p, (size_t)n * 7);
    s.dp_dec = carve<uint8_t>(p, (size_t)n * 8);
    s.task_node = carve<int>(p, maxw);
    s.task_parent = carve<uint32_t>(p, maxw);
    s.refit_counters = carve<uint32_t>(p, maxw);
    s.sah_tree = carve<uint32_t>(p, 16 + (size_t)(n / SAH_TREE_BIG_MIN + 2) * (SAH_BIG_WORDS + 8));
    s.shard_hist = carve<uint32_t>(p, (1u << RT_SHARD_BITS) + 64);
    s.cub_temp = carve<char>(p, cub_bytes);
    return (size_t)(p - base) + 256;
}

}  // namespace

BvhBuilder::~BvhBuilder() {
    if (scratch_) cudaFree(scratch_);
}

cudaError_t BvhBuilder::reserve(uint32_t n) {
    if (n < 2) n = 2;
    if (n <= cap_) return cudaSuccess;
    uint32_t cap = n + n / 8 + 64;
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)cap, 0, 63);
    if (e != cudaSuccess) return e;
    Scratch s;
    size_t bytes = layout(nullptr, cap, cub_bytes, s);
    if (scratch_) cudaFree(scratch_);
    scratch_ = nullptr;
    cap_ = 0;
    e = cudaMalloc(&scratch_, bytes);
    if (e != cudaSuccess) return e;
    scratch_bytes_ = bytes;
    cub_bytes_ = cub_bytes;
    cap_ = cap;
    if (coop_blocks_ == 0) {
        int dev = 0, sms = 0, per_sm = 0, coop = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_collapse_coop, 128, 0);
        coop_blocks_ = sms * (per_sm > 0 ? per_sm : 1);
        coop_ok_ = coop != 0 && per_sm > 0;
        int tree_per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tree_per_sm, k_sah_tree, SAH_TB, 0);
        if (tree_per_sm > SAH_BLOCKS_PER_SM) tree_per_sm = SAH_BLOCKS_PER_SM;  // grid.sync gets dearer with the grid; 128 x 4 / 256 x 4 / 512 x 2 measured within 15 % (profiles/r04gh_sah_build_time.txt)
        sah_tree_blocks_ = coop != 0 && tree_per_sm > 0 ? sms * tree_per_sm : 0;
    }
    return cudaSuccess;
}

cudaError_t BvhBuilder::build(const Aabb* d_boxes, uint32_t n, uint32_t max_leaf, Node8* nodes_pool, uint32_t node_offset,
                              uint32_t prim_offset, uint32_t* d_leaf_order, uint32_t* d_node_count, bool fast_sort, bool sah_collapse, cudaStream_t stream,
                              int sah_splits) {
    cudaError_t e = reserve(n);
    if (e != cudaSuccess) return e;
    Scratch s;
    layout(static_cast<char*>(scratch_), cap_, cub_bytes_, s);
    const int TB = 256;
    k_init_state<<<1, 32, 0, stream>>>(s.bounds, s.state, s.task_node, s.task_parent, n >= 2 ? 0 : ~0);
    note_launch();
    if (n == 0) {
        k_empty_root<<<1, 1, 0, stream>>>(nodes_pool, node_offset, prim_offset, s.state);
        note_launch();
        if (d_node_count) { k_copy_u32<<<1, 1, 0, stream>>>(&s.state[ST_WIDE_COUNT], d_node_count); note_launch(); }
        return cudaGetLastError();
    }
    uint32_t blocks = (n + TB - 1) / TB;
    bool radix_tree = true;
    if (sah_splits != SAH_NEVER && n >= 2 && sah_tree_blocks_ > 0) {
        // top-down binned SAH (step 4b), every level inside one cooperative launch whatever the size: nothing for the host to follow
        SahTreeArgs P;
        P.A.boxes = d_boxes; P.A.order = s.vals_out; P.A.tmp = s.vals_in;
        P.A.counters = s.state + 4;
        P.A.left = s.left; P.A.right = s.right; P.A.parent_int = s.parent_int; P.A.parent_leaf = s.parent_leaf; P.A.first = s.first; P.A.last = s.last;
        P.qs[0] = reinterpret_cast<uint4*>(s.keys_in); P.qs[1] = reinterpret_cast<uint4*>(s.keys_out);  // <= n / 2 entries of 16 bytes
        const size_t big_cap = (size_t)cap_ / SAH_TREE_BIG_MIN + 2;
        P.words = s.sah_tree;
        P.qb[0] = reinterpret_cast<uint4*>(s.sah_tree + 16);
        P.qb[1] = P.qb[0] + big_cap;
        P.big = s.sah_tree + 16 + 8 * big_cap;
        k_sah_tree_init<<<blocks, TB, 0, stream>>>(P, n);
        note_launch();
        void* args[] = {&P};
        uint32_t want = (n + 3u) / 4u;
        uint32_t grid = want < (uint32_t)sah_tree_blocks_ ? want : (uint32_t)sah_tree_blocks_;
        cudaError_t ce = cudaLaunchCooperativeKernel((void*)k_sah_tree, dim3(grid), dim3(SAH_TB), args, 0, stream);
        if (ce == cudaSuccess) { radix_tree = false; note_launch(); }
        else { (void)cudaGetLastError(); sah_tree_blocks_ = 0; }  // no cooperative launch: the radix tree below
    }
    if (radix_tree) {
    k_centroid_bounds<<<blocks, TB, 0, stream>>>(d_boxes, n, s.bounds);
    // Morton bits per axis: all 21 for a static build; for per-frame rebuilds (fast_sort) enough cells to
    // separate n primitives with 4 bits to spare (10..21), which saves radix-sort passes
    uint32_t axis_bits = fast_sort ? 10 : 21;
    while (axis_bits < 21 && (1ull << (3 * (axis_bits - 4))) < n) axis_bits++;
    const uint32_t key_bits = 3 * axis_bits;
    k_morton<<<blocks, TB, 0, stream>>>(d_boxes, n, s.bounds, 63u - key_bits, s.keys_in, s.vals_in);
    note_launch(2);
    size_t cub_bytes = cub_bytes_;
    e = cub::DeviceRadixSort::SortPairs(s.cub_temp, cub_bytes, s.keys_in, s.keys_out, s.vals_in, s.vals_out, (int)n, 0, (int)key_bits,
                                        stream);
    if (e != cudaSuccess) return e;
    }
    Tree2 T;
    T.keys = s.keys_out; T.vals = s.vals_out;
    T.left = s.left; T.right = s.right; T.parent_int = s.parent_int; T.parent_leaf = s.parent_leaf;
    T.first = s.first; T.last = s.last; T.ibox = s.ibox; T.flags = s.flags;
    T.boxes = d_boxes; T.n = (int)n;
    T.dp_cost = sah_collapse ? s.dp_cost : nullptr; T.dp_dec = s.dp_dec;
    T.max_leaf = max_leaf;
    // one node visit costs about three triangle tests (220 vs 70 instructions, profiles/r01_notes.md); entering an
    // instance (TLAS leaves, max_leaf == 1) is a constant per primitive and does not influence the cut
    T.c_node = 3.0f; T.c_prim = 1.0f;
    if (n >= 2) {
        cudaMemsetAsync(s.flags, 0, sizeof(uint32_t) * n, stream);
        if (radix_tree) { k_hierarchy<<<blocks, TB, 0, stream>>>(T); note_launch(); }
        k_fit<<<blocks, TB, 0, stream>>>(T);
        note_launch();
    }
    CollapseArgs A;
    A.T = T; A.task_node = s.task_node; A.task_parent = s.task_parent; A.state = s.state;
    A.nodes = nodes_pool; A.node_offset = node_offset; A.prim_offset = prim_offset; A.max_leaf = max_leaf;
    A.leaf_order = d_leaf_order;
    A.node_count_out = d_node_count;
    A.greedy = !sah_collapse;
    bool done = false;
    bool count_written = false;
    if (coop_ok_) {
        void* args[] = {&A};
        uint32_t want = (max_wide_nodes(n) + 127) / 128;
        uint32_t grid = want < (uint32_t)coop_blocks_ ? (want ? want : 1u) : (uint32_t)coop_blocks_;
        cudaError_t ce = cudaLaunchCooperativeKernel((void*)k_collapse_coop, dim3(grid), dim3(128), args, 0, stream);
        if (ce == cudaSuccess) { done = true; count_written = true; note_launch(); }
        else { (void)cudaGetLastError(); coop_ok_ = false; }
    }
    if (!done) {
        // level-synchronous fallback (host reads the level range back)
        uint32_t begin = 0, end = 1;
        while (begin < end) {
            k_collapse_level<<<(end - begin + 127) / 128, 128, 0, stream>>>(A, begin, end);
            note_launch();
            uint32_t cnt = 0;
            e = cudaMemcpyAsync(&cnt, &s.state[ST_WIDE_COUNT], sizeof(cnt), cudaMemcpyDeviceToHost, stream);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) return e;
            begin = end;
            end = cnt;
        }
    }
    if (d_node_count && !count_written) { k_copy_u32<<<1, 1, 0, stream>>>(&s.state[ST_WIDE_COUNT], d_node_count); note_launch(); }
    return cudaGetLastError();
}

cudaError_t BvhBuilder::shard_select(const Aabb* d_boxes, uint32_t n, uint32_t n_shards, uint32_t shard, Aabb* d_sel_boxes, uint32_t* d_sel_index,
                                     uint32_t* d_counts, cudaStream_t stream) {
    if (n_shards < 1 || n_shards > 32 || shard >= n_shards) return cudaErrorInvalidValue;
    cudaError_t e = reserve(n);
    if (e != cudaSuccess) return e;
    Scratch s;
    layout(static_cast<char*>(scratch_), cap_, cub_bytes_, s);
    const int TB = 256;
    const uint32_t blocks = (n + TB - 1) / TB;
    k_init_state<<<1, 32, 0, stream>>>(s.bounds, s.state, s.task_node, s.task_parent, 0);
    cudaMemsetAsync(s.shard_hist, 0, sizeof(uint32_t) * ((1u << RT_SHARD_BITS) + 64), stream);
    if (n) {
        k_centroid_bounds<<<blocks, TB, 0, stream>>>(d_boxes, n, s.bounds);
        k_morton<<<blocks, TB, 0, stream>>>(d_boxes, n, s.bounds, 0u, s.keys_in, s.vals_in);
        k_shard_hist<<<blocks, TB, 0, stream>>>(s.keys_in, n, s.shard_hist);
    }
    uint32_t* bounds = s.shard_hist + (1u << RT_SHARD_BITS);
    k_shard_split<<<1, 1, 0, stream>>>(s.shard_hist, n, n_shards, bounds, d_counts);
    if (n) k_shard_select<<<blocks, TB, 0, stream>>>(d_boxes, s.keys_in, n, bounds, shard, d_sel_boxes, d_sel_index);
    note_launch(n ? 6 : 2);
    return cudaGetLastError();
}

cudaError_t launch_map_order(const uint32_t* treelet_order, const uint32_t* sel_index, uint32_t count, uint32_t* leaf_order_out, cudaStream_t stream) {
    if (count) { k_map_order<<<(count + 255) / 256, 256, 0, stream>>>(treelet_order, sel_index, count, leaf_order_out); note_launch(); }
    return cudaGetLastError();
}
cudaError_t launch_treelet_rebase(Node8* nodes, uint32_t root_at, uint32_t rest_at, uint32_t count, uint32_t slot, cudaStream_t stream) {
    if (count) { k_treelet_rebase<<<(count + 255) / 256, 256, 0, stream>>>(nodes, root_at, rest_at, count, slot); note_launch(); }
    return cudaGetLastError();
}
cudaError_t launch_tlas_top(Node8* nodes, uint32_t k, uint32_t total_nodes, uint32_t* d_node_count, cudaStream_t stream) {
    k_tlas_top<<<1, 1, 0, stream>>>(nodes, k, total_nodes, d_node_count);
    note_launch();
    return cudaGetLastError();
}

cudaError_t BvhBuilder::refit(const Aabb* d_boxes_leaf_order, uint32_t n, Node8* nodes_pool, uint32_t node_offset,
                              uint32_t prim_offset, const uint32_t* d_node_count, uint32_t node_capacity,
                              cudaStream_t stream) {
    cudaError_t e = reserve(n);
    if (e != cudaSuccess) return e;
    Scratch s;
    layout(static_cast<char*>(scratch_), cap_, cub_bytes_, s);
    uint32_t maxw = max_wide_nodes(cap_);
    if (node_capacity > maxw) node_capacity = maxw;
    cudaMemsetAsync(s.refit_counters, 0, sizeof(uint32_t) * node_capacity, stream);
    RefitArgs A;
    A.nodes = nodes_pool; A.node_offset = node_offset; A.prim_offset = prim_offset; A.node_count = d_node_count;
    A.boxes = d_boxes_leaf_order; A.counters = s.refit_counters;
    k_refit<<<(node_capacity + 127) / 128, 128, 0, stream>>>(A);
    note_launch();
    return cudaGetLastError();
}

}  // namespace b200rt
