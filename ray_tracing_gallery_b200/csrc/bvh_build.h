// bvh_build.h — GPU builder of the compressed 8-wide BVH (used for every BLAS and for the TLAS).
//
// Replaces what the Vulkan driver does behind vkCmdBuildAccelerationStructuresKHR
// (reference call sites: src/util_structs.rs:269-274 BLAS/TLAS build, :345-354 TLAS update).
#pragma once
#include "rt_types.h"

namespace b200rt {

// Upper bound on wide nodes for n primitives: every wide node is rooted at its own binary internal node
// (n - 1 of them) and the cost-driven cut gives no tighter guarantee.
inline uint32_t max_wide_nodes(uint32_t n) { return n + 8; }

// BvhBuilder::build, sah_splits (the two SAH values are the same thing since the build is one stream-ordered launch)
enum { SAH_NEVER = 0, SAH_IF_STREAM_ORDERED = 1, SAH_ALWAYS = 2 };

class BvhBuilder {
public:
    BvhBuilder() = default;
    ~BvhBuilder();
    BvhBuilder(const BvhBuilder&) = delete;
    BvhBuilder& operator=(const BvhBuilder&) = delete;

    // Grow scratch for up to n primitives (no-op if already large enough).
    cudaError_t reserve(uint32_t n);

    // Build over `n` primitive boxes (device memory).  Writes wide nodes to
    // nodes_pool[node_offset ...] (root = node_offset; at most max_wide_nodes(n) are written),
    // child indices are absolute pool indices, prim_base values are prim_offset + position in
    // leaf order.  d_leaf_order[i] = index of the input primitive stored at leaf position i.
    // d_node_count (device, optional) receives the number of wide nodes written.
    // sah_collapse: children of each wide node from the SAH-optimal cut table (else greedy largest-area expansion).
    // fast_sort: Morton keys keep only as many bits as n needs (fewer radix passes; per-frame TLAS rebuilds).
    // Everything is enqueued on `stream`; no host synchronisation.
    // sah_splits: the binary tree is grown top-down with binned-SAH splits (one cooperative launch, stream-ordered like
    // everything else) instead of the Morton radix tree, which remains for devices without cooperative launches.
    cudaError_t build(const Aabb* d_boxes, uint32_t n, uint32_t max_leaf, Node8* nodes_pool, uint32_t node_offset,
                      uint32_t prim_offset, uint32_t* d_leaf_order, uint32_t* d_node_count, bool fast_sort, bool sah_collapse, cudaStream_t stream,
                      int sah_splits = SAH_NEVER);

    // Refit in place: recompute boxes bottom-up for a tree built by build() whose leaf order is
    // unchanged.  d_boxes_leaf_order[i] = new box of the primitive at leaf position i.
    cudaError_t refit(const Aabb* d_boxes_leaf_order, uint32_t n, Node8* nodes_pool, uint32_t node_offset,
                      uint32_t prim_offset, const uint32_t* d_node_count, uint32_t node_capacity, cudaStream_t stream);

    // Sharded builds (rt_group_build_tlas): the 63-bit Morton keys of all n boxes, a histogram of their top RT_SHARD_BITS bits,
    // `n_shards` contiguous key ranges of (nearly) equal population, and the boxes of range `shard` compacted into
    // d_sel_boxes (any order) with their input indices in d_sel_index.  d_counts[n_shards] (device) receives every
    // shard's population; identical on every GPU that runs this over the same boxes.
    cudaError_t shard_select(const Aabb* d_boxes, uint32_t n, uint32_t n_shards, uint32_t shard, Aabb* d_sel_boxes, uint32_t* d_sel_index,
                             uint32_t* d_counts, cudaStream_t stream);

private:
    uint32_t cap_ = 0;
    void* scratch_ = nullptr;
    size_t scratch_bytes_ = 0;
    size_t cub_bytes_ = 0;
    int coop_blocks_ = 0;
    bool coop_ok_ = true;
    int sah_tree_blocks_ = 0;
};

// ---- assembling a TLAS from treelets built on different GPUs (bvh_build.cu)
#define RT_SHARD_BITS 15u
// leaf_order_out[i] = sel_index[treelet_order[i]]: treelet-local primitive numbers -> input primitive numbers
cudaError_t launch_map_order(const uint32_t* treelet_order, const uint32_t* sel_index, uint32_t count, uint32_t* leaf_order_out, cudaStream_t stream);
// A treelet arrives built at node offset 0: its root is stored at nodes[root_at], its other `count - 1` nodes at
// nodes[rest_at ...] in their original order.  Child indices and parent links are moved accordingly; the root gets
// parent 0 / parent_slot `slot`.
cudaError_t launch_treelet_rebase(Node8* nodes, uint32_t root_at, uint32_t rest_at, uint32_t count, uint32_t slot, cudaStream_t stream);
// nodes[0] = the top node over `k` treelet roots stored at nodes[1 .. k] (their exact bounds are read from the nodes).
cudaError_t launch_tlas_top(Node8* nodes, uint32_t k, uint32_t total_nodes, uint32_t* d_node_count, cudaStream_t stream);

}  // namespace b200rt
