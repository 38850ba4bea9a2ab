// group.cu — one frame on several GPUs of one box: the rt_group_* exports of include/b200rt.h.
//
// The reference is single-GPU; its two seams a multi-GPU host goes through are device creation (src/main.rs:157-204)
// and the per-frame scene update (src/scene.rs:167-204).  One process per GPU, one RtContext each, the whole scene on
// every rank.  What crosses GPUs:
//   * instance records on a TLAS change: ONE ncclBroadcast over NVLink, landing directly in the staging instance buffer
//     the TLAS builder reads (no intermediate copy), then refit / rebuild on every rank;
//   * the frame: either the render kernels of every rank store their rows straight into rank 0's device frame through
//     NVLink peer memory (cudaIpc mapping of rank 0's allocation) and a one-thread kernel raises the rank's arrival flag
//     next to it (st.release.sys), rank 0 waits for the flags in a kernel on its stream (ld.acquire.sys) — no
//     collective, no gather kernel, no host round trip; or every rank copies its own strips over ITS OWN PCIe link into
//     one page-locked host frame shared by the ranks (POSIX shared memory, cudaHostRegister on every rank), so that
//     the host-resident frame rank 0 hands out is assembled by 2/4/8 DMA engines in parallel instead of rank 0's one.
// NCCL is resolved with dlopen at rt_group_create (the copy already loaded in the process if there is one).
#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>

#include <vector>

#include "context.h"
#include "launch_count.h"

using namespace b200rt;

namespace b200rt {
// hooks into api.cu (not part of the C ABI)
RtInstance* internal_stage_instances(RtContext* ctx, uint32_t first, uint32_t count);
int internal_begin_full_build(RtContext* ctx, uint32_t count);
int internal_build_tlas_replicated(RtContext* ctx);
cudaStream_t internal_stream(RtContext* ctx);
int internal_device(RtContext* ctx);
}  // namespace b200rt

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (handle) return true;
        // One NCCL per process: the loader identifies libraries by SONAME, so a second copy cannot be loaded next to the
        // first, and whoever comes second gets the first one's symbols.  Order: B200RT_NCCL_LIB (explicit path), the copy
        // the process already uses (e.g. the one bundled with PyTorch — ray_tracing_gallery_b200/native.py preloads it
        // when it exists, because libtorch needs its newer symbols), the system library.
        if (const char* path = getenv("B200RT_NCCL_LIB")) handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!handle) { err = std::string("NCCL not found: ") + dlerror(); return false; }
#define RT_NCCL_SYM(field, name)                                                    \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));                 \
    if (!field) { err = std::string("NCCL symbol missing: ") + name; return false; }
        RT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        RT_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        RT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        RT_NCCL_SYM(Broadcast, "ncclBroadcast")
        RT_NCCL_SYM(AllReduce, "ncclAllReduce")
        RT_NCCL_SYM(AllGather, "ncclAllGather")
        RT_NCCL_SYM(GroupStart, "ncclGroupStart")
        RT_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        RT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef RT_NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;
std::string g_group_error;

// What lives at the start of rank 0's device allocation, before the frame slots.
struct DevFlags {
    uint32_t arrived[RT_GROUP_MAX_RANKS];  // frame number whose rows of rank r are complete in rank 0's frame
    uint32_t released;                      // frames <= released may be overwritten
    uint32_t timed_out;                     // a bounded wait gave up
};
constexpr size_t kDevHeader = 4096;

// What lives at the start of the shared host segment, before the frame slots.
struct ShmHeader {
    uint32_t magic, n_ranks, width, height;
    uint64_t frame_bytes;
    std::atomic<uint64_t> arrived[RT_GROUP_MAX_RANKS];
    std::atomic<uint64_t> released;
    uint64_t ray_counts[RT_GROUP_FRAME_SLOTS][RT_GROUP_MAX_RANKS][2];
};
constexpr size_t kShmHeader = 4096;
static_assert(sizeof(ShmHeader) <= kShmHeader, "shared header fits its page");
static_assert(sizeof(ncclUniqueId) <= RT_GROUP_ID_BYTES, "id fits the ABI's 128 bytes");

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// This rank's rows of frame `seq` are in rank 0's frame: every store of the preceding kernels on this stream has been
// performed (kernel boundary), the release store orders them before the flag for the acquiring reader on rank 0.
__global__ void k_group_signal(uint32_t* flag, uint32_t seq) {
    __threadfence_system();
    st_release_sys(flag, seq);
}
// One thread per awaited word: spin until it reaches `seq` (wrap-safe), give up after `timeout_ns` and say so.
__global__ void k_group_wait(const uint32_t* words, uint32_t n, uint32_t seq, uint32_t* timed_out, unsigned long long timeout_ns) {
    if (threadIdx.x >= n) return;
    const unsigned long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(words + threadIdx.x) - seq) < 0) {
        if (global_ns() - t0 > timeout_ns) {
            *timed_out = 1u;
            break;
        }
        __nanosleep(200);
    }
}

}  // namespace

struct RtGroup {
    RtContext* ctx = nullptr;
    int n = 1, rank = 0, device = 0;
    uint32_t width = 0, height = 0;
    size_t frame_bytes = 0;
    ncclComm_t comm = nullptr;
    std::string err;
    // device path
    uint8_t* d_base = nullptr;      // rank 0: own allocation; others: cudaIpc mapping of it
    bool d_base_is_ipc = false;
    // host path
    uint8_t* h_base = nullptr;
    size_t h_bytes = 0;
    bool h_registered = false;
    // host path: two sets of {compact rows, ray counts}, so that the copy of frame k (on copy_stream, over this rank's PCIe
    // link) overlaps the rendering of frame k+1
    uint8_t* d_local[2] = {nullptr, nullptr};
    size_t d_local_bytes = 0;
    uint64_t* d_counts_host[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t rendered[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    bool copy_pending[2] = {false, false};
    uint64_t* d_counts = nullptr;   // {ray-gen segments, shadow rays} of this rank's share (device path)
    uint32_t* d_word = nullptr;     // scratch for barriers / the handle exchange (64 B)

    DevFlags* flags() const { return reinterpret_cast<DevFlags*>(d_base); }
    uint8_t* dev_frame(uint64_t seq) const { return d_base + kDevHeader + (seq % RT_GROUP_FRAME_SLOTS) * frame_bytes; }
    ShmHeader* shm() const { return reinterpret_cast<ShmHeader*>(h_base); }
    uint8_t* host_frame(uint64_t seq) const { return h_base + kShmHeader + (seq % RT_GROUP_FRAME_SLOTS) * frame_bytes; }
};

namespace {

int gfail(RtGroup* g, int code, const std::string& msg) {
    if (g) g->err = msg;
    else g_group_error = msg;
    return code;
}
#define GCK(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t _e = (call);                                                                          \
        if (_e != cudaSuccess) {                                                                          \
            (void)cudaGetLastError();                                                                     \
            return gfail(g, RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));             \
        }                                                                                                 \
    } while (0)
#define GNCCL(call)                                                                                       \
    do {                                                                                                  \
        ncclResult_t _r = (call);                                                                         \
        if (_r != ncclSuccess) return gfail(g, RT_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

struct Strips {
    uint32_t total, own, rows_last, rows_own;
    bool owns_last;
};
Strips strips_of(uint32_t height, int n, int rank) {
    Strips s;
    s.total = (height + RT_GROUP_STRIP_ROWS - 1) / RT_GROUP_STRIP_ROWS;
    s.own = s.total / n + ((uint32_t)rank < s.total % n ? 1u : 0u);
    s.rows_last = height - (s.total - 1) * RT_GROUP_STRIP_ROWS;
    s.owns_last = s.total > 0 && (s.total - 1) % n == (uint32_t)rank;
    s.rows_own = s.own * RT_GROUP_STRIP_ROWS - (s.owns_last ? RT_GROUP_STRIP_ROWS - s.rows_last : 0u);
    return s;
}

std::string shm_name(const void* id) {
    // every rank derives the same name from the id it was given (FNV-1a over the 128 bytes)
    uint64_t h = 1469598103934665603ull;
    const unsigned char* p = static_cast<const unsigned char*>(id);
    for (int i = 0; i < RT_GROUP_ID_BYTES; i++) h = (h ^ p[i]) * 1099511628211ull;
    char buf[64];
    snprintf(buf, sizeof(buf), "/b200rt_%016llx", (unsigned long long)h);
    return buf;
}

int barrier(RtGroup* g) {
    cudaStream_t st = internal_stream(g->ctx);
    GNCCL(g_nccl.AllReduce(g->d_word, g->d_word, 1, ncclUint32, ncclSum, g->comm, st));
    GCK(cudaStreamSynchronize(st));
    return RT_OK;
}

struct Publish {
    ShmHeader* hdr;
    int rank;
    uint64_t seq;
};
void CUDART_CB publish_arrival(void* p) {
    Publish* a = static_cast<Publish*>(p);
    a->hdr->arrived[a->rank].store(a->seq, std::memory_order_release);
    delete a;
}

}  // namespace

extern "C" {

const char* rt_group_last_error(const RtGroup* g) { return g ? g->err.c_str() : g_group_error.c_str(); }

int rt_group_unique_id(void* out_id) {
    RtGroup* g = nullptr;
    if (!out_id) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_unique_id: NULL");
    std::string e;
    if (!g_nccl.load(e)) return gfail(g, RT_ERR_CUDA, e);
    ncclUniqueId id;
    GNCCL(g_nccl.GetUniqueId(&id));
    memset(out_id, 0, RT_GROUP_ID_BYTES);
    memcpy(out_id, &id, sizeof(id));
    return RT_OK;
}

void rt_group_destroy(RtGroup* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->ctx) cudaStreamSynchronize(internal_stream(g->ctx));
    if (g->comm) g_nccl.CommDestroy(g->comm);
    if (g->d_base) {
        if (g->d_base_is_ipc) cudaIpcCloseMemHandle(g->d_base);
        else cudaFree(g->d_base);
    }
    if (g->h_base) {
        if (g->h_registered) cudaHostUnregister(g->h_base);
        munmap(g->h_base, g->h_bytes);
    }
    if (g->copy_stream) { cudaStreamSynchronize(g->copy_stream); cudaStreamDestroy(g->copy_stream); }
    for (int b = 0; b < 2; b++) {
        cudaFree(g->d_local[b]); cudaFree(g->d_counts_host[b]);
        if (g->rendered[b]) cudaEventDestroy(g->rendered[b]);
        if (g->copied[b]) cudaEventDestroy(g->copied[b]);
    }
    cudaFree(g->d_counts); cudaFree(g->d_word);
    (void)cudaGetLastError();
    delete g;
}

int rt_group_create(RtContext* ctx, int n_ranks, int rank, const void* id_bytes, uint32_t width, uint32_t height, RtGroup** out) {
    RtGroup* g = nullptr;
    if (!out) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_create: out is NULL");
    *out = nullptr;
    if (!ctx || !id_bytes || n_ranks < 1 || n_ranks > RT_GROUP_MAX_RANKS || rank < 0 || rank >= n_ranks || !width || !height)
        return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_create: bad argument");
    std::string e;
    if (!g_nccl.load(e)) return gfail(g, RT_ERR_CUDA, e);
    RtGroup* grp = new (std::nothrow) RtGroup();
    if (!grp) return gfail(g, RT_ERR_CUDA, "rt_group_create: out of host memory");
    grp->ctx = ctx; grp->n = n_ranks; grp->rank = rank; grp->device = internal_device(ctx);
    grp->width = width; grp->height = height; grp->frame_bytes = (size_t)width * height * 4;
    auto bail = [&](int code, const std::string& msg) {
        g_group_error = msg;
        rt_group_destroy(grp);
        return code;
    };
#define BCK(call)                                                                                     \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) { (void)cudaGetLastError(); return bail(RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } \
    } while (0)
    BCK(cudaSetDevice(grp->device));
    cudaStream_t st = internal_stream(ctx);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclResult_t nr = g_nccl.CommInitRank(&grp->comm, n_ranks, id, rank);
    if (nr != ncclSuccess) return bail(RT_ERR_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(nr));
    BCK(cudaMalloc(&grp->d_word, 64));
    BCK(cudaMemset(grp->d_word, 0, 64));
    BCK(cudaMalloc(&grp->d_counts, 16));
    BCK(cudaMemset(grp->d_counts, 0, 16));

    // ---- rank 0's device frame, mapped by every other rank through NVLink peer memory
    const size_t dev_bytes = kDevHeader + RT_GROUP_FRAME_SLOTS * grp->frame_bytes;
    cudaIpcMemHandle_t handle;
    memset(&handle, 0, sizeof(handle));
    if (rank == 0) {
        BCK(cudaMalloc(&grp->d_base, dev_bytes));
        BCK(cudaMemset(grp->d_base, 0, dev_bytes));
        if (n_ranks > 1) BCK(cudaIpcGetMemHandle(&handle, grp->d_base));
    }
    if (n_ranks > 1) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the exchange buffer holds one IPC handle");
        BCK(cudaMemcpyAsync(grp->d_word, &handle, 64, cudaMemcpyHostToDevice, st));
        nr = g_nccl.Broadcast(grp->d_word, grp->d_word, 64, ncclUint8, 0, grp->comm, st);
        if (nr != ncclSuccess) return bail(RT_ERR_CUDA, std::string("ncclBroadcast(ipc handle): ") + g_nccl.GetErrorString(nr));
        BCK(cudaMemcpyAsync(&handle, grp->d_word, 64, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        BCK(cudaMemsetAsync(grp->d_word, 0, 64, st));
        if (rank != 0) {
            void* p = nullptr;
            BCK(cudaIpcOpenMemHandle(&p, handle, cudaIpcMemLazyEnablePeerAccess));
            grp->d_base = static_cast<uint8_t*>(p);
            grp->d_base_is_ipc = true;
        }
    }

    // ---- the shared page-locked host frame
    grp->h_bytes = kShmHeader + RT_GROUP_FRAME_SLOTS * grp->frame_bytes;
    const std::string name = shm_name(id_bytes);
    int fd = -1;
    if (rank == 0) {
        shm_unlink(name.c_str());
        fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)grp->h_bytes) != 0) {
            if (fd >= 0) close(fd);
            return bail(RT_ERR_CUDA, "rt_group_create: cannot create the shared host frame " + name);
        }
    }
    if (n_ranks > 1 && barrier(grp) != RT_OK) return bail(RT_ERR_CUDA, std::string(grp->err));  // the segment exists
    if (rank != 0) {
        fd = shm_open(name.c_str(), O_RDWR, 0600);
        if (fd < 0) return bail(RT_ERR_CUDA, "rt_group_create: cannot open the shared host frame " + name);
    }
    void* hp = mmap(nullptr, grp->h_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (hp == MAP_FAILED) return bail(RT_ERR_CUDA, "rt_group_create: mmap of the shared host frame failed");
    grp->h_base = static_cast<uint8_t*>(hp);
    if (rank == 0) {
        ShmHeader* h = new (grp->h_base) ShmHeader();
        h->magic = 0xB200F7A3u; h->n_ranks = (uint32_t)n_ranks; h->width = width; h->height = height; h->frame_bytes = grp->frame_bytes;
        for (auto& a : h->arrived) a.store(0);
        h->released.store(0);
        memset(h->ray_counts, 0, sizeof(h->ray_counts));
    }
    BCK(cudaHostRegister(grp->h_base, grp->h_bytes, cudaHostRegisterPortable));
    grp->h_registered = true;
    if (n_ranks > 1 && barrier(grp) != RT_OK) return bail(RT_ERR_CUDA, std::string(grp->err));  // everyone has it mapped
    if (rank == 0) shm_unlink(name.c_str());  // the mappings keep it alive; nothing is left behind in /dev/shm

    const Strips s = strips_of(height, n_ranks, rank);
    grp->d_local_bytes = (size_t)(s.rows_own ? s.rows_own : 1u) * width * 4;
    BCK(cudaStreamCreateWithFlags(&grp->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        BCK(cudaMalloc(&grp->d_local[b], grp->d_local_bytes));
        BCK(cudaMalloc(&grp->d_counts_host[b], 16));
        BCK(cudaEventCreateWithFlags(&grp->rendered[b], cudaEventDisableTiming));
        BCK(cudaEventCreateWithFlags(&grp->copied[b], cudaEventDisableTiming));
    }
#undef BCK
    *out = grp;
    return RT_OK;
}

uint32_t rt_group_partition(const RtGroup* g, RtRenderParams* p) {
    if (!g) return 0;
    const Strips s = strips_of(g->height, g->n, g->rank);
    if (p) {
        p->width = g->width; p->height = g->height;
        p->tile_x0 = p->tile_y0 = p->tile_w = p->tile_h = 0;
        p->strip_height = g->n > 1 ? RT_GROUP_STRIP_ROWS : 0;
        p->strip_count = g->n > 1 ? (uint32_t)g->n : 0;
        p->strip_index = g->n > 1 ? (uint32_t)g->rank : 0;
    }
    return s.rows_own;
}

static int group_update(RtGroup* g, int root, uint32_t first, uint32_t count, const void* records, cudaMemcpyKind kind, uint32_t mode) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    if (root < 0 || root >= g->n) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_update_instances: bad root");
    if (g->rank == root && count && !records) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_update_instances: the root needs the records");
    GCK(cudaSetDevice(g->device));
    cudaStream_t st = internal_stream(g->ctx);
    if (count) {
        // the staging instance buffer of the context: what the refit / rebuild that follows reads
        RtInstance* dst = internal_stage_instances(g->ctx, first, count);
        if (!dst) return gfail(g, RT_ERR_OUT_OF_RANGE, std::string("rt_group_update_instances: ") + rt_last_error(g->ctx));
        if (g->rank == root) GCK(cudaMemcpyAsync(dst, records, sizeof(RtInstance) * (size_t)count, kind, st));
        if (g->n > 1) GNCCL(g_nccl.Broadcast(dst, dst, sizeof(RtInstance) * (size_t)count, ncclUint8, root, g->comm, st));
    }
    int rc = rt_update_tlas(g->ctx, mode);
    if (rc) return gfail(g, rc, std::string("rt_group_update_instances: ") + rt_last_error(g->ctx));
    return RT_OK;
}
int rt_group_update_instances(RtGroup* g, int root, uint32_t first, uint32_t count, const RtInstance* host_records, uint32_t mode) {
    return group_update(g, root, first, count, host_records, cudaMemcpyHostToDevice, mode);
}
int rt_group_update_instances_device(RtGroup* g, int root, uint32_t first, uint32_t count, const void* device_records, uint32_t mode) {
    return group_update(g, root, first, count, device_records, cudaMemcpyDeviceToDevice, mode);
}

int rt_group_build_tlas(RtGroup* g, int root, const RtInstance* host_records, uint32_t count, uint32_t flags) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    if (root < 0 || root >= g->n) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_build_tlas: bad root");
    if (g->rank == root && count && !host_records) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_build_tlas: the root needs the records");
    GCK(cudaSetDevice(g->device));
    RtContext* ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    const int N = g->n;
    // ---- the records: root -> every rank, straight into the instance buffer of the set the build writes
    int rc = internal_begin_full_build(ctx, count);
    if (rc) return gfail(g, rc, std::string("rt_group_build_tlas: ") + rt_last_error(ctx));
    RtContext::TlasSet& D = ctx->sets[ctx->cur];
    if (count) {
        if (g->rank == root) GCK(cudaMemcpyAsync(D.d_instances, host_records, sizeof(RtInstance) * (size_t)count, cudaMemcpyHostToDevice, st));
        if (N > 1) GNCCL(g_nccl.Broadcast(D.d_instances, D.d_instances, sizeof(RtInstance) * (size_t)count, ncclUint8, root, g->comm, st));
    }
    const bool sharded = (flags & RT_GROUP_BUILD_FORCE_SHARDED) ? (N <= 8) : (N > 1 && N <= 8 && count >= 1024u * (uint32_t)N);
    if (!sharded) {
        rc = internal_build_tlas_replicated(ctx);
        if (rc) return gfail(g, rc, std::string("rt_group_build_tlas: ") + rt_last_error(ctx));
        GCK(cudaStreamSynchronize(st));
        return RT_OK;
    }
    GCK(cudaEventRecord(ctx->ev[2], st));
    // ---- every rank: traversal records + world boxes of ALL instances (cheap, and the keys need the boxes), then the shard
    const uint32_t n = count;
    GCK(launch_prepare_instances(D.d_instances, n, ctx->d_blas_info.ptr, (uint32_t)ctx->models.size(), nullptr, ctx->d_inst_unsorted, ctx->d_inst_boxes, st));
    Aabb* d_sel_boxes = nullptr;
    uint32_t *d_sel_index = nullptr, *d_counts = nullptr, *d_treelet_order = nullptr, *d_cnt = nullptr;
    Node8* d_treelet = nullptr;
    auto cleanup = [&]() { cudaFree(d_sel_boxes); cudaFree(d_sel_index); cudaFree(d_counts); cudaFree(d_treelet_order); cudaFree(d_cnt); cudaFree(d_treelet); };
#define SCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess) { (void)cudaGetLastError(); cleanup(); return gfail(g, RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } \
    } while (0)
#define SNCCL(call)                                                                                        \
    do {                                                                                                   \
        ncclResult_t _r = (call);                                                                          \
        if (_r != ncclSuccess) { cleanup(); return gfail(g, RT_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(_r)); } \
    } while (0)
    SCK(cudaMalloc(&d_sel_boxes, sizeof(Aabb) * (size_t)(n ? n : 1)));
    SCK(cudaMalloc(&d_sel_index, sizeof(uint32_t) * (size_t)(n ? n : 1)));
    SCK(cudaMalloc(&d_counts, sizeof(uint32_t) * 64));
    SCK(cudaMalloc(&d_cnt, sizeof(uint32_t) * 64));
    SCK(ctx->builder.shard_select(ctx->d_inst_boxes, n, (uint32_t)N, (uint32_t)g->rank, d_sel_boxes, d_sel_index, d_counts, st));
    uint32_t counts[32] = {0};
    SCK(cudaMemcpyAsync(counts, d_counts, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    uint32_t first[33] = {0};
    for (int r = 0; r < N; r++) first[r + 1] = first[r] + counts[r];
    if (first[N] != n) { cleanup(); return gfail(g, RT_ERR_CUDA, "rt_group_build_tlas: shard populations do not add up"); }
    // ---- this rank's treelet: the ordinary builder over its key range, leaf positions offset to its place in the global order
    const uint32_t mine = counts[g->rank];
    SCK(cudaMalloc(&d_treelet, sizeof(Node8) * (size_t)max_wide_nodes(mine)));
    SCK(cudaMalloc(&d_treelet_order, sizeof(uint32_t) * (size_t)(mine ? mine : 1)));
    uint32_t my_nodes = 0;
    if (mine) {
        SCK(ctx->builder.build(d_sel_boxes, mine, 1, d_treelet, 0, first[g->rank], d_treelet_order, d_cnt + 32, true, /*sah_collapse=*/true, st, /*sah_splits=*/RT_TLAS_SAH ? SAH_ALWAYS : SAH_NEVER));
        SCK(launch_map_order(d_treelet_order, d_sel_index, mine, D.d_leaf_order + first[g->rank], st));
    } else {
        SCK(cudaMemsetAsync(d_cnt + 32, 0, sizeof(uint32_t), st));
    }
    // ---- node counts of all treelets (each rank only knows its own, and only on the device)
    if (N > 1) SNCCL(g_nccl.AllGather(d_cnt + 32, d_cnt, 1, ncclUint32, g->comm, st));
    else SCK(cudaMemcpyAsync(d_cnt, d_cnt + 32, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    uint32_t node_counts[32] = {0};
    SCK(cudaMemcpyAsync(node_counts, d_cnt, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    my_nodes = node_counts[g->rank];
    // ---- layout: nodes[0] top, nodes[1..K] the roots of the K non-empty treelets, then each treelet's other nodes, compact
    uint32_t K = 0, root_at[32], rest_at[32], cursor = 1;
    for (int r = 0; r < N; r++) if (counts[r]) root_at[r] = 1 + K++;
    cursor = 1 + K;
    for (int r = 0; r < N; r++) { rest_at[r] = cursor; if (counts[r]) cursor += node_counts[r] - 1; }
    const uint32_t total_nodes = cursor;
    if (total_nodes > ctx->tlas_node_cap) { cleanup(); return gfail(g, RT_ERR_OUT_OF_RANGE, "rt_group_build_tlas: node pool too small for the assembled tree"); }
    // ---- exchange over NVLink: exactly the nodes in use and the leaf-order ranges, one broadcast per treelet part
    if (N > 1) SNCCL(g_nccl.GroupStart());
    for (int r = 0; r < N; r++) {
        if (!counts[r]) continue;
        const bool me = r == g->rank;
        if (N > 1) {
            SNCCL(g_nccl.Broadcast(me ? (const void*)d_treelet : (const void*)(D.d_tlas_nodes + root_at[r]), D.d_tlas_nodes + root_at[r], sizeof(Node8), ncclUint8, r, g->comm, st));
            if (node_counts[r] > 1)
                SNCCL(g_nccl.Broadcast(me ? (const void*)(d_treelet + 1) : (const void*)(D.d_tlas_nodes + rest_at[r]), D.d_tlas_nodes + rest_at[r],
                                       sizeof(Node8) * (size_t)(node_counts[r] - 1), ncclUint8, r, g->comm, st));
            SNCCL(g_nccl.Broadcast(D.d_leaf_order + first[r], D.d_leaf_order + first[r], sizeof(uint32_t) * (size_t)counts[r], ncclUint8, r, g->comm, st));
        } else {
            SCK(cudaMemcpyAsync(D.d_tlas_nodes + root_at[r], d_treelet, sizeof(Node8), cudaMemcpyDeviceToDevice, st));
            if (my_nodes > 1) SCK(cudaMemcpyAsync(D.d_tlas_nodes + rest_at[r], d_treelet + 1, sizeof(Node8) * (size_t)(my_nodes - 1), cudaMemcpyDeviceToDevice, st));
        }
    }
    if (N > 1) SNCCL(g_nccl.GroupEnd());
    // ---- every rank: move the links to the assembled layout, put the top node over the treelet roots, records into leaf order
    uint32_t slot = 0;
    for (int r = 0; r < N; r++)
        if (counts[r]) SCK(launch_treelet_rebase(D.d_tlas_nodes, root_at[r], rest_at[r], node_counts[r], slot++, st));
    SCK(launch_tlas_top(D.d_tlas_nodes, K, total_nodes, D.d_node_count, st));
    SCK(launch_gather_instances(ctx->d_inst_unsorted, D.d_leaf_order, n, D.d_inst_rt, st));
    SCK(cudaEventRecord(ctx->ev[3], st));
    SCK(cudaStreamSynchronize(st));
#undef SCK
#undef SNCCL
    cleanup();
    ctx->tlas_timed = true;
    ctx->tlas_built = true;
    ctx->writes_since_build = 0;
    return RT_OK;
}

int rt_group_render_device(RtGroup* g, uint64_t seq, const RtUniforms* u, const RtRenderParams* params) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    if (!u || !params || !seq) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_render_device: bad argument (frame numbers start at 1)");
    GCK(cudaSetDevice(g->device));
    cudaStream_t st = internal_stream(g->ctx);
    DevFlags* f = g->flags();
    if (seq > RT_GROUP_FRAME_SLOTS) {  // the slot's previous frame must have been released by rank 0
        k_group_wait<<<1, 32, 0, st>>>(&f->released, 1, (uint32_t)(seq - RT_GROUP_FRAME_SLOTS), &f->timed_out, 10000000000ull);
        note_launch();
    }
    RtRenderParams p = *params;
    rt_group_partition(g, &p);
    p.flags |= RT_RENDER_OUTPUT_IMAGE_ROWS;
    RtFrameOutputs out;
    memset(&out, 0, sizeof(out));
    out.rgba8 = g->dev_frame(seq);
    out.ray_counts = g->d_counts;
    int rc = rt_render_device(g->ctx, u, &p, &out);
    if (rc) return gfail(g, rc, std::string("rt_group_render_device: ") + rt_last_error(g->ctx));
    k_group_signal<<<1, 1, 0, st>>>(&f->arrived[g->rank], (uint32_t)seq);
    note_launch();

    GCK(cudaGetLastError());
    return RT_OK;
}

int rt_group_acquire_device(RtGroup* g, uint64_t seq, uint8_t** out_rgba8) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    if (g->rank != 0) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_acquire_device: rank 0 owns the frame");
    GCK(cudaSetDevice(g->device));
    DevFlags* f = g->flags();
    k_group_wait<<<1, 32, 0, internal_stream(g->ctx)>>>(f->arrived, (uint32_t)g->n, (uint32_t)seq, &f->timed_out, 10000000000ull);
    note_launch();
    GCK(cudaGetLastError());
    if (out_rgba8) *out_rgba8 = g->dev_frame(seq);
    return RT_OK;
}

int rt_group_render_host(RtGroup* g, uint64_t seq, const RtUniforms* u, const RtRenderParams* params) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    if (!u || !params || !seq) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_render_host: bad argument (frame numbers start at 1)");
    GCK(cudaSetDevice(g->device));
    cudaStream_t st = internal_stream(g->ctx);
    ShmHeader* h = g->shm();
    if (seq > RT_GROUP_FRAME_SLOTS) {  // bounded wait for the slot
        const auto t0 = std::chrono::steady_clock::now();
        while (h->released.load(std::memory_order_acquire) + RT_GROUP_FRAME_SLOTS < seq) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10))
                return gfail(g, RT_ERR_OUT_OF_RANGE, "rt_group_render_host: rank 0 did not release the frame slot within 10 s");
            std::this_thread::yield();
        }
    }
    RtRenderParams p = *params;
    rt_group_partition(g, &p);
    p.flags &= ~(uint32_t)RT_RENDER_OUTPUT_IMAGE_ROWS;
    const int b2 = (int)(seq & 1);
    if (g->copy_pending[b2]) GCK(cudaStreamWaitEvent(st, g->copied[b2], 0));  // the set's previous frame has left the device
    RtFrameOutputs out;
    memset(&out, 0, sizeof(out));
    out.rgba8 = g->d_local[b2];
    out.ray_counts = g->d_counts_host[b2];
    int rc = rt_render_device(g->ctx, u, &p, &out);
    if (rc) return gfail(g, rc, std::string("rt_group_render_host: ") + rt_last_error(g->ctx));
    GCK(cudaEventRecord(g->rendered[b2], st));
    // this rank's strips -> their rows of the shared host frame, over this rank's own PCIe link, on the copy stream
    cudaStream_t cs = g->copy_stream;
    GCK(cudaStreamWaitEvent(cs, g->rendered[b2], 0));
    const Strips s = strips_of(g->height, g->n, g->rank);
    const size_t strip_bytes = (size_t)RT_GROUP_STRIP_ROWS * g->width * 4;
    uint8_t* dst = g->host_frame(seq) + (size_t)g->rank * strip_bytes;
    const uint32_t full = s.own - ((s.owns_last && s.rows_last < RT_GROUP_STRIP_ROWS) ? 1u : 0u);
    if (full) GCK(cudaMemcpy2DAsync(dst, strip_bytes * g->n, g->d_local[b2], strip_bytes, strip_bytes, full, cudaMemcpyDeviceToHost, cs));
    if (full < s.own)
        GCK(cudaMemcpyAsync(dst + (size_t)full * strip_bytes * g->n, g->d_local[b2] + (size_t)full * strip_bytes, (size_t)s.rows_last * g->width * 4,
                            cudaMemcpyDeviceToHost, cs));
    GCK(cudaMemcpyAsync(&h->ray_counts[seq % RT_GROUP_FRAME_SLOTS][g->rank][0], g->d_counts_host[b2], 16, cudaMemcpyDeviceToHost, cs));
    Publish* pub = new (std::nothrow) Publish{h, g->rank, seq};
    if (!pub) return gfail(g, RT_ERR_CUDA, "rt_group_render_host: out of host memory");
    GCK(cudaLaunchHostFunc(cs, publish_arrival, pub));
    GCK(cudaEventRecord(g->copied[b2], cs));
    g->copy_pending[b2] = true;
    return RT_OK;
}

int rt_group_acquire_host(RtGroup* g, uint64_t seq, uint32_t timeout_ms, const uint8_t** out_rgba8, uint64_t* ray_counts) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    if (g->rank != 0) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_acquire_host: rank 0 owns the frame");
    ShmHeader* h = g->shm();
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < g->n; r++) {
        while (h->arrived[r].load(std::memory_order_acquire) < seq) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms ? timeout_ms : 10000u))
                return gfail(g, RT_ERR_OUT_OF_RANGE, "rt_group_acquire_host: rank " + std::to_string(r) + " did not deliver frame " + std::to_string(seq));
            std::this_thread::yield();
        }
    }
    if (ray_counts) {
        ray_counts[0] = ray_counts[1] = 0;
        for (int r = 0; r < g->n; r++) {
            ray_counts[0] += h->ray_counts[seq % RT_GROUP_FRAME_SLOTS][r][0];
            ray_counts[1] += h->ray_counts[seq % RT_GROUP_FRAME_SLOTS][r][1];
        }
    }
    if (out_rgba8) *out_rgba8 = g->host_frame(seq);
    return RT_OK;
}

int rt_group_release(RtGroup* g, uint64_t seq) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    if (g->rank != 0) return gfail(g, RT_ERR_INVALID_ARGUMENT, "rt_group_release: rank 0 owns the frame");
    // Frames are numbered across both paths and released in order, so both "released" words advance together: the host
    // word at once (a frame of the device path never occupied a host slot), the device word in stream order behind
    // whatever rank 0 enqueued to consume the frame.
    g->shm()->released.store(seq, std::memory_order_release);
    GCK(cudaSetDevice(g->device));
    k_group_signal<<<1, 1, 0, internal_stream(g->ctx)>>>(&g->flags()->released, (uint32_t)seq);
    note_launch();
    GCK(cudaGetLastError());
    return RT_OK;
}

int rt_group_readback(RtGroup* g, uint64_t seq, void* host_rgba8, size_t capacity_bytes) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    if (!host_rgba8 || capacity_bytes < g->frame_bytes) return gfail(g, RT_ERR_OUT_OF_RANGE, "rt_group_readback: destination too small");
    uint8_t* src = nullptr;
    int rc = rt_group_acquire_device(g, seq, &src);
    if (rc) return rc;
    cudaStream_t st = internal_stream(g->ctx);
    GCK(cudaMemcpyAsync(host_rgba8, src, g->frame_bytes, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    uint32_t t = 0;
    GCK(cudaMemcpy(&t, &g->flags()->timed_out, 4, cudaMemcpyDeviceToHost));
    if (t) return gfail(g, RT_ERR_OUT_OF_RANGE, "rt_group_readback: a rank did not deliver its rows within 10 s");
    return RT_OK;
}

int rt_group_local_ray_counts(RtGroup* g, uint64_t** out) {
    if (!g || !out) return RT_ERR_INVALID_ARGUMENT;
    *out = g->d_counts;
    return RT_OK;
}

int rt_group_barrier(RtGroup* g) {
    if (!g) return RT_ERR_INVALID_ARGUMENT;
    GCK(cudaSetDevice(g->device));
    if (g->copy_stream) GCK(cudaStreamSynchronize(g->copy_stream));
    if (g->n > 1) {
        int rc = barrier(g);
        if (rc) return rc;
    } else {
        GCK(cudaStreamSynchronize(internal_stream(g->ctx)));
    }
    if (g->rank == 0 && g->d_base) {
        uint32_t t = 0;
        GCK(cudaMemcpy(&t, &g->flags()->timed_out, 4, cudaMemcpyDeviceToHost));
        if (t) {
            GCK(cudaMemset(&g->flags()->timed_out, 0, 4));
            return gfail(g, RT_ERR_OUT_OF_RANGE, "rt_group: a rank did not deliver its rows within 10 s");
        }
    }
    return RT_OK;
}

}  // extern "C"
