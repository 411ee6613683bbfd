/*
 * rt_comm.cu — multi-GPU behind the C ABI (include/rt_b200.h: rt_comm_*, rt_scene_broadcast, rt_gather_framebuffer).
 *
 * The reference is single-GPU (optimized.cu:774-884 drives one device). Rendering shards naturally: pixels are independent,
 * so the only exchanges are one scene broadcast before and one framebuffer gather after a frame:
 *
 *   rt_scene_broadcast      the packed scene blob (rt_layout.h) of the root rank, ncclBroadcast on the scenes' streams, adopted
 *                           in place by the other ranks — the scene is built (OBJ, BVH, repack) once, on the root only
 *   rt_gather_framebuffer   the row-interleaved bands (row % nranks == rank) to the root rank: grouped ncclSend / ncclRecv
 *                           into a staging area + one strided device copy per source rank (NCCL has no native gather), or
 *                           direct peer stores when all ranks live in one process (rt_comm_init_all, peer access enabled)
 *
 * Two ways to form a communicator: one process (or thread) per GPU with an application-transported unique id
 * (rt_comm_unique_id + rt_comm_init — what torchrun / MPI launchers use), or one process driving N devices
 * (rt_comm_init_all — what `rt_render --gpus N` uses).
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2): a process that already carries an NCCL (a PyTorch process) keeps
 * using that one copy, and librtb200.so has no link-time dependency for single-GPU callers. No CPU path.
 */
#include "host_common.h"
#include "rt_rows.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> /* types and enums only; nothing here links against libnccl */

#include <cstring>
#include <mutex>
#include <vector>

namespace rtb {
cudaStream_t scene_stream(rt_scene* s);
int scene_device(rt_scene* s);
int scene_blob_reserve(rt_scene* s, size_t bytes, void** device_ptr);
} // namespace rtb

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        bool all = true;
        auto bind = [&](auto& fn, const char* sym) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.handle, sym));
            all = all && fn != nullptr;
        };
        bind(api.GetUniqueId, "ncclGetUniqueId");
        bind(api.CommInitRank, "ncclCommInitRank");
        bind(api.CommInitAll, "ncclCommInitAll");
        bind(api.CommDestroy, "ncclCommDestroy");
        bind(api.Broadcast, "ncclBroadcast");
        bind(api.AllGather, "ncclAllGather");
        bind(api.Send, "ncclSend");
        bind(api.Recv, "ncclRecv");
        bind(api.GroupStart, "ncclGroupStart");
        bind(api.GroupEnd, "ncclGroupEnd");
        bind(api.GetErrorString, "ncclGetErrorString");
        bind(api.GetVersion, "ncclGetVersion");
        api.ok = all;
    });
    return api;
}

#define NCCL_TRY(expr)                                                                                                  \
    do {                                                                                                                \
        ncclResult_t r__ = (expr);                                                                                      \
        if (r__ != ncclSuccess) return rtb::fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)
#define CUDA_TRY(expr)                                                                                                  \
    do {                                                                                                                \
        cudaError_t e__ = (expr);                                                                                       \
        if (e__ != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) {
        cudaGetDevice(&prev);
        cudaSetDevice(dev);
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

} // namespace

struct rt_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    bool same_process = false;       /* formed by rt_comm_init_all: peer stores may replace the send / recv gather */
    unsigned long long* d_word = nullptr; /* 8-byte device scratch (blob size) */
    unsigned char* staging = nullptr;     /* root: the bands of the other ranks as they arrive */
    size_t staging_bytes = 0;
};

extern "C" {

int rt_comm_available(int* nccl_version) {
    NcclApi& n = nccl();
    if (!n.ok) return rtb::fail(RT_ERR_UNSUPPORTED, "rt_comm: libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbol missing");
    int v = 0;
    if (n.GetVersion(&v) != ncclSuccess) v = 0;
    if (nccl_version) *nccl_version = v;
    return RT_OK;
}

int rt_comm_unique_id(uint8_t id[128]) {
    if (!id) return rtb::fail(RT_ERR_INVALID, "rt_comm_unique_id: NULL id");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    int rc = rt_comm_available(nullptr);
    if (rc != RT_OK) return rc;
    ncclUniqueId u;
    NCCL_TRY(nccl().GetUniqueId(&u));
    memcpy(id, &u, 128);
    return RT_OK;
}

static int finish_init(rt_comm* c) {
    DevGuard g(c->device);
    CUDA_TRY(cudaMalloc(&c->d_word, sizeof(unsigned long long)));
    return RT_OK;
}

int rt_comm_init(rt_comm** out, int nranks, int rank, const uint8_t id[128], int device) {
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return rtb::fail(RT_ERR_INVALID, "rt_comm_init: bad argument");
    *out = nullptr;
    int rc = rt_comm_available(nullptr);
    if (rc != RT_OK) return rc;
    DevGuard g(device);
    ncclUniqueId u;
    memcpy(&u, id, 128);
    rt_comm* c = new rt_comm();
    c->nranks = nranks;
    c->rank = rank;
    c->device = device;
    ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) {
        delete c;
        return rtb::fail(RT_ERR_CUDA, "rt_comm_init: ncclCommInitRank: %s", nccl().GetErrorString(r));
    }
    rc = finish_init(c);
    if (rc != RT_OK) {
        rt_comm_destroy(c);
        return rc;
    }
    *out = c;
    return RT_OK;
}

int rt_comm_init_all(rt_comm** comms, int ndev, const int* devices) {
    if (!comms || ndev < 1) return rtb::fail(RT_ERR_INVALID, "rt_comm_init_all: bad argument");
    for (int k = 0; k < ndev; k++) comms[k] = nullptr;
    int rc = rt_comm_available(nullptr);
    if (rc != RT_OK) return rc;
    std::vector<int> devs(ndev);
    for (int k = 0; k < ndev; k++) devs[k] = devices ? devices[k] : k;
    std::vector<ncclComm_t> nc(ndev);
    NCCL_TRY(nccl().CommInitAll(nc.data(), ndev, devs.data()));
    for (int k = 0; k < ndev; k++) {
        rt_comm* c = new rt_comm();
        c->comm = nc[k];
        c->nranks = ndev;
        c->rank = k;
        c->device = devs[k];
        c->same_process = true;
        comms[k] = c;
        rc = finish_init(c);
        if (rc != RT_OK) return rc;
    }
    /* one process: rank r may store straight into the root's frame when the devices can address each other */
    for (int a = 0; a < ndev; a++) {
        DevGuard g(devs[a]);
        for (int b = 0; b < ndev; b++) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devs[a], devs[b]) != cudaSuccess || !can) {
                for (int k = 0; k < ndev; k++) comms[k]->same_process = false;
                continue;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                for (int k = 0; k < ndev; k++) comms[k]->same_process = false;
            cudaGetLastError();
        }
    }
    return RT_OK;
}

void rt_comm_destroy(rt_comm* c) {
    if (!c) return;
    DevGuard g(c->device);
    if (c->d_word) cudaFree(c->d_word);
    if (c->staging) cudaFree(c->staging);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

int rt_comm_rank(const rt_comm* c, int* rank, int* nranks) {
    if (!c) return rtb::fail(RT_ERR_INVALID, "rt_comm_rank: NULL communicator");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return RT_OK;
}

int rt_scene_broadcast(rt_scene* s, rt_comm* c, int root, size_t* bytes_out) {
    if (!s || !c || root < 0 || root >= c->nranks) return rtb::fail(RT_ERR_INVALID, "rt_scene_broadcast: bad argument");
    if (rtb::scene_device(s) != c->device) return rtb::fail(RT_ERR_INVALID, "rt_scene_broadcast: scene on device %d, communicator on %d", rtb::scene_device(s), c->device);
    DevGuard g(c->device);
    cudaStream_t stream = rtb::scene_stream(s);
    void* src = nullptr;
    size_t bytes = 0;
    if (c->rank == root) {
        int rc = rt_scene_blob_export(s, &src, &bytes);
        if (rc != RT_OK) return rc;
    }
    unsigned long long word = bytes;
    if (c->rank == root) CUDA_TRY(cudaMemcpyAsync(c->d_word, &word, sizeof word, cudaMemcpyHostToDevice, stream));
    NCCL_TRY(nccl().Broadcast(c->d_word, c->d_word, sizeof word, ncclChar, root, c->comm, stream));
    CUDA_TRY(cudaMemcpyAsync(&word, c->d_word, sizeof word, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    bytes = (size_t)word;
    void* dst = src;
    if (c->rank != root) {
        int rc = rtb::scene_blob_reserve(s, bytes, &dst); /* received straight into the scene's own blob */
        if (rc != RT_OK) return rc;
    }
    NCCL_TRY(nccl().Broadcast(dst, dst, bytes, ncclChar, root, c->comm, stream));
    if (c->rank != root) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        int rc = rt_scene_blob_import(s, dst, bytes); /* same pointer: header adopted, nothing copied */
        if (rc != RT_OK) return rc;
    }
    if (bytes_out) *bytes_out = bytes;
    return RT_OK;
}

int rt_gather_framebuffer(rt_scene* s, rt_comm* c, const void* band, int32_t W, int32_t H, int32_t bytes_per_pixel, void* frame, int root) {
    return rt_gather_framebuffer_groups(s, c, band, W, H, bytes_per_pixel, 1, frame, root);
}

/* row_group G: rank r rendered the groups r, r + nranks, ... of G consecutive rows each (rt_params: row_begin = r G, row_step = nranks G,
 * row_group = G) */
int rt_gather_framebuffer_groups(rt_scene* s, rt_comm* c, const void* band, int32_t W, int32_t H, int32_t bytes_per_pixel, int32_t row_group, void* frame,
                                 int root) {
    if (!s || !c || !band || W <= 0 || H <= 0 || bytes_per_pixel <= 0 || root < 0 || root >= c->nranks || row_group < 1 || row_group > 64 ||
        (row_group & (row_group - 1)) != 0)
        return rtb::fail(RT_ERR_INVALID, "rt_gather_framebuffer: bad argument");
    if (c->rank == root && !frame) return rtb::fail(RT_ERR_INVALID, "rt_gather_framebuffer: the root rank needs a frame buffer");
    DevGuard g(c->device);
    cudaStream_t stream = rtb::scene_stream(s);
    const size_t line = (size_t)W * bytes_per_pixel;
    const int n = c->nranks, G = row_group;
    if (c->rank != root) {
        const size_t mine = (size_t)rtb::shard_rows(H, c->rank, n, G) * line;
        if (mine) NCCL_TRY(nccl().Send(band, mine, ncclChar, root, c->comm, stream));
        return RT_OK;
    }
    /* root: its own band goes straight to its rows, the others arrive in the staging area */
    size_t need = 0;
    for (int r = 0; r < n; r++)
        if (r != root) need += (size_t)rtb::shard_rows(H, r, n, G) * line;
    if (c->staging_bytes < need) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (c->staging) cudaFree(c->staging);
        c->staging = nullptr;
        c->staging_bytes = 0;
        CUDA_TRY(cudaMalloc(&c->staging, need));
        c->staging_bytes = need;
    }
    NCCL_TRY(nccl().GroupStart());
    size_t off = 0;
    for (int r = 0; r < n; r++) {
        if (r == root) continue;
        const size_t sz = (size_t)rtb::shard_rows(H, r, n, G) * line;
        if (sz) NCCL_TRY(nccl().Recv(c->staging + off, sz, ncclChar, r, c->comm, stream));
        off += sz;
    }
    NCCL_TRY(nccl().GroupEnd());
    off = 0;
    for (int r = 0; r < n; r++) {
        const int rows = rtb::shard_rows(H, r, n, G);
        if (!rows) continue;
        const void* src = (r == root) ? band : (const void*)(c->staging + off);
        CUDA_TRY(rtb::scatter_band(frame, src, line, r * G, n * G, G, rows, cudaMemcpyDeviceToDevice, stream));
        if (r != root) off += (size_t)rows * line;
    }
    return RT_OK;
}

} /* extern "C" */
