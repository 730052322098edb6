// Batch split / merge across GPUs (one process per GPU) over NCCL -- NVLink 5 / NVSwitch on the
// 8 x B200 box.  The transforms themselves have no exchange step (every clip is independent,
// SURVEY.md section 8e), so the only collectives on the path are the ones that move a batch:
//
//   zafb_dist_scatter_rows   root holds n_rows rows; rank r receives rows [r n/R, (r+1) n/R)
//   zafb_dist_gather_rows    the inverse (merge the per-rank results on root)
//   zafb_dist_allgather_rows every rank ends up with all rows
//   zafb_dist_broadcast      tables / operators from root
//
// Rows are opaque byte strings (a clip, or a clip's spectrogram), so one set of entry points
// serves every transform.  Transfers are grouped ncclSend / ncclRecv on the caller's stream: they
// are stream-ordered with the kernels and overlap with compute on other streams.
//
// NCCL is resolved with dlopen at the first zafb_dist_* call, so libzafb200.so itself has no
// link-time dependency on it and the single-GPU path never loads it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <cstring>
#include <mutex>

#include "common.cuh"

using namespace zafb;

struct zafb_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double* d_scalar = nullptr;  // device word of zafb_dist_max_f64 (allocated once: cudaMalloc / cudaFree synchronise the device)
};

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    std::string error;
};

NcclApi& api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("ZAFB_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        // a copy the process already holds (e.g. the one a launcher loaded) is reused before a second one is mapped
        for (const char* n : names)
            if (n && *n && !a.handle) a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
        for (const char* n : names)
            if (n && *n && !a.handle) a.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (!a.handle) {
            const char* why = dlerror();  // one call: dlerror() clears the error state it returns
            a.error = std::string("cannot load libnccl.so.2: ") + (why ? why : "unknown error");
            return;
        }
        bool ok = true;
        auto sym = [&](auto& fn, const char* name) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(a.handle, name));
            if (!fn) {
                ok = false;
                a.error = std::string("libnccl is missing the symbol ") + name;
            }
        };
        sym(a.GetVersion, "ncclGetVersion");
        sym(a.GetUniqueId, "ncclGetUniqueId");
        sym(a.CommInitRank, "ncclCommInitRank");
        sym(a.CommDestroy, "ncclCommDestroy");
        sym(a.GetErrorString, "ncclGetErrorString");
        sym(a.Send, "ncclSend");
        sym(a.Recv, "ncclRecv");
        sym(a.Broadcast, "ncclBroadcast");
        sym(a.AllGather, "ncclAllGather");
        sym(a.AllReduce, "ncclAllReduce");
        sym(a.GroupStart, "ncclGroupStart");
        sym(a.GroupEnd, "ncclGroupEnd");
        if (!ok) {
            dlclose(a.handle);
            a.handle = nullptr;
        }
    });
    return a;
}

#define ZAFB_NCCL_READY()                                                         \
    NcclApi& nc = api();                                                          \
    if (!nc.handle) return fail(ZAFB_E_NCCL, "NCCL unavailable: %s", nc.error.c_str())

#define ZAFB_NCCL(expr)                                                           \
    do {                                                                          \
        ncclResult_t _r = (expr);                                                 \
        if (_r != ncclSuccess)                                                    \
            return fail(ZAFB_E_NCCL, "%s failed: %s (%s:%d)", #expr, nc.GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

// Inside ncclGroupStart / ncclGroupEnd an early return would leave the group open: the first failure is remembered,
// the remaining calls are skipped and the group is always closed before the error is reported.
#define ZAFB_NCCL_IN_GROUP(expr)                                                  \
    do {                                                                          \
        if (_grp == ncclSuccess) {                                                \
            _grp = (expr);                                                        \
            if (_grp != ncclSuccess) _grp_what = #expr;                           \
        }                                                                         \
    } while (0)

#define ZAFB_NCCL_GROUP_BEGIN()                                                   \
    ncclResult_t _grp = ncclSuccess;                                              \
    const char* _grp_what = "";                                                   \
    ZAFB_NCCL(nc.GroupStart())

#define ZAFB_NCCL_GROUP_END()                                                     \
    do {                                                                          \
        ncclResult_t _e = nc.GroupEnd();                                          \
        if (_grp != ncclSuccess)                                                  \
            return fail(ZAFB_E_NCCL, "%s failed: %s (%s:%d)", _grp_what, nc.GetErrorString(_grp), __FILE__, __LINE__); \
        if (_e != ncclSuccess)                                                    \
            return fail(ZAFB_E_NCCL, "ncclGroupEnd failed: %s (%s:%d)", nc.GetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

inline int64_t shard_begin(int64_t n, int r, int world) { return (int64_t(r) * n) / world; }

}  // namespace

extern "C" {

int zafb_dist_shard_range(int64_t n_rows, int rank, int world, int64_t* begin, int64_t* end) {
    ZAFB_REQUIRE(world >= 1 && rank >= 0 && rank < world && n_rows >= 0, "bad rank %d / world %d / n_rows %lld", rank, world,
                 (long long)n_rows);
    if (begin) *begin = shard_begin(n_rows, rank, world);
    if (end) *end = shard_begin(n_rows, rank + 1, world);
    return ZAFB_OK;
}

int zafb_dist_nccl_version(int* version) {
    ZAFB_REQUIRE(version != nullptr, "version is NULL");
    ZAFB_NCCL_READY();
    ZAFB_NCCL(nc.GetVersion(version));
    return ZAFB_OK;
}

int zafb_dist_unique_id(void* id128) {
    ZAFB_REQUIRE(id128 != nullptr, "id buffer is NULL");
    ZAFB_NCCL_READY();
    ncclUniqueId id;
    ZAFB_NCCL(nc.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, sizeof(id));
    return ZAFB_OK;
}

int zafb_dist_init(zafb_comm** out, const void* id128, int rank, int world) {
    ZAFB_REQUIRE(out != nullptr && id128 != nullptr, "comm/id is NULL");
    ZAFB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    ZAFB_NCCL_READY();
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    zafb_comm* c = new zafb_comm();
    c->rank = rank;
    c->world = world;
    ncclResult_t r = nc.CommInitRank(&c->comm, world, id, rank);  // binds to the current device (zafb_init)
    if (r != ncclSuccess) {
        delete c;
        return fail(ZAFB_E_NCCL, "ncclCommInitRank(rank %d of %d) failed: %s", rank, world, nc.GetErrorString(r));
    }
    *out = c;
    return ZAFB_OK;
}

int zafb_dist_destroy(zafb_comm* c) {
    if (!c) return ZAFB_OK;
    NcclApi& nc = api();
    if (nc.handle && c->comm) nc.CommDestroy(c->comm);
    cudaFree(c->d_scalar);
    delete c;
    return ZAFB_OK;
}

int zafb_dist_broadcast(zafb_comm* c, void* buf, size_t bytes, int root, void* stream) {
    ZAFB_REQUIRE(c != nullptr && root >= 0 && root < c->world, "bad comm/root");
    if (bytes == 0) return ZAFB_OK;
    ZAFB_REQUIRE(buf != nullptr, "buf is NULL");
    ZAFB_NCCL_READY();
    ZAFB_NCCL(nc.Broadcast(buf, buf, bytes, ncclChar, root, c->comm, static_cast<cudaStream_t>(stream)));
    return ZAFB_OK;
}

int zafb_dist_scatter_rows(zafb_comm* c, const void* src_root, void* dst, int64_t n_rows, int64_t row_bytes, int root,
                           void* stream) {
    ZAFB_REQUIRE(c != nullptr && root >= 0 && root < c->world, "bad comm/root");
    ZAFB_REQUIRE(n_rows >= 0 && row_bytes >= 0, "bad row geometry");
    ZAFB_NCCL_READY();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t b = shard_begin(n_rows, c->rank, c->world), e = shard_begin(n_rows, c->rank + 1, c->world);
    const size_t mine = size_t(e - b) * size_t(row_bytes);
    ZAFB_REQUIRE(mine == 0 || dst != nullptr, "dst is NULL");
    if (c->rank == root) {
        ZAFB_REQUIRE(n_rows * row_bytes == 0 || src_root != nullptr, "src is NULL on root");
        const char* s = static_cast<const char*>(src_root);
        ZAFB_NCCL_GROUP_BEGIN();
        for (int r = 0; r < c->world; ++r) {
            if (r == root) continue;
            const int64_t rb = shard_begin(n_rows, r, c->world), re = shard_begin(n_rows, r + 1, c->world);
            const size_t bytes = size_t(re - rb) * size_t(row_bytes);
            if (bytes) ZAFB_NCCL_IN_GROUP(nc.Send(s + size_t(rb) * row_bytes, bytes, ncclChar, r, c->comm, st));
        }
        ZAFB_NCCL_GROUP_END();
        if (mine && dst != s + size_t(b) * row_bytes)
            ZAFB_CUDA(cudaMemcpyAsync(dst, s + size_t(b) * row_bytes, mine, cudaMemcpyDeviceToDevice, st));
    } else if (mine) {
        ZAFB_NCCL(nc.Recv(dst, mine, ncclChar, root, c->comm, st));
    }
    return ZAFB_OK;
}

int zafb_dist_gather_rows(zafb_comm* c, const void* src, void* dst_root, int64_t n_rows, int64_t row_bytes, int root,
                          void* stream) {
    ZAFB_REQUIRE(c != nullptr && root >= 0 && root < c->world, "bad comm/root");
    ZAFB_REQUIRE(n_rows >= 0 && row_bytes >= 0, "bad row geometry");
    ZAFB_NCCL_READY();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t b = shard_begin(n_rows, c->rank, c->world), e = shard_begin(n_rows, c->rank + 1, c->world);
    const size_t mine = size_t(e - b) * size_t(row_bytes);
    ZAFB_REQUIRE(mine == 0 || src != nullptr, "src is NULL");
    if (c->rank == root) {
        ZAFB_REQUIRE(n_rows * row_bytes == 0 || dst_root != nullptr, "dst is NULL on root");
        char* d = static_cast<char*>(dst_root);
        ZAFB_NCCL_GROUP_BEGIN();
        for (int r = 0; r < c->world; ++r) {
            if (r == root) continue;
            const int64_t rb = shard_begin(n_rows, r, c->world), re = shard_begin(n_rows, r + 1, c->world);
            const size_t bytes = size_t(re - rb) * size_t(row_bytes);
            if (bytes) ZAFB_NCCL_IN_GROUP(nc.Recv(d + size_t(rb) * row_bytes, bytes, ncclChar, r, c->comm, st));
        }
        ZAFB_NCCL_GROUP_END();
        if (mine && src != d + size_t(b) * row_bytes)
            ZAFB_CUDA(cudaMemcpyAsync(d + size_t(b) * row_bytes, src, mine, cudaMemcpyDeviceToDevice, st));
    } else if (mine) {
        ZAFB_NCCL(nc.Send(src, mine, ncclChar, root, c->comm, st));
    }
    return ZAFB_OK;
}

int zafb_dist_allgather_rows(zafb_comm* c, const void* src, void* dst, int64_t n_rows, int64_t row_bytes, void* stream) {
    ZAFB_REQUIRE(c != nullptr, "comm is NULL");
    ZAFB_REQUIRE(n_rows >= 0 && row_bytes >= 0, "bad row geometry");
    if (n_rows * row_bytes == 0) return ZAFB_OK;
    ZAFB_REQUIRE(dst != nullptr, "dst is NULL");
    ZAFB_NCCL_READY();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* d = static_cast<char*>(dst);
    if (n_rows % c->world == 0) {  // equal shards: the native collective
        const size_t bytes = size_t(n_rows / c->world) * size_t(row_bytes);
        ZAFB_NCCL(nc.AllGather(src, d, bytes, ncclChar, c->comm, st));
        return ZAFB_OK;
    }
    // ragged shards: one broadcast per owner inside a group
    const int64_t b = shard_begin(n_rows, c->rank, c->world), e = shard_begin(n_rows, c->rank + 1, c->world);
    if (e > b && src != d + size_t(b) * row_bytes)
        ZAFB_CUDA(cudaMemcpyAsync(d + size_t(b) * row_bytes, src, size_t(e - b) * row_bytes, cudaMemcpyDeviceToDevice, st));
    ZAFB_NCCL_GROUP_BEGIN();
    for (int r = 0; r < c->world; ++r) {
        const int64_t rb = shard_begin(n_rows, r, c->world), re = shard_begin(n_rows, r + 1, c->world);
        const size_t bytes = size_t(re - rb) * size_t(row_bytes);
        if (bytes) ZAFB_NCCL_IN_GROUP(nc.Broadcast(d + size_t(rb) * row_bytes, d + size_t(rb) * row_bytes, bytes, ncclChar, r, c->comm, st));
    }
    ZAFB_NCCL_GROUP_END();
    return ZAFB_OK;
}

// max over ranks of a device-timed duration (bench: "time every multi-GPU number on the device
// as the max over ranks"); `value` is a host double, the reduction runs on `stream` and synchronises it.
int zafb_dist_max_f64(zafb_comm* c, double* value, void* stream) {
    ZAFB_REQUIRE(c != nullptr && value != nullptr, "comm/value is NULL");
    ZAFB_NCCL_READY();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (c->d_scalar == nullptr) ZAFB_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_scalar), sizeof(double)));
    double* d = c->d_scalar;
    cudaError_t e = cudaMemcpyAsync(d, value, sizeof(double), cudaMemcpyHostToDevice, st);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess) r = nc.AllReduce(d, d, 1, ncclDouble, ncclMax, c->comm, st);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpyAsync(value, d, sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaStreamSynchronize(st);
    if (r != ncclSuccess) return fail(ZAFB_E_NCCL, "ncclAllReduce failed: %s", nc.GetErrorString(r));
    if (e != cudaSuccess) return fail(ZAFB_E_CUDA, "max over ranks failed: %s", cudaGetErrorString(e));
    return ZAFB_OK;
}

// ------------------------------------------------------------------ peer memory (same node, one process per GPU)
// The merge without a separate collective: the root exports its result buffer, every peer maps it
// (CUDA IPC; NVLink peer access is enabled lazily by the driver) and passes the mapped address as the
// OUTPUT pointer of its transform -- the kernel's own stores travel over NVLink / NVSwitch, so the
// transfer overlaps the math frame by frame and the spectrum never touches the sender's HBM.
int zafb_dist_peer_export(const void* dev_ptr, void* handle64) {
    ZAFB_REQUIRE(dev_ptr != nullptr && handle64 != nullptr, "pointer/handle is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    ZAFB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle64, &h, sizeof(h));
    return ZAFB_OK;
}

int zafb_dist_peer_open(const void* handle64, void** mapped) {
    ZAFB_REQUIRE(handle64 != nullptr && mapped != nullptr, "handle/mapped is NULL");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    ZAFB_CUDA(cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess));
    return ZAFB_OK;
}

int zafb_dist_peer_close(void* mapped) {
    if (!mapped) return ZAFB_OK;
    ZAFB_CUDA(cudaIpcCloseMemHandle(mapped));
    return ZAFB_OK;
}

int zafb_dist_rank(const zafb_comm* c, int* rank, int* world) {
    ZAFB_REQUIRE(c != nullptr, "comm is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return ZAFB_OK;
}

}  // extern "C"
