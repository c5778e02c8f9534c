// comm.cuh -- the multi-GPU plumbing of the prover, inside the library (no reference call site: the
// reference is single-threaded).  One prover replica per GPU; the stages that shard exchange data in two ways only:
//   * a symmetric, peer-readable ARENA per rank: every rank's kernels read the other ranks' arena directly over
//     NVLink (column-sharded coefficients / LDE columns feeding row-sharded leaf hashing, Merkle subtree nodes
//     feeding the authentication paths) -- the transfer happens inside the consuming kernel, no exchange pass;
//   * tiny collectives for ordering and for digests: barrier (stream-ordered) and all-gather (32 bytes x ranks).
// Two backends behind one interface:
//   NcclComm   one process per GPU.  NCCL (dlopen'ed: libnccl.so.2, the copy already in the process if the host
//              loaded one) for barrier / all-gather, CUDA IPC for the arenas.  Bootstrapped by a 128-byte
//              ncclUniqueId that the host distributes by any means (ms_comm_unique_id / ms_comm_init_nccl).
//   LocalComm  one process, one host thread per rank (ms_comm_init_local): host-side rendezvous, plain device
//              pointers (peer access enabled between different devices).  Also runs G "virtual ranks" on ONE
//              GPU, which is how the 1-GPU test box exercises every sharded code path.
#pragma once
#include <dlfcn.h>

#include <condition_variable>
#include <memory>
#include <mutex>

#include "common.cuh"

namespace ms {

struct Comm {
    int rank = 0, world = 1;
    void* arena = nullptr;          // this rank's peer-readable buffer
    size_t arena_bytes = 0;
    std::vector<void*> bases;       // bases[g]: rank g's arena as seen from this rank's kernels
    virtual ~Comm() {}
    virtual const char* backend() const = 0;
    // after this returns, work queued later on c->stream sees everything every rank queued on its stream before its own call
    virtual int barrier(Ctx* c) = 0;
    // d_recv[g*bytes ..] = rank g's d_send[0..bytes); stream-ordered like barrier
    virtual int all_gather(Ctx* c, const void* d_send, void* d_recv, size_t bytes) = 0;
    // every rank has drained its stream and reached this point on the host
    virtual int host_barrier(Ctx* c) = 0;
    // collective: make sure every rank's arena holds at least `bytes` (all ranks pass the same value)
    virtual int ensure_arena(Ctx* c, size_t bytes) = 0;
    virtual int release(Ctx* c) = 0;  // collective teardown of the arena mappings
};

// ---------------------------------------------------------------------------------------------- NCCL (dlopen)
typedef struct ncclComm* ms_ncclComm_t;
struct ms_ncclUniqueId { char internal[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ms_ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ms_ncclComm_t*, int, ms_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ms_ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ms_ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ms_ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        const char* env = getenv("MINISTARK_NCCL_LIB");  // when set, the only candidate
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        if (env && *env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        else
            for (const char* n : names) {
                lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                if (lib) break;
            }
        if (!lib) { err = std::string("NCCL library not loadable: ") + (env && *env ? env : "libnccl.so.2") + " (MINISTARK_NCCL_LIB overrides)"; return false; }
#define MS_SYM(field, name)                                           \
    *reinterpret_cast<void**>(&field) = dlsym(lib, name);             \
    if (!field) { err = std::string("NCCL symbol missing: ") + name; lib = nullptr; return false; }
        MS_SYM(GetUniqueId, "ncclGetUniqueId");
        MS_SYM(CommInitRank, "ncclCommInitRank");
        MS_SYM(CommDestroy, "ncclCommDestroy");
        MS_SYM(AllGather, "ncclAllGather");
        MS_SYM(AllReduce, "ncclAllReduce");
        MS_SYM(GetErrorString, "ncclGetErrorString");
        MS_SYM(GetVersion, "ncclGetVersion");
#undef MS_SYM
        return true;
    }
};
inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

#define MS_NCCL(c, expr)                                                                                   \
    do {                                                                                                   \
        int r__ = (expr);                                                                                  \
        if (r__ != 0)                                                                                      \
            return ms::fail((c), MS_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, nccl_api().GetErrorString(r__), \
                            __FILE__, __LINE__);                                                           \
    } while (0)

struct NcclComm : Comm {
    ms_ncclComm_t comm = nullptr;
    int* d_flag = nullptr;
    const char* backend() const override { return "nccl"; }
    ~NcclComm() override {
        if (comm) nccl_api().CommDestroy(comm);
        if (d_flag) cudaFree(d_flag);
    }
    int init(Ctx* c, const uint8_t id128[128], int rank_, int world_) {
        NcclApi& api = nccl_api();
        if (!api.load()) return fail(c, MS_ERR_NCCL, "%s", api.err.c_str());
        rank = rank_;
        world = world_;
        ms_ncclUniqueId id;
        memcpy(id.internal, id128, 128);
        MS_CUDA(c, cudaSetDevice(c->device));
        MS_NCCL(c, api.CommInitRank(&comm, world, id, rank));
        MS_CUDA(c, cudaMalloc(&d_flag, 2 * sizeof(int)));
        MS_CUDA(c, cudaMemset(d_flag, 0, 2 * sizeof(int)));
        return MS_OK;
    }
    int barrier(Ctx* c) override {
        MS_NCCL(c, nccl_api().AllReduce(d_flag, d_flag + 1, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, comm, c->stream));
        return MS_OK;
    }
    int all_gather(Ctx* c, const void* d_send, void* d_recv, size_t bytes) override {
        MS_NCCL(c, nccl_api().AllGather(d_send, d_recv, bytes, /*ncclInt8*/ 0, comm, c->stream));
        return MS_OK;
    }
    int host_barrier(Ctx* c) override {
        MS_TRY(barrier(c));
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        return MS_OK;
    }
    int release(Ctx* c) override {
        if (!arena) return MS_OK;
        MS_TRY(host_barrier(c));  // nobody reads any more
        for (int g = 0; g < world; g++)
            if (g != rank && bases[g]) cudaIpcCloseMemHandle(bases[g]);
        MS_TRY(host_barrier(c));  // every mapping is gone before the owner frees
        cudaFree(arena);
        arena = nullptr;
        arena_bytes = 0;
        bases.clear();
        return MS_OK;
    }
    int ensure_arena(Ctx* c, size_t bytes) override {
        if (arena && arena_bytes >= bytes) return MS_OK;
        MS_TRY(release(c));
        bytes = (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
        MS_CUDA(c, cudaSetDevice(c->device));
        MS_CUDA(c, cudaMalloc(&arena, bytes));
        arena_bytes = bytes;
        cudaIpcMemHandle_t mine;
        MS_CUDA(c, cudaIpcGetMemHandle(&mine, arena));
        Scratch ds(c), dr(c);
        MS_TRY(ds.alloc(64));
        MS_TRY(dr.alloc(64 * (size_t)world));
        MS_CUDA(c, cudaMemcpyAsync(ds.p, &mine, 64, cudaMemcpyHostToDevice, c->stream));
        MS_TRY(all_gather(c, ds.p, dr.p, 64));
        std::vector<cudaIpcMemHandle_t> all(world);
        MS_CUDA(c, cudaMemcpyAsync(all.data(), dr.p, 64 * (size_t)world, cudaMemcpyDeviceToHost, c->stream));
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        bases.assign(world, nullptr);
        for (int g = 0; g < world; g++) {
            if (g == rank) { bases[g] = arena; continue; }
            MS_CUDA(c, cudaIpcOpenMemHandle(&bases[g], all[g], cudaIpcMemLazyEnablePeerAccess));
        }
        return MS_OK;
    }
};

// ---------------------------------------------------------------------------------------------- local threads
struct LocalGroup {
    std::mutex mu;
    std::condition_variable cv;
    int world = 1, arrived = 0;
    uint64_t gen = 0;
    std::vector<const void*> slot;
    std::vector<int> device;
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        const uint64_t g = gen;
        if (++arrived == world) {
            arrived = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};
struct LocalComm : Comm {
    std::shared_ptr<LocalGroup> grp;
    const char* backend() const override { return "local"; }
    int barrier(Ctx* c) override {
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        grp->wait();
        return MS_OK;
    }
    int host_barrier(Ctx* c) override { return barrier(c); }
    int all_gather(Ctx* c, const void* d_send, void* d_recv, size_t bytes) override {
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        grp->slot[rank] = d_send;
        grp->wait();
        for (int g = 0; g < world; g++)
            MS_CUDA(c, cudaMemcpyAsync(static_cast<uint8_t*>(d_recv) + (size_t)g * bytes, grp->slot[g], bytes, cudaMemcpyDefault, c->stream));
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        grp->wait();
        return MS_OK;
    }
    int release(Ctx* c) override {
        if (!arena) return MS_OK;
        MS_TRY(barrier(c));
        cudaFree(arena);
        arena = nullptr;
        arena_bytes = 0;
        bases.clear();
        return MS_OK;
    }
    int ensure_arena(Ctx* c, size_t bytes) override {
        if (arena && arena_bytes >= bytes) return MS_OK;
        MS_TRY(release(c));
        bytes = (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
        MS_CUDA(c, cudaSetDevice(c->device));
        MS_CUDA(c, cudaMalloc(&arena, bytes));
        arena_bytes = bytes;
        grp->slot[rank] = arena;
        grp->wait();
        bases.assign(world, nullptr);
        for (int g = 0; g < world; g++) bases[g] = const_cast<void*>(grp->slot[g]);
        grp->wait();
        return MS_OK;
    }
};

// ---------------------------------------------------------------------------------------------- shard plans
// contiguous shares, as even as possible: the first `n % world` ranks get one more (sharded.py column_ranges)
inline void shard_range(uint64_t n, int world, int rank, uint64_t* a, uint64_t* b) {
    const uint64_t base = n / (uint64_t)world, extra = n % (uint64_t)world, r = (uint64_t)rank;
    *a = r * base + (r < extra ? r : extra);
    *b = *a + base + (r < extra ? 1 : 0);
}
inline int shard_owner(uint64_t n, int world, uint64_t i) {
    for (int g = 0; g < world; g++) {
        uint64_t a, b;
        shard_range(n, world, g, &a, &b);
        if (i >= a && i < b) return g;
    }
    return world - 1;
}
// How a tree with `groups` leaf groups and arity k splits over `world` ranks (sharded.py SubtreePlan): every rank
// hashes groups/world leaf groups and climbs while whole groups of k digests remain; `left` (< k, or 1) digests per
// rank are gathered and joined.  false: the split does not exist (the caller then builds the tree locally).
inline bool subtree_plan(uint64_t groups, uint64_t k, int world, uint64_t* per_rank, uint64_t* left) {
    if (world < 1 || groups == 0 || groups % (uint64_t)world) return false;
    const uint64_t per = groups / (uint64_t)world;
    if (!is_pow2(per) || !is_pow2(k) || k < 2) return false;
    uint64_t lv = per;
    while (lv > 1 && lv % k == 0) lv /= k;
    uint64_t t = (uint64_t)world * lv;
    while (t > 1) {
        if (t % k) return false;
        t /= k;
    }
    *per_rank = per;
    *left = lv;
    return true;
}

}  // namespace ms
