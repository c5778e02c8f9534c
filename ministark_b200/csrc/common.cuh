// common.cuh -- context, error plumbing and stream-ordered scratch memory shared by all stages.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/ministark.h"

namespace ms {

struct Comm;  // comm.cuh

struct Ctx {
    int field = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    cudaStream_t copy_stream = nullptr;  // proof download (overlaps the query-phase kernels)
    cudaEvent_t copy_event = nullptr;
    std::string err;
    int zero_display_empty = 0;  // ark-ff 0.4 printed "" for zero; 0.5.0 prints "0" 
    void* wtab[2] = {nullptr, nullptr};  // plain DIT twiddles, forward / inverse (see ntt.cuh)
    void* tw16_plain[2][16] = {};        // plain block-twiddle tables per direction and transform size (ntt.cuh)
    uint32_t* dec4 = nullptr;            // ASCII of 4-digit groups (see merkle.cuh)
    // small results that the host needs while the copy engine is busy with the proof download: kernels store
    // them straight into mapped pinned memory (no D2H copy that would queue behind the multi-MB transfers)
    uint8_t* hstage = nullptr;           // host address
    uint8_t* hstage_dev = nullptr;       // the same memory as seen from the device
    size_t hstage_bytes = 0;             // [0, HSTAGE_OUT): device -> host results; the rest: host -> device inputs
    size_t hstage_in_used = 0;           // bump allocator over the input region, reset at stream syncs it forces
    unsigned long long launches = 0;     // kernels launched by this library (bench gpu_launches)
    // optional per-kernel device timing (bench.py roofline): event pairs collected by ms_profile_collect
    bool profile = false;
    struct ProfEntry { const char* name; cudaEvent_t a, b; };
    std::vector<ProfEntry> prof;
    // transcript switches (host prover)
    uint8_t bridge_masks[3] = {0x00, 0x01, 0x02};
    int leftover_as_published = 1;  // nimue DigestBridge leftovers branch as published (see transcript.hpp)
    // multi-GPU: the communicator this context is a rank of (nullptr: single GPU) and which stages shard
    Comm* comm = nullptr;
    int shard_mask = MS_SHARD_ALL;
    struct NttTables {  // a two-pass transform's inter-pass factor table and coset block twiddles (ntt.cuh lde_batch)
        void* ft = nullptr;
        void* tw = nullptr;
        size_t ft_bytes = 0, tw_bytes = 0;
        int logN = -1, logB = 0, inverse = 0, tile = 0, field = -1;
        uint64_t shift = 0;
        uint64_t stamp = 0;  // last use (least recently used entries leave first)
    };
    // One entry per (n, blowup, shift, direction, tile size): a proof runs the trace iNTT, the LDE and one codeword transform
    // per FRI round, each with its own tables, and the FRI ones (shift 1) are the same for every proof of a shape.
    std::vector<NttTables> ntt_tables;
    uint64_t ntt_tables_clock = 0;
    size_t ntt_tables_budget = (size_t)4 << 30;  // bytes kept across calls (MINISTARK_NTT_TABLE_MB); one entry always stays
    // Device workspace: blocks handed out by Scratch come from, and go back to, this per-context cache (see Scratch).
    std::multimap<size_t, void*> block_cache;
    size_t cached_bytes = 0, cache_misses = 0;
    int dl_max_ranks = 0;   // sharded proof download: at most this many ranks take a share (0: all; MINISTARK_DL_MAX_RANKS)
    int dl_skip_rank0 = 0;  // sharded proof download: leave rank 0 out of the split from 4 ranks on (MINISTARK_DL_SKIP_RANK0)
    int lde_linearity = 1;  // prover: constraint columns of the LDE by linearity when the constraint matrix is sparse (prover.cuh)
    int ntt_log_tile = 13;  // log2 elements of an NTT tile (ntt.cuh NTT_LOG_TILE_PREF); MINISTARK_NTT_TILE=12 selects half tiles
};

inline int fail(Ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define MS_CUDA(c, expr)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return ms::fail((c), MS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                                 \
    } while (0)

#define MS_TRY(expr)                  \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != MS_OK) return rc__; \
    } while (0)

#define MS_LAUNCH_CHECK(c)                                                                   \
    do {                                                                                     \
        (c)->launches++;                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return ms::fail((c), MS_ERR_CUDA, "kernel launch failed: %s (%s:%d)",            \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)

// Device temporaries.  Every kernel of a context runs on its one stream, so a block can be handed to its next user the
// moment the previous one lets go of it (the stream orders the two uses): Scratch takes blocks from a per-context cache
// (best fit within 25 %, sizes rounded to 512 B / 2 MiB) and returns them there, and cudaMalloc only runs the first time a
// shape is seen.  (r01 used the driver's stream-ordered pool, cudaMallocAsync: its remapping of freed ranges between
// differently sized requests cost up to hundreds of ms per proof whenever the allocation pattern changed -- 2^24 x 64 proofs
// took 327 ms in a fresh process and 1044 ms after other shapes had run, profiles/r02_d_bench_1gpu.json.)
// The copy stream only ever reads blocks that stay alive until the call that queued the copies has synchronised it.
inline size_t block_round(size_t bytes) {
    if (bytes == 0) bytes = 16;
    return bytes < (1u << 20) ? (bytes + 511) & ~(size_t)511 : (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
}
inline void block_cache_release(Ctx* c) {  // give everything back to the driver (ctx teardown, or out of memory)
    for (auto& kv : c->block_cache) cudaFree(kv.second);
    c->block_cache.clear();
    c->cached_bytes = 0;
}
inline int block_alloc(Ctx* c, size_t bytes, void** p, size_t* got) {
    const size_t want = block_round(bytes);
    auto it = c->block_cache.lower_bound(want);
    if (it != c->block_cache.end() && it->first <= want + want / 4) {
        *p = it->second;
        *got = it->first;
        c->cached_bytes -= it->first;
        c->block_cache.erase(it);
        return MS_OK;
    }
    c->cache_misses++;
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {  // out of memory: drop the cache and try once more
        cudaGetLastError();
        cudaStreamSynchronize(c->stream);
        block_cache_release(c);
        e = cudaMalloc(p, want);
    }
    if (e != cudaSuccess) {
        *p = nullptr;
        return fail(c, MS_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    *got = want;
    return MS_OK;
}
inline void block_free(Ctx* c, void* p, size_t size) {
    if (!p) return;
    c->block_cache.emplace(size, p);
    c->cached_bytes += size;
}
struct Scratch {
    Ctx* c;
    void* p = nullptr;
    size_t size = 0;
    Scratch(Ctx* ctx) : c(ctx) {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    Scratch(Scratch&& o) noexcept : c(o.c), p(o.p), size(o.size) { o.p = nullptr; }
    int alloc(size_t bytes) {
        release();
        return block_alloc(c, bytes, &p, &size);
    }
    void release() {
        if (p) block_free(c, p, size);
        p = nullptr;
    }
    ~Scratch() { release(); }
    template <class T>
    T* as() { return reinterpret_cast<T*>(p); }
};

inline void prof_begin(Ctx* c, const char* name) {
    if (!c->profile) return;
    Ctx::ProfEntry e{name, nullptr, nullptr};
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, c->stream);
    c->prof.push_back(e);
}
inline void prof_end(Ctx* c) {
    if (!c->profile || c->prof.empty()) return;
    cudaEventRecord(c->prof.back().b, c->stream);
}

constexpr size_t HSTAGE_OUT = 6u << 20, HSTAGE_TOTAL = 12u << 20;
__global__ void k_store_words(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
// queue "copy `bytes` (multiple of 4) from device memory to offset `off` of the mapped staging buffer" on the
// stream; the host reads c->hstage + off after synchronising the stream
inline int ensure_hstage(Ctx* c) {
    if (c->hstage) return MS_OK;
    MS_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&c->hstage), HSTAGE_TOTAL, cudaHostAllocMapped));
    MS_CUDA(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->hstage_dev), c->hstage, 0));
    c->hstage_bytes = HSTAGE_TOTAL;
    return MS_OK;
}
// Small host -> device inputs (index lists, scan multipliers) the same way in the other direction: the host
// writes them into the mapped buffer and a kernel copies them to `d_dst`, so they never wait on a copy engine
// that is busy with bulk transfers.  Chunks are handed out by a bump allocator; when the region is full the
// stream is drained first (everything queued so far has then consumed its chunk).
inline int stage_from_host(Ctx* c, const void* h_src, size_t bytes, void* d_dst) {
    MS_TRY(ensure_hstage(c));
    if (bytes == 0) return MS_OK;
    const size_t need = (bytes + 15) & ~(size_t)15, cap = c->hstage_bytes - HSTAGE_OUT;
    if ((bytes & 3) || need > cap) return fail(c, MS_ERR_UNSUPPORTED, "staged input of %zu bytes", bytes);
    if (c->hstage_in_used + need > cap) {
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        c->hstage_in_used = 0;
    }
    const size_t off = HSTAGE_OUT + c->hstage_in_used;
    c->hstage_in_used += need;
    memcpy(c->hstage + off, h_src, bytes);
    const uint64_t n = bytes / 4;
    k_store_words<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<uint32_t*>(d_dst),
                                                                     reinterpret_cast<const uint32_t*>(c->hstage_dev + off), n);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}
inline int stage_to_host(Ctx* c, size_t off, const void* d_src, size_t bytes) {
    MS_TRY(ensure_hstage(c));
    if (off + bytes > HSTAGE_OUT || (bytes & 3) || (off & 3)) return fail(c, MS_ERR_UNSUPPORTED, "staging buffer too small (%zu + %zu bytes)", off, bytes);
    const uint64_t n = bytes / 4;
    if (n == 0) return MS_OK;
    k_store_words<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<uint32_t*>(c->hstage_dev + off),
                                                                     reinterpret_cast<const uint32_t*>(d_src), n);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

inline int ilog2(uint64_t v) {
    int l = 0;
    while ((1ULL << l) < v) l++;
    return l;
}
inline bool is_pow2(uint64_t v) { return v && !(v & (v - 1)); }

}  // namespace ms
