// common.cuh -- context, error plumbing and stream-ordered scratch memory shared by all stages.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/ministark.h"

namespace ms {

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
    size_t hstage_bytes = 0;
    unsigned long long launches = 0;     // kernels launched by this library (bench gpu_launches)
    // optional per-kernel device timing (bench.py roofline): event pairs collected by ms_profile_collect
    bool profile = false;
    struct ProfEntry { const char* name; cudaEvent_t a, b; };
    std::vector<ProfEntry> prof;
    // transcript switches (host prover)
    uint8_t bridge_masks[3] = {0x00, 0x01, 0x02};
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

// Stream-ordered temporary (cudaMallocAsync pool; release threshold is raised at ctx creation so
// steady-state allocations never reach the driver).
struct Scratch {
    Ctx* c;
    void* p = nullptr;
    Scratch(Ctx* ctx) : c(ctx) {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    Scratch(Scratch&& o) noexcept : c(o.c), p(o.p) { o.p = nullptr; }
    int alloc(size_t bytes) {
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocAsync(&p, bytes, c->stream);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(c, MS_ERR_CUDA, "cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
        }
        return MS_OK;
    }
    ~Scratch() {
        if (p) cudaFreeAsync(p, c->stream);
    }
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

__global__ void k_store_words(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
// queue "copy `bytes` (multiple of 4) from device memory to offset `off` of the mapped staging buffer" on the
// stream; the host reads c->hstage + off after synchronising the stream
inline int stage_to_host(Ctx* c, size_t off, const void* d_src, size_t bytes) {
    if (!c->hstage) {
        const size_t cap = 8u << 20;
        MS_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&c->hstage), cap, cudaHostAllocMapped));
        MS_CUDA(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->hstage_dev), c->hstage, 0));
        c->hstage_bytes = cap;
    }
    if (off + bytes > c->hstage_bytes || (bytes & 3) || (off & 3)) return fail(c, MS_ERR_UNSUPPORTED, "staging buffer too small (%zu + %zu bytes)", off, bytes);
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
