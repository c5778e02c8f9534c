// trace.cuh -- trace generation on the device: the N x W trace of TraceTable::new + add_row
// (src/air.rs:73-112) written straight into the column-major layout the prover reads, so that a synthetic or
// recurrent AIR never crosses the PCIe bus (8 GiB at 2^24 x 64).
//   trace_synth       the synthetic benchmark trace of SURVEY.md 8d: splitmix64(seed ^ (row*W + col)) mod p
//   trace_recurrence  rows [0, steps): row_{i+1} = M * row_i (a W x W matrix over the base field: the Fibonacci AIR of
//                     tests/e2e_goldilocks.rs:36-41 is M = [[0,1,0],[0,0,1],[0,1,1]]); rows [steps, N): the padding
//                     value in every cell (src/air.rs:80-82: one fresh test_rng per cell, so all cells are equal)
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace ms {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

template <class F>
__global__ void k_trace_synth(typename F::T* __restrict__ out, uint64_t n, uint64_t w, uint64_t seed) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // column-major index: col * n + row
    if (i >= n * w) return;
    const uint64_t col = i / n, row = i % n;
    out[i] = (typename F::T)(splitmix64((row * w + col) ^ seed) % (uint64_t)F::P);
}
template <class F>
int trace_synth(Ctx* c, uint64_t seed, uint64_t n, uint64_t w, typename F::T* d_out_cm) {
    if (n == 0 || w == 0) return MS_OK;
    k_trace_synth<F><<<(unsigned)((n * w + 255) / 256), 256, 0, c->stream>>>(d_out_cm, n, w, seed);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

constexpr int TRACE_CHUNK = 64;   // rows a thread produces from its chunk's first row
constexpr int TRACE_MAXW = 16;
// starts: [chunks][W] first row of every chunk (host: M^CHUNK applied repeatedly); mat: W x W row-major
template <class F>
__global__ void k_trace_recurrence(typename F::T* __restrict__ out, uint64_t n, int w, uint64_t steps, const typename F::T* __restrict__ starts,
                                   const typename F::T* __restrict__ mat, typename F::T padding) {
    using T = typename F::T;
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t r0 = chunk * TRACE_CHUNK;
    if (r0 >= n) return;
    T row[TRACE_MAXW], nxt[TRACE_MAXW];
    for (int j = 0; j < w; j++) row[j] = starts[chunk * w + j];
    for (int i = 0; i < TRACE_CHUNK && r0 + i < n; i++) {
        const bool live = r0 + i < steps;
        for (int j = 0; j < w; j++) out[(uint64_t)j * n + r0 + i] = live ? row[j] : padding;
        for (int a = 0; a < w; a++) {
            T acc = 0;
            for (int b = 0; b < w; b++) {
                const T m = mat[a * w + b];
                if (m) acc = F::add(acc, m == 1 ? row[b] : F::mul(m, row[b]));
            }
            nxt[a] = acc;
        }
        for (int j = 0; j < w; j++) row[j] = nxt[j];
    }
}
template <class F>
int trace_recurrence(Ctx* c, const typename F::T* mat_host, const typename F::T* row0_host, uint64_t w, uint64_t steps, uint64_t n,
                     typename F::T padding, typename F::T* d_out_cm) {
    using T = typename F::T;
    if (w == 0 || w > TRACE_MAXW) return fail(c, MS_ERR_UNSUPPORTED, "recurrence width %llu (max %d)", (unsigned long long)w, TRACE_MAXW);
    if (!is_pow2(n) || steps > n) return fail(c, MS_ERR_BAD_SHAPE, "trace rows %llu must be a power of two >= steps (air.rs:74)", (unsigned long long)n);
    std::vector<T> m(mat_host, mat_host + w * w), mc(w * w), tmp(w * w);
    for (auto& v : m) v = (T)((uint64_t)v % (uint64_t)F::P);
    // mc = M^TRACE_CHUNK by repeated squaring (TRACE_CHUNK is a power of two)
    mc = m;
    for (int s = 1; s < TRACE_CHUNK; s <<= 1) {
        for (uint64_t a = 0; a < w; a++)
            for (uint64_t b = 0; b < w; b++) {
                T acc = 0;
                for (uint64_t k = 0; k < w; k++) acc = F::add(acc, F::mul(mc[a * w + k], mc[k * w + b]));
                tmp[a * w + b] = acc;
            }
        mc = tmp;
    }
    const uint64_t chunks = (n + TRACE_CHUNK - 1) / TRACE_CHUNK;
    std::vector<T> starts(chunks * w);
    for (uint64_t j = 0; j < w; j++) starts[j] = (T)((uint64_t)row0_host[j] % (uint64_t)F::P);
    for (uint64_t k = 1; k < chunks; k++)
        for (uint64_t a = 0; a < w; a++) {
            T acc = 0;
            for (uint64_t b = 0; b < w; b++) acc = F::add(acc, F::mul(mc[a * w + b], starts[(k - 1) * w + b]));
            starts[k * w + a] = acc;
        }
    Scratch ds(c), dm(c);
    MS_TRY(ds.alloc(starts.size() * sizeof(T)));
    MS_TRY(dm.alloc(m.size() * sizeof(T)));
    MS_CUDA(c, cudaMemcpyAsync(ds.p, starts.data(), starts.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    MS_CUDA(c, cudaMemcpyAsync(dm.p, m.data(), m.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    k_trace_recurrence<F><<<(unsigned)((chunks + 127) / 128), 128, 0, c->stream>>>(d_out_cm, n, (int)w, steps, ds.as<T>(), dm.as<T>(),
                                                                                  (T)((uint64_t)padding % (uint64_t)F::P));
    MS_LAUNCH_CHECK(c);
    MS_CUDA(c, cudaStreamSynchronize(c->stream));  // the host vectors above back the async copies
    return MS_OK;
}

}  // namespace ms
