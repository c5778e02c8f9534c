// ntt.cuh -- batched radix-2 (i)NTT / coset low-degree extension over column-major matrices.
//
// One primitive serves trace interpolation (src/air.rs:147-160), the LDE loop
// (src/starks.rs:82-91) and every FRI codeword (src/fri.rs:345-351):
//
//   lde_batch:  out[c][B*k + j] = sum_{m<N} in[c][m] * (s_j * w^k)^m ,   s_j = shift * w_L^j ,
//               k in [0,N), j in [0,B)          (B = 1, shift = 1, w = w_N^-1, scale = 1/N: iNTT)
//
// i.e. the size-L evaluation on the coset is done as B coset transforms of size N (the zero padded
// size-L transform's first log2 B stages are pure replication).  N = n1 * n2 is split in at most
// two passes over HBM; inside a pass a CTA owns a shared-memory tile and runs decimation-in-time
// butterflies on it:
//
//   pass 1  (stages 0..a-1, n2 = 2^a):  tile = [n2 strided rows] x [R consecutive m1] x [B cosets].
//           The coset shift is folded into the stage twiddles t1[j][2^u+q] = s_j^(N/2^(u+1)) w_(2^(u+1))^q
//           (no separate "distribute powers" pass), inputs are replicated B times on load, and the
//           inter-pass factor ft = scale * (s_j w_N^k2)^m1 is applied on store.
//   pass 2  (stages a..logN-1, n1 = 2^b): tile = contiguous block of R2 * n1 * B elements of the
//           intermediate; plain twiddles; stores runs of R2*B consecutive output rows.
//
// Intermediate and output share one index map, ((m1 + n1*k2)*B + j), which degenerates to the
// natural output order when n1 = 1 (single pass).  Bit reversal is never a pass of its own: tiles
// are loaded into shared memory in bit-reversed transform order.
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace ms {

constexpr int NTT_MAXLOG = 14;       // largest in-tile transform (plain twiddle table size)
constexpr int NTT_LOG_TILE_PREF = 13;  // preferred tile: 8192 elements (64 KB Goldilocks)
constexpr int NTT_LOG_TILE_MAX = 14;   // 128 KB Goldilocks tile, one CTA per SM
constexpr int NTT_THREADS = 512;

__device__ __forceinline__ uint32_t brev_bits(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0; }

// plain DIT twiddles W[2^u + q] = g_(2^(u+1))^q, u < NTT_MAXLOG, g_(2^(u+1)) = (inverse) root of unity
template <class F>
__global__ void k_build_wtab(typename F::T* w, const typename F::T* roots /*[NTT_MAXLOG]*/) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (1 << NTT_MAXLOG)) return;
    if (idx == 0) { w[0] = 0; return; }
    int u = 31 - __clz(idx);
    int q = idx - (1 << u);
    w[idx] = fpow<F>(roots[u], (uint64_t)q);
}

// t1[j*n2 + 2^u + q] = sbase[j*a + u] * W[2^u + q]
template <class F>
__global__ void k_build_t1(typename F::T* t1, const typename F::T* wtab, const typename F::T* sbase, int a, int B) {
    int n2 = 1 << a;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n2 * B) return;
    int j = idx >> a, e = idx & (n2 - 1);
    if (e == 0) { t1[idx] = 0; return; }
    int u = 31 - __clz(e);
    t1[idx] = F::mul(sbase[j * a + u], wtab[e]);
}

// ft[((m1 + n1*k2) << logB) + j] = scale * s_j^m1 * wN^(m1*k2)
template <class F>
__global__ void k_build_ft(typename F::T* ft, const typename F::T* shifts /*[B]*/, typename F::T wN, typename F::T scale,
                           int a, int b, int logB, int chunk_log) {
    using T = typename F::T;
    const uint64_t n1 = 1ULL << b;
    const uint64_t lanes = n1 << logB;
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lane = gid % lanes, chunk = gid / lanes;
    uint64_t nchunks = (1ULL << a) >> chunk_log;
    if (chunk >= nchunks) return;
    uint64_t m1 = lane >> logB;
    int j = (int)(lane & ((1u << logB) - 1));
    uint64_t k2 = chunk << chunk_log;
    T rho = fpow<F>(wN, m1);
    T v = F::mul(F::mul(scale, fpow<F>(shifts[j], m1)), fpow<F>(rho, k2));
    for (uint64_t i = 0; i < (1ULL << chunk_log); i++, k2++) {
        ft[((m1 + n1 * k2) << logB) + j] = v;
        v = F::mul(v, rho);
    }
}

// shared-memory DIT stages [0, nst) on a tile laid out as S[(P << logRB) | beta]; the twiddle for
// butterfly (stage u, index q, batch beta) is tw[(beta & jmask) * tw_jstride + 2^u + q].
template <class F>
__device__ __forceinline__ void tile_dit(typename F::T* S, int nst, int logRB, const typename F::T* __restrict__ tw,
                                         uint32_t jmask, uint32_t tw_jstride) {
    using T = typename F::T;
    const uint32_t half = (1u << nst) >> 1;
    const uint32_t total = half << logRB;
    const uint32_t RBm = (1u << logRB) - 1;
    for (int u = 0; u < nst; u++) {
        __syncthreads();
        const uint32_t h = 1u << u;
        for (uint32_t x = threadIdx.x; x < total; x += blockDim.x) {
            uint32_t beta = x & RBm, bf = x >> logRB;
            uint32_t q = bf & (h - 1);
            uint32_t P = ((bf >> u) << (u + 1)) | q;
            uint32_t i0 = (P << logRB) | beta, i1 = i0 + (h << logRB);
            T w = __ldg(&tw[(beta & jmask) * tw_jstride + h + q]);
            T A = S[i0];
            T Bv = F::mul(S[i1], w);
            S[i0] = F::add(A, Bv);
            S[i1] = F::sub(A, Bv);
        }
    }
    __syncthreads();
}

template <class F>
__global__ void __launch_bounds__(NTT_THREADS)
k_lde_pass1(const typename F::T* __restrict__ in, uint64_t in_stride, typename F::T* __restrict__ out, uint64_t out_stride,
            const typename F::T* __restrict__ t1, const typename F::T* __restrict__ ft, int a, int b, int logB, int logR,
            typename F::T scale) {
    using T = typename F::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);
    const int logRB = logR + logB;
    const uint32_t n2 = 1u << a;
    const uint64_t n1 = 1ULL << b;
    const uint64_t m1_0 = (uint64_t)blockIdx.y << logR;
    const T* src = in + (uint64_t)blockIdx.x * in_stride;
    T* dst = out + (uint64_t)blockIdx.x * out_stride;
    const uint32_t tile = n2 << logRB;
    const uint32_t Rm = (1u << logR) - 1;
    // load: every coefficient is replicated over the B cosets (trivial first log2 B DIT levels);
    // lanes of one replication group read the same address (single broadcast transaction)
    for (uint32_t s = threadIdx.x; s < tile; s += blockDim.x) {
        uint32_t r = (s >> logB) & Rm, P = s >> logRB;
        uint64_t m2 = brev_bits(P, a);
        S[s] = src[m1_0 + r + n1 * m2];
    }
    tile_dit<F>(S, a, logRB, t1, (1u << logB) - 1, n2);
    const bool has_ft = ft != nullptr;
    for (uint32_t s = threadIdx.x; s < tile; s += blockDim.x) {
        uint32_t beta = s & ((1u << logRB) - 1);
        uint64_t k2 = s >> logRB;
        uint64_t g = ((m1_0 + n1 * k2) << logB) + beta;
        T v = S[s];
        if (has_ft) v = F::mul(v, __ldg(&ft[g]));
        else if (scale != 1) v = F::mul(v, scale);
        dst[g] = v;
    }
}

template <class F>
__global__ void __launch_bounds__(NTT_THREADS)
k_lde_pass2(const typename F::T* __restrict__ in, uint64_t in_stride, typename F::T* __restrict__ out, uint64_t out_stride,
            const typename F::T* __restrict__ wtab, int a, int b, int logB, int logR) {
    using T = typename F::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);
    const int logRB = logR + logB;
    const uint64_t n1 = 1ULL << b, n2 = 1ULL << a;
    const uint64_t k2_0 = (uint64_t)blockIdx.y << logR;
    const T* src = in + (uint64_t)blockIdx.x * in_stride + ((n1 * k2_0) << logB);
    T* dst = out + (uint64_t)blockIdx.x * out_stride;
    const uint32_t tile = (uint32_t)(n1 << logRB);
    const uint32_t Bm = (1u << logB) - 1, RBm = (1u << logRB) - 1;
    for (uint32_t s = threadIdx.x; s < tile; s += blockDim.x) {
        uint32_t beta = s & RBm, P = s >> logRB;
        uint32_t r2 = beta >> logB, j = beta & Bm;
        uint64_t m1 = brev_bits(P, b);
        S[s] = src[(((uint64_t)r2 * n1 + m1) << logB) + j];
    }
    tile_dit<F>(S, b, logRB, wtab, 0u, 0u);
    for (uint32_t s = threadIdx.x; s < tile; s += blockDim.x) {
        uint32_t beta = s & RBm;
        uint64_t k1 = s >> logRB;
        dst[((k1 * n2 + k2_0) << logB) + beta] = S[s];
    }
}

template <class F>
__global__ void k_transpose(const typename F::T* __restrict__ in, typename F::T* __restrict__ out, uint64_t rows,
                            uint64_t width, int to_colmajor) {
    // in row-major [rows][width] <-> out column-major [width][rows]; 32x32 tiles through shared memory
    __shared__ typename F::T tile[32][33];
    uint64_t r0 = (uint64_t)blockIdx.x * 32, c0 = (uint64_t)blockIdx.y * 32;
    if (to_colmajor) {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            uint64_t r = r0 + i, c = c0 + threadIdx.x;
            if (r < rows && c < width) tile[i][threadIdx.x] = in[r * width + c];
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            uint64_t c = c0 + i, r = r0 + threadIdx.x;
            if (r < rows && c < width) out[c * rows + r] = tile[threadIdx.x][i];
        }
    } else {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            uint64_t c = c0 + i, r = r0 + threadIdx.x;
            if (r < rows && c < width) tile[threadIdx.x][i] = in[c * rows + r];
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            uint64_t r = r0 + i, c = c0 + threadIdx.x;
            if (r < rows && c < width) out[r * width + c] = tile[i][threadIdx.x];
        }
    }
}

// -------------------------------------------------------------------------------------------------
template <class F>
int ensure_wtab(Ctx* c, int inverse) {
    using T = typename F::T;
    if (c->wtab[inverse]) return MS_OK;
    T roots[NTT_MAXLOG];
    for (int u = 0; u < NTT_MAXLOG; u++) {
        T g = (u + 1 <= F::TWO_ADICITY) ? root_of_unity<F>(u + 1) : (T)1;
        roots[u] = inverse ? finv<F>(g) : g;
    }
    T* d_roots;
    MS_CUDA(c, cudaMalloc(&d_roots, sizeof roots));
    MS_CUDA(c, cudaMemcpyAsync(d_roots, roots, sizeof roots, cudaMemcpyHostToDevice, c->stream));
    T* w;
    MS_CUDA(c, cudaMalloc(&w, sizeof(T) << NTT_MAXLOG));
    k_build_wtab<F><<<(1 << NTT_MAXLOG) / 256, 256, 0, c->stream>>>(w, d_roots);
    MS_LAUNCH_CHECK(c);
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_roots);
    c->wtab[inverse] = w;
    return MS_OK;
}

struct NttPlan {
    int a, b, logR1, logR2;
};
inline bool ntt_plan(int logN, int logB, NttPlan* p) {
    if (logN + logB <= NTT_LOG_TILE_PREF) {
        *p = {logN, 0, 0, 0};
        return true;
    }
    int a = (logN + 1) / 2, b = logN - a;
    if (a + logB > NTT_LOG_TILE_MAX) return false;
    int r1 = NTT_LOG_TILE_PREF - a - logB;
    if (r1 < 0) r1 = 0;
    if (r1 > b) r1 = b;
    int r2 = NTT_LOG_TILE_PREF - b - logB;
    if (r2 < 0) r2 = 0;
    if (r2 > a) r2 = a;
    *p = {a, b, r1, r2};
    return true;
}

// See the header comment.  `d_tmp` (cols * (N<<logB) elements) is only needed for two-pass sizes;
// pass nullptr to have it allocated from the stream-ordered pool.
template <class F>
int lde_batch(Ctx* c, const typename F::T* d_in, uint64_t in_stride, uint64_t cols, int logN, int logB,
              typename F::T shift, bool inverse, typename F::T* d_out, uint64_t out_stride) {
    using T = typename F::T;
    if (cols == 0) return MS_OK;
    if (logN + logB > F::TWO_ADICITY) return fail(c, MS_ERR_BAD_SHAPE, "domain 2^%d exceeds the field's two-adicity", logN + logB);
    if (inverse && logB != 0) return fail(c, MS_ERR_UNSUPPORTED, "inverse transform with blowup");
    NttPlan pl;
    if (!ntt_plan(logN, logB, &pl)) return fail(c, MS_ERR_UNSUPPORTED, "transform 2^%d x blowup 2^%d too large", logN, logB);
    MS_TRY(ensure_wtab<F>(c, inverse ? 1 : 0));
    const T* wtab = reinterpret_cast<const T*>(c->wtab[inverse ? 1 : 0]);
    const int B = 1 << logB;
    const uint64_t N = 1ULL << logN;
    T wN = root_of_unity<F>(logN);
    if (inverse) wN = finv<F>(wN);
    T scale = inverse ? finv<F>((T)(N % (uint64_t)F::P)) : (T)1;
    T wL = root_of_unity<F>(logN + logB);
    // host-side per-coset constants
    std::vector<T> hbuf((size_t)B * (pl.a ? pl.a : 1) + B);
    T* sbase = hbuf.data();
    T* shifts = hbuf.data() + (size_t)B * (pl.a ? pl.a : 1);
    T sj = shift;
    for (int j = 0; j < B; j++) {
        shifts[j] = sj;
        for (int u = 0; u < pl.a; u++) sbase[j * pl.a + u] = fpow<F>(sj, N >> (u + 1));
        sj = F::mul(sj, wL);
    }
    Scratch consts(c), t1(c), ft(c), tmp(c);
    MS_TRY(consts.alloc(hbuf.size() * sizeof(T)));
    MS_CUDA(c, cudaMemcpyAsync(consts.p, hbuf.data(), hbuf.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    // the H2D source must stay alive until the copy ran (pageable memory: the call is synchronous
    // with respect to the host buffer for pageable sources, so this is safe)
    const T* d_sbase = consts.as<T>();
    const T* d_shifts = d_sbase + (size_t)B * (pl.a ? pl.a : 1);
    MS_TRY(t1.alloc(((size_t)B << pl.a) * sizeof(T)));
    if (pl.a > 0) {
        int n = B << pl.a;
        prof_begin(c, "k_build_t1");
        k_build_t1<F><<<(n + 255) / 256, 256, 0, c->stream>>>(t1.as<T>(), wtab, d_sbase, pl.a, B);
        prof_end(c);
        MS_LAUNCH_CHECK(c);
    }
    const bool two = pl.b > 0;
    if (two) {
        MS_TRY(ft.alloc(((size_t)N << logB) * sizeof(T)));
        int chunk_log = pl.a < 6 ? pl.a : 6;
        uint64_t threads = (((uint64_t)1 << pl.b) << logB) * ((1ULL << pl.a) >> chunk_log);
        prof_begin(c, "k_build_ft");
        k_build_ft<F><<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(ft.as<T>(), d_shifts, wN, scale, pl.a, pl.b, logB, chunk_log);
        prof_end(c);
        MS_LAUNCH_CHECK(c);
        MS_TRY(tmp.alloc(cols * ((size_t)N << logB) * sizeof(T)));
    }
    {
        size_t smem = ((size_t)sizeof(T) << pl.a) << (pl.logR1 + logB);
        MS_CUDA(c, cudaFuncSetAttribute(k_lde_pass1<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)cols, (unsigned)((1ULL << pl.b) >> pl.logR1));
        T* o = two ? tmp.as<T>() : d_out;
        uint64_t os = two ? ((uint64_t)N << logB) : out_stride;
        prof_begin(c, "k_lde_pass1");
        k_lde_pass1<F><<<grid, NTT_THREADS, smem, c->stream>>>(d_in, in_stride, o, os, t1.as<T>(), two ? ft.as<T>() : nullptr,
                                                               pl.a, pl.b, logB, pl.logR1, scale);
        prof_end(c);
        MS_LAUNCH_CHECK(c);
    }
    if (two) {
        size_t smem = ((size_t)sizeof(T) << pl.b) << (pl.logR2 + logB);
        MS_CUDA(c, cudaFuncSetAttribute(k_lde_pass2<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)cols, (unsigned)((1ULL << pl.a) >> pl.logR2));
        prof_begin(c, "k_lde_pass2");
        k_lde_pass2<F><<<grid, NTT_THREADS, smem, c->stream>>>(tmp.as<T>(), (uint64_t)N << logB, d_out, out_stride, wtab,
                                                               pl.a, pl.b, logB, pl.logR2);
        prof_end(c);
        MS_LAUNCH_CHECK(c);
    }
    return MS_OK;
}

template <class F>
int transpose(Ctx* c, const typename F::T* d_in, typename F::T* d_out, uint64_t rows, uint64_t width, bool to_colmajor) {
    if (rows == 0 || width == 0) return MS_OK;
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((width + 31) / 32));
    k_transpose<F><<<grid, dim3(32, 8), 0, c->stream>>>(d_in, d_out, rows, width, to_colmajor ? 1 : 0);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

}  // namespace ms
