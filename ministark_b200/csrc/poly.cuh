// poly.cuh -- coefficient-space stages of the prover:
//   a3  linear transition constraints            (src/air.rs:130-134, closures of tests/e2e_*.rs:48-59)
//   a6  constraint mixing  g = sum r^i f_i        (src/starks.rs:108-117)
//   a7  DEEP-ALI openings  f_c(z_q), z in Ext      (src/starks.rs:140-151)
// plus the two generic device primitives the FRI stages share:
//   * chunked Horner evaluation at extension points (block tree with powers z^(SEG*2^l))
//   * suffix scan  T(j) = e_j + mu * T(j+1)  (= synthetic division), three launches, exact
//     arithmetic so any association order is bit-identical to the reference's serial loops.
#pragma once
#include "common.cuh"
#include "field.cuh"
#include "ntt.cuh"

namespace ms {

constexpr int EV_THREADS = 256;
constexpr int EV_SEG = 8;
constexpr int EV_BLOCK = EV_THREADS * EV_SEG;  // coefficients per block

// ------------------------------------------------------------------------------------------ a3
template <class F>
__global__ void k_linear_constraints(const typename F::T* __restrict__ coef, uint64_t stride, uint64_t n, int w,
                                     const typename F::T* __restrict__ mat, int t, typename F::T* __restrict__ out,
                                     uint64_t out_stride) {
    using T = typename F::T;
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    for (int r = 0; r < t; r++) {
        T acc = 0;
        for (int j = 0; j < w; j++) {
            T s = mat[r * w + j];  // warp-uniform
            if (s == 0) continue;
            T v = coef[(uint64_t)j * stride + m];
            acc = F::add(acc, s == 1 ? v : F::mul(s, v));
        }
        out[(uint64_t)r * out_stride + m] = acc;
    }
}

template <class F>
int linear_constraints(Ctx* c, const typename F::T* d_coef, uint64_t stride, uint64_t n, uint64_t w, const typename F::T* mat_host,
                       uint64_t t, typename F::T* d_out, uint64_t out_stride) {
    using T = typename F::T;
    if (t == 0 || n == 0) return MS_OK;
    std::vector<T> m(mat_host, mat_host + t * w);
    for (auto& v : m) v = (T)((uint64_t)v % (uint64_t)F::P);
    Scratch dm(c);
    MS_TRY(dm.alloc(t * w * sizeof(T)));
    MS_CUDA(c, cudaMemcpyAsync(dm.p, m.data(), t * w * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    k_linear_constraints<F><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_coef, stride, n, (int)w, dm.as<T>(), (int)t, d_out, out_stride);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

// Column-sharded form: trace coefficient column j lives at cols[j] (possibly in a peer GPU's arena, read over
// NVLink); only the columns a row of the matrix actually uses are touched (zero entries are skipped), so the
// bidiagonal matrices of the e2e / synthetic AIRs read two columns per constraint.
template <class F>
__global__ void k_linear_constraints_gather(const typename F::T* const* __restrict__ cols, uint64_t n, int w,
                                            const typename F::T* __restrict__ mat, int t, typename F::T* __restrict__ out,
                                            uint64_t out_stride) {
    using T = typename F::T;
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    for (int r = 0; r < t; r++) {
        T acc = 0;
        for (int j = 0; j < w; j++) {
            T s = mat[r * w + j];  // warp-uniform
            if (s == 0) continue;
            T v = cols[j][m];
            acc = F::add(acc, s == 1 ? v : F::mul(s, v));
        }
        out[(uint64_t)r * out_stride + m] = acc;
    }
}
// rows [t0, t0 + t) of the T x W host matrix applied to the columns in the device pointer table d_cols
template <class F>
int linear_constraints_gather(Ctx* c, const typename F::T* const* d_cols, uint64_t n, uint64_t w, const typename F::T* mat_rows_host,
                              uint64_t t, typename F::T* d_out, uint64_t out_stride) {
    using T = typename F::T;
    if (t == 0 || n == 0) return MS_OK;
    std::vector<T> m(mat_rows_host, mat_rows_host + t * w);
    for (auto& v : m) v = (T)((uint64_t)v % (uint64_t)F::P);
    Scratch dm(c);
    MS_TRY(dm.alloc(t * w * sizeof(T)));
    MS_TRY(stage_from_host(c, m.data(), ((t * w * sizeof(T) + 3) / 4) * 4, dm.p));
    k_linear_constraints_gather<F><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_cols, n, (int)w, dm.as<T>(), (int)t, d_out, out_stride);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

// Additive constants of affine constraints (f_{W+t} = sum_w M[t][w] f_w + c_t): the constant polynomial c_t adds c_t to
// coefficient 0 of the column -- or to every evaluation of it, when the column is built in evaluation space.
template <class F>
__global__ void k_add_consts(typename F::T* __restrict__ cols, uint64_t stride, uint64_t n, const typename F::T* __restrict__ consts, int t,
                             int everywhere) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (everywhere ? n : 1)) return;
    for (int r = 0; r < t; r++) {
        const typename F::T cst = consts[r];
        if (cst) cols[(uint64_t)r * stride + m] = F::add(cols[(uint64_t)r * stride + m], cst);
    }
}
template <class F>
int add_consts(Ctx* c, typename F::T* d_cols, uint64_t stride, uint64_t n, const typename F::T* consts_host, uint64_t t, bool everywhere) {
    using T = typename F::T;
    if (!consts_host || t == 0 || n == 0) return MS_OK;
    std::vector<T> cs(t);
    bool any = false;
    for (uint64_t r = 0; r < t; r++) { cs[r] = (T)((uint64_t)consts_host[r] % (uint64_t)F::P); any = any || cs[r]; }
    if (!any) return MS_OK;
    Scratch dc(c);
    MS_TRY(dc.alloc(t * sizeof(T)));
    MS_TRY(stage_from_host(c, cs.data(), t * sizeof(T), dc.p));
    const uint64_t work = everywhere ? n : 1;
    k_add_consts<F><<<(unsigned)((work + 255) / 256), 256, 0, c->stream>>>(d_cols, stride, n, dc.as<T>(), (int)t, everywhere ? 1 : 0);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

// Sparse rows (at most 4 non-zero entries each: every AIR of the reference's tests, and the synthetic one): a row is a
// short list of (column, scalar) pairs, +1 / -1 scalars cost an add / a sub instead of a product, and a thread handles two
// consecutive elements with 16-byte loads.  Same values as k_linear_constraints (exact arithmetic).
constexpr int LIN_MAX_NNZ = 4;
template <class F>
struct SparseRow {
    int nnz;
    int col[LIN_MAX_NNZ];
    typename F::T val[LIN_MAX_NNZ];
};
template <class F>
__global__ void k_linear_sparse(const typename F::T* const* __restrict__ cols, uint64_t n, const SparseRow<F>* __restrict__ rows, int t,
                                typename F::T* __restrict__ out, uint64_t out_stride) {
    using T = typename F::T;
    const uint64_t m = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (m >= n) return;
    const bool pair = m + 1 < n;
    for (int r = 0; r < t; r++) {
        const SparseRow<F> row = rows[r];  // warp-uniform
        T a0 = 0, a1 = 0;
        for (int e = 0; e < row.nnz; e++) {
            const T* __restrict__ src = cols[row.col[e]];
            T v0 = src[m], v1 = pair ? src[m + 1] : (T)0;
            const T sc = row.val[e];
            if (sc == 1) { a0 = F::add(a0, v0); a1 = F::add(a1, v1); }
            else if (sc == (T)(F::P - 1)) { a0 = F::sub(a0, v0); a1 = F::sub(a1, v1); }
            else { a0 = F::add(a0, F::mul(sc, v0)); a1 = F::add(a1, F::mul(sc, v1)); }
        }
        out[(uint64_t)r * out_stride + m] = a0;
        if (pair) out[(uint64_t)r * out_stride + m + 1] = a1;
    }
}
// rows of the host matrix applied to the columns of the device pointer table; false when a row has too many entries
template <class F>
bool sparse_rows(const typename F::T* mat_rows_host, uint64_t t, uint64_t w, std::vector<SparseRow<F>>* out) {
    out->assign(t, SparseRow<F>{});
    for (uint64_t r = 0; r < t; r++) {
        SparseRow<F>& row = (*out)[r];
        for (uint64_t j = 0; j < w; j++) {
            const typename F::T v = (typename F::T)((uint64_t)mat_rows_host[r * w + j] % (uint64_t)F::P);
            if (!v) continue;
            if (row.nnz == LIN_MAX_NNZ) return false;
            row.col[row.nnz] = (int)j;
            row.val[row.nnz++] = v;
        }
    }
    return true;
}
template <class F>
int linear_sparse(Ctx* c, const typename F::T* const* d_cols, uint64_t n, const std::vector<SparseRow<F>>& rows, typename F::T* d_out,
                  uint64_t out_stride) {
    if (rows.empty() || n == 0) return MS_OK;
    Scratch dr(c);
    const size_t bytes = rows.size() * sizeof(SparseRow<F>);
    MS_TRY(dr.alloc(bytes));
    MS_TRY(stage_from_host(c, rows.data(), bytes, dr.p));
    k_linear_sparse<F><<<(unsigned)((n / 2 + 256) / 256), 256, 0, c->stream>>>(d_cols, n, dr.as<SparseRow<F>>(), (int)rows.size(), d_out, out_stride);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}
// pointer table of `w` columns of a contiguous column-major matrix
template <class F>
int column_table(Ctx* c, const typename F::T* base, uint64_t stride, uint64_t w, Scratch* tab) {
    std::vector<const typename F::T*> cols(w);
    for (uint64_t j = 0; j < w; j++) cols[j] = base + j * stride;
    MS_TRY(tab->alloc(w * sizeof(void*)));
    return stage_from_host(c, cols.data(), w * sizeof(void*), tab->p);
}

// ------------------------------------------------------------------------------------------ a6
// A rank's share of the mix: out[m] = r0a * sum_{i<na} r^i a_i[m] + r0b * sum_{i<nb} r^i b_i[m]  (two runs of
// consecutive columns with their starting powers r0 = r^(global index of the run's first column)).
template <class F>
__global__ void k_mix_parts(const typename F::T* __restrict__ a, uint64_t sa, int na, typename F::T r0a,
                            const typename F::T* __restrict__ b, uint64_t sb, int nb, typename F::T r0b, uint64_t n,
                            typename F::T r, typename F::T* __restrict__ out) {
    using T = typename F::T;
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    T tot = 0;
    if (na > 0) {
        T acc = a[(uint64_t)(na - 1) * sa + m];
        for (int i = na - 2; i >= 0; i--) acc = F::add(F::mul(acc, r), a[(uint64_t)i * sa + m]);
        tot = F::mul(acc, r0a);
    }
    if (nb > 0) {
        T acc = b[(uint64_t)(nb - 1) * sb + m];
        for (int i = nb - 2; i >= 0; i--) acc = F::add(F::mul(acc, r), b[(uint64_t)i * sb + m]);
        tot = F::add(tot, F::mul(acc, r0b));
    }
    out[m] = tot;
}
constexpr int MS_MAX_RANKS = 16;
struct PeerTable {
    const void* p[MS_MAX_RANKS];
};
// out[m] = sum_g part_g[m]: the ranks' partial mixes, read from their arenas (exact arithmetic: any order)
template <class F>
__global__ void k_sum_peers(PeerTable parts, int world, uint64_t n, typename F::T* __restrict__ out) {
    using T = typename F::T;
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    T acc = 0;
    for (int g = 0; g < world; g++) acc = F::add(acc, static_cast<const T*>(parts.p[g])[m]);
    out[m] = acc;
}

template <class F>
__global__ void k_mix(const typename F::T* __restrict__ coef, uint64_t stride, uint64_t n, int cols, typename F::T r,
                      typename F::T* __restrict__ out) {
    using T = typename F::T;
    uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    T acc = coef[(uint64_t)(cols - 1) * stride + m];
    for (int i = cols - 2; i >= 0; i--) acc = F::add(F::mul(acc, r), coef[(uint64_t)i * stride + m]);
    out[m] = acc;
}

template <class F>
int mix(Ctx* c, const typename F::T* d_coef, uint64_t stride, uint64_t n, uint64_t cols, typename F::T r, typename F::T* d_out) {
    if (cols == 0 || n == 0) return fail(c, MS_ERR_BAD_SHAPE, "mix needs at least one column");
    k_mix<F><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_coef, stride, n, (int)cols, r, d_out);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

// ------------------------------------------------------------------------------------------ evaluation
// Sequence element i of polynomial `poly`: coordinate d at coef[(poly*CD + d)*stride + off + i*step].
// f(z) = sum_i c_i z^i is computed without a dependency chain: block b, thread t, k < EV_SEG owns
// i = b*EV_BLOCK + k*EV_THREADS + t (coalesced), so
//   f(z) = sum_b Zb[b] * sum_t zt[t] * sum_k c_i * z256[k],   zt[t] = z^t, z256[k] = z^(256k), Zb[b] = z^(2048b)
// with the three power tables built once per call (k_eval_powers).  For base-field coefficients
// (trace / constraint polynomials, src/starks.rs:140-151) the inner products are base x extension,
// i.e. D base multiplications each; one coefficient load serves all Q points.
template <class F, int CD>
__device__ __forceinline__ Ext<F> load_coef(const typename F::T* __restrict__ coef, uint64_t stride, uint64_t poly, uint64_t idx) {
    Ext<F> r = ext_zero<F>();
#pragma unroll
    for (int d = 0; d < CD; d++) r.c[d] = coef[(poly * CD + d) * stride + idx];
    return r;
}

// tables for Q points: [q][0..EV_THREADS) z^t | [q][EV_THREADS..+EV_SEG) z^(256 k) | [q][..+nblk) z^(2048 b)
template <class F>
__global__ void k_eval_powers(const Ext<F>* __restrict__ z, int Q, uint64_t nblk, Ext<F>* __restrict__ tab) {
    const uint64_t per_q = EV_THREADS + EV_SEG + nblk;
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_q * Q) return;
    const int q = (int)(gid / per_q);
    const uint64_t i = gid % per_q;
    Ext<F> v;
    if (i < EV_THREADS) v = ext_pow<F>(z[q], i);
    else if (i < EV_THREADS + EV_SEG) v = ext_pow<F>(z[q], (uint64_t)EV_THREADS * (i - EV_THREADS));
    else v = ext_pow<F>(ext_pow<F>(z[q], EV_BLOCK), i - EV_THREADS - EV_SEG);
    tab[gid] = v;
}

template <class F, int CD>
__device__ __forceinline__ Ext<F> coef_times(const Ext<F>& c, const Ext<F>& p) {
    if (CD == 1) return ext_mul_base(p, c.c[0]);
    return ext_mul(c, p);
}

template <class F, int CD, int QMAX>
__global__ void __launch_bounds__(EV_THREADS)
k_eval_partial(const typename F::T* __restrict__ coef, uint64_t stride, uint64_t off, uint64_t step, uint64_t n,
               const Ext<F>* __restrict__ tab, int Q, Ext<F>* __restrict__ partial, uint64_t poly_off, int shared_planes,
               uint64_t limit) {
    __shared__ Ext<F> sm[EV_THREADS / 32];
    const uint64_t poly = blockIdx.y, npoly = gridDim.y, nblk = gridDim.x;
    const uint64_t per_q = EV_THREADS + EV_SEG + nblk;
    const uint64_t i0 = (uint64_t)blockIdx.x * EV_BLOCK + threadIdx.x;
    Ext<F> cf[EV_SEG];
#pragma unroll
    for (int k = 0; k < EV_SEG; k++) {
        const uint64_t i = i0 + (uint64_t)k * EV_THREADS;
        // polynomial `poly` starts `poly * poly_off` entries further; with shared_planes all polynomials are
        // subsequences of the same planes (the even / odd halves of one FRI polynomial, src/fri.rs:329-343)
        const uint64_t idx = off + poly * poly_off + i * step;
        cf[k] = (i < n && idx < limit) ? load_coef<F, CD>(coef, stride, shared_planes ? 0 : poly, idx) : ext_zero<F>();
    }
    for (int q = 0; q < Q; q++) {
        const Ext<F>* tq = tab + (uint64_t)q * per_q;
        Ext<F> acc = cf[0];
        if (CD == 1 && sizeof(typename F::T) == 8) {
            // base coefficient x extension power, Goldilocks: lazy multiply-accumulate per coordinate
            // (Fast<GL>: 17 + 5 instructions instead of 33 + 10 for the canonical ops)
#pragma unroll
            for (int k = 1; k < EV_SEG; k++) {
                const Ext<F> pw = tq[EV_THREADS + k];
#pragma unroll
                for (int d = 0; d < F::D; d++) acc.c[d] = Fast<F>::add(acc.c[d], Fast<F>::mul(cf[k].c[0], pw.c[d]));
            }
#pragma unroll
            for (int d = 0; d < F::D; d++) acc.c[d] = Fast<F>::canon(acc.c[d]);
        } else {
#pragma unroll
            for (int k = 1; k < EV_SEG; k++) acc = ext_add(acc, coef_times<F, CD>(cf[k], tq[EV_THREADS + k]));
        }
        acc = ext_mul(acc, tq[threadIdx.x]);
        // block sum: warp shuffles, then one value per warp through shared memory
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            Ext<F> o;
#pragma unroll
            for (int d = 0; d < F::D; d++) o.c[d] = __shfl_down_sync(0xffffffffu, acc.c[d], ofs);
            acc = ext_add(acc, o);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            Ext<F> r = sm[0];
#pragma unroll
            for (int w = 1; w < EV_THREADS / 32; w++) r = ext_add(r, sm[w]);
            partial[((uint64_t)q * npoly + poly) * nblk + blockIdx.x] = ext_mul(r, tq[EV_THREADS + EV_SEG + blockIdx.x]);
        }
    }
}

// out[pq] = sum_b partial[pq*nblk + b]
template <class F>
__global__ void __launch_bounds__(EV_THREADS)
k_eval_final(const Ext<F>* __restrict__ partial, uint64_t nblk, Ext<F>* __restrict__ out) {
    __shared__ Ext<F> sm[EV_THREADS / 32];
    const uint64_t pq = blockIdx.x;
    Ext<F> acc = ext_zero<F>();
    for (uint64_t b = threadIdx.x; b < nblk; b += EV_THREADS) acc = ext_add(acc, partial[pq * nblk + b]);
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
        Ext<F> o;
#pragma unroll
        for (int d = 0; d < F::D; d++) o.c[d] = __shfl_down_sync(0xffffffffu, acc.c[d], ofs);
        acc = ext_add(acc, o);
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        Ext<F> r = sm[0];
        for (int w = 1; w < EV_THREADS / 32; w++) r = ext_add(r, sm[w]);
        out[pq] = r;
    }
}

// Evaluate `npoly` polynomials (coefficient degree CD = 1 base / F::D extension) at Q extension
// points; out_host[q*npoly + poly].
template <class F>
int eval_points(Ctx* c, const typename F::T* d_coef, uint64_t stride, uint64_t off, uint64_t step, uint64_t n, int coefdeg,
                uint64_t npoly, const Ext<F>* z_host, int Q, Ext<F>* out_host, uint64_t poly_off = 0, bool shared_planes = false,
                uint64_t limit = ~0ULL) {
    if (Q == 0 || npoly == 0) return MS_OK;
    if (n == 0) {
        for (uint64_t i = 0; i < (uint64_t)Q * npoly; i++) out_host[i] = ext_zero<F>();
        return MS_OK;
    }
    const uint64_t nblk = (n + EV_BLOCK - 1) / EV_BLOCK;
    const uint64_t per_q = EV_THREADS + EV_SEG + nblk;
    Scratch dz(c), tab(c), part(c), dout(c);
    MS_TRY(dz.alloc(Q * sizeof(Ext<F>)));
    MS_TRY(tab.alloc((size_t)Q * per_q * sizeof(Ext<F>)));
    MS_TRY(part.alloc((size_t)Q * npoly * nblk * sizeof(Ext<F>)));
    MS_TRY(dout.alloc((size_t)Q * npoly * sizeof(Ext<F>)));
    MS_CUDA(c, cudaMemcpyAsync(dz.p, z_host, Q * sizeof(Ext<F>), cudaMemcpyHostToDevice, c->stream));
    prof_begin(c, "k_eval_powers");
    k_eval_powers<F><<<(unsigned)((per_q * Q + 127) / 128), 128, 0, c->stream>>>(dz.as<Ext<F>>(), Q, nblk, tab.as<Ext<F>>());
    prof_end(c);
    MS_LAUNCH_CHECK(c);
    dim3 grid((unsigned)nblk, (unsigned)npoly);
    prof_begin(c, "k_eval_partial");
    if (coefdeg == 1) k_eval_partial<F, 1, 0><<<grid, EV_THREADS, 0, c->stream>>>(d_coef, stride, off, step, n, tab.as<Ext<F>>(), Q, part.as<Ext<F>>(), poly_off, shared_planes ? 1 : 0, limit);
    else k_eval_partial<F, F::D, 0><<<grid, EV_THREADS, 0, c->stream>>>(d_coef, stride, off, step, n, tab.as<Ext<F>>(), Q, part.as<Ext<F>>(), poly_off, shared_planes ? 1 : 0, limit);
    prof_end(c);
    MS_LAUNCH_CHECK(c);
    k_eval_final<F><<<(unsigned)(Q * npoly), EV_THREADS, 0, c->stream>>>(part.as<Ext<F>>(), nblk, dout.as<Ext<F>>());
    MS_LAUNCH_CHECK(c);
    MS_CUDA(c, cudaMemcpyAsync(out_host, dout.p, (size_t)Q * npoly * sizeof(Ext<F>), cudaMemcpyDeviceToHost, c->stream));
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    return MS_OK;
}

// ------------------------------------------------------------------------------------------ a7
template <class F>
int deep_open(Ctx* c, const typename F::T* d_coef, uint64_t stride, uint64_t n, uint64_t cols, const typename F::T* z_host, uint64_t Q,
              typename F::T* out_host) {
    std::vector<Ext<F>> z(Q);
    for (uint64_t q = 0; q < Q; q++)
        for (int d = 0; d < F::D; d++) z[q].c[d] = (typename F::T)((uint64_t)z_host[q * F::D + d] % (uint64_t)F::P);
    return eval_points<F>(c, d_coef, stride, 0, 1, n, 1, cols, z.data(), (int)Q, reinterpret_cast<Ext<F>*>(out_host));
}

// ------------------------------------------------------------------------------------------ degree
// out = 1 + (largest index with a non-zero coordinate in any of `planes` planes), 0 for the zero poly
template <class F>
__global__ void k_poly_len(const typename F::T* __restrict__ p, uint64_t stride, int planes, uint64_t n, unsigned long long* out) {
    __shared__ unsigned long long sm[8];
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool nz = false;
    if (i < n)
        for (int d = 0; d < planes; d++) nz = nz || (p[(uint64_t)d * stride + i] != 0);
    // one atomic per block: the highest non-zero lane of the highest non-empty warp
    const unsigned ballot = __ballot_sync(0xffffffffu, nz);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sm[warp] = ballot ? (unsigned long long)(i + (31 - __clz(ballot)) + 1) : 0ULL;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = sm[w] > m ? sm[w] : m;
        if (m) atomicMax(out, m);
    }
}

// ------------------------------------------------------------------------------------------ suffix scan
// Given a sequence e_0..e_{len-1} (functor Src), a multiplier mu and an external tail T(len) = tail_in,
// produce dst(i) = T(i+1) where T(j) = e_j + mu*T(j+1): exactly the quotient coefficients of the
// synthetic division by (x - mu) (src/fri.rs:101) / by (x^2 - mu) per parity (src/fri.rs:159-167).
template <class F>
struct ExtOps {
    using V = Ext<F>;
    using M = Ext<F>;
    static MS_HD V zero() { return ext_zero<F>(); }
    static MS_HD V add(const V& a, const V& b) { return ext_add(a, b); }
    static MS_HD V mulm(const V& a, const M& m) { return ext_mul(a, m); }
    static MS_HD M mm(const M& a, const M& b) { return ext_mul(a, b); }
    static MS_HD M one() { return ext_from_base<F>(1); }
};
template <class F>
struct BaseOps {
    using V = typename F::T;
    using M = typename F::T;
    static MS_HD V zero() { return 0; }
    static MS_HD V add(V a, V b) { return F::add(a, b); }
    static MS_HD V mulm(V a, M m) { return F::mul(a, m); }
    static MS_HD M mm(M a, M b) { return F::mul(a, b); }
    static MS_HD M one() { return 1; }
};
template <class Ops>
__device__ __forceinline__ typename Ops::M mpow(typename Ops::M b, uint64_t e) {
    typename Ops::M r = Ops::one();
    while (e) {
        if (e & 1) r = Ops::mm(r, b);
        b = Ops::mm(b, b);
        e >>= 1;
    }
    return r;
}

constexpr int SC_THREADS = 256;

// One block handles SC_THREADS*seg consecutive elements of batch blockIdx.y (multiplier mus[batch]).
//   blocksum != nullptr : write H_b = sum_{i in block} e_i mu^(i - start_b)         (phase A)
//   otherwise           : tails[b] (or zero) is T(end_b); write dst(i) = T(i+1)       (phase B)
// Elements past a batch's own length must read as zero (Src's job); Dst drops what it does not want.
template <class Ops, class Src, class Dst>
__global__ void __launch_bounds__(SC_THREADS)
k_suffix_scan(Src src, uint64_t len, const typename Ops::M* __restrict__ mus, uint32_t seg,
              const typename Ops::V* __restrict__ tails, Dst dst, typename Ops::V* __restrict__ blocksum) {
    using V = typename Ops::V;
    using M = typename Ops::M;
    __shared__ V sm[SC_THREADS + 1];
    const int t = threadIdx.x;
    const uint32_t batch = blockIdx.y;
    const M mu = mus[batch];
    const uint64_t start = ((uint64_t)blockIdx.x * SC_THREADS + t) * seg;
    // local Horner value of this thread's segment
    V h = Ops::zero();
    for (uint32_t k = seg; k-- > 0;) {
        uint64_t i = start + k;
        V e = i < len ? src(batch, i) : Ops::zero();
        h = Ops::add(Ops::mulm(h, mu), e);
    }
    sm[t] = h;
    if (t == 0) sm[SC_THREADS] = (!blocksum && tails) ? tails[(uint64_t)batch * gridDim.x + blockIdx.x] : Ops::zero();
    // inclusive suffix scan over SC_THREADS + 1 items, all with span mu^seg
    M pw = mpow<Ops>(mu, seg);
    for (int ofs = 1; ofs <= SC_THREADS; ofs <<= 1) {
        __syncthreads();
        V add_v = Ops::zero();
        bool has = t + ofs <= SC_THREADS;
        if (has) add_v = sm[t + ofs];
        __syncthreads();
        if (has) sm[t] = Ops::add(sm[t], Ops::mulm(add_v, pw));
        pw = Ops::mm(pw, pw);
    }
    __syncthreads();
    if (blocksum) {
        if (t == 0) blocksum[(uint64_t)batch * gridDim.x + blockIdx.x] = sm[0];
        return;
    }
    V r = sm[t + 1];  // T(end of this thread's segment)
    for (uint32_t k = seg; k-- > 0;) {
        uint64_t i = start + k;
        if (i < len) {
            dst(batch, i, r);
            r = Ops::add(src(batch, i), Ops::mulm(r, mu));
        }
    }
}

template <class V>
struct ArraySrc {  // [batch][n]
    const V* p;
    uint64_t n;
    __device__ __forceinline__ V operator()(uint32_t b, uint64_t i) const { return p[(uint64_t)b * n + i]; }
};
template <class V>
struct ArrayDst {
    V* p;
    uint64_t n;
    __device__ __forceinline__ void operator()(uint32_t b, uint64_t i, const V& v) const { p[(uint64_t)b * n + i] = v; }
};
template <class V>
struct NullDst {
    __device__ __forceinline__ void operator()(uint32_t, uint64_t, const V&) const {}
};

// `nbatch` independent scans of (padded) length len; mus_host[b] is batch b's multiplier.
template <class Ops, class Src, class Dst>
int suffix_scan(Ctx* c, Src src, uint64_t len, uint32_t nbatch, const typename Ops::M* mus_host, Dst dst) {
    using V = typename Ops::V;
    using M = typename Ops::M;
    if (len == 0 || nbatch == 0) return MS_OK;
    const uint32_t seg = 8;
    const uint64_t per_block = (uint64_t)SC_THREADS * seg;  // a power of two
    const uint64_t nblk = (len + per_block - 1) / per_block;
    if (nblk > (uint64_t)SC_THREADS * 4096) return fail(c, MS_ERR_UNSUPPORTED, "suffix scan too long");
    std::vector<M> hm(2 * (size_t)nbatch);
    for (uint32_t b = 0; b < nbatch; b++) {
        M r = mus_host[b];
        hm[b] = r;
        for (uint64_t e = per_block; e > 1; e >>= 1) r = Ops::mm(r, r);  // mu^per_block
        hm[nbatch + b] = r;
    }
    Scratch dmu(c);
    MS_TRY(dmu.alloc(hm.size() * sizeof(M)));
    MS_TRY(stage_from_host(c, hm.data(), hm.size() * sizeof(M), dmu.p));  // not a copy-engine H2D: see common.cuh
    const M* d_mu = dmu.as<M>();
    const M* d_big = d_mu + nbatch;
    if (nblk == 1) {
        k_suffix_scan<Ops, Src, Dst><<<dim3(1, nbatch), SC_THREADS, 0, c->stream>>>(src, len, d_mu, seg, nullptr, dst, nullptr);
        MS_LAUNCH_CHECK(c);
        return MS_OK;
    }
    Scratch sums(c), tails(c);
    MS_TRY(sums.alloc(nblk * nbatch * sizeof(V)));
    MS_TRY(tails.alloc(nblk * nbatch * sizeof(V)));
    k_suffix_scan<Ops, Src, NullDst<V>><<<dim3((unsigned)nblk, nbatch), SC_THREADS, 0, c->stream>>>(src, len, d_mu, seg, nullptr, NullDst<V>{}, sums.as<V>());
    MS_LAUNCH_CHECK(c);
    // tails[b] = T(end_b): the same scan over the block sums with multiplier mu^per_block
    const uint32_t seg2 = (uint32_t)((nblk + SC_THREADS - 1) / SC_THREADS);
    k_suffix_scan<Ops, ArraySrc<V>, ArrayDst<V>><<<dim3(1, nbatch), SC_THREADS, 0, c->stream>>>(
        ArraySrc<V>{sums.as<V>(), nblk}, nblk, d_big, seg2, nullptr, ArrayDst<V>{tails.as<V>(), nblk}, nullptr);
    MS_LAUNCH_CHECK(c);
    k_suffix_scan<Ops, Src, Dst><<<dim3((unsigned)nblk, nbatch), SC_THREADS, 0, c->stream>>>(src, len, d_mu, seg, tails.as<V>(), dst, nullptr);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

}  // namespace ms
