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

// ------------------------------------------------------------------------------------------ a6
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
template <class F, int CD>
__device__ __forceinline__ Ext<F> load_coef(const typename F::T* __restrict__ coef, uint64_t stride, uint64_t poly, uint64_t idx) {
    Ext<F> r = ext_zero<F>();
#pragma unroll
    for (int d = 0; d < CD; d++) r.c[d] = coef[(poly * CD + d) * stride + idx];
    return r;
}

// block-level value sum_t v_t * z^(seg*t) via a tree; result valid in thread 0
template <class F>
__device__ __forceinline__ Ext<F> block_tree(Ext<F> v, Ext<F> zseg, Ext<F>* sm) {
    const int t = threadIdx.x;
    sm[t] = v;
    Ext<F> pw = zseg;
    for (int ofs = 1; ofs < EV_THREADS; ofs <<= 1) {
        __syncthreads();
        if ((t & (2 * ofs - 1)) == 0) sm[t] = ext_add(sm[t], ext_mul(sm[t + ofs], pw));
        pw = ext_mul(pw, pw);
    }
    __syncthreads();
    Ext<F> r = sm[0];
    __syncthreads();
    return r;
}

template <class F, int CD>
__global__ void __launch_bounds__(EV_THREADS)
k_eval_partial(const typename F::T* __restrict__ coef, uint64_t stride, uint64_t off, uint64_t step, uint64_t n,
               const Ext<F>* __restrict__ z, int Q, Ext<F>* __restrict__ partial) {
    __shared__ Ext<F> sm[EV_THREADS];
    const uint64_t poly = blockIdx.y, npoly = gridDim.y, nblk = gridDim.x;
    const uint64_t i0 = (uint64_t)blockIdx.x * EV_BLOCK + (uint64_t)threadIdx.x * EV_SEG;
    Ext<F> cf[EV_SEG];
#pragma unroll
    for (int k = 0; k < EV_SEG; k++) cf[k] = (i0 + k < n) ? load_coef<F, CD>(coef, stride, poly, off + (i0 + k) * step) : ext_zero<F>();
    for (int q = 0; q < Q; q++) {
        const Ext<F> zq = z[q];
        Ext<F> acc = cf[EV_SEG - 1];
#pragma unroll
        for (int k = EV_SEG - 2; k >= 0; k--) acc = ext_add(ext_mul(acc, zq), cf[k]);
        Ext<F> zseg = zq;
#pragma unroll
        for (int k = 1; k < EV_SEG; k <<= 1) zseg = ext_mul(zseg, zseg);
        Ext<F> r = block_tree<F>(acc, zseg, sm);
        if (threadIdx.x == 0) partial[((uint64_t)q * npoly + poly) * nblk + blockIdx.x] = r;
    }
}

// out[q*npoly + poly] = sum_b partial[(q*npoly+poly)*nblk + b] * (z_q^EV_BLOCK)^b
template <class F>
__global__ void __launch_bounds__(EV_THREADS)
k_eval_final(const Ext<F>* __restrict__ partial, uint64_t nblk, uint64_t npoly, const Ext<F>* __restrict__ z, Ext<F>* __restrict__ out) {
    __shared__ Ext<F> sm[EV_THREADS];
    const uint64_t pq = blockIdx.x;
    const uint64_t q = pq / npoly;
    const Ext<F> Z = ext_pow<F>(z[q], EV_BLOCK);
    const uint64_t seg = (nblk + EV_THREADS - 1) / EV_THREADS;
    const uint64_t b0 = (uint64_t)threadIdx.x * seg;
    Ext<F> acc = ext_zero<F>();
    for (uint64_t k = seg; k-- > 0;) {
        uint64_t b = b0 + k;
        Ext<F> v = b < nblk ? partial[pq * nblk + b] : ext_zero<F>();
        acc = ext_add(ext_mul(acc, Z), v);
    }
    Ext<F> r = block_tree<F>(acc, ext_pow<F>(Z, seg), sm);
    if (threadIdx.x == 0) out[pq] = r;
}

// Evaluate `npoly` polynomials (coefficient degree CD = 1 base / F::D extension) at Q extension
// points; out_host[q*npoly + poly].
template <class F>
int eval_points(Ctx* c, const typename F::T* d_coef, uint64_t stride, uint64_t off, uint64_t step, uint64_t n, int coefdeg,
                uint64_t npoly, const Ext<F>* z_host, int Q, Ext<F>* out_host) {
    if (Q == 0 || npoly == 0) return MS_OK;
    if (n == 0) {
        for (uint64_t i = 0; i < (uint64_t)Q * npoly; i++) out_host[i] = ext_zero<F>();
        return MS_OK;
    }
    const uint64_t nblk = (n + EV_BLOCK - 1) / EV_BLOCK;
    Scratch dz(c), part(c), dout(c);
    MS_TRY(dz.alloc(Q * sizeof(Ext<F>)));
    MS_TRY(part.alloc((size_t)Q * npoly * nblk * sizeof(Ext<F>)));
    MS_TRY(dout.alloc((size_t)Q * npoly * sizeof(Ext<F>)));
    MS_CUDA(c, cudaMemcpyAsync(dz.p, z_host, Q * sizeof(Ext<F>), cudaMemcpyHostToDevice, c->stream));
    dim3 grid((unsigned)nblk, (unsigned)npoly);
    if (coefdeg == 1) k_eval_partial<F, 1><<<grid, EV_THREADS, 0, c->stream>>>(d_coef, stride, off, step, n, dz.as<Ext<F>>(), Q, part.as<Ext<F>>());
    else k_eval_partial<F, F::D><<<grid, EV_THREADS, 0, c->stream>>>(d_coef, stride, off, step, n, dz.as<Ext<F>>(), Q, part.as<Ext<F>>());
    MS_LAUNCH_CHECK(c);
    k_eval_final<F><<<(unsigned)(Q * npoly), EV_THREADS, 0, c->stream>>>(part.as<Ext<F>>(), nblk, npoly, dz.as<Ext<F>>(), dout.as<Ext<F>>());
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
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool nz = false;
    for (int d = 0; d < planes; d++) nz = nz || (p[(uint64_t)d * stride + i] != 0);
    if (nz) atomicMax(out, (unsigned long long)(i + 1));
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
    MS_CUDA(c, cudaMemcpyAsync(dmu.p, hm.data(), hm.size() * sizeof(M), cudaMemcpyHostToDevice, c->stream));
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
