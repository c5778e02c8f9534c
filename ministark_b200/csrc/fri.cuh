// fri.cuh -- DEEP-FRI stages (src/fri.rs).  Extension polynomials / codewords live on the device
// as D coordinate planes of base-field elements, zero padded to a power of two.
//   a8   codeword + (2,2) Merkle tree per round       src/fri.rs:314-352
//   a9   d = [f_even(z), f_odd(z)]                      src/fri.rs:354-359
//   a10  next = (f_even + alpha f_odd - d(alpha)) / (x - z)   src/fri.rs:97-101, 361-372
//   (f)  query phase: openings by value search, per-query quotients   src/fri.rs:132-172
#pragma once
#include "common.cuh"
#include "field.cuh"
#include "merkle.cuh"
#include "ntt.cuh"
#include "poly.cuh"

namespace ms {

// a8.  ark embeds the base-field root into the extension, so the extension FFT is
// D independent base-field transforms: the poly's D planes go through lde_batch as D columns with
// shift 1 (plain subgroup, fri.rs:315), then leaves are pairs of adjacent evaluations (fri.rs:351
// with leafs_per_node = 2, starks.rs:290-295).
template <class F>
int fri_commit(Ctx* c, const typename F::T* d_poly, uint64_t poly_stride, uint64_t domain, uint64_t blowup,
               typename F::T* d_cw, uint64_t cw_stride, uint32_t* d_nodes, uint8_t* root32) {
    if (!is_pow2(domain) || !is_pow2(blowup) || domain < blowup || domain < 2)
        return fail(c, MS_ERR_BAD_SHAPE, "FRI domain %llu / blowup %llu", (unsigned long long)domain, (unsigned long long)blowup);
    const uint64_t npad = domain / blowup;
    MS_TRY(lde_batch<F>(c, d_poly, poly_stride, F::D, ilog2(npad), ilog2(blowup), (typename F::T)1, false, d_cw, cw_stride));
    return merkle_commit<F>(c, d_cw, cw_stride, domain, 1, F::D, 2, 2, d_nodes, root32);
}

template <class F>
inline Ext<F> ext_from_host(const typename F::T* p) {
    Ext<F> r;
    for (int d = 0; d < F::D; d++) r.c[d] = (typename F::T)((uint64_t)p[d] % (uint64_t)F::P);
    return r;
}

// a9.  split_poly (fri.rs:329-343) is the even / odd coefficient subsequence: off = parity, step = 2.
template <class F>
int fri_deep_coeffs(Ctx* c, const typename F::T* d_poly, uint64_t stride, uint64_t n, const typename F::T* z_host,
                    typename F::T* d_out_host) {
    Ext<F> z = ext_from_host<F>(z_host);
    Ext<F>* out = reinterpret_cast<Ext<F>*>(d_out_host);
    // both halves in one pass: polynomial p (0 = even, 1 = odd) is the subsequence off = p, step = 2 of the same planes
    return eval_points<F>(c, d_poly, stride, 0, 2, (n + 1) / 2, F::D, 2, &z, 1, out, /*poly_off=*/1, /*shared_planes=*/true, /*limit=*/n);
}

// a10.  element i of the folded, DEEP-adjusted sequence
template <class F>
struct FoldSrc {
    const typename F::T* p;
    uint64_t stride, n;
    Ext<F> alpha, deep_value;
    __device__ __forceinline__ Ext<F> operator()(uint32_t, uint64_t i) const {
        Ext<F> e = ext_zero<F>(), o = ext_zero<F>();
        const uint64_t i0 = 2 * i, i1 = 2 * i + 1;
#pragma unroll
        for (int d = 0; d < F::D; d++) {
            if (i0 < n) e.c[d] = p[(uint64_t)d * stride + i0];
            if (i1 < n) o.c[d] = p[(uint64_t)d * stride + i1];
        }
        Ext<F> f = ext_add(e, ext_mul(alpha, o));
        if (i == 0) f = ext_sub(f, deep_value);
        return f;
    }
};
template <class F>
struct PlanesDst {
    typename F::T* p;
    uint64_t stride;
    __device__ __forceinline__ void operator()(uint32_t, uint64_t i, const Ext<F>& v) const {
#pragma unroll
        for (int d = 0; d < F::D; d++) p[(uint64_t)d * stride + i] = v.c[d];
    }
};

// next[k] = q_k for k < ceil(n/2) (q_{last} = 0 keeps the zero padding), planes of stride next_stride
template <class F>
int fri_fold(Ctx* c, const typename F::T* d_poly, uint64_t stride, uint64_t n, const typename F::T* z_host,
             const typename F::T* alpha_host, const typename F::T* d_host, typename F::T* d_next, uint64_t next_stride) {
    Ext<F> z = ext_from_host<F>(z_host), alpha = ext_from_host<F>(alpha_host);
    Ext<F> d0 = ext_from_host<F>(d_host), d1 = ext_from_host<F>(d_host + F::D);
    FoldSrc<F> src{d_poly, stride, n, alpha, ext_add(d0, ext_mul(d1, alpha))};
    const uint64_t len = (n + 1) / 2;
    if (len == 0) return MS_OK;
    return suffix_scan<ExtOps<F>, FoldSrc<F>, PlanesDst<F>>(c, src, len, 1, &z, PlanesDst<F>{d_next, next_stride});
}

// ------------------------------------------------------------------------------------------ query phase
// Quotient by (x - x1)(x - x2) = x^2 - s, s = x1^2 (x2 = -x1, fri.rs:149).  q_k only involves
// coefficients of index >= k + 2, so the linear interpolant a x + b (fri.rs:159-161) never enters a
// quotient coefficient (it only cancels the remainder): q_k = T(k+2) with T(j) = c_j + s T(j+2),
// an independent base-field scan per coordinate plane and parity.  Output is AoS (D coordinates
// per coefficient), i.e. already the little-endian byte layout of the proof dump.
// batch b = (query k, plane d, parity par): b = (k*D + d)*2 + par; all queries read the same polynomial
template <class F>
struct ParitySrc {
    const typename F::T* p;
    uint64_t stride, len;
    __device__ __forceinline__ typename F::T operator()(uint32_t b, uint64_t i) const {
        uint32_t par = b & 1u, d = (b >> 1) % F::D;
        uint64_t j = 2 * i + par;
        return j < len ? p[(uint64_t)d * stride + j] : (typename F::T)0;
    }
};
template <class F>
struct QuotDst {
    typename F::T* out;  // [query][nq][D] AoS
    uint64_t nq;
    // sequence position i <-> coefficient index j = 2i + parity; the scan's value T(j + 2) is q_j
    __device__ __forceinline__ void operator()(uint32_t b, uint64_t i, typename F::T v) const {
        uint32_t par = b & 1u, d = (b >> 1) % F::D, k = (b >> 1) / F::D;
        uint64_t j = 2 * i + par;
        if (j < nq) out[((uint64_t)k * nq + j) * F::D + d] = v;
    }
};

// quotients of one FRI round for `nq_queries` queries; s_host[k] = x1_k^2; out: [queries][len-2][D]
template <class F>
int fri_query_quotients(Ctx* c, const typename F::T* d_poly, uint64_t stride, uint64_t len, const typename F::T* s_host,
                        uint32_t queries, typename F::T* d_out_aos) {
    if (len < 3 || queries == 0) return MS_OK;
    const uint64_t nq = len - 2;
    std::vector<typename F::T> mus((size_t)queries * F::D * 2);
    for (uint32_t k = 0; k < queries; k++)
        for (int i = 0; i < F::D * 2; i++) mus[(size_t)k * F::D * 2 + i] = s_host[k];
    return suffix_scan<BaseOps<F>, ParitySrc<F>, QuotDst<F>>(c, ParitySrc<F>{d_poly, stride, len}, (len + 1) / 2, queries * F::D * 2,
                                                             mus.data(), QuotDst<F>{d_out_aos, nq});
}

// first index (natural order) whose extension value equals each target (merkle.rs:216-225)
template <class F>
__global__ void k_find_first(const typename F::T* __restrict__ cw, uint64_t stride, uint64_t n, const Ext<F>* __restrict__ targets,
                             int nt, unsigned long long* __restrict__ found) {
    extern __shared__ unsigned char sm_raw[];
    Ext<F>* tg = reinterpret_cast<Ext<F>*>(sm_raw);
    for (int i = threadIdx.x; i < nt; i += blockDim.x) tg[i] = targets[i];
    __syncthreads();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ext<F> v;
#pragma unroll
    for (int d = 0; d < F::D; d++) v.c[d] = cw[(uint64_t)d * stride + i];
    for (int k = 0; k < nt; k++)
        if (ext_eq(v, tg[k]) && (unsigned long long)i < found[k]) atomicMin(&found[k], (unsigned long long)i);
}

// gather: values[k] = codeword[idx[k]] (AoS), for the y1/y2/y3 look-ups and leaf neighbours
template <class F>
__global__ void k_gather_ext(const typename F::T* __restrict__ cw, uint64_t stride, const unsigned long long* __restrict__ idx, int n,
                             Ext<F>* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Ext<F> v;
#pragma unroll
    for (int d = 0; d < F::D; d++) v.c[d] = cw[(uint64_t)d * stride + idx[k]];
    out[k] = v;
}

// authentication path of a binary (2,2) tree: for level l = 0.. the sibling PAIR containing node
// (leaf_idx/2) >> l of that level (merkle.rs:241-265).  out: [nt][levels-1][2][8] words.
__global__ void k_gather_paths(const uint32_t* __restrict__ nodes, uint64_t n_groups, int path_len,
                               const unsigned long long* __restrict__ leaf_idx, int nt, uint32_t* __restrict__ out) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    int per = path_len * 16;
    if (tid >= nt * per) return;
    int k = tid / per, rem = tid % per, l = rem / 16, w = rem % 16;
    uint64_t node = (uint64_t)(leaf_idx[k] >> 1) >> l;
    uint64_t off = 0, lv = n_groups;
    for (int i = 0; i < l; i++) {
        off += lv;
        lv >>= 1;
    }
    uint64_t pair = node & ~1ULL;
    out[tid] = nodes[(off + pair) * 8 + w];
}

// The same for a tree whose nodes are spread over the ranks (row-sharded FRI tree, prover.cuh): level l holds
// n_groups >> l digests; while that is more than `world`, rank g owns the contiguous share g of the level, stored in
// its arena at arena_off in local level order (n_groups/world, n_groups/(2 world), ..., 1 digests); the levels from
// `world` digests upward are replicated in `top` (world, world/2, ..., 1 digests).  Peer digests come over NVLink.
__global__ void k_gather_paths_sharded(PeerTable arenas, uint64_t arena_off, int world, const uint32_t* __restrict__ top,
                                       uint64_t n_groups, int path_len, const unsigned long long* __restrict__ leaf_idx, int nt,
                                       uint32_t* __restrict__ out) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    int per = path_len * 16;
    if (tid >= nt * per) return;
    int k = tid / per, rem = tid % per, l = rem / 16, w = rem % 16;
    const uint64_t node = (uint64_t)(leaf_idx[k] >> 1) >> l, pair = node & ~1ULL, idx = pair + (uint64_t)(w >> 3);
    const uint64_t lv = n_groups >> l;
    uint32_t v;
    if (lv > (uint64_t)world) {
        const uint64_t share = lv / (uint64_t)world, g = idx / share, loc = idx % share;
        uint64_t off = 0, s = n_groups / (uint64_t)world;
        for (int i = 0; i < l; i++) { off += s; s >>= 1; }
        v = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(arenas.p[g]) + arena_off)[(off + loc) * 8 + (w & 7)];
    } else {
        uint64_t off = 0, s = (uint64_t)world;
        while (s > lv) { off += s; s >>= 1; }
        v = top[(off + idx) * 8 + (w & 7)];
    }
    out[tid] = v;
}

// One round of Fri::query_phase (src/fri.rs:132-176) for the betas of the transcript: the three points per query
// (y1, y2 = prev.poly at x1 = g^beta, x2 = -x1; y3 = next.poly at x3 = x1^2: evaluations at domain points are codeword
// entries, so they are gathered, not re-evaluated), and the two Merkle openings by value search (first leaf equal to y,
// src/merkle.rs:216-225) with leaf neighbours and sibling pairs (src/merkle.rs:230-265).  Results arrive in host memory
// through the mapped staging buffer; the stream is synchronised once per step (look-ups, then neighbours + paths).
template <class F>
struct QueryLookups {
    std::vector<typename F::T> x1, x2, x3, s2;     // [q] domain points, s2 = x1^2
    std::vector<Ext<F>> ys;                          // [3q]: y1,y2 interleaved per query (2k, 2k+1), then the q y3's
    std::vector<unsigned long long> found;           // [2q] leaf index of y1, y2
    std::vector<Ext<F>> neigh;                       // [4q] the two leaves of each found leaf's group
    std::vector<uint32_t> paths;                     // [2q][path_len][2][8] digest words
    int path_len = 0;
};
struct ShardedNodes {  // where a row-sharded tree's digests live (k_gather_paths_sharded); world == 0: not sharded
    PeerTable arenas{};
    uint64_t arena_off = 0;
    int world = 0;
    const uint32_t* top = nullptr;
};
// leaf neighbours of the found leaves, indices taken on the device: out[2m], out[2m+1] = codeword[found[m] & ~1], [found[m] | 1]
// (a target that was not found reads index 0: the host reports MS_ERR_LEAF_NOT_FOUND from `found` itself)
template <class F>
__global__ void k_gather_neigh(const typename F::T* __restrict__ cw, uint64_t stride, uint64_t n, const unsigned long long* __restrict__ found,
                               int nt, Ext<F>* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * nt) return;
    unsigned long long f = found[k >> 1];
    if (f >= n) f = 0;
    const unsigned long long i = (k & 1) ? (f | 1ULL) : (f & ~1ULL);
    Ext<F> v;
#pragma unroll
    for (int d = 0; d < F::D; d++) v.c[d] = cw[(uint64_t)d * stride + i];
    out[k] = v;
}
__global__ void k_clamp_found(unsigned long long* found, int nt, unsigned long long n, const unsigned long long* __restrict__ raw) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nt) found[k] = raw[k] >= n ? 0 : raw[k];
}

// Bytes one round's look-up results take in the device result buffer / the staging buffer:
// [ys 3q E | found 2q u64 | neigh 4q E | paths 2q * path_len * 64]
template <class F>
inline size_t fri_lookup_bytes(uint64_t nd, uint64_t QF) {
    return 7 * QF * sizeof(Ext<F>) + 2 * QF * 8 + (size_t)2 * QF * (size_t)ilog2(nd / 2) * 64;
}
// Queue one round's look-ups on the stream, results into d_res (fri_lookup_bytes long); no host round trip: the
// neighbour and path gathers take the found indices from device memory.  Host-side values (points, s2) go to `out`.
template <class F>
int fri_query_lookups_enqueue(Ctx* c, const typename F::T* d_prev_cw, uint64_t prev_stride, uint64_t nd, const uint32_t* d_prev_nodes,
                              const typename F::T* d_next_cw, uint64_t next_stride, uint64_t next_domain, const uint64_t* betas, uint64_t QF,
                              QueryLookups<F>* out, const ShardedNodes* sharded, uint8_t* d_res) {
    using T = typename F::T;
    using E = Ext<F>;
    if (!is_pow2(nd) || nd < 2 || next_domain * 2 != nd) return fail(c, MS_ERR_BAD_SHAPE, "FRI query: domains %llu -> %llu", (unsigned long long)nd, (unsigned long long)next_domain);
    const T g_prev = root_of_unity<F>(ilog2(nd)), g_next = root_of_unity<F>(ilog2(next_domain));
    std::vector<unsigned long long> idx3(3 * QF);
    out->x1.resize(QF); out->x2.resize(QF); out->x3.resize(QF); out->s2.resize(QF);
    for (uint64_t k = 0; k < QF; k++) {
        uint64_t beta = betas[k];
        if (beta > nd) beta %= nd;                                                   // fri.rs:144 (strict >)
        out->x1[k] = fpow<F>(g_prev, beta);                                          // fri.rs:148
        out->x2[k] = fpow<F>(g_prev, next_domain + beta);                            // fri.rs:149
        out->x3[k] = fpow<F>(g_next, beta);                                          // fri.rs:150
        out->s2[k] = F::mul(out->x1[k], out->x1[k]);
        idx3[2 * k] = beta % nd;
        idx3[2 * k + 1] = (next_domain + beta) % nd;
        idx3[2 * QF + k] = beta % next_domain;
    }
    const int path_len = ilog2(nd / 2);
    out->path_len = path_len;
    E* d_ys = reinterpret_cast<E*>(d_res);
    unsigned long long* d_found = reinterpret_cast<unsigned long long*>(d_res + 3 * QF * sizeof(E));
    E* d_neigh = reinterpret_cast<E*>(d_res + 3 * QF * sizeof(E) + 2 * QF * 8);
    uint32_t* d_paths = reinterpret_cast<uint32_t*>(d_res + 7 * QF * sizeof(E) + 2 * QF * 8);
    Scratch d_idx(c), d_fc(c);
    MS_TRY(d_idx.alloc(3 * QF * 8));
    MS_TRY(d_fc.alloc(2 * QF * 8));
    MS_TRY(stage_from_host(c, idx3.data(), idx3.size() * 8, d_idx.p));
    // y1, y2 = prev.poly(x1), prev.poly(x2); y3 = next.poly(x3): evaluations at domain points are codeword entries (exact
    // arithmetic), so they are gathered instead of re-evaluated (fri.rs:151-153)
    k_gather_ext<F><<<(unsigned)((2 * QF + 127) / 128), 128, 0, c->stream>>>(d_prev_cw, prev_stride, d_idx.as<unsigned long long>(), (int)(2 * QF), d_ys);
    MS_LAUNCH_CHECK(c);
    k_gather_ext<F><<<(unsigned)((QF + 127) / 128), 128, 0, c->stream>>>(d_next_cw, next_stride, d_idx.as<unsigned long long>() + 2 * QF, (int)QF, d_ys + 2 * QF);
    MS_LAUNCH_CHECK(c);
    // openings by value search: first leaf equal to y (merkle.rs:216-225), for y1 and y2 of each query
    MS_CUDA(c, cudaMemsetAsync(d_found, 0xff, 2 * QF * 8, c->stream));
    k_find_first<F><<<(unsigned)((nd + 255) / 256), 256, 2 * QF * sizeof(E), c->stream>>>(d_prev_cw, prev_stride, nd, d_ys, (int)(2 * QF), d_found);
    MS_LAUNCH_CHECK(c);
    k_gather_neigh<F><<<(unsigned)((4 * QF + 127) / 128), 128, 0, c->stream>>>(d_prev_cw, prev_stride, nd, d_found, (int)(2 * QF), d_neigh);
    MS_LAUNCH_CHECK(c);
    if (path_len) {
        k_clamp_found<<<(unsigned)((2 * QF + 127) / 128), 128, 0, c->stream>>>(d_fc.as<unsigned long long>(), (int)(2 * QF), nd, d_found);
        MS_LAUNCH_CHECK(c);
        int total = (int)(2 * QF) * path_len * 16;
        if (sharded && sharded->world > 1)
            k_gather_paths_sharded<<<(total + 255) / 256, 256, 0, c->stream>>>(sharded->arenas, sharded->arena_off, sharded->world, sharded->top, nd / 2,
                                                                              path_len, d_fc.as<unsigned long long>(), (int)(2 * QF), d_paths);
        else
            k_gather_paths<<<(total + 255) / 256, 256, 0, c->stream>>>(d_prev_nodes, nd / 2, path_len, d_fc.as<unsigned long long>(), (int)(2 * QF), d_paths);
        MS_LAUNCH_CHECK(c);
    }
    return MS_OK;
}
// the staged copy of a round's results (host memory) -> out; MS_ERR_LEAF_NOT_FOUND like generate_proof (merkle.rs:216-225)
template <class F>
int fri_query_lookups_parse(Ctx* c, const uint8_t* h_res, uint64_t nd, uint64_t QF, QueryLookups<F>* out) {
    using E = Ext<F>;
    out->ys.resize(3 * QF); out->found.resize(2 * QF); out->neigh.resize(4 * QF);
    out->paths.assign((size_t)2 * QF * out->path_len * 16, 0);
    memcpy(out->ys.data(), h_res, 3 * QF * sizeof(E));
    memcpy(out->found.data(), h_res + 3 * QF * sizeof(E), 2 * QF * 8);
    memcpy(out->neigh.data(), h_res + 3 * QF * sizeof(E) + 2 * QF * 8, 4 * QF * sizeof(E));
    if (out->path_len) memcpy(out->paths.data(), h_res + 7 * QF * sizeof(E) + 2 * QF * 8, out->paths.size() * 4);
    for (auto f : out->found)
        if (f >= nd) return fail(c, MS_ERR_LEAF_NOT_FOUND, "leaf is not included in the tree");
    return MS_OK;
}
// one round, synchronously (the stage export ms_fri_query)
template <class F>
int fri_query_lookups(Ctx* c, const typename F::T* d_prev_cw, uint64_t prev_stride, uint64_t nd, const uint32_t* d_prev_nodes,
                      const typename F::T* d_next_cw, uint64_t next_stride, uint64_t next_domain, const uint64_t* betas, uint64_t QF,
                      QueryLookups<F>* out, const ShardedNodes* sharded = nullptr) {
    const size_t bytes = fri_lookup_bytes<F>(nd, QF);
    Scratch res(c);
    MS_TRY(res.alloc(bytes));
    MS_TRY(fri_query_lookups_enqueue<F>(c, d_prev_cw, prev_stride, nd, d_prev_nodes, d_next_cw, next_stride, next_domain, betas, QF, out, sharded,
                                        res.as<uint8_t>()));
    MS_TRY(stage_to_host(c, 0, res.p, bytes));
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    return fri_query_lookups_parse<F>(c, c->hstage, nd, QF, out);
}

}  // namespace ms
