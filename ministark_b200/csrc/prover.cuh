// prover.cuh -- Stark::prove behind the AIR (src/starks.rs:59-169) and Fri::prove (src/fri.rs:53-189):
// the host walks the reference's sequence of transcript operations; between two transcript steps
// all bulk data stays on the device.  What crosses the PCIe bus per step is a 32-byte root, a few
// field elements, and at the very end the proof (whose size is dominated by the per-query FRI
// quotient polynomials the reference puts in it, fri.rs:167).
//
// Multi-GPU: with a communicator bound to the context (comm.cuh; ms_comm_init_nccl /
// ms_comm_init_local) every rank runs this same function as one replica of the prover and the stages that shard
// are split over the ranks INSIDE it -- no host callbacks:
//   trace tree        row ranges -> subtree digests -> all-gather -> join                      (starks.rs:70-72)
//   iNTT              trace columns [a_g, b_g) per rank, into the rank's peer-readable arena    (air.rs:147-160)
//   constraints       constraint rows [ta_g, tb_g) per rank; the trace columns a row uses are read from their owners'
//                     arenas inside the kernel                                                  (air.rs:130-134)
//   LDE + its tree    each rank extends the columns it owns into its arena; rank h hashes rows [h L/G, (h+1) L/G) with
//                     one pointer per column into the owners' arenas (NVLink loads under the SHA-256 arithmetic, no
//                     exchange pass), climbs its subtree; digests all-gathered and joined        (starks.rs:82-94)
//   mixing            partial sums over the owned columns, summed from the arenas               (starks.rs:108-117)
//   DEEP openings     owned columns only; Q*C extension elements all-gathered                   (starks.rs:140-151)
//   FRI round trees   codeword replicated (one polynomial), leaf groups row-sharded for rounds with >= 2^16 leaves:
//                     subtree per rank kept in its arena, the G roots all-gathered, top log2 G levels replicated;
//                     authentication paths are gathered from the owners' arenas                 (fri.rs:345-352)
//   proof download    every rank computes and downloads its share of the quotient polynomials into one shared host
//                     buffer over its own PCIe link (MS_PROOF_SHARED)                            (fri.rs:167)
// The transcript, the fold and the deep coefficients run identically on every rank (one polynomial, serial
// challenges), so all ranks derive the same challenges without a broadcast.  Whenever a split does not exist for a
// shape (tree not divisible over the ranks, leaf groups spanning several rows ...) that stage falls back to the
// replicated computation on every rank: the proof bytes never depend on the number of GPUs.
#pragma once
#include <algorithm>
#include <chrono>
#include <utility>

#include "comm.cuh"
#include "common.cuh"
#include "field.cuh"
#include "fri.cuh"
#include "merkle.cuh"
#include "ntt.cuh"
#include "poly.cuh"
#include "transcript.hpp"

namespace ms {

struct ProverState {
    std::vector<std::pair<const char*, float>> timings;
};

// Per-stage device times without stalling the pipeline: event pairs are recorded on the stream and only read
// back once, after the proof is complete (a cudaEventSynchronize per stage would cost a host round trip each).
struct StageTimer {
    Ctx* c;
    ProverState* ps;
    struct Span { const char* name; cudaEvent_t a, b; };
    std::vector<Span> spans;
    StageTimer(Ctx* ctx, ProverState* p) : c(ctx), ps(p) {}
    ~StageTimer() {
        for (auto& s : spans) {
            cudaEventDestroy(s.a);
            if (s.b) cudaEventDestroy(s.b);
        }
    }
    void begin(const char* n) {
        Span s{n, nullptr, nullptr};
        cudaEventCreate(&s.a);
        cudaEventRecord(s.a, c->stream);
        spans.push_back(s);
    }
    void end() {
        if (spans.empty() || spans.back().b) return;
        cudaEventCreate(&spans.back().b);
        cudaEventRecord(spans.back().b, c->stream);
    }
    void collect() {  // after the stream has been synchronised
        for (auto& s : spans) {
            if (!s.b) continue;
            float ms_ = 0;
            if (cudaEventElapsedTime(&ms_, s.a, s.b) != cudaSuccess) continue;
            bool found = false;
            for (auto& e : ps->timings)
                if (e.first == s.name) { e.second += ms_; found = true; break; }
            if (!found) ps->timings.emplace_back(s.name, ms_);
        }
    }
};

struct ProofWriter {
    uint8_t* out;
    uint64_t cap, pos = 0;
    bool mute = false;  // advance only (another rank writes these bytes into the shared buffer)
    ProofWriter(uint8_t* o, uint64_t c) : out(o), cap(c) {}
    void bytes(const void* p, size_t n) {
        if (out && !mute && pos + n <= cap) memcpy(out + pos, p, n);
        pos += n;
    }
    void u64(uint64_t v) { bytes(&v, 8); }  // little-endian host
    void u32(uint32_t v) { bytes(&v, 4); }
    uint64_t reserve(size_t n) {  // returns the offset of a region filled later (D2H)
        uint64_t at = pos;
        pos += n;
        return at;
    }
    bool fits() const { return mute || (out && pos <= cap); }
};

template <class F>
struct FriRoundDev {
    typename F::T* poly = nullptr;   // D planes x npad
    typename F::T* cw = nullptr;     // D planes x domain
    uint32_t* nodes = nullptr;       // (2,2) tree, domain - 1 digests (replicated rounds)
    ShardedNodes sh;                 // row-sharded rounds: where the digests live
    uint64_t npad = 0, domain = 0, len = 0;
    uint8_t root[32];
};

template <class F>
static void ser_ext(ProofWriter& w, const Ext<F>& e) {
    w.bytes(e.c, sizeof(typename F::T) * F::D);  // ark compressed: LE, ceil(bits/8) bytes per coordinate
}

// upper bound of the proof dump size for a given shape (so callers can allocate once)
template <class F>
uint64_t proof_size_bound(const ms_stark_params& p, const StarkDerived& d, uint64_t n, uint64_t cols) {
    const uint64_t E = sizeof(typename F::T) * F::D;
    uint64_t sz = 8 + 8 + 8 + 64 * d.rounds + 64 + 16 + d.constrain_queries * cols * E + 8 + d.constrain_queries * E + 8;
    uint64_t npad = n, domain = n * p.blowup_factor;
    for (uint64_t i = 0; i + 1 < d.rounds; i++) {
        uint64_t path_len = domain >= 2 ? (uint64_t)ilog2(domain / 2) : 0;
        uint64_t per_q = 6 * E + 2 * (8 + 2 * E + 8 + path_len * (8 + 64)) + 8 + npad * E;
        sz += 8 + d.fri_queries * per_q;
        npad = npad > 1 ? npad / 2 : 1;
        domain /= 2;
    }
    return sz;
}

constexpr uint64_t FRI_SHARD_MIN_DOMAIN = 1ULL << 16;

// digest words on the device -> root bytes on the host (one small D2H + sync)
inline int read_root(Ctx* c, const uint32_t* d_digest, uint8_t* root32) {
    uint32_t w[8];
    MS_CUDA(c, cudaMemcpyAsync(w, d_digest, 32, cudaMemcpyDeviceToHost, c->stream));
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    digest_words_to_bytes(w, root32);
    return MS_OK;
}

// A rank's rows of a tree: leaf groups of the aligned row range -> subtree digests -> all-gather -> join.
// `gather`: d_data is a device table of plane pointers (possibly peer memory), each at the first row of the range.
template <class F>
int sharded_tree_root(Ctx* c, Comm* cm, const typename F::T* d_data, uint64_t stride, uint64_t rows, uint64_t width, uint64_t lpn,
                      uint64_t k, uint64_t left, bool gather, uint8_t* root32) {
    Scratch mine(c), all(c);
    MS_TRY(mine.alloc((left ? left : 1) * 32));
    MS_TRY(all.alloc((size_t)cm->world * left * 32));
    uint64_t got = 0;
    MS_TRY(merkle_subtree<F>(c, d_data, stride, rows, width, 1, lpn, k, mine.as<uint32_t>(), &got, gather));
    if (got != left) return fail(c, MS_ERR_BAD_SHAPE, "subtree left %llu digests, plan says %llu", (unsigned long long)got, (unsigned long long)left);
    MS_TRY(cm->all_gather(c, mine.p, all.p, left * 32));
    return merkle_reduce(c, all.as<uint32_t>(), (uint64_t)cm->world * left, k, root32);
}

template <class F>
int stark_prove(Ctx* c, ProverState* ps, const ms_stark_params& p, const void* trace_rm_host, const void* d_trace_cm_in,
                uint64_t n, uint64_t w, const typename F::T* cmat_host, uint64_t t, uint8_t* proof_out, uint64_t* proof_len,
                const ms_commit_hooks* hooks = nullptr, int32_t flags = 0, const typename F::T* cconst_host = nullptr) {
    using T = typename F::T;
    using E = Ext<F>;
    constexpr int D = F::D;
    ps->timings.clear();
    if (!proof_len) return fail(c, MS_ERR_BAD_SHAPE, "proof_len is null");
    StarkDerived der;
    if (stark_derive(F::ID, p, &der) != MS_OK) return fail(c, MS_ERR_BAD_SHAPE, "bad STARK parameters (starks.rs:317-320)");
    // TraceTable::new pads to next_pow2(steps + 1) rows (air.rs:74)
    uint64_t want_n = 1;
    while (want_n < p.steps + 1) want_n <<= 1;
    if (n != want_n || w == 0) return fail(c, MS_ERR_BAD_SHAPE, "trace must have next_pow2(steps+1) = %llu rows, got %llu", (unsigned long long)want_n, (unsigned long long)n);
    const uint64_t C = w + t, B = p.blowup_factor, L = B * n;
    const uint64_t lpn = p.trace_columns, kk = p.inner_children ? p.inner_children : 2;
    const uint64_t R = der.rounds, Q = der.constrain_queries, QF = der.fri_queries;
    const uint64_t bound = proof_size_bound<F>(p, der, n, C);
    // ---- multi-GPU plan: identical on every rank (it only depends on the shape), decided before any collective
    Comm* cm = (c->comm && c->comm->world > 1 && !hooks) ? c->comm : nullptr;
    const int G = cm ? cm->world : 1, rank = cm ? cm->rank : 0;
    if (G > MS_MAX_RANKS) return fail(c, MS_ERR_UNSUPPORTED, "at most %d ranks", MS_MAX_RANKS);
    const int mask = cm ? c->shard_mask : 0;
    uint64_t tt_per = 0, tt_left = 0, lt_per = 0, lt_left = 0;
    const bool sh_trace = (mask & MS_SHARD_TRACE_TREE) && lpn && (n * w) % lpn == 0 && subtree_plan(n * w / lpn, kk, G, &tt_per, &tt_left) &&
                          (tt_per * lpn) % w == 0;
    const bool sh_cols = (mask & MS_SHARD_COLUMNS) && lpn == C && subtree_plan(L, kk, G, &lt_per, &lt_left);
    const bool sh_fri = (mask & MS_SHARD_FRI_TREES) && is_pow2((uint64_t)G);
    const bool dl_hooks = hooks && hooks->download_world > 1;
    const bool dl_native = cm && (flags & MS_PROOF_SHARED) && (mask & MS_SHARD_DOWNLOAD);
    const bool dl_sharded = dl_hooks || dl_native;
    const uint64_t dl_rank = dl_hooks ? (uint64_t)hooks->download_rank : (dl_native ? (uint64_t)rank : 0);
    uint64_t dl_world = dl_hooks ? (uint64_t)hooks->download_world : (dl_native ? (uint64_t)G : 1);
    // the shared host buffer sits behind one memory system: more than a handful of GPUs writing into it at once get less
    // aggregate bandwidth than four do (profiles/r02_e_*): cap the number of downloading ranks (ranks beyond it compute nothing)
    if (dl_native && c->dl_max_ranks > 0 && dl_world > (uint64_t)c->dl_max_ranks) dl_world = (uint64_t)c->dl_max_ranks;
    const bool dl_idle = dl_native && (uint64_t)rank >= dl_world;  // this rank takes no share
    // ranks whose proof bytes nobody reads still run every stage that feeds the transcript (lock step) but compute
    // and download no quotient polynomials
    const bool replica_only = !dl_sharded && ((hooks && hooks->replica_only) || (cm && rank != 0 && !(flags & MS_PROOF_ALL_RANKS)));
    if (!(replica_only && cm) && (!proof_out || *proof_len < bound)) {
        *proof_len = bound;
        return fail(c, MS_ERR_BUFFER_TOO_SMALL, "proof buffer needs %llu bytes", (unsigned long long)bound);
    }
    uint64_t wa = 0, wb = w, ta = 0, tb = t;  // owned trace columns / constraint rows
    uint64_t max_w = w, max_c = C;
    size_t off_a = 0, off_b = 0, off_m = 0;  // arena regions: trace coefficients | LDE columns, later FRI nodes | mix partial
    if (cm) {
        if (sh_cols) {
            shard_range(w, G, rank, &wa, &wb);
            shard_range(t, G, rank, &ta, &tb);
            max_w = (w + G - 1) / G;
            max_c = max_w + (t + G - 1) / G;
        }
        auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
        size_t fri_nodes = 0;
        if (sh_fri)
            for (uint64_t dom = L; dom >= FRI_SHARD_MIN_DOMAIN && dom / 2 >= (uint64_t)G; dom >>= 1) fri_nodes += align((dom / G) * 32);
        const size_t a_bytes = sh_cols ? align(max_w * n * sizeof(T)) : 0;
        const size_t b_bytes = std::max(sh_cols ? align(max_c * L * sizeof(T)) : (size_t)0, fri_nodes);
        const size_t m_bytes = sh_cols ? align(n * sizeof(T)) : 0;
        off_a = 0;
        off_b = a_bytes;
        off_m = a_bytes + b_bytes;
        if (a_bytes + b_bytes + m_bytes) MS_TRY(cm->ensure_arena(c, a_bytes + b_bytes + m_bytes));
    }
    auto arena_of = [&](int g, size_t off) -> uint8_t* { return static_cast<uint8_t*>(cm->bases[g]) + off; };
    auto w_owner = [&](uint64_t col, uint64_t* first) -> int {
        const int g = shard_owner(w, G, col);
        uint64_t a, b;
        shard_range(w, G, g, &a, &b);
        *first = a;
        return g;
    };
    // The LDE is linear, and every constraint column is a fixed linear combination of trace columns (the T x W matrix):
    // when the matrix is sparse (the e2e / synthetic AIRs use 2-3 trace columns per constraint) the constraint columns of the
    // LDE are that same combination of the trace columns' evaluations -- one streaming pass instead of T more transforms,
    // bit-identical because the arithmetic is exact.  (Dense matrices keep the transforms.)
    std::vector<SparseRow<F>> srows;
    const bool sparse = t > 0 && sparse_rows<F>(cmat_host, t, w, &srows);
    const bool lde_by_linearity = sparse && c->lde_linearity;
    StageTimer tm(c, ps);
    IOPattern io = stark_iopattern(F::BITS, D, R, Q, QF);
    Merlin merlin(io, c->bridge_masks, c->leftover_as_published != 0);
    auto challenge_base = [&](T* out) -> bool {  // let [x]: [F::Base; 1] = merlin.challenge_scalars()
        uint64_t v;
        if (!merlin.challenge_scalars(F::BITS, (uint64_t)F::P, 1, 1, &v)) return false;
        *out = (T)v;
        return true;
    };
    auto challenge_exts = [&](E* out, size_t count) -> bool {  // one fill_challenge_scalars call for `count` scalars
        std::vector<uint64_t> v(count * D);
        if (!merlin.challenge_scalars(F::BITS, (uint64_t)F::P, D, count, v.data())) return false;
        for (size_t i = 0; i < count; i++)
            for (int d = 0; d < D; d++) out[i].c[d] = (T)v[i * D + d];
        return true;
    };
    auto challenge_ext = [&](E* out) -> bool { return challenge_exts(out, 1); };
#define TR(expr) do { if (!(expr)) return fail(c, MS_ERR_TRANSCRIPT, "transcript pattern violated at %s", #expr); } while (0)

    // ---- 1.1 trace on the device (column-major) and its commitment           starks.rs:68-73
    tm.begin("upload+transpose");
    Scratch trace_cm(c), trace_rm(c);
    const T* d_trace = reinterpret_cast<const T*>(d_trace_cm_in);
    if (!d_trace) {
        MS_TRY(trace_rm.alloc(n * w * sizeof(T)));
        MS_TRY(trace_cm.alloc(n * w * sizeof(T)));
        MS_CUDA(c, cudaMemcpyAsync(trace_rm.p, trace_rm_host, n * w * sizeof(T), cudaMemcpyHostToDevice, c->stream));
        MS_TRY(transpose<F>(c, trace_rm.as<T>(), trace_cm.as<T>(), n, w, true));
        d_trace = trace_cm.as<T>();
    }
    tm.end();
    tm.begin("trace_commit");
    uint8_t trace_root[32], lde_root[32];
    if (hooks && hooks->trace_commit) {
        int rc = hooks->trace_commit(hooks->user, d_trace, n, w, trace_root);
        if (rc != MS_OK) return fail(c, rc, "trace_commit hook failed");
    } else if (cm && sh_trace) {
        const uint64_t rows = tt_per * lpn / w;
        MS_TRY(sharded_tree_root<F>(c, cm, d_trace + (uint64_t)rank * rows, n, rows, w, lpn, kk, tt_left, false, trace_root));
    } else {
        MS_TRY(merkle_commit<F>(c, d_trace, n, n, w, 1, lpn, kk, nullptr, trace_root));
    }
    tm.end();
    TR(merlin.add_bytes(trace_root, 32));
    // ---- 1.2 LDE of trace + constraint polynomials and its commitment            starks.rs:80-95
    T shift;
    TR(challenge_base(&shift));
    if (shift == 0) return fail(c, MS_ERR_BAD_SHAPE, "random_shift is zero (get_coset unwrap, starks.rs:84-85)");
    Scratch coeffs(c), lde(c), ptab(c);
    T* d_mixed = nullptr;
    // column-sharded: the owned trace coefficient columns live in the arena (peers read them), the owned constraint
    // columns and the mixed polynomial in local memory
    T* my_trace_coef = nullptr;  // [wb - wa][n]
    T* my_cons_coef = nullptr;   // [tb - ta][n]
    const uint64_t nw = wb - wa, nt = tb - ta;
    if (cm && sh_cols) {
        my_trace_coef = reinterpret_cast<T*>(arena_of(rank, off_a));
        MS_TRY(coeffs.alloc((nt + 1) * n * sizeof(T)));
        my_cons_coef = coeffs.as<T>();
        d_mixed = coeffs.as<T>() + nt * n;
        tm.begin("intt");
        MS_TRY(lde_batch<F>(c, d_trace + wa * n, n, nw, ilog2(n), 0, (T)1, true, my_trace_coef, n));  // air.rs:147-160, own columns
        tm.end();
        tm.begin("constraints");
        MS_TRY(cm->barrier(c));  // every rank's trace coefficients are in its arena
        if (nt) {
            std::vector<const T*> cols(w);
            for (uint64_t j = 0; j < w; j++) {
                uint64_t first;
                const int g = w_owner(j, &first);
                cols[j] = reinterpret_cast<const T*>(arena_of(g, off_a)) + (j - first) * n;
            }
            MS_TRY(ptab.alloc(w * sizeof(void*)));
            MS_TRY(stage_from_host(c, cols.data(), w * sizeof(void*), ptab.p));
            if (sparse) MS_TRY(linear_sparse<F>(c, ptab.as<const T*>(), n, std::vector<SparseRow<F>>(srows.begin() + ta, srows.begin() + tb), my_cons_coef, n));
            else MS_TRY(linear_constraints_gather<F>(c, ptab.as<const T*>(), n, w, cmat_host + ta * w, nt, my_cons_coef, n));  // air.rs:130-134, own rows
            MS_TRY(add_consts<F>(c, my_cons_coef, n, n, cconst_host ? cconst_host + ta : nullptr, nt, false));
        }
        tm.end();
    } else {
        MS_TRY(coeffs.alloc((C + 1) * n * sizeof(T)));  // C constraint columns + the mixed polynomial
        d_mixed = coeffs.as<T>() + C * n;
        tm.begin("intt");
        MS_TRY(lde_batch<F>(c, d_trace, n, w, ilog2(n), 0, (T)1, true, coeffs.as<T>(), n));  // air.rs:147-160
        tm.end();
        tm.begin("constraints");
        if (sparse) {
            MS_TRY(column_table<F>(c, coeffs.as<T>(), n, w, &ptab));
            MS_TRY(linear_sparse<F>(c, ptab.as<const T*>(), n, srows, coeffs.as<T>() + w * n, n));  // air.rs:130-134
        } else {
            MS_TRY(linear_constraints<F>(c, coeffs.as<T>(), n, n, w, cmat_host, t, coeffs.as<T>() + w * n, n));  // air.rs:130-134
        }
        MS_TRY(add_consts<F>(c, coeffs.as<T>() + w * n, n, n, cconst_host, t, false));
        tm.end();
    }
    if (hooks && hooks->lde_commit) {
        tm.begin("lde+commit(hook)");
        int rc = hooks->lde_commit(hooks->user, coeffs.as<T>(), n, C, B, (uint64_t)shift, lde_root);
        if (rc != MS_OK) return fail(c, rc, "lde_commit hook failed");
        tm.end();
    } else if (cm && sh_cols) {
        T* my_lde = reinterpret_cast<T*>(arena_of(rank, off_b));  // [nw + nt][L]: own trace columns, then own constraint columns
        tm.begin("lde");
        MS_TRY(lde_batch<F>(c, my_trace_coef, n, nw, ilog2(n), ilog2(B), shift, false, my_lde, L));                 // starks.rs:87-91
        if (lde_by_linearity) MS_TRY(cm->barrier(c));  // (every rank: collectives are never conditional on a rank's share)
        if (lde_by_linearity && nt) {
            // the trace columns' evaluations of every rank are complete: combine them pointwise
            std::vector<const T*> cols(w);
            for (uint64_t j = 0; j < w; j++) {
                uint64_t first;
                const int g = w_owner(j, &first);
                cols[j] = reinterpret_cast<const T*>(arena_of(g, off_b)) + (j - first) * L;
            }
            Scratch etab(c);
            MS_TRY(etab.alloc(w * sizeof(void*)));
            MS_TRY(stage_from_host(c, cols.data(), w * sizeof(void*), etab.p));
            MS_TRY(linear_sparse<F>(c, etab.as<const T*>(), L, std::vector<SparseRow<F>>(srows.begin() + ta, srows.begin() + tb), my_lde + nw * L, L));
            MS_TRY(add_consts<F>(c, my_lde + nw * L, L, L, cconst_host ? cconst_host + ta : nullptr, nt, true));
        } else if (!lde_by_linearity) {
            MS_TRY(lde_batch<F>(c, my_cons_coef, n, nt, ilog2(n), ilog2(B), shift, false, my_lde + nw * L, L));
        }
        tm.end();
        tm.begin("lde_commit");
        MS_TRY(cm->barrier(c));  // every rank's columns are complete before anyone reads them
        const uint64_t rows = lt_per;  // lpn == C: one row per leaf group
        std::vector<const T*> planes(C);
        for (uint64_t col = 0; col < C; col++) {
            int g;
            uint64_t local;
            if (col < w) {
                uint64_t first;
                g = w_owner(col, &first);
                local = col - first;
            } else {
                g = shard_owner(t, G, col - w);
                uint64_t a, b, a2, b2;
                shard_range(t, G, g, &a, &b);
                shard_range(w, G, g, &a2, &b2);
                local = (b2 - a2) + (col - w - a);
            }
            planes[col] = reinterpret_cast<const T*>(arena_of(g, off_b)) + local * L + (uint64_t)rank * rows;
        }
        Scratch ltab(c);
        MS_TRY(ltab.alloc(C * sizeof(void*)));
        MS_TRY(stage_from_host(c, planes.data(), C * sizeof(void*), ltab.p));
        // the all-gather inside also tells every rank that its columns have been read (the region is reused below)
        MS_TRY(sharded_tree_root<F>(c, cm, ltab.as<T>(), 0, rows, C, lpn, kk, lt_left, true, lde_root));  // starks.rs:92-94
        tm.end();
    } else {
        MS_TRY(lde.alloc(C * L * sizeof(T)));
        tm.begin("lde");
        if (lde_by_linearity) {
            MS_TRY(lde_batch<F>(c, coeffs.as<T>(), n, w, ilog2(n), ilog2(B), shift, false, lde.as<T>(), L));  // starks.rs:87-91, trace columns
            Scratch etab(c);
            MS_TRY(column_table<F>(c, lde.as<T>(), L, w, &etab));
            MS_TRY(linear_sparse<F>(c, etab.as<const T*>(), L, srows, lde.as<T>() + w * L, L));               // constraint columns, pointwise
            MS_TRY(add_consts<F>(c, lde.as<T>() + w * L, L, L, cconst_host, t, true));
        } else {
            MS_TRY(lde_batch<F>(c, coeffs.as<T>(), n, C, ilog2(n), ilog2(B), shift, false, lde.as<T>(), L));  // starks.rs:87-91
        }
        tm.end();
        tm.begin("lde_commit");
        MS_TRY(merkle_commit<F>(c, lde.as<T>(), L, L, C, 1, lpn, kk, nullptr, lde_root));  // starks.rs:92-94
        tm.end();
        lde.release();  // the LDE tree is never opened by the reference
    }
    TR(merlin.add_bytes(lde_root, 32));
    // ---- 1.3 mixing                                                                  starks.rs:108-119
    T r;
    TR(challenge_base(&r));
    tm.begin("mix");
    if (cm && sh_cols) {
        T* part = reinterpret_cast<T*>(arena_of(rank, off_m));
        k_mix_parts<F><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(my_trace_coef, n, (int)nw, fpow<F>(r, wa), my_cons_coef, n, (int)nt,
                                                                          fpow<F>(r, w + ta), n, r, part);
        MS_LAUNCH_CHECK(c);
        MS_TRY(cm->barrier(c));
        PeerTable pt{};
        for (int g = 0; g < G; g++) pt.p[g] = arena_of(g, off_m);
        k_sum_peers<F><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(pt, G, n, d_mixed);
        MS_LAUNCH_CHECK(c);
    } else {
        MS_TRY(mix<F>(c, coeffs.as<T>(), n, n, C, r, d_mixed));
    }
    tm.end();
    // divide_by_vanishing_poly yields (quotient, remainder); the reference asserts quotient == 0,
    // which holds for any polynomial with <= N coefficients, and carries the remainder (= mixed) on.
    // ---- 2. DEEP-ALI queries                                                         starks.rs:124-151
    std::vector<E> zq(Q);
    TR(challenge_exts(zq.data(), Q));  // merlin.fill_challenge_scalars(&mut queries), starks.rs:124-125
    std::vector<E> opens(Q * (C + 1));
    tm.begin("deep_open");
    if (cm && sh_cols) {
        // own columns, then every rank's results gathered: rank g contributes [Q][max_c] slots (its trace columns, then
        // its constraint columns); the mixed polynomial is evaluated by everyone (one polynomial)
        std::vector<E> mine(Q * max_c, ext_zero<F>()), tmp(Q * std::max<uint64_t>(std::max(nw, nt), 1)), vq(Q);
        if (nw) {
            MS_TRY(eval_points<F>(c, my_trace_coef, n, 0, 1, n, 1, nw, zq.data(), (int)Q, tmp.data()));
            for (uint64_t q = 0; q < Q; q++)
                for (uint64_t j = 0; j < nw; j++) mine[q * max_c + j] = tmp[q * nw + j];
        }
        if (nt) {
            MS_TRY(eval_points<F>(c, my_cons_coef, n, 0, 1, n, 1, nt, zq.data(), (int)Q, tmp.data()));
            for (uint64_t q = 0; q < Q; q++)
                for (uint64_t j = 0; j < nt; j++) mine[q * max_c + nw + j] = tmp[q * nt + j];
        }
        MS_TRY(eval_points<F>(c, d_mixed, n, 0, 1, n, 1, 1, zq.data(), (int)Q, vq.data()));
        Scratch ds(c), dr(c);
        const size_t blk = Q * max_c * sizeof(E);
        MS_TRY(ds.alloc(blk));
        MS_TRY(dr.alloc(blk * G));
        MS_CUDA(c, cudaMemcpyAsync(ds.p, mine.data(), blk, cudaMemcpyHostToDevice, c->stream));
        MS_TRY(cm->all_gather(c, ds.p, dr.p, blk));
        std::vector<E> all(Q * max_c * G);
        MS_CUDA(c, cudaMemcpyAsync(all.data(), dr.p, blk * G, cudaMemcpyDeviceToHost, c->stream));
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int g = 0; g < G; g++) {
            uint64_t a, b, a2, b2;
            shard_range(w, G, g, &a, &b);
            shard_range(t, G, g, &a2, &b2);
            for (uint64_t q = 0; q < Q; q++) {
                for (uint64_t j = a; j < b; j++) opens[q * (C + 1) + j] = all[((uint64_t)g * Q + q) * max_c + (j - a)];
                for (uint64_t j = a2; j < b2; j++) opens[q * (C + 1) + w + j] = all[((uint64_t)g * Q + q) * max_c + (b - a) + (j - a2)];
            }
        }
        for (uint64_t q = 0; q < Q; q++) opens[q * (C + 1) + C] = vq[q];
    } else {
        MS_TRY(eval_points<F>(c, coeffs.as<T>(), n, 0, 1, n, 1, C + 1, zq.data(), (int)Q, opens.data()));
    }
    tm.end();
    // ---- 3. FRI commit phase                                                        fri.rs:64-113
    tm.begin("fri_commit_phase");
    std::vector<FriRoundDev<F>> rounds(R);
    std::vector<Scratch> keep;
    keep.reserve(5 * R + 8);
    auto dev_alloc = [&](size_t bytes, void** out) -> int {
        keep.emplace_back(c);
        MS_TRY(keep.back().alloc(bytes));
        *out = keep.back().p;
        return MS_OK;
    };
    // coefficient counts of all rounds: computed on the device as the rounds are produced, read back once (the
    // commit phase itself only needs the padded sizes; `len` matters for the quotient lengths of the query phase)
    Scratch d_lens(c);
    MS_TRY(d_lens.alloc(8 * (R + 1)));
    MS_CUDA(c, cudaMemsetAsync(d_lens.p, 0, 8 * (R + 1), c->stream));
    auto poly_len_async = [&](const T* planes, uint64_t stride, int nplanes, uint64_t npad, uint64_t slot) -> int {
        k_poly_len<F><<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(planes, stride, nplanes, npad, d_lens.as<unsigned long long>() + slot);
        MS_LAUNCH_CHECK(c);
        return MS_OK;
    };
    {
        // round 0: the mixed polynomial lifted to the extension (field.rs:23-32): upper planes zero.
        MS_TRY(poly_len_async(d_mixed, n, 1, n, 0));
        unsigned long long len0 = 0;
        MS_CUDA(c, cudaMemcpyAsync(&len0, d_lens.p, 8, cudaMemcpyDeviceToHost, c->stream));
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        uint64_t deg = len0 ? len0 - 1 : 0;  // ark: degree() of the zero polynomial is 0
        uint64_t npad = 1;
        while (npad < deg + 1) npad <<= 1;  // Radix2EvaluationDomain::new((deg+1)*B) rounds up (fri.rs:74,315)
        FriRoundDev<F>& r0 = rounds[0];
        r0.npad = npad;
        r0.domain = npad * B;
        r0.len = len0;
        MS_TRY(dev_alloc(D * npad * sizeof(T), (void**)&r0.poly));
        MS_CUDA(c, cudaMemsetAsync(r0.poly, 0, D * npad * sizeof(T), c->stream));
        MS_CUDA(c, cudaMemcpyAsync(r0.poly, d_mixed, npad * sizeof(T), cudaMemcpyDeviceToDevice, c->stream));
    }
    size_t fri_arena_used = 0;
    for (uint64_t i = 0; i < R; i++) {
        FriRoundDev<F>& cur = rounds[i];
        if (i > 0) {
            FriRoundDev<F>& prev = rounds[i - 1];
            E z, alpha, deep[2];
            TR(challenge_ext(&z));                                                      // fri.rs:89
            MS_TRY(fri_deep_coeffs<F>(c, prev.poly, prev.npad, prev.npad, z.c, deep[0].c));  // fri.rs:90
            uint8_t dbytes[2 * 4 * 8];
            memcpy(dbytes, deep, 2 * sizeof(E));
            TR(merlin.add_bytes(dbytes, 2 * sizeof(E)));                                // fri.rs:94
            TR(challenge_ext(&alpha));                                                  // fri.rs:96
            cur.npad = prev.npad > 1 ? prev.npad / 2 : 1;
            cur.domain = prev.domain / 2;                                               // fri.rs:374-376
            if (cur.domain < 2) return fail(c, MS_ERR_BAD_SHAPE, "FRI domain exhausted at round %llu (merkle.rs:93-104 would panic)", (unsigned long long)i);
            MS_TRY(dev_alloc(D * cur.npad * sizeof(T), (void**)&cur.poly));
            MS_CUDA(c, cudaMemsetAsync(cur.poly, 0, D * cur.npad * sizeof(T), c->stream));
            MS_TRY(fri_fold<F>(c, prev.poly, prev.npad, prev.npad, z.c, alpha.c, deep[0].c, cur.poly, cur.npad));  // fri.rs:97-101
            MS_TRY(poly_len_async(cur.poly, cur.npad, D, cur.npad, i));
        }
        MS_TRY(dev_alloc(D * cur.domain * sizeof(T), (void**)&cur.cw));
        const uint64_t groups = cur.domain / 2;
        if (cm && sh_fri && cur.domain >= FRI_SHARD_MIN_DOMAIN && groups >= (uint64_t)G) {
            // codeword replicated, tree row-sharded (fri.rs:345-352): this rank's leaf groups -> its subtree (kept in the
            // arena for the authentication paths) -> the G subtree roots all-gathered -> the top levels on every rank
            const uint64_t per = groups / G, npadc = cur.npad;
            MS_TRY(lde_batch<F>(c, cur.poly, npadc, D, ilog2(npadc), ilog2(cur.domain / npadc), (T)1, false, cur.cw, cur.domain));
            uint32_t* mine = reinterpret_cast<uint32_t*>(arena_of(rank, off_b + fri_arena_used));
            MS_TRY(merkle_leaf_level<F>(c, cur.cw + (uint64_t)rank * 2 * per, cur.domain, 1, D, 2, per, mine));
            uint64_t top_at = 0;
            MS_TRY(merkle_climb(c, mine, per, 2, 1, &top_at));
            uint32_t* top = nullptr;
            MS_TRY(dev_alloc((size_t)(2 * G) * 32, (void**)&top));
            MS_TRY(cm->all_gather(c, mine + top_at * 8, top, 32));
            MS_TRY(merkle_upper_levels(c, top, (uint64_t)G, 2));
            MS_TRY(read_root(c, top + (size_t)(2 * G - 2) * 8, cur.root));
            cur.sh.world = G;
            cur.sh.arena_off = off_b + fri_arena_used;
            cur.sh.top = top;
            for (int g = 0; g < G; g++) cur.sh.arenas.p[g] = cm->bases[g];
            fri_arena_used += ((2 * per) * 32 + 255) & ~(size_t)255;
        } else {
            MS_TRY(dev_alloc((cur.domain - 1) * 32, (void**)&cur.nodes));
            MS_TRY(fri_commit<F>(c, cur.poly, cur.npad, cur.domain, cur.domain / cur.npad, cur.cw, cur.domain, cur.nodes, cur.root));  // fri.rs:345-352
        }
        if (i > 0) TR(merlin.add_bytes(cur.root, 32));                                  // fri.rs:107-108 (round 0 is not absorbed)
    }
    {
        std::vector<unsigned long long> lens(R);
        MS_CUDA(c, cudaMemcpyAsync(lens.data(), d_lens.p, 8 * R, cudaMemcpyDeviceToHost, c->stream));
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        for (uint64_t i = 1; i < R; i++) rounds[i].len = lens[i];
    }
    tm.end();
    // ---- FRI query phase                                                            fri.rs:115-189
    tm.begin("fri_query_phase");
    std::vector<uint8_t> braw(8 * QF, 0);  // vec![0u8; 8 * queries], fri.rs:121
    TR(merlin.challenge_bytes(braw.data(), braw.size()));                               // fri.rs:121-122
    std::vector<uint64_t> betas(QF);
    for (uint64_t k = 0; k < QF; k++) memcpy(&betas[k], &braw[8 * k], 8);               // usize::from_le_bytes
    // ---- serialise the fixed part (field order of starks.rs:21-28)
    // sharded download: every rank computes the same offsets, rank 0 alone writes the fixed part
    uint64_t copy_seq = 0;
    // Which rank downloads (and therefore computes) the seq-th quotient polynomial: round robin over all ranks.  (r01 left
    // rank 0 out from 4 ranks on, because its look-ups then queued behind its own downloads; they go through mapped memory
    // now, and an even split shortens the tail every other rank adds.  c->dl_skip_rank0 restores the old split.)
    auto owns = [&](uint64_t seq) -> bool {
        if (dl_idle) return false;
        if (c->dl_skip_rank0 && dl_world >= 4) return dl_rank != 0 && seq % (dl_world - 1) + 1 == dl_rank;
        return seq % dl_world == dl_rank;
    };
    ProofWriter pw(proof_out, proof_out ? *proof_len : 0);
    pw.mute = (dl_sharded && dl_rank != 0) || (replica_only && (!proof_out || *proof_len < bound));
    pw.bytes("MSTARKP1", 8);
    pw.u32((uint32_t)F::ID);
    pw.u32((uint32_t)D);
    pw.u64(merlin.transcript.size());
    pw.bytes(merlin.transcript.data(), merlin.transcript.size());
    pw.bytes(trace_root, 32);
    pw.bytes(lde_root, 32);
    pw.u64(Q);
    pw.u64(Q ? C : 0);
    for (uint64_t q = 0; q < Q; q++)
        for (uint64_t col = 0; col < C; col++) ser_ext<F>(pw, opens[q * (C + 1) + col]);
    pw.u64(Q);
    for (uint64_t q = 0; q < Q; q++) ser_ext<F>(pw, opens[q * (C + 1) + C]);
    pw.u64(R - 1);
    // The query phase in two batches (r02; r01 made two host round trips per round, 46 per proof, each of them queued behind
    // the multi-MB quotient downloads of the rounds before):
    //   1. look-ups of ALL rounds (gathers, value search, neighbours, paths: the found indices never leave the device),
    //      staged to the host in one copy, one event;
    //   2. quotient kernels and their downloads of all rounds, queued without a host round trip: the offsets in the dump
    //      only depend on sizes, so they are known before any value is;
    // the host serialises the fixed part while the quotients stream out.
    cudaEvent_t dl[2] = {nullptr, nullptr};  // first / last proof-download copy on the copy stream (timing only)
    cudaEvent_t ev_lk = nullptr;
    // every exit path below (errors included) waits for the downloads already queued into proof_out and frees the events
    struct CopyGuard {
        Ctx* c;
        cudaEvent_t* dl;
        cudaEvent_t* lk;
        ~CopyGuard() {
            cudaStreamSynchronize(c->copy_stream);
            for (int i = 0; i < 2; i++)
                if (dl[i]) { cudaEventDestroy(dl[i]); dl[i] = nullptr; }
            if (*lk) { cudaEventDestroy(*lk); *lk = nullptr; }
        }
    } copy_guard{c, dl, &ev_lk};
    uint64_t dl_bytes = 0;
    double q_host[4] = {0, 0, 0, 0};  // wall ms: queue look-ups | queue quotients + copies | wait for the look-ups | serialise
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_q = now_ms();
    const bool lookups = !pw.mute;
    const uint64_t E_ = sizeof(E);
    // ---- 1. look-ups of every round
    std::vector<QueryLookups<F>> lks(R ? R - 1 : 0);
    std::vector<size_t> lk_off(R, 0);
    size_t lk_total = 0;
    for (uint64_t i = 0; i + 1 < R; i++) {
        lk_off[i] = lk_total;
        lk_total += fri_lookup_bytes<F>(rounds[i].domain, QF);
    }
    Scratch d_lk(c);
    if (lookups && lk_total) {
        if (lk_total > HSTAGE_OUT) return fail(c, MS_ERR_UNSUPPORTED, "query look-ups need %zu staged bytes (limit %zu)", lk_total, (size_t)HSTAGE_OUT);
        MS_TRY(d_lk.alloc(lk_total));
        for (uint64_t i = 0; i + 1 < R; i++) {
            FriRoundDev<F>& prev = rounds[i];
            FriRoundDev<F>& nxt = rounds[i + 1];
            MS_TRY(fri_query_lookups_enqueue<F>(c, prev.cw, prev.domain, prev.domain, prev.nodes, nxt.cw, nxt.domain, nxt.domain, betas.data(), QF,
                                                &lks[i], &prev.sh, d_lk.as<uint8_t>() + lk_off[i]));
        }
        MS_TRY(stage_to_host(c, 0, d_lk.p, lk_total));
        MS_CUDA(c, cudaEventCreateWithFlags(&ev_lk, cudaEventDisableTiming));
        MS_CUDA(c, cudaEventRecord(ev_lk, c->stream));
    } else {
        for (uint64_t i = 0; i + 1 < R; i++) {  // ranks that only contribute quotients: the scan multipliers x1^2
            const uint64_t nd = rounds[i].domain;
            const T g_prev = root_of_unity<F>(ilog2(nd));
            lks[i].s2.resize(QF);
            lks[i].path_len = ilog2(nd / 2);
            for (uint64_t k = 0; k < QF; k++) {
                uint64_t beta = betas[k];
                if (beta > nd) beta %= nd;
                const T x1 = fpow<F>(g_prev, beta);
                lks[i].s2[k] = F::mul(x1, x1);
            }
        }
    }
    q_host[0] = now_ms() - t_q; t_q = now_ms();
    // ---- 2. quotients (fri.rs:157-167): only the ones this rank downloads (all of them unless the download is sharded),
    // each to its final offset in the dump
    uint64_t pos = pw.pos;
    for (uint64_t i = 0; i + 1 < R; i++) {
        FriRoundDev<F>& prev = rounds[i];
        const uint64_t nq = prev.len >= 3 ? prev.len - 2 : 0, path_len = (uint64_t)lks[i].path_len;
        const uint64_t per_q = 6 * E_ + 2 * (8 + 2 * E_ + 8 + path_len * (8 + 64)) + 8 + nq * E_;
        pos += 8;  // the round's Vec length
        T* d_quot = nullptr;
        std::vector<int> slot(QF, -1);  // query k's quotient is the slot[k]-th polynomial of d_quot
        if (nq) {
            std::vector<T> s2_own;
            for (uint64_t k = 0; k < QF; k++)
                if (owns(copy_seq + k) && !replica_only) {
                    slot[k] = (int)s2_own.size();
                    s2_own.push_back(lks[i].s2[k]);
                }
            copy_seq += QF;
            if (!s2_own.empty()) {
                MS_TRY(dev_alloc(s2_own.size() * nq * sizeof(E), (void**)&d_quot));
                // large rounds go query by query, so that the first download starts after one polynomial's scan instead of
                // after the whole round's (the copy engine is the bottleneck of this phase: keep it fed from the start)
                const size_t step = nq * E_ >= (8u << 20) ? 1 : s2_own.size();
                std::vector<uint64_t> owner_of_slot(s2_own.size());
                for (uint64_t k = 0; k < QF; k++)
                    if (slot[k] >= 0) owner_of_slot[slot[k]] = k;
                for (size_t s0 = 0; s0 < s2_own.size(); s0 += step) {
                    const size_t cnt = std::min(step, s2_own.size() - s0);
                    MS_TRY(fri_query_quotients<F>(c, prev.poly, prev.npad, prev.len, s2_own.data() + s0, (uint32_t)cnt, d_quot + s0 * nq * D));
                    MS_CUDA(c, cudaEventRecord(c->copy_event, c->stream));
                    MS_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->copy_event, 0));
                    if (!dl[0]) {
                        cudaEventCreate(&dl[0]);
                        cudaEventCreate(&dl[1]);
                        cudaEventRecord(dl[0], c->copy_stream);
                    }
                    for (size_t sl = s0; sl < s0 + cnt; sl++) {
                        const uint64_t at = pos + owner_of_slot[sl] * per_q + (per_q - nq * E_);
                        MS_CUDA(c, cudaMemcpyAsync(proof_out + at, d_quot + sl * nq * D, nq * E_, cudaMemcpyDeviceToHost, c->copy_stream));
                        dl_bytes += nq * E_;
                    }
                }
            }
        }
        pos += QF * per_q;
    }
    q_host[1] = now_ms() - t_q; t_q = now_ms();
    // ---- 3. the look-up results, then the dump's fixed bytes (fri.rs:18-22, merkle.rs:293-298)
    if (ev_lk) {
        MS_CUDA(c, cudaEventSynchronize(ev_lk));
        for (uint64_t i = 0; i + 1 < R; i++) MS_TRY(fri_query_lookups_parse<F>(c, c->hstage + lk_off[i], rounds[i].domain, QF, &lks[i]));
    }
    q_host[2] = now_ms() - t_q; t_q = now_ms();
    for (uint64_t i = 0; i + 1 < R; i++) {
        QueryLookups<F>& lk = lks[i];
        const uint64_t nq = rounds[i].len >= 3 ? rounds[i].len - 2 : 0;
        const int path_len = lk.path_len;
        if (!lookups) { lk.ys.assign(3 * QF, ext_zero<F>()); lk.neigh.assign(4 * QF, ext_zero<F>()); lk.paths.assign((size_t)2 * QF * path_len * 16, 0); lk.x1.assign(QF, 0); lk.x2.assign(QF, 0); lk.x3.assign(QF, 0); }
        pw.u64(QF);
        for (uint64_t k = 0; k < QF; k++) {
            E pts[6] = {ext_from_base<F>(lk.x1[k]), lk.ys[2 * k], ext_from_base<F>(lk.x2[k]), lk.ys[2 * k + 1], ext_from_base<F>(lk.x3[k]), lk.ys[2 * QF + k]};
            for (int e = 0; e < 6; e++) ser_ext<F>(pw, pts[e]);
            for (int which = 0; which < 2; which++) {
                uint64_t m = 2 * k + which;
                pw.u64(2);
                ser_ext<F>(pw, lk.neigh[2 * m]);
                ser_ext<F>(pw, lk.neigh[2 * m + 1]);
                pw.u64(path_len);
                for (int l = 0; l < path_len; l++) {
                    pw.u64(2);
                    for (int sb = 0; sb < 2; sb++) {
                        uint8_t dg[32];
                        digest_words_to_bytes(&lk.paths[((size_t)m * path_len + l) * 16 + sb * 8], dg);
                        pw.bytes(dg, 32);
                    }
                }
            }
            pw.u64(nq);
            if (nq) pw.reserve(nq * sizeof(E));  // filled by the download queued above
        }
    }
    if (pw.pos != pos) return fail(c, MS_ERR_UNSUPPORTED, "proof layout mismatch (%llu vs %llu)", (unsigned long long)pw.pos, (unsigned long long)pos);
    q_host[3] = now_ms() - t_q;
    ps->timings.emplace_back("(query host: queue look-ups)", (float)q_host[0]);
    ps->timings.emplace_back("(query host: queue quotients + copies)", (float)q_host[1]);
    ps->timings.emplace_back("(query host: wait for look-ups)", (float)q_host[2]);
    ps->timings.emplace_back("(query host: serialise)", (float)q_host[3]);
    if (dl[0]) cudaEventRecord(dl[1], c->copy_stream);
    MS_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    if (dl[0]) {
        float ms_ = 0;
        cudaEventElapsedTime(&ms_, dl[0], dl[1]);
        ps->timings.emplace_back("(proof download, copy stream)", ms_);
        ps->timings.emplace_back("(proof download GB)", (float)(dl_bytes * 1e-9));
        cudaEventDestroy(dl[0]);
        cudaEventDestroy(dl[1]);
        dl[0] = dl[1] = nullptr;
    }
    if (!pw.fits()) {
        *proof_len = pw.pos;
        return fail(c, MS_ERR_BUFFER_TOO_SMALL, "proof buffer needs %llu bytes", (unsigned long long)pw.pos);
    }
    tm.end();
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    // every rank's share of a shared proof buffer has landed, and nobody still reads this rank's arena
    if (cm) MS_TRY(cm->host_barrier(c));
    tm.collect();
    *proof_len = pw.pos;
#undef TR
    return MS_OK;
}

}  // namespace ms
