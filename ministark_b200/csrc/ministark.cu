// ministark.cu -- C ABI (include/ministark.h) over the sm_100a kernels.  No torch types, no CPU
// fallback: every entry point either runs the CUDA path or returns an error code.
#include "common.cuh"
#include "field.cuh"
#include "ntt.cuh"
#include "merkle.cuh"
#include "poly.cuh"
#include "fri.cuh"
#include "prover.cuh"
#include "trace.cuh"
#include "verifier.cuh"

using namespace ms;

struct ms_ctx : public ms::Ctx {
    ms::ProverState* prover = nullptr;
};

#define FIELD_DISPATCH(ctx, CALL)                                            \
    ((ctx)->field == MS_FIELD_GOLDILOCKS ? CALL(ms::GL) : CALL(ms::BB))

// Host-buffer LDE (the reference-facing call: coefficients in, row-major Matrix out, src/starks.rs:87-91).
// PCIe is the bound (N*C*s in, L*C*s out), so the copies are pipelined against the kernels on the
// context's copy stream: column groups are extended while the next group uploads, and the row-major
// result leaves in row chunks through two staging buffers while the next chunk is transposed.
template <class F>
static int coset_lde_host(Ctx* c, const void* coeffs_host, uint64_t n, uint64_t cols, uint64_t blowup, uint64_t shift,
                          void* out_rm_host) {
    using T = typename F::T;
    const uint64_t L = n * blowup;
    if (n == 0 || cols == 0) return MS_OK;
    const uint64_t ngroups = (n * cols * sizeof(T) >= (64u << 20)) ? (cols < 8 ? cols : 8) : 1;
    const uint64_t nchunks = (L * cols * sizeof(T) >= (256u << 20) && L >= 1024) ? 16 : 1;
    const uint64_t chunk_rows = (L + nchunks - 1) / nchunks;
    struct Events {
        std::vector<cudaEvent_t> v;
        ~Events() { for (auto e : v) cudaEventDestroy(e); }
        cudaEvent_t make() {
            cudaEvent_t e = nullptr;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            v.push_back(e);
            return e;
        }
    } events;
    Scratch din(c), dout(c), stage0(c), stage1(c);
    MS_TRY(din.alloc(n * cols * sizeof(T)));
    MS_TRY(dout.alloc(L * cols * sizeof(T)));
    MS_TRY(stage0.alloc(chunk_rows * cols * sizeof(T)));
    MS_TRY(stage1.alloc(chunk_rows * cols * sizeof(T)));
    cudaEvent_t ready = events.make();
    MS_CUDA(c, cudaEventRecord(ready, c->stream));
    MS_CUDA(c, cudaStreamWaitEvent(c->copy_stream, ready, 0));
    const T* h_in = static_cast<const T*>(coeffs_host);
    T* h_out = static_cast<T*>(out_rm_host);
    for (uint64_t g = 0; g < ngroups; g++) {
        const uint64_t c0 = cols * g / ngroups, c1 = cols * (g + 1) / ngroups;
        if (c1 == c0) continue;
        MS_CUDA(c, cudaMemcpyAsync(din.as<T>() + c0 * n, h_in + c0 * n, (c1 - c0) * n * sizeof(T), cudaMemcpyHostToDevice,
                                   c->copy_stream));
        cudaEvent_t up = events.make();
        MS_CUDA(c, cudaEventRecord(up, c->copy_stream));
        MS_CUDA(c, cudaStreamWaitEvent(c->stream, up, 0));
        MS_TRY(lde_batch<F>(c, din.as<T>() + c0 * n, n, c1 - c0, ilog2(n), ilog2(blowup), (T)(shift % (uint64_t)F::P), false,
                            dout.as<T>() + c0 * L, L));
    }
    cudaEvent_t drained[2] = {nullptr, nullptr};
    for (uint64_t i = 0; i < nchunks; i++) {
        const uint64_t r0 = i * chunk_rows;
        if (r0 >= L) break;
        const uint64_t rows = (L - r0 < chunk_rows) ? L - r0 : chunk_rows;
        T* stage = (i & 1) ? stage1.as<T>() : stage0.as<T>();
        if (drained[i & 1]) MS_CUDA(c, cudaStreamWaitEvent(c->stream, drained[i & 1], 0));
        MS_TRY(transpose<F>(c, dout.as<T>() + r0, stage, rows, cols, false, L));
        cudaEvent_t done = events.make();
        MS_CUDA(c, cudaEventRecord(done, c->stream));
        MS_CUDA(c, cudaStreamWaitEvent(c->copy_stream, done, 0));
        MS_CUDA(c, cudaMemcpyAsync(h_out + r0 * cols, stage, rows * cols * sizeof(T), cudaMemcpyDeviceToHost, c->copy_stream));
        drained[i & 1] = events.make();
        MS_CUDA(c, cudaEventRecord(drained[i & 1], c->copy_stream));
    }
    MS_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    return MS_OK;
}

template <class F>
static int fri_query_export(Ctx* c, const void* d_prev_poly, uint64_t poly_stride, uint64_t prev_len, const void* d_prev_cw,
                            uint64_t prev_cw_stride, uint64_t prev_domain, const uint32_t* d_prev_nodes, const void* d_next_cw,
                            uint64_t next_cw_stride, const uint64_t* betas, uint64_t q, void* points_host, uint64_t* found_host,
                            void* neigh_host, uint8_t* paths_host, void* quot_host) {
    using T = typename F::T;
    using E = Ext<F>;
    if (q == 0) return MS_OK;
    QueryLookups<F> lk;
    MS_TRY(fri_query_lookups<F>(c, (const T*)d_prev_cw, prev_cw_stride, prev_domain, d_prev_nodes, (const T*)d_next_cw, next_cw_stride,
                                prev_domain / 2, betas, q, &lk));
    if (points_host) {
        E* pts = reinterpret_cast<E*>(points_host);
        for (uint64_t k = 0; k < q; k++) {
            pts[6 * k + 0] = ext_from_base<F>(lk.x1[k]); pts[6 * k + 1] = lk.ys[2 * k];
            pts[6 * k + 2] = ext_from_base<F>(lk.x2[k]); pts[6 * k + 3] = lk.ys[2 * k + 1];
            pts[6 * k + 4] = ext_from_base<F>(lk.x3[k]); pts[6 * k + 5] = lk.ys[2 * q + k];
        }
    }
    if (found_host) for (uint64_t k = 0; k < 2 * q; k++) found_host[k] = lk.found[k];
    if (neigh_host) memcpy(neigh_host, lk.neigh.data(), 4 * q * sizeof(E));
    if (paths_host)
        for (size_t i = 0; i < lk.paths.size() / 8; i++) digest_words_to_bytes(&lk.paths[8 * i], paths_host + 32 * i);
    if (quot_host && prev_len >= 3) {
        const uint64_t nq = prev_len - 2;
        Scratch dq(c);
        MS_TRY(dq.alloc(q * nq * sizeof(E)));
        MS_TRY(fri_query_quotients<F>(c, (const T*)d_prev_poly, poly_stride, prev_len, lk.s2.data(), (uint32_t)q, dq.as<T>()));
        MS_CUDA(c, cudaMemcpyAsync(quot_host, dq.p, q * nq * sizeof(E), cudaMemcpyDeviceToHost, c->stream));
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return MS_OK;
}

extern "C" {

int32_t ms_version(void) { return 1; }

int32_t ms_ctx_create(int32_t field, int32_t device, void* stream, ms_ctx** out) {
    if (!out || (field != MS_FIELD_GOLDILOCKS && field != MS_FIELD_BABYBEAR)) return MS_ERR_BAD_SHAPE;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return MS_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return MS_ERR_CUDA;
    ms_ctx* c = new ms_ctx();
    c->field = field;
    c->device = device;
    if (const char* e = getenv("MINISTARK_DL_MAX_RANKS")) c->dl_max_ranks = atoi(e);
    if (const char* e = getenv("MINISTARK_DL_SKIP_RANK0")) c->dl_skip_rank0 = atoi(e) ? 1 : 0;
    if (const char* e = getenv("MINISTARK_LDE_LINEARITY")) c->lde_linearity = atoi(e) ? 1 : 0;
    if (const char* e = getenv("MINISTARK_NTT_TABLE_MB")) c->ntt_tables_budget = (size_t)atoll(e) << 20;
    if (const char* e = getenv("MINISTARK_NTT_TILE")) {
        const int v = atoi(e);
        if (v == NTT_LOG_TILE_PREF || v == NTT_LOG_TILE_PREF - 1) c->ntt_log_tile = v;
    }
    // NULL = the legacy default stream (what torch uses unless told otherwise), so that the caller's
    // copies and this library's kernels are ordered without extra synchronisation
    c->stream = reinterpret_cast<cudaStream_t>(stream);
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->copy_event, cudaEventDisableTiming) != cudaSuccess) {
        delete c;
        return MS_ERR_CUDA;
    }
    *out = c;
    return MS_OK;
}

void ms_ctx_destroy(ms_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 2; i++)
        if (c->wtab[i]) cudaFree(c->wtab[i]);
    for (int i = 0; i < 2; i++)
        for (int a = 0; a < 16; a++)
            if (c->tw16_plain[i][a]) cudaFree(c->tw16_plain[i][a]);
    for (auto& t : c->ntt_tables) {
        if (t.ft) cudaFree(t.ft);
        if (t.tw) cudaFree(t.tw);
    }
    block_cache_release(c);
    if (c->dec4) cudaFree(c->dec4);
    if (c->hstage) cudaFreeHost(c->hstage);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->copy_event) cudaEventDestroy(c->copy_event);
    delete c->prover;
    if (c->comm) {  // not collective here: mappings are dropped without waiting for the peers (use ms_comm_destroy first)
        if (c->comm->arena) cudaFree(c->comm->arena);
        c->comm->arena = nullptr;
        delete c->comm;
        c->comm = nullptr;
    }
    if (c->owns_stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* ms_last_error(const ms_ctx* c) { return c ? c->err.c_str() : "null context"; }

int32_t ms_sync(ms_ctx* c) {
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    return MS_OK;
}
uint64_t ms_launch_count(const ms_ctx* c) { return c ? c->launches : 0; }
int32_t ms_set_profiling(ms_ctx* c, int32_t on) {
    c->profile = on != 0;
    return MS_OK;
}
int32_t ms_profile_collect(ms_ctx* c, const char** names, float* total_ms, uint32_t* counts, int32_t cap) {
    cudaStreamSynchronize(c->stream);
    int n = 0;
    for (auto& e : c->prof) {
        float t = 0;
        cudaEventElapsedTime(&t, e.a, e.b);
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
        int k = 0;
        for (; k < n; k++)
            if (names[k] == e.name) break;
        if (k == n) {
            if (n >= cap) continue;
            names[n] = e.name;
            total_ms[n] = 0;
            counts[n] = 0;
            n++;
        }
        total_ms[k] += t;
        counts[k]++;
    }
    c->prof.clear();
    return n;
}
int32_t ms_set_zero_display(ms_ctx* c, int32_t empty) {
    c->zero_display_empty = empty ? 1 : 0;
    return MS_OK;
}

int32_t ms_set_transcript_option(ms_ctx* c, int32_t option, int32_t value) {
    if (option >= MS_OPT_MASK_ABSORB && option <= MS_OPT_MASK_SQUEEZE_END) c->bridge_masks[option] = (uint8_t)value;
    else if (option == MS_OPT_LEFTOVER_AS_PUBLISHED) c->leftover_as_published = value ? 1 : 0;
    else return fail(c, MS_ERR_UNSUPPORTED, "unknown transcript option %d", option);
    return MS_OK;
}

int32_t ms_selftest_field_ops(ms_ctx* c, uint64_t n_random, uint64_t* n_bad) {
#define CALL(F) selftest_ops<F>(c, n_random, n_bad)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_dev_alloc(ms_ctx* c, size_t bytes, void** d_out) {
    MS_CUDA(c, cudaSetDevice(c->device));
    MS_CUDA(c, cudaMalloc(d_out, bytes ? bytes : 16));
    return MS_OK;
}
int32_t ms_dev_free(ms_ctx* c, void* p) {
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    MS_CUDA(c, cudaFree(p));
    return MS_OK;
}
int32_t ms_h2d(ms_ctx* c, void* d, const void* s, size_t bytes) {
    MS_CUDA(c, cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, c->stream));
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    return MS_OK;
}
int32_t ms_d2h(ms_ctx* c, void* d, const void* s, size_t bytes) {
    MS_CUDA(c, cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, c->stream));
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    return MS_OK;
}

int32_t ms_transpose_rm_to_cm(ms_ctx* c, const void* d_rm, uint64_t rows, uint64_t width, void* d_cm) {
#define CALL(F) transpose<F>(c, (const F::T*)d_rm, (F::T*)d_cm, rows, width, true)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}
int32_t ms_transpose_cm_to_rm(ms_ctx* c, const void* d_cm, uint64_t rows, uint64_t width, void* d_rm) {
#define CALL(F) transpose<F>(c, (const F::T*)d_cm, (F::T*)d_rm, rows, width, false)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_trace_synth(ms_ctx* c, uint64_t seed, uint64_t n, uint64_t w, void* d_trace_cm) {
#define CALL(F) trace_synth<F>(c, seed, n, w, (F::T*)d_trace_cm)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}
int32_t ms_trace_recurrence(ms_ctx* c, const void* matrix_host, const void* row0_host, uint64_t w, uint64_t steps, uint64_t n, uint64_t padding,
                            void* d_trace_cm) {
#define CALL(F) trace_recurrence<F>(c, (const F::T*)matrix_host, (const F::T*)row0_host, w, steps, n, (F::T)(padding % (uint64_t)F::P), (F::T*)d_trace_cm)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

uint64_t ms_merkle_node_count(uint64_t n_groups, uint64_t k) { return merkle_node_count(n_groups, k); }

int32_t ms_merkle_commit(ms_ctx* c, const void* d_data, uint64_t stride, uint64_t rows, uint64_t width, int32_t deg,
                         uint64_t lpn, uint64_t k, uint32_t* d_nodes, uint8_t* root32) {
#define CALL(F) merkle_commit<F>(c, (const F::T*)d_data, stride, rows, width, deg, lpn, k, d_nodes, root32)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_intt_columns(ms_ctx* c, const void* d_evals, uint64_t in_stride, uint64_t n, uint64_t cols, void* d_coeffs,
                        uint64_t out_stride) {
    if (!is_pow2(n)) return fail(c, MS_ERR_BAD_SHAPE, "trace length %llu is not a power of two (air.rs:23)", (unsigned long long)n);
#define CALL(F) lde_batch<F>(c, (const F::T*)d_evals, in_stride, cols, ilog2(n), 0, (F::T)1, true, (F::T*)d_coeffs, out_stride)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_coset_lde(ms_ctx* c, const void* d_coeffs, uint64_t in_stride, uint64_t n, uint64_t cols, uint64_t blowup,
                     uint64_t shift, void* d_out, uint64_t out_stride) {
    if (!is_pow2(n) || !is_pow2(blowup)) return fail(c, MS_ERR_BAD_SHAPE, "n and blowup must be powers of two");
#define CALL(F)                                                                                             \
    ((shift % (uint64_t)F::P) == 0 ? fail(c, MS_ERR_BAD_SHAPE, "coset offset is zero (starks.rs:84 unwrap)") \
                                   : lde_batch<F>(c, (const F::T*)d_coeffs, in_stride, cols, ilog2(n), ilog2(blowup), \
                                                  (F::T)(shift % (uint64_t)F::P), false, (F::T*)d_out, out_stride))
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_coset_lde_host(ms_ctx* c, const void* coeffs_host, uint64_t n, uint64_t cols, uint64_t blowup, uint64_t shift,
                          void* out_host_rowmajor) {
    if (!is_pow2(n) || !is_pow2(blowup)) return fail(c, MS_ERR_BAD_SHAPE, "n and blowup must be powers of two");
#define CALL(F) coset_lde_host<F>(c, coeffs_host, n, cols, blowup, shift, out_host_rowmajor)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_linear_constraints(ms_ctx* c, const void* d_coeffs, uint64_t stride, uint64_t n, uint64_t w, const void* matrix_host,
                              uint64_t t, void* d_out, uint64_t out_stride) {
#define CALL(F) linear_constraints<F>(c, (const F::T*)d_coeffs, stride, n, w, (const F::T*)matrix_host, t, (F::T*)d_out, out_stride)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_mix(ms_ctx* c, const void* d_coeffs, uint64_t stride, uint64_t n, uint64_t cols, uint64_t r, void* d_out) {
#define CALL(F) mix<F>(c, (const F::T*)d_coeffs, stride, n, cols, (F::T)(r % (uint64_t)F::P), (F::T*)d_out)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_deep_open(ms_ctx* c, const void* d_coeffs, uint64_t stride, uint64_t n, uint64_t cols, const void* z_host, uint64_t q,
                     void* out_host) {
#define CALL(F) deep_open<F>(c, (const F::T*)d_coeffs, stride, n, cols, (const F::T*)z_host, q, (F::T*)out_host)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_fri_commit(ms_ctx* c, const void* d_poly, uint64_t poly_stride, uint64_t domain, uint64_t blowup, void* d_codeword,
                      uint64_t cw_stride, uint32_t* d_nodes, uint8_t* root32) {
#define CALL(F) fri_commit<F>(c, (const F::T*)d_poly, poly_stride, domain, blowup, (F::T*)d_codeword, cw_stride, d_nodes, root32)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_fri_deep_coeffs(ms_ctx* c, const void* d_poly, uint64_t stride, uint64_t n_coeffs, const void* z_host, void* d_out_host) {
#define CALL(F) fri_deep_coeffs<F>(c, (const F::T*)d_poly, stride, n_coeffs, (const F::T*)z_host, (F::T*)d_out_host)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_fri_fold(ms_ctx* c, const void* d_poly, uint64_t stride, uint64_t n_coeffs, const void* z_host, const void* alpha_host,
                    const void* d_host, void* d_next, uint64_t next_stride) {
#define CALL(F) fri_fold<F>(c, (const F::T*)d_poly, stride, n_coeffs, (const F::T*)z_host, (const F::T*)alpha_host, (const F::T*)d_host, (F::T*)d_next, next_stride)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_fri_query(ms_ctx* c, const void* d_prev_poly, uint64_t poly_stride, uint64_t prev_len, const void* d_prev_cw,
                     uint64_t prev_cw_stride, uint64_t prev_domain, const uint32_t* d_prev_nodes, const void* d_next_cw,
                     uint64_t next_cw_stride, const uint64_t* betas_host, uint64_t q, void* points_host, uint64_t* found_host,
                     void* neigh_host, uint8_t* paths_host, void* quot_host) {
#define CALL(F) fri_query_export<F>(c, d_prev_poly, poly_stride, prev_len, d_prev_cw, prev_cw_stride, prev_domain, d_prev_nodes, d_next_cw, next_cw_stride, betas_host, q, points_host, found_host, neigh_host, paths_host, quot_host)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_stark_derive(int32_t field, const ms_stark_params* p, uint64_t* rounds, uint64_t* cq, uint64_t* fq) {
    ms::StarkDerived d;
    if (!p || (field != MS_FIELD_GOLDILOCKS && field != MS_FIELD_BABYBEAR)) return MS_ERR_BAD_SHAPE;
    int rc = ms::stark_derive(field, *p, &d);
    if (rc != MS_OK) return rc;
    if (rounds) *rounds = d.rounds;
    if (cq) *cq = d.constrain_queries;
    if (fq) *fq = d.fri_queries;
    return MS_OK;
}

uint64_t ms_stark_proof_bound(int32_t field, const ms_stark_params* p, uint64_t n, uint64_t cols) {
    ms::StarkDerived d;
    if (!p || (field != MS_FIELD_GOLDILOCKS && field != MS_FIELD_BABYBEAR)) return 0;
    if (ms::stark_derive(field, *p, &d) != MS_OK) return 0;
    return field == MS_FIELD_GOLDILOCKS ? ms::proof_size_bound<ms::GL>(*p, d, n, cols) : ms::proof_size_bound<ms::BB>(*p, d, n, cols);
}

int32_t ms_stark_prove(ms_ctx* c, const ms_stark_params* p, const void* trace_rm_host, uint64_t n, uint64_t w,
                       const void* cmat_host, uint64_t t, uint8_t* proof_out, uint64_t* proof_len) {
    if (!c->prover) c->prover = new ms::ProverState();
#define CALL(F) stark_prove<F>(c, c->prover, *p, trace_rm_host, nullptr, n, w, (const F::T*)cmat_host, t, proof_out, proof_len)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}
int32_t ms_stark_prove_affine(ms_ctx* c, const ms_stark_params* p, const void* trace_rm_host, uint64_t n, uint64_t w, const void* cmat_host,
                              const void* cconst_host, uint64_t t, uint8_t* proof_out, uint64_t* proof_len) {
    if (!c->prover) c->prover = new ms::ProverState();
#define CALL(F) stark_prove<F>(c, c->prover, *p, trace_rm_host, nullptr, n, w, (const F::T*)cmat_host, t, proof_out, proof_len, nullptr, 0, (const F::T*)cconst_host)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}
int32_t ms_stark_prove_device(ms_ctx* c, const ms_stark_params* p, const void* d_trace_cm, uint64_t n, uint64_t w,
                              const void* cmat_host, uint64_t t, uint8_t* proof_out, uint64_t* proof_len) {
    if (!c->prover) c->prover = new ms::ProverState();
#define CALL(F) stark_prove<F>(c, c->prover, *p, nullptr, d_trace_cm, n, w, (const F::T*)cmat_host, t, proof_out, proof_len)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}
int32_t ms_stark_prove_hooked(ms_ctx* c, const ms_stark_params* p, const void* d_trace_cm, uint64_t n, uint64_t w,
                              const void* cmat_host, uint64_t t, const ms_commit_hooks* hooks, uint8_t* proof_out,
                              uint64_t* proof_len) {
    if (!c->prover) c->prover = new ms::ProverState();
#define CALL(F) stark_prove<F>(c, c->prover, *p, nullptr, d_trace_cm, n, w, (const F::T*)cmat_host, t, proof_out, proof_len, hooks)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}
int32_t ms_stark_verify(ms_ctx* c, const ms_stark_params* p, const void* d_constrains_cm, uint64_t stride, uint64_t n, uint64_t cols,
                        const uint8_t* proof, uint64_t proof_len, int32_t strict, int32_t* accepted, int32_t* failed_check) {
#define CALL(F) stark_verify<F>(c, *p, (const F::T*)d_constrains_cm, stride, n, cols, proof, proof_len, strict, accepted, failed_check)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

/* ---- multi-GPU: communicators (csrc/comm.cuh) -------------------------------------------------- */
int32_t ms_comm_unique_id(uint8_t id128[128]) {
    NcclApi& api = nccl_api();
    if (!id128 || !api.load()) return MS_ERR_NCCL;
    ms_ncclUniqueId id;
    if (api.GetUniqueId(&id) != 0) return MS_ERR_NCCL;
    memcpy(id128, id.internal, 128);
    return MS_OK;
}
int32_t ms_comm_init_nccl(ms_ctx* c, const uint8_t id128[128], int32_t rank, int32_t world) {
    if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(c, MS_ERR_BAD_SHAPE, "ms_comm_init_nccl: rank %d of %d", rank, world);
    if (c->comm) return fail(c, MS_ERR_UNSUPPORTED, "context already belongs to a communicator");
    NcclComm* nc = new NcclComm();
    int rc = nc->init(c, id128, rank, world);
    if (rc != MS_OK) {
        delete nc;
        return rc;
    }
    c->comm = nc;
    return MS_OK;
}
int32_t ms_comm_init_local(ms_ctx* const* ctxs, int32_t world) {
    if (!ctxs || world < 1) return MS_ERR_BAD_SHAPE;
    for (int g = 0; g < world; g++)
        if (!ctxs[g] || ctxs[g]->comm) return MS_ERR_UNSUPPORTED;
    auto grp = std::make_shared<LocalGroup>();
    grp->world = world;
    grp->slot.assign(world, nullptr);
    grp->device.resize(world);
    for (int g = 0; g < world; g++) grp->device[g] = ctxs[g]->device;
    for (int g = 0; g < world; g++) {
        // kernels of rank g read the arenas of the ranks on other devices
        cudaSetDevice(ctxs[g]->device);
        for (int h = 0; h < world; h++)
            if (grp->device[h] != grp->device[g]) {
                cudaError_t e = cudaDeviceEnablePeerAccess(grp->device[h], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctxs[g], MS_ERR_CUDA, "no peer access %d -> %d: %s", grp->device[g], grp->device[h], cudaGetErrorString(e));
                cudaGetLastError();
            }
        LocalComm* lc = new LocalComm();
        lc->rank = g;
        lc->world = world;
        lc->grp = grp;
        ctxs[g]->comm = lc;
    }
    return MS_OK;
}
int32_t ms_comm_destroy(ms_ctx* c) {
    if (!c->comm) return MS_OK;
    cudaSetDevice(c->device);
    int rc = c->comm->release(c);
    delete c->comm;
    c->comm = nullptr;
    return rc;
}
int32_t ms_comm_info(const ms_ctx* c, int32_t* rank, int32_t* world, const char** backend) {
    if (rank) *rank = c->comm ? c->comm->rank : 0;
    if (world) *world = c->comm ? c->comm->world : 1;
    if (backend) *backend = c->comm ? c->comm->backend() : "none";
    return MS_OK;
}
int32_t ms_set_shard_mask(ms_ctx* c, int32_t mask) {
    c->shard_mask = mask & MS_SHARD_ALL;
    return MS_OK;
}
int32_t ms_shard_plan(uint64_t cols, uint64_t groups, uint64_t k, int32_t world, int32_t rank, uint64_t* a, uint64_t* b, uint64_t* per_rank,
                      uint64_t* left) {
    if (world < 1 || rank < 0 || rank >= world) return MS_ERR_BAD_SHAPE;
    uint64_t aa, bb, pr = 0, lf = 0;
    shard_range(cols, world, rank, &aa, &bb);
    if (a) *a = aa;
    if (b) *b = bb;
    const bool ok = subtree_plan(groups, k, world, &pr, &lf);
    if (per_rank) *per_rank = pr;
    if (left) *left = lf;
    return ok ? MS_OK : MS_ERR_BAD_SHAPE;
}
int32_t ms_stark_prove_multi(ms_ctx* c, const ms_stark_params* p, const void* d_trace_cm, uint64_t n, uint64_t w, const void* cmat_host,
                             uint64_t t, uint8_t* proof_out, uint64_t* proof_len, int32_t flags) {
    if (!c->prover) c->prover = new ms::ProverState();
    MS_CUDA(c, cudaSetDevice(c->device));
#define CALL(F) stark_prove<F>(c, c->prover, *p, nullptr, d_trace_cm, n, w, (const F::T*)cmat_host, t, proof_out, proof_len, nullptr, flags)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}

int32_t ms_merkle_subtree(ms_ctx* c, const void* d_data, uint64_t stride, uint64_t rows, uint64_t width, int32_t deg,
                          uint64_t lpn, uint64_t k, uint32_t* d_out, uint64_t* n_out) {
#define CALL(F) merkle_subtree<F>(c, (const F::T*)d_data, stride, rows, width, deg, lpn, k, d_out, n_out)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}
int32_t ms_merkle_subtree_gather(ms_ctx* c, const void* const* plane_ptrs_host, uint64_t rows, uint64_t width, int32_t deg,
                                 uint64_t lpn, uint64_t k, uint32_t* d_out, uint64_t* n_out) {
    if (!plane_ptrs_host || deg < 1 || width == 0) return fail(c, MS_ERR_BAD_SHAPE, "ms_merkle_subtree_gather: null plane table");
    Scratch tab(c);
    const size_t bytes = (size_t)width * (size_t)deg * sizeof(void*);
    MS_TRY(tab.alloc(bytes));
    MS_CUDA(c, cudaMemcpyAsync(tab.p, plane_ptrs_host, bytes, cudaMemcpyHostToDevice, c->stream));
#define CALL(F) merkle_subtree<F>(c, (const F::T*)tab.p, 0, rows, width, deg, lpn, k, d_out, n_out, true)
    return FIELD_DISPATCH(c, CALL);
#undef CALL
}
/* ---- peer memory (CUDA IPC): buffers other ranks' kernels read over NVLink ------------------- */
int32_t ms_peer_alloc(ms_ctx* c, uint64_t bytes, void** d_out) {
    if (!d_out) return fail(c, MS_ERR_BAD_SHAPE, "ms_peer_alloc: null out");
    MS_CUDA(c, cudaSetDevice(c->device));
    MS_CUDA(c, cudaMalloc(d_out, bytes ? bytes : 16));
    return MS_OK;
}
int32_t ms_peer_free(ms_ctx* c, void* d_ptr) {
    if (d_ptr) MS_CUDA(c, cudaFree(d_ptr));
    return MS_OK;
}
int32_t ms_peer_export(ms_ctx* c, void* d_ptr, uint8_t* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    MS_CUDA(c, cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64, &h, 64);
    return MS_OK;
}
int32_t ms_peer_open(ms_ctx* c, const uint8_t* handle64, void** d_out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    MS_CUDA(c, cudaSetDevice(c->device));
    MS_CUDA(c, cudaIpcOpenMemHandle(d_out, h, cudaIpcMemLazyEnablePeerAccess));
    return MS_OK;
}
int32_t ms_peer_close(ms_ctx* c, void* d_ptr) {
    if (d_ptr) MS_CUDA(c, cudaIpcCloseMemHandle(d_ptr));
    return MS_OK;
}
int32_t ms_host_register(ms_ctx* c, void* host_ptr, uint64_t bytes) {
    MS_CUDA(c, cudaSetDevice(c->device));
    MS_CUDA(c, cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
    return MS_OK;
}
int32_t ms_host_unregister(ms_ctx* c, void* host_ptr) {
    MS_CUDA(c, cudaHostUnregister(host_ptr));
    return MS_OK;
}
int32_t ms_merkle_reduce(ms_ctx* c, const uint32_t* d_digests, uint64_t n, uint64_t k, uint8_t* root32) {
    return merkle_reduce(c, d_digests, n, k, root32);
}
int32_t ms_stark_last_timings(ms_ctx* c, const char** names, float* msv, int32_t cap) {
    if (!c->prover) return 0;
    int n = 0;
    for (auto& e : c->prover->timings) {
        if (n >= cap) break;
        names[n] = e.first;
        msv[n] = e.second;
        n++;
    }
    return n;
}

}  // extern "C"
