/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the data-parallel stages of mini-stark's prover, the path
 * BASELINE.json:north_star names.  It is the checker for the CUDA path and the CPU baseline of
 * bench.py; nothing in ministark_b200/ links, imports or executes it.
 *
 * PARITY STATUS: the reference is Rust and cannot be built in this image (no cargo/rustc), and
 * its arithmetic lives in crates that are not vendored (ark-ff/ark-poly/ark-std 0.5.0,
 * sha2 0.10.8, nimue@0e584985 -- Cargo.lock).  The stages restated HERE are pinned by mathematics
 * and standards: exact field arithmetic (unique results), FIPS 180-4 SHA-256 (checked against
 * hashlib in tests/), decimal Display of non-zero values, and the tree shape of src/merkle.rs
 * (in the repo).  The items that are "parity unpinned" (Display of zero / QuadExtField, nimue's
 * DigestBridge, test_rng) are isolated behind switches: or_set_zero_display() here, the rest in
 * oracle/pyref.py.  See DESIGN.md section "Oracle".
 *
 * All elements are canonical uint64_t (see fields.h).  field: 0 = Goldilocks, 1 = BabyBear.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <pthread.h>
#include "fields.h"

#define PFX(x) gl_##x
#define F_P GL_P
#define F_ROOT GL_ROOT
#define F_TWO_ADICITY GL_TWO_ADICITY
#define EXT_D 2
#include "oracle_impl.inc"
#undef PFX
#undef F_P
#undef F_ROOT
#undef F_TWO_ADICITY
#undef EXT_D

#define PFX(x) bb_##x
#define F_P BB_P
#define F_ROOT BB_ROOT
#define F_TWO_ADICITY BB_TWO_ADICITY
#define EXT_D 4
#include "oracle_impl.inc"
#undef PFX
#undef F_P
#undef F_ROOT
#undef F_TWO_ADICITY
#undef EXT_D

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ SHA-256 (FIPS 180-4) */
static const u32 K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

typedef struct { u32 h[8]; u64 len; unsigned char buf[64]; unsigned fill; } sha256_ctx;

static inline u32 rotr(u32 x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha256_block(u32 *h, const unsigned char *p) {
    u32 w[64];
    for (int i = 0; i < 16; i++) w[i] = (u32)p[4 * i] << 24 | (u32)p[4 * i + 1] << 16 | (u32)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        u32 s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        u32 s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    u32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        u32 t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
        u32 t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha256_init(sha256_ctx *c) {
    static const u32 iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(c->h, iv, sizeof iv); c->len = 0; c->fill = 0;
}
static void sha256_update(sha256_ctx *c, const void *data, size_t n) {
    const unsigned char *p = (const unsigned char *)data;
    c->len += n;
    while (n) {
        size_t take = 64 - c->fill; if (take > n) take = n;
        memcpy(c->buf + c->fill, p, take); c->fill += take; p += take; n -= take;
        if (c->fill == 64) { sha256_block(c->h, c->buf); c->fill = 0; }
    }
}
static void sha256_final(sha256_ctx *c, unsigned char *out) {
    u64 bits = c->len * 8;
    unsigned char pad = 0x80; sha256_update(c, &pad, 1);
    unsigned char z = 0; while (c->fill != 56) sha256_update(c, &z, 1);
    unsigned char lb[8]; for (int i = 0; i < 8; i++) lb[i] = (unsigned char)(bits >> (56 - 8 * i));
    sha256_update(c, lb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = c->h[i] >> 24; out[4 * i + 1] = c->h[i] >> 16; out[4 * i + 2] = c->h[i] >> 8; out[4 * i + 3] = c->h[i]; }
}
API void or_sha256(const void *data, u64 n, unsigned char *out32) {
    sha256_ctx c; sha256_init(&c); sha256_update(&c, data, n); sha256_final(&c, out32);
}

/* ------------------------------------------------------------------ Display (to_string) */
/* ark-ff 0.5.0 Fp Display prints the canonical integer in decimal; zero prints "0" (0.4.x printed
 * the empty string) -- UNPINNED, switchable.  QuadExtField prints "QuadExtField(c0 + c1 * u)",
 * nested for the BabyBear quartic tower.  Used by src/merkle.rs:165. */
static int g_zero_display_empty = 0;
API void or_set_zero_display(int empty) { g_zero_display_empty = empty; }

static size_t fmt_base(u64 v, char *o) {
    if (v == 0) { if (g_zero_display_empty) return 0; o[0] = '0'; return 1; }
    char t[24]; int n = 0;
    while (v) { t[n++] = '0' + (char)(v % 10); v /= 10; }
    for (int i = 0; i < n; i++) o[i] = t[n - 1 - i];
    return n;
}
static size_t fmt_quad(const u64 *c, int deg, char *o) {
    if (deg == 1) return fmt_base(c[0], o);
    size_t n = 0;
    memcpy(o + n, "QuadExtField(", 13); n += 13;
    n += fmt_quad(c, deg / 2, o + n);
    memcpy(o + n, " + ", 3); n += 3;
    n += fmt_quad(c + deg / 2, deg / 2, o + n);
    memcpy(o + n, " * u)", 5); n += 5;
    return n;
}
/* Display of one element with `deg` prime-field coordinates (1 = base field). */
API u64 or_leaf_string(const u64 *elem, int deg, char *out) { return fmt_quad(elem, deg, out); }

/* ------------------------------------------------------------------ Merkle tree */
/* MerkleTree::new (src/merkle.rs:81-148): leaf groups of `lpn` consecutive elements hashed as
 * SHA256(concat(to_string(e))) (calculate_from_leafs, :162-168), inner nodes SHA256(concat of
 * `k` child digests) (:171-177), all nodes in level order in one vector (:119-140).
 * data: n_elems elements of `deg` coordinates each.  nodes_out may be NULL (root only).
 * Returns the node count, or -1 for the reference's panics (:93-104). */
/* ---- tiny pthread fan-out used by the multi-threaded legs (the reference itself is single-threaded) */
static double or_now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
typedef void (*or_item_fn)(void *ctx, u64 item, int tid);
typedef struct { or_item_fn fn; void *ctx; u64 n; u64 *next; pthread_mutex_t *mu; int tid; } or_pf_job;
static void *or_pf_worker(void *p) {
    or_pf_job *j = (or_pf_job *)p;
    for (;;) {
        pthread_mutex_lock(j->mu);
        u64 i = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (i >= j->n) break;
        j->fn(j->ctx, i, j->tid);
    }
    return 0;
}
/* run fn(ctx, i) for i in [0, n) on up to `threads` pthreads (dynamic, one item at a time: items are coarse) */
static void or_parallel_items(u64 n, int threads, or_item_fn fn, void *ctx) {
    if (threads < 1) threads = 1;
    if ((u64)threads > n) threads = (int)n;
    if (threads <= 1) { for (u64 i = 0; i < n; i++) fn(ctx, i, 0); return; }
    pthread_t *th = (pthread_t *)malloc(threads * sizeof(pthread_t));
    or_pf_job *jobs = (or_pf_job *)malloc(threads * sizeof(or_pf_job));
    pthread_mutex_t mu; pthread_mutex_init(&mu, 0);
    u64 next = 0;
    for (int t = 0; t < threads; t++) { jobs[t] = (or_pf_job){fn, ctx, n, &next, &mu, t}; pthread_create(&th[t], 0, or_pf_worker, &jobs[t]); }
    for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
    pthread_mutex_destroy(&mu);
    free(th); free(jobs);
}

typedef struct { const u64 *data; int deg; u64 lpn, n1, chunk; unsigned char *nodes; } mk_leaf_job;
static void mk_leaf_chunk(void *ctx, u64 ci, int tid) {
    (void)tid;
    mk_leaf_job *j = (mk_leaf_job *)ctx;
    char s[512];
    u64 g0 = ci * j->chunk, g1 = g0 + j->chunk; if (g1 > j->n1) g1 = j->n1;
    for (u64 g = g0; g < g1; g++) {
        sha256_ctx c; sha256_init(&c);
        for (u64 e = 0; e < j->lpn; e++) {
            size_t n = fmt_quad(j->data + (g * j->lpn + e) * j->deg, j->deg, s);
            sha256_update(&c, s, n);
        }
        sha256_final(&c, j->nodes + 32 * g);
    }
}
typedef struct { unsigned char *nodes; u64 k, src0, dst0, count, chunk; } mk_node_job;
static void mk_node_chunk(void *ctx, u64 ci, int tid) {
    (void)tid;
    mk_node_job *j = (mk_node_job *)ctx;
    u64 i0 = ci * j->chunk, i1 = i0 + j->chunk; if (i1 > j->count) i1 = j->count;
    for (u64 i = i0; i < i1; i++) or_sha256(j->nodes + 32 * (j->src0 + i * j->k), 32 * j->k, j->nodes + 32 * (j->dst0 + i));
}
static int64_t or_merkle_mt(const u64 *data, int deg, u64 n_elems, u64 lpn, u64 k, unsigned char *nodes_out, unsigned char *root_out, int threads) {
    if (lpn == 0 || k < 2 || (k & (k - 1)) || n_elems % lpn) return -1;
    u64 n1 = n_elems / lpn;
    if (n1 == 0 || (n1 & (n1 - 1))) return -1;
    unsigned lgk = __builtin_ctzll(k), lg1 = __builtin_ctzll(n1);
    if (lg1 % lgk) return -1;
    unsigned levels = lg1 / lgk + 1;
    u64 total = 0, lv = n1;
    for (unsigned l = 0; l < levels; l++) { total += lv; lv /= k; }
    unsigned char *nodes = nodes_out ? nodes_out : (unsigned char *)malloc(total * 32);
    {
        mk_leaf_job j = {data, deg, lpn, n1, 1024, nodes};
        or_parallel_items((n1 + j.chunk - 1) / j.chunk, threads, mk_leaf_chunk, &j);
    }
    /* inner levels in level order (src/merkle.rs:131-140): level l+1 hashes groups of k level-l digests */
    u64 src = 0, dst = n1, cnt = n1 / k;
    while (dst < total) {
        mk_node_job j = {nodes, k, src, dst, cnt, 4096};
        or_parallel_items((cnt + j.chunk - 1) / j.chunk, threads, mk_node_chunk, &j);
        src += cnt * k; dst += cnt; cnt /= k;
    }
    if (root_out) memcpy(root_out, nodes + 32 * (total - 1), 32);
    if (!nodes_out) free(nodes);
    return (int64_t)total;
}
API int64_t or_merkle(const u64 *data, int deg, u64 n_elems, u64 lpn, u64 k, unsigned char *nodes_out, unsigned char *root_out) {
    return or_merkle_mt(data, deg, n_elems, lpn, k, nodes_out, root_out, 1);
}
API int64_t or_merkle_threads(const u64 *data, int deg, u64 n_elems, u64 lpn, u64 k, unsigned char *nodes_out, unsigned char *root_out, int threads) {
    return or_merkle_mt(data, deg, n_elems, lpn, k, nodes_out, root_out, threads);
}

/* ------------------------------------------------------------------ Keccak-f[1600] (for the nimue tag) */
static const u64 KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
API void or_keccak_f1600(u64 *st) {
    static const int rotc[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int piln[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int r = 0; r < 24; r++) {
        u64 bc[5], t;
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            t = bc[(i + 4) % 5] ^ ((bc[(i + 1) % 5] << 1) | (bc[(i + 1) % 5] >> 63));
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = piln[i]; u64 b = st[j];
            st[j] = (t << rotc[i]) | (t >> (64 - rotc[i])); t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= KRC[r];
    }
}

/* ------------------------------------------------------------------ field-dispatched stage API */
#define DISPATCH(field, call_gl, call_bb) do { if ((field) == 0) { call_gl; } else { call_bb; } } while (0)

API u64 or_modulus(int field) { return field == 0 ? GL_P : BB_P; }
API int or_ext_degree(int field) { return field == 0 ? 2 : 4; }
API u64 or_root_of_unity(int field, unsigned log_n) { return field == 0 ? gl_root_of_unity(log_n) : bb_root_of_unity(log_n); }
API u64 or_fmul(int field, u64 a, u64 b) { return field == 0 ? gl_mul(a, b) : bb_mul(a, b); }
API u64 or_fpow(int field, u64 a, u64 e) { return field == 0 ? gl_pow(a, e) : bb_pow(a, e); }
API u64 or_finv(int field, u64 a) { return field == 0 ? gl_inv(a) : bb_inv(a); }
API void or_ext_mul(int field, const u64 *a, const u64 *b, u64 *r) {
    if (field == 0) { gl_ext x, y; memcpy(&x, a, sizeof x); memcpy(&y, b, sizeof y); x = gl_ext_mul(x, y); memcpy(r, &x, sizeof x); }
    else { bb_ext x, y; memcpy(&x, a, sizeof x); memcpy(&y, b, sizeof y); x = bb_ext_mul(x, y); memcpy(r, &x, sizeof x); }
}

/* natural-order (i)NTT on the size-n subgroup; inverse includes 1/n (domain.fft / domain.ifft) */
API void or_ntt(int field, u64 *a, u64 n, int inverse) {
    unsigned lg = 0; while ((1ULL << lg) < n) lg++;
    if (field == 0) {
        u64 w = gl_root_of_unity(lg);
        if (inverse) w = gl_inv(w);
        gl_ntt_inplace(a, n, w);
        if (inverse) { u64 ni = gl_inv(n % GL_P); for (u64 i = 0; i < n; i++) a[i] = gl_mul(a[i], ni); }
    } else {
        u64 w = bb_root_of_unity(lg);
        if (inverse) w = bb_inv(w);
        bb_ntt_inplace(a, n, w);
        if (inverse) { u64 ni = bb_inv(n % BB_P); for (u64 i = 0; i < n; i++) a[i] = bb_mul(a[i], ni); }
    }
}
API void or_eval_domain_naive(int field, const u64 *coef, u64 n_coeffs, u64 n, u64 offset, u64 *out) {
    DISPATCH(field, gl_eval_domain_naive(coef, n_coeffs, n, offset, out), bb_eval_domain_naive(coef, n_coeffs, n, offset, out));
}
API void or_trace_polys(int field, const u64 *trace_rm, u64 N, u64 W, u64 *out_cm) {
    DISPATCH(field, gl_trace_polys(trace_rm, N, W, out_cm), bb_trace_polys(trace_rm, N, W, out_cm));
}

/* LDE loop of src/starks.rs:87-91 over C polynomials (poly-major coefficients, stride N) into the
 * row-major L x C Matrix of src/air.rs:15-59.  `threads` > 1 splits the independent columns over
 * pthreads (the reference is single-threaded; used only by bench.py's reference arm). */
typedef struct { int field; const u64 *polys; u64 N, C, L, shift; u64 *out; u64 c0, c1; } lde_job;
static void *lde_worker(void *p) {
    lde_job *j = (lde_job *)p;
    u64 *scratch = (u64 *)malloc(j->N * sizeof(u64));
    for (u64 c = j->c0; c < j->c1; c++) {
        if (j->field == 0) gl_lde_column(j->polys + c * j->N, j->N, j->L, j->shift, j->out, j->C, c, scratch);
        else bb_lde_column(j->polys + c * j->N, j->N, j->L, j->shift, j->out, j->C, c, scratch);
    }
    free(scratch);
    return 0;
}
API void or_coset_lde(int field, const u64 *polys_cm, u64 N, u64 C, u64 L, u64 shift, u64 *out_rm, int threads) {
    if (threads < 1) threads = 1;
    if ((u64)threads > C) threads = (int)C;
    pthread_t *th = (pthread_t *)malloc(threads * sizeof(pthread_t));
    lde_job *jobs = (lde_job *)malloc(threads * sizeof(lde_job));
    for (int t = 0; t < threads; t++) {
        jobs[t] = (lde_job){field, polys_cm, N, C, L, shift, out_rm, C * t / threads, C * (t + 1) / threads};
        if (threads == 1) lde_worker(&jobs[t]); else pthread_create(&th[t], 0, lde_worker, &jobs[t]);
    }
    if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
    free(th); free(jobs);
}
API void or_mix(int field, const u64 *polys_cm, u64 N, u64 C, u64 r, u64 *out) {
    DISPATCH(field, gl_mix(polys_cm, N, C, r, out), bb_mix(polys_cm, N, C, r, out));
}
API void or_eval_base_at_ext(int field, const u64 *coef, u64 n, const u64 *z, u64 *out) {
    if (field == 0) { gl_ext zz; memcpy(&zz, z, sizeof zz); zz = gl_eval_base_at_ext(coef, n, zz); memcpy(out, &zz, sizeof zz); }
    else { bb_ext zz; memcpy(&zz, z, sizeof zz); zz = bb_eval_base_at_ext(coef, n, zz); memcpy(out, &zz, sizeof zz); }
}
API void or_eval_ext_at_ext(int field, const u64 *coef, u64 n, const u64 *z, u64 *out) {
    if (field == 0) { gl_ext zz; memcpy(&zz, z, sizeof zz); zz = gl_eval_ext_at_ext((const gl_ext *)coef, n, 1, zz); memcpy(out, &zz, sizeof zz); }
    else { bb_ext zz; memcpy(&zz, z, sizeof zz); zz = bb_eval_ext_at_ext((const bb_ext *)coef, n, 1, zz); memcpy(out, &zz, sizeof zz); }
}
API void or_fri_codeword(int field, const u64 *poly, u64 n_coeffs, u64 n, u64 *out) {
    DISPATCH(field, gl_fri_codeword((const gl_ext *)poly, n_coeffs, n, (gl_ext *)out), bb_fri_codeword((const bb_ext *)poly, n_coeffs, n, (bb_ext *)out));
}
API u64 or_fri_fold(int field, const u64 *poly, u64 n, const u64 *z, const u64 *alpha, u64 *d_out, u64 *next) {
    if (field == 0) { gl_ext zz, aa; memcpy(&zz, z, sizeof zz); memcpy(&aa, alpha, sizeof aa); return gl_fri_fold((const gl_ext *)poly, n, zz, aa, (gl_ext *)d_out, (gl_ext *)next); }
    bb_ext zz, aa; memcpy(&zz, z, sizeof zz); memcpy(&aa, alpha, sizeof aa);
    return bb_fri_fold((const bb_ext *)poly, n, zz, aa, (bb_ext *)d_out, (bb_ext *)next);
}
API u64 or_fri_query_quotient(int field, const u64 *poly, u64 n, u64 x1, u64 x2, const u64 *y1, const u64 *y2, u64 *q) {
    if (field == 0) { gl_ext a, b; memcpy(&a, y1, sizeof a); memcpy(&b, y2, sizeof b); return gl_fri_query_quotient((const gl_ext *)poly, n, x1, x2, a, b, (gl_ext *)q); }
    bb_ext a, b; memcpy(&a, y1, sizeof a); memcpy(&b, y2, sizeof b);
    return bb_fri_query_quotient((const bb_ext *)poly, n, x1, x2, a, b, (bb_ext *)q);
}

/* ------------------------------------------------------------------ transcript + whole prover / verifier */
#include "transcript.inc"

#define PFX(x) gl_##x
#define F_P GL_P
#define F_ROOT GL_ROOT
#define F_TWO_ADICITY GL_TWO_ADICITY
#define EXT_D 2
#define F_BITS 64
#define F_SER 8
#define F_ID 0
#include "prover.inc"
#undef PFX
#undef F_P
#undef F_ROOT
#undef F_TWO_ADICITY
#undef EXT_D
#undef F_BITS
#undef F_SER
#undef F_ID

#define PFX(x) bb_##x
#define F_P BB_P
#define F_ROOT BB_ROOT
#define F_TWO_ADICITY BB_TWO_ADICITY
#define EXT_D 4
#define F_BITS 31
#define F_SER 4
#define F_ID 1
#include "prover.inc"
#undef PFX
#undef F_P
#undef F_ROOT
#undef F_TWO_ADICITY
#undef EXT_D
#undef F_BITS
#undef F_SER
#undef F_ID

/* Stark::prove (src/starks.rs:59-169).  trace_rm: row-major N x W canonical elements (uint64 for both fields),
 * cmat: T x W.  stage_ms (optional): 8 doubles, see OR_T_*.  Returns the proof length or a negative error. */
API int64_t or_stark_prove(int field, const or_stark_params *p, const u64 *trace_rm, u64 N, u64 W, const u64 *cmat, const u64 *cconst,
                           u64 T, unsigned char *out, u64 cap, int threads, double *stage_ms) {
    return field == 0 ? gl_stark_prove(p, trace_rm, N, W, cmat, cconst, T, out, cap, threads, stage_ms)
                      : bb_stark_prove(p, trace_rm, N, W, cmat, cconst, T, out, cap, threads, stage_ms);
}
/* upper bound of the proof dump for an n x cols problem (same formula as ms_stark_proof_bound) */
API u64 or_stark_proof_bound(int field, const or_stark_params *p, u64 n, u64 cols) {
    u64 R, Q, QF;
    if (or_stark_derive(field, p, &R, &Q, &QF)) return 0;
    const u64 E = field == 0 ? 16 : 16;
    u64 sz = 8 + 8 + 8 + 64 * R + 64 + 16 + Q * cols * E + 8 + Q * E + 8;
    u64 npad = n, domain = n * p->blowup_factor;
    for (u64 i = 0; i + 1 < R; i++) {
        u64 path_len = 0; while ((2ULL << path_len) < domain) path_len++;
        if (domain < 2) path_len = 0;
        u64 per_q = 6 * E + 2 * (8 + 2 * E + 8 + path_len * (8 + 64)) + 8 + npad * E;
        sz += 8 + QF * per_q;
        npad = npad > 1 ? npad / 2 : 1;
        domain /= 2;
    }
    return sz;
}
/* Stark::verify (src/starks.rs:171-235): coeffs = the Constrains polynomials, poly-major [C][N]. */
API int or_stark_verify(int field, const or_stark_params *p, const u64 *coeffs, u64 N, u64 C, const unsigned char *proof, u64 len,
                        int strict, int *why) {
    return field == 0 ? gl_stark_verify(p, coeffs, N, C, proof, len, strict, why) : bb_stark_verify(p, coeffs, N, C, proof, len, strict, why);
}
/* the Constrains object of TraceTable::derive_constrains (src/air.rs:127-144) for a linear AIR: [W + T][N] */
API void or_derive_constrains(int field, const u64 *trace_rm, u64 N, u64 W, const u64 *cmat, const u64 *cconst, u64 T, u64 *out_cm, int threads) {
    if (field == 0) {
        gl_intt_job a = {trace_rm, N, W, out_cm}; or_parallel_items(W, threads, gl_intt_item, &a);
        gl_cons_job b = {cconst, N, W, T, cmat, out_cm}; or_parallel_items(T, threads, gl_cons_item, &b);
    } else {
        bb_intt_job a = {trace_rm, N, W, out_cm}; or_parallel_items(W, threads, bb_intt_item, &a);
        bb_cons_job b = {cconst, N, W, T, cmat, out_cm}; or_parallel_items(T, threads, bb_cons_item, &b);
    }
}
