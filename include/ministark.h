/*
 * ministark.h -- C ABI of libministark.so, the B200 (sm_100a) implementation of mini-stark's
 * data-parallel prover core.
 *
 * The reference (alv-around/mini-stark, pure Rust) has no FFI: its boundary for this path is the
 * crate's public API (`StarkConfig::new`, `Stark::prove`, src/starks.rs:59-63,268-273).  Each entry
 * point below replaces one call site *behind* `Stark::prove`; the `Replaces:` line cites it
 * (paths relative to the reference root).  INTEGRATION.md shows the Rust `extern "C"` block and
 * the shim a maintainer would add.
 *
 * Conventions
 *  - every function returns an int32 status (MS_OK = 0); nothing unwinds across the boundary.
 *    Shape violations the reference turns into panics (src/merkle.rs:95,99-104, src/air.rs:23,
 *    src/starks.rs:119) come back as MS_ERR_BAD_SHAPE / MS_ERR_QUOTIENT_NONZERO.
 *  - field elements are canonical integers: uint64_t for Goldilocks, uint32_t for BabyBear.
 *    Extension elements are D consecutive base elements in ark's tower order (D = 2 / 4).
 *  - `d_*` arguments are DEVICE pointers (cudaMalloc'ed by the caller, e.g. a torch tensor's
 *    data_ptr, or ms_dev_alloc); all other pointers are HOST memory owned by the caller.
 *  - device matrices are column-major ("poly-major"): column c starts at base + c*stride
 *    elements and holds `rows` contiguous elements.  Extension vectors are D coordinate planes.
 *  - digests on the device are 8 uint32_t SHA-256 state words; on the host they are the usual 32
 *    big-endian bytes.
 *  - a context is bound to one device and one stream and is not thread-safe; calls are
 *    asynchronous on that stream unless they return host data.
 */
#ifndef MINISTARK_H
#define MINISTARK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    MS_OK = 0,
    MS_ERR_BAD_SHAPE = 1,        /* reference panics: non-full tree, non power-of-two length ... */
    MS_ERR_CUDA = 2,
    MS_ERR_NCCL = 3,
    MS_ERR_QUOTIENT_NONZERO = 4, /* src/starks.rs:119 assert_eq!(rest, zero) */
    MS_ERR_TRANSCRIPT = 5,       /* nimue IOPatternError / ProofError (src/error.rs:5-8) */
    MS_ERR_UNSUPPORTED = 6,
    MS_ERR_LEAF_NOT_FOUND = 7,   /* src/error.rs:13-16 MerkleProofError::LeafNotFound */
    MS_ERR_BUFFER_TOO_SMALL = 8
};

enum { MS_FIELD_GOLDILOCKS = 0, MS_FIELD_BABYBEAR = 1 };

typedef struct ms_ctx ms_ctx;

/* ---- context --------------------------------------------------------------------------- */
int32_t ms_version(void);
/* stream: the cudaStream_t to run on (e.g. torch's current stream); NULL = the legacy default stream. */
int32_t ms_ctx_create(int32_t field, int32_t device, void* stream, ms_ctx** out);
void ms_ctx_destroy(ms_ctx* ctx);
const char* ms_last_error(const ms_ctx* ctx);
int32_t ms_sync(ms_ctx* ctx);
/* kernels launched by this library on this context since creation (bench.py gpu_launches) */
uint64_t ms_launch_count(const ms_ctx* ctx);
/* Per-kernel device timing for roofline reports: when on, the LDE / Merkle kernels are bracketed by
 * CUDA events on the context's stream; ms_profile_collect synchronises, sums the elapsed time and the
 * launch count per kernel name (static strings) since the last collect, and returns the entry count. */
int32_t ms_set_profiling(ms_ctx* ctx, int32_t on);
int32_t ms_profile_collect(ms_ctx* ctx, const char** names, float* total_ms, uint32_t* counts, int32_t cap);
/* Display of the zero element: 0 -> "0" (ark-ff 0.5.0, default), 1 -> "" (ark-ff 0.4.x) */
int32_t ms_set_zero_display(ms_ctx* ctx, int32_t empty);
/* Switches of the host transcript's nimue restatement (third-party crate, not in the reference tree; csrc/transcript.hpp):
 * MS_OPT_MASK_ABSORB / _SQUEEZE / _SQUEEZE_END = first byte of DigestBridge's three domain-separation blocks (0, 1, 2);
 * MS_OPT_LEFTOVER_AS_PUBLISHED = 1 (default): squeeze consumes digest bytes left over from the previous squeeze call
 * without writing them to the output, as nimue@0e584985's leftovers branch does; 0: the intended behaviour.
 * Replaces: nothing in the reference tree -- nimue call sites src/starks.rs:81,108,125, src/fri.rs:89,96,122. */
enum { MS_OPT_MASK_ABSORB = 0, MS_OPT_MASK_SQUEEZE = 1, MS_OPT_MASK_SQUEEZE_END = 2, MS_OPT_LEFTOVER_AS_PUBLISHED = 3 };
int32_t ms_set_transcript_option(ms_ctx* ctx, int32_t option, int32_t value);

/* Device self-test of the lazy / Montgomery butterfly arithmetic (csrc/ntt.cuh Fast<F>) against exact host
 * arithmetic on edge values and n_random random operand pairs; *n_bad = mismatches (0 expected). */
int32_t ms_selftest_field_ops(ms_ctx* ctx, uint64_t n_random, uint64_t* n_bad);

int32_t ms_dev_alloc(ms_ctx* ctx, size_t bytes, void** d_out);
int32_t ms_dev_free(ms_ctx* ctx, void* d_ptr);
int32_t ms_h2d(ms_ctx* ctx, void* d_dst, const void* src, size_t bytes);
int32_t ms_d2h(ms_ctx* ctx, void* dst, const void* d_src, size_t bytes);

/* ---- a1: trace layout ------------------------------------------------------------------- */
/* Row-major N x W trace (the reference's Matrix, src/air.rs:15-59) -> column-major W x N.
 * Replaces: the stride-W gather of src/air.rs:151-153. */
int32_t ms_transpose_rm_to_cm(ms_ctx* ctx, const void* d_rowmajor, uint64_t rows, uint64_t width, void* d_colmajor);
/* Column-major -> row-major (for callers that want the reference's Matrix layout back). */
int32_t ms_transpose_cm_to_rm(ms_ctx* ctx, const void* d_colmajor, uint64_t rows, uint64_t width, void* d_rowmajor);

/* Trace generation on the device (column-major W x N, stride n), so that synthetic / recurrent AIRs never upload a trace:
 * ms_trace_synth: element (row, col) = splitmix64(seed ^ (row*W + col)) mod p, the benchmark trace of SURVEY.md 8d
 * (ministark_b200/synth.py synth_trace).  ms_trace_recurrence: rows [0, steps) follow row_{i+1} = M row_i from row0
 * (M: host, W x W row-major canonical scalars, W <= 16), rows [steps, n) hold `padding` in every cell.
 * Replaces: TraceTable::new + add_row as driven by a recurrent `Provable::trace`, src/air.rs:73-112
 * (e.g. tests/e2e_goldilocks.rs:22-41 with M = [[0,1,0],[0,0,1],[0,1,1]], row0 = (1, b, 1 + b)). */
int32_t ms_trace_synth(ms_ctx* ctx, uint64_t seed, uint64_t n, uint64_t w, void* d_trace_colmajor);
int32_t ms_trace_recurrence(ms_ctx* ctx, const void* matrix_host, const void* row0_host, uint64_t w, uint64_t steps, uint64_t n,
                            uint64_t padding, void* d_trace_colmajor);

/* ---- a1 / a5 / a8: Merkle commitment ------------------------------------------------------ */
/* MerkleTree::new over the ROW-MAJOR flattening of a column-major device matrix of `rows` x `width`
 * elements of `deg` coordinates each (deg = 1 base field, deg = D extension planes): leaf group g
 * hashes SHA256(concat(to_string(e))) of flat elements [g*lpn, (g+1)*lpn), inner nodes hash `k`
 * child digests, nodes in level order.  d_nodes (optional) receives all (k^levels-1)/(k-1)
 * digests (8 words each); root32 (optional, host) the root bytes.  Limits of the leaf kernel's
 * 32-bit counters: leafs_per_node <= 2^27, width < 2^32 (MS_ERR_UNSUPPORTED beyond).
 * Replaces: MerkleTree::new, src/merkle.rs:81-148 with :162-177 (call sites src/starks.rs:70-72,
 * 92-94, src/fri.rs:351). */
int32_t ms_merkle_commit(ms_ctx* ctx, const void* d_data, uint64_t stride, uint64_t rows, uint64_t width,
                         int32_t deg, uint64_t leafs_per_node, uint64_t inner_children,
                         uint32_t* d_nodes, uint8_t* root32);
/* number of digests MerkleTree::new produces for n_groups leaf groups (0 if the tree is not full) */
uint64_t ms_merkle_node_count(uint64_t n_groups, uint64_t inner_children);

/* ---- a2: trace interpolation ---------------------------------------------------------------- */
/* Per column: coefficients of the degree < N interpolant over the size-N subgroup (iNTT incl. 1/N),
 * natural order in and out.  Replaces: TraceTable::get_trace_polys, src/air.rs:147-160. */
int32_t ms_intt_columns(ms_ctx* ctx, const void* d_evals, uint64_t in_stride, uint64_t n, uint64_t cols,
                        void* d_coeffs, uint64_t out_stride);

/* ---- a3: linear transition constraints -------------------------------------------------------- */
/* out column t = sum_w M[t*W + w] * coeff column w  (M: host, T x W canonical scalars).  Only
 * constraints linear in the trace polynomials are provable by the reference.
 * Replaces: the closures run by TraceTable::derive_constrains, src/air.rs:130-134. */
int32_t ms_linear_constraints(ms_ctx* ctx, const void* d_coeffs, uint64_t stride, uint64_t n, uint64_t w,
                              const void* matrix_host, uint64_t t, void* d_out, uint64_t out_stride);

/* ---- a4: coset low-degree extension ----------------------------------------------------------- */
/* out[c][i] = f_c(shift * w_L^i), L = blowup*n, natural order, for `cols` coefficient columns of
 * length n.  Replaces: the LDE loop src/starks.rs:82-91 (get_coset + evaluate_over_domain +
 * Matrix::add_col). */
int32_t ms_coset_lde(ms_ctx* ctx, const void* d_coeffs, uint64_t in_stride, uint64_t n, uint64_t cols,
                     uint64_t blowup, uint64_t shift, void* d_out, uint64_t out_stride);
/* Host-buffer form of the same call (row-major in/out like the reference's Vec<DensePolynomial> ->
 * Matrix path is NOT assumed: coeffs_host is poly-major [cols][n], out_host is row-major [L][cols]
 * exactly as src/starks.rs:87-91 leaves `constrain_trace`).  H2D + LDE + D2H; used for `e2e`. */
int32_t ms_coset_lde_host(ms_ctx* ctx, const void* coeffs_host, uint64_t n, uint64_t cols, uint64_t blowup,
                          uint64_t shift, void* out_host_rowmajor);

/* ---- a6: constraint mixing ----------------------------------------------------------------------- */
/* out[m] = sum_i r^i f_i[m].  Replaces: src/starks.rs:108-117. */
int32_t ms_mix(ms_ctx* ctx, const void* d_coeffs, uint64_t stride, uint64_t n, uint64_t cols, uint64_t r, void* d_out);

/* ---- a7: DEEP-ALI openings ----------------------------------------------------------------------- */
/* out[q][c] = f_c(z_q) in the extension field (z: host, Q x D; out: host, Q x cols x D).
 * Replaces: src/starks.rs:140-151 (+ extend_poly src/field.rs:23-32). */
int32_t ms_deep_open(ms_ctx* ctx, const void* d_coeffs, uint64_t stride, uint64_t n, uint64_t cols,
                     const void* z_host, uint64_t q, void* out_host);

/* ---- a8-a11: DEEP-FRI ------------------------------------------------------------------------------ */
/* Codeword of an extension polynomial (D planes of n_coeffs_padded = domain/blowup coefficients,
 * zero padded) on the size-`domain` subgroup, natural order, plus its (2,2) Merkle tree.
 * Replaces: FriRound::codeword_commit, src/fri.rs:345-352. */
int32_t ms_fri_commit(ms_ctx* ctx, const void* d_poly, uint64_t poly_stride, uint64_t domain, uint64_t blowup,
                      void* d_codeword, uint64_t cw_stride, uint32_t* d_nodes, uint8_t* root32);
/* d = [f_even(z), f_odd(z)] (host out, 2 x D).  Replaces: get_deep_coeffs, src/fri.rs:354-359. */
int32_t ms_fri_deep_coeffs(ms_ctx* ctx, const void* d_poly, uint64_t stride, uint64_t n_coeffs, const void* z_host,
                           void* d_out_host);
/* next = (f_even + alpha f_odd - (d0 + d1 alpha)) / (x - z), zero padded to n_coeffs/2 entries.
 * Replaces: src/fri.rs:97-101 with fold_poly :361-372. */
int32_t ms_fri_fold(ms_ctx* ctx, const void* d_poly, uint64_t stride, uint64_t n_coeffs, const void* z_host,
                    const void* alpha_host, const void* d_host, void* d_next, uint64_t next_stride);

/* One round of the query phase for q betas (usize values as squeezed, src/fri.rs:121-126): prev / next are two consecutive
 * committed rounds (codewords as D planes, prev's (2,2) tree nodes from ms_fri_commit, prev's polynomial with prev_len
 * coefficients).  Host outputs: points [q][6][D] = x1,y1,x2,y2,x3,y3 (src/fri.rs:148-154); found [2q] = leaf index the value
 * search returns for y1, y2 (first match, src/merkle.rs:216-225); neigh [2q][2][D] leaf neighbours; paths
 * [2q][log2(prev_domain/2)][2][32] sibling pairs, digest bytes (src/merkle.rs:238-265); quot [q][prev_len-2][D] quotient
 * coefficients (src/fri.rs:157-167; nothing written when prev_len < 3).  Any output pointer may be NULL.
 * Replaces: the body of the round loop of Fri::query_phase, src/fri.rs:132-176, with MerkleTree::generate_proof,
 * src/merkle.rs:272-288. */
int32_t ms_fri_query(ms_ctx* ctx, const void* d_prev_poly, uint64_t poly_stride, uint64_t prev_len, const void* d_prev_cw,
                     uint64_t prev_cw_stride, uint64_t prev_domain, const uint32_t* d_prev_nodes, const void* d_next_cw,
                     uint64_t next_cw_stride, const uint64_t* betas_host, uint64_t q, void* points_host, uint64_t* found_host,
                     void* neigh_host, uint8_t* paths_host, void* quot_host);

/* ---- whole prover ------------------------------------------------------------------------------------ */
typedef struct ms_stark_params {
    uint64_t security_bits;  /* StarkConfig::new arguments, src/starks.rs:268-273 */
    uint64_t blowup_factor;
    uint64_t steps;
    uint64_t trace_columns;  /* leafs_per_node of the trace / LDE trees (src/starks.rs:297-302) */
    uint64_t inner_children; /* 2 in the reference (src/starks.rs:299); 4 / 8 = BASELINE configs 3, 5 */
} ms_stark_params;

/* Parameters StarkConfig::new derives (src/starks.rs:274-277, 312-332).  Needs no context (runs without a GPU).  MS_ERR_BAD_SHAPE: fewer
 * than 20 security bits (the reference's assert, src/starks.rs:317-320), steps = 0, a blowup that is not a power of two >= 2, an unknown
 * field id.  ms_stark_proof_bound returns 0 in the same cases. */
int32_t ms_stark_derive(int32_t field, const ms_stark_params* p, uint64_t* rounds, uint64_t* constrain_queries,
                        uint64_t* fri_queries);

/* Upper bound (bytes) of the proof dump for an n x (cols = W + T) problem; 0 on bad parameters. */
uint64_t ms_stark_proof_bound(int32_t field, const ms_stark_params* p, uint64_t n, uint64_t cols);

/* Stark::prove behind the AIR: the caller supplies what `air.trace(&witness)` built -- the row-major
 * padded N x W trace (src/air.rs:73-96) -- and the transition constraints as a T x W matrix of
 * canonical scalars (row t: f_{W+t} = sum_w M[t][w] f_w).  The proof comes back in the canonical
 * dump described in DESIGN.md ("Proof bytes").  If *proof_len is too small the needed size is
 * written back and MS_ERR_BUFFER_TOO_SMALL returned.
 * Replaces: Stark::prove, src/starks.rs:59-169 (host transcript: src/fiatshamir.rs:48-64,96-116). */
int32_t ms_stark_prove(ms_ctx* ctx, const ms_stark_params* p, const void* trace_rowmajor_host, uint64_t n, uint64_t w,
                       const void* constraint_matrix_host, uint64_t t, uint8_t* proof_out, uint64_t* proof_len);
/* Affine transition constraints: row t of the AIR is f_{W+t} = sum_w M[t][w] f_w + c_t (constants_host: T canonical scalars, or
 * NULL).  The reference proves such an AIR like any other (an affine combination still has at most N coefficients, so the
 * `assert_eq!(rest, zero)` of src/starks.rs:119 holds); its closures are free to add a constant polynomial (src/air.rs:61). */
int32_t ms_stark_prove_affine(ms_ctx* ctx, const ms_stark_params* p, const void* trace_rowmajor_host, uint64_t n, uint64_t w,
                              const void* constraint_matrix_host, const void* constants_host, uint64_t t, uint8_t* proof_out,
                              uint64_t* proof_len);
/* Same with the trace already resident on the device (column-major W x N, stride n). */
int32_t ms_stark_prove_device(ms_ctx* ctx, const ms_stark_params* p, const void* d_trace_colmajor, uint64_t n, uint64_t w,
                              const void* constraint_matrix_host, uint64_t t, uint8_t* proof_out, uint64_t* proof_len);
/* Multi-GPU commitments.  The prover runs as one replica per GPU; the two stages whose
 * work shards -- the trace tree (row ranges) and the LDE + its tree (column shards -> all-to-all -> row
 * ranges -> subtree roots -> all-gather) -- are delegated to the host through these hooks, which is
 * where the NCCL exchange lives (ministark_b200/sharded.py drives it with torch.distributed).  A hook
 * returns an MS_* status and must leave the 32 root bytes in root32.  Pointers are device pointers
 * valid on the context's stream.  Replaces the same call sites as ms_merkle_commit / ms_coset_lde
 * (src/starks.rs:70-72 and src/starks.rs:82-94). */
typedef struct ms_commit_hooks {
    void* user;
    int32_t (*trace_commit)(void* user, const void* d_trace_colmajor, uint64_t n, uint64_t w, uint8_t* root32);
    int32_t (*lde_commit)(void* user, const void* d_coeffs_colmajor, uint64_t n, uint64_t cols, uint64_t blowup,
                          uint64_t shift, uint8_t* root32);
    /* non-zero on ranks whose proof bytes nobody reads (every rank but one): every stage that feeds the
     * transcript still runs, so all ranks stay in lock step, but the per-query quotient polynomials -- ~all
     * of the proof bytes, src/fri.rs:167 -- are neither computed nor downloaded (N simultaneous multi-GB
     * PCIe reads would only slow rank 0 down).  proof_out then holds the fixed part and *proof_len the full
     * length. */
    int32_t replica_only;
    /* Sharded proof download (download_world > 1): proof_out is then ONE host buffer shared by all ranks
     * (e.g. POSIX shared memory, page-locked in every process with ms_host_register).  Every replica holds the
     * codewords, so each rank computes and downloads only its share of the quotient polynomials over its own
     * PCIe link, each to its final offset (round robin over all ranks; from 4 ranks on over ranks 1..world-1,
     * because rank 0 alone runs the look-ups and writes the fixed part).  The caller adds a barrier after the
     * call.  replica_only is ignored in this mode. */
    int32_t download_rank;
    int32_t download_world;
} ms_commit_hooks;
/* ms_stark_prove_device with the two commitments routed through `hooks` (NULL members = local). */
int32_t ms_stark_prove_hooked(ms_ctx* ctx, const ms_stark_params* p, const void* d_trace_colmajor, uint64_t n, uint64_t w,
                              const void* constraint_matrix_host, uint64_t t, const ms_commit_hooks* hooks,
                              uint8_t* proof_out, uint64_t* proof_len);
/* Stark::verify on the canonical proof dump.  d_constrains_colmajor: what the reference's caller passes as `constrains`
 * (the Constrains object: the W trace polynomials followed by the T constraint polynomials), as `cols` coefficient columns of
 * length n on the device.  The constraint polynomials are re-evaluated at the Q query points on the device; the transcript
 * replay, the FRI consistency checks and the Merkle paths run on the host.  *accepted = 1 / 0; *failed_check (optional) = the
 * reference line of the first failed check (0 accepted, -2 malformed dump).  strict != 0 also enforces the Merkle paths, which
 * the reference computes and discards (src/fri.rs:237,239).
 * Replaces: Stark::verify, src/starks.rs:171-235, with Fri::verify, src/fri.rs:191-281, and MerkleRoot::check_proof,
 * src/merkle.rs:312-338. */
int32_t ms_stark_verify(ms_ctx* ctx, const ms_stark_params* p, const void* d_constrains_colmajor, uint64_t stride, uint64_t n,
                        uint64_t cols, const uint8_t* proof, uint64_t proof_len, int32_t strict, int32_t* accepted,
                        int32_t* failed_check);

/* ---- multi-GPU inside the library (csrc/comm.cuh, csrc/prover.cuh) ----------------------------------------
 * One prover replica per GPU.  Bind every rank's context to a communicator, then call ms_stark_prove_multi (or any
 * ms_stark_prove*) on all ranks with the same arguments: the trace tree, iNTT, constraints, LDE + its tree, mixing, DEEP
 * openings, the large FRI round trees and the proof download are split over the ranks inside the call; proofs are byte-identical
 * for every number of ranks.  Two backends:
 *   NCCL  (one process per GPU): rank 0 makes a 128-byte id with ms_comm_unique_id, the host distributes it by any means
 *         (MPI, a file, torch.distributed), every rank calls ms_comm_init_nccl.  libnccl.so.2 is dlopen'ed (the copy already
 *         loaded in the process, if any; MINISTARK_NCCL_LIB overrides); failures return MS_ERR_NCCL.  Bulk data never goes
 *         through NCCL: kernels read the peers' buffers over NVLink (CUDA IPC); NCCL carries barriers and 32-byte digests.
 *   local (one process, one host thread per rank): ms_comm_init_local binds `world` contexts (same or different devices;
 *         several ranks may share one GPU) into a group; each context must then be driven by its own thread.
 * No reference call site: the reference is single-threaded (README.md:33). */
int32_t ms_comm_unique_id(uint8_t id128[128]);
int32_t ms_comm_init_nccl(ms_ctx* ctx, const uint8_t id128[128], int32_t rank, int32_t world);
int32_t ms_comm_init_local(ms_ctx* const* ctxs, int32_t world);
/* collective: every rank of the group calls it, concurrently (local backend: on each rank's own thread -- a loop over the
 * contexts on one thread waits for the other ranks forever).  ms_ctx_destroy alone also drops a communicator, without
 * waiting for the peers. */
int32_t ms_comm_destroy(ms_ctx* ctx);
int32_t ms_comm_info(const ms_ctx* ctx, int32_t* rank, int32_t* world, const char** backend);
/* which stages shard (default MS_SHARD_ALL); the proof does not depend on it */
enum { MS_SHARD_TRACE_TREE = 1, MS_SHARD_COLUMNS = 2, MS_SHARD_FRI_TREES = 4, MS_SHARD_DOWNLOAD = 8, MS_SHARD_ALL = 15 };
int32_t ms_set_shard_mask(ms_ctx* ctx, int32_t mask);
/* the split the library uses, for hosts that want to size buffers (and for tests): [a, b) = rank's share of `cols` columns;
 * per_rank / left = leaf groups per rank and digests each rank contributes for a tree of `groups` leaf groups with arity k
 * (returns MS_ERR_BAD_SHAPE when the tree does not split over `world` ranks: the library then builds it on every rank). */
int32_t ms_shard_plan(uint64_t cols, uint64_t groups, uint64_t k, int32_t world, int32_t rank, uint64_t* a, uint64_t* b,
                      uint64_t* per_rank, uint64_t* left);
/* flags of ms_stark_prove_multi: MS_PROOF_SHARED = proof_out is ONE host buffer shared by all ranks (shm / the same pointer),
 * page-locked in every process (ms_host_register): every rank downloads its share of the quotient polynomials to the final
 * offsets, rank 0 writes the rest, and the call returns on every rank when the proof is complete.  Without it rank 0 gets the
 * proof (all ranks with MS_PROOF_ALL_RANKS) and the other ranks may pass proof_out = NULL. */
enum { MS_PROOF_SHARED = 1, MS_PROOF_ALL_RANKS = 2 };
int32_t ms_stark_prove_multi(ms_ctx* ctx, const ms_stark_params* p, const void* d_trace_colmajor, uint64_t n, uint64_t w,
                             const void* constraint_matrix_host, uint64_t t, uint8_t* proof_out, uint64_t* proof_len, int32_t flags);

/* Upper levels of a tree whose level-`n` digests already exist (8 words each, device): hashes groups of
 * `inner_children` digests until one is left (src/merkle.rs:133-140).  Used to join the subtree roots
 * the ranks gathered.  n must be a power of inner_children. */
int32_t ms_merkle_reduce(ms_ctx* ctx, const uint32_t* d_digests, uint64_t n, uint64_t inner_children, uint8_t* root32);

/* One rank's share of MerkleTree::new: leaf groups of an aligned, contiguous row range (rows x width
 * elements starting at d_data, same layout as ms_merkle_commit) hashed and reduced while whole groups of
 * inner_children digests remain; *n_out (< inner_children) digests are left in d_digests_out. */
int32_t ms_merkle_subtree(ms_ctx* ctx, const void* d_data, uint64_t stride, uint64_t rows, uint64_t width, int32_t deg,
                          uint64_t leafs_per_node, uint64_t inner_children, uint32_t* d_digests_out, uint64_t* n_out);

/* ms_merkle_subtree over a row range whose coordinate planes are given one pointer each
 * (plane_ptrs_host: HOST array of width*deg DEVICE pointers, plane p = column p/deg, coordinate p%deg,
 * each pointing at the first row of the range).  The planes may be peer-GPU memory opened with
 * ms_peer_open: the leaf kernel then reads the other ranks' LDE columns over NVLink while it hashes,
 * which replaces the column-to-row exchange in front of the row-sharded LDE tree (src/starks.rs:92-94
 * on a column-sharded matrix).  Same digests as ms_merkle_subtree on the gathered block. */
int32_t ms_merkle_subtree_gather(ms_ctx* ctx, const void* const* plane_ptrs_host, uint64_t rows, uint64_t width, int32_t deg,
                                 uint64_t leafs_per_node, uint64_t inner_children, uint32_t* d_digests_out, uint64_t* n_out);

/* Peer memory for the gather above (CUDA IPC, one process per GPU on one node): a buffer from
 * ms_peer_alloc can be exported as a 64-byte handle, sent to the other ranks by any means, and opened
 * there; the returned pointer is valid in kernels of the opening context's device.  No reference call
 * site: the reference is single-process. */
int32_t ms_peer_alloc(ms_ctx* ctx, uint64_t bytes, void** d_out);
int32_t ms_peer_free(ms_ctx* ctx, void* d_ptr);
int32_t ms_peer_export(ms_ctx* ctx, void* d_ptr, uint8_t* handle64);
int32_t ms_peer_open(ms_ctx* ctx, const uint8_t* handle64, void** d_out);
int32_t ms_peer_close(ms_ctx* ctx, void* d_ptr);

/* Page-lock / unlock caller-owned host memory (cudaHostRegister) so that downloads into it run at PCIe speed;
 * used for the shared proof buffer of the sharded download. */
int32_t ms_host_register(ms_ctx* ctx, void* host_ptr, uint64_t bytes);
int32_t ms_host_unregister(ms_ctx* ctx, void* host_ptr);

/* per-stage device times (ms) of the last ms_stark_prove* call: fills up to `cap` entries, returns count.
 * names[i] points to static strings. */
int32_t ms_stark_last_timings(ms_ctx* ctx, const char** names, float* ms, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif
