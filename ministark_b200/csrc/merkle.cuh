// merkle.cuh -- the reference's Merkle tree, bit for bit (src/merkle.rs:81-177).
//
// Leaf groups: SHA256(concat(to_string(e))) over `lpn` consecutive elements of the ROW-MAJOR
// flattening (calculate_from_leafs, :162-168).  `to_string` is ark-ff's Display: the canonical value
// in decimal, and "QuadExtField(c0 + c1 * u)" (nested for the quartic tower) for extension elements
// (SURVEY.md App. A item 4).  Inner nodes: SHA256(concat of k child digests) (:171-177), level
// order in one array (:119-140).
//
// Device mapping: one thread per leaf group (or per parent node).  The decimal conversion and the
// variable-length concatenation are byte stores into a per-thread 128-byte ring in shared memory
// (word-interleaved across the block: bank-conflict free for equal positions), compressions are
// issued warp-uniformly whenever a thread has >= 64 bytes pending.
#pragma once
#include "common.cuh"
#include "field.cuh"
#include "sha256.cuh"

namespace ms {

constexpr int LEAF_THREADS = 128;
constexpr int LEAF_WORDS = 32;  // per-thread message buffer: two SHA-256 blocks

// ASCII of a 4-digit group: DEC4[c] = "0000".."9999" packed big-endian (first character in the top byte)
__global__ void k_build_dec4(uint32_t* tab) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 10000) return;
    uint32_t d0 = c / 1000, d1 = (c / 100) % 10, d2 = (c / 10) % 10, d3 = c % 10;
    tab[c] = 0x30303030u + ((d0 << 24) | (d1 << 16) | (d2 << 8) | d3);
}

// The message of one leaf group is assembled as big-endian 32-bit words in a per-thread buffer in
// shared memory (word i of thread t at buf[i*LEAF_THREADS + t]: conflict free).  `pos` counts the bytes
// in the buffer: pos >> 2 complete words, then the word with the pos & 3 pending bytes (left aligned,
// zeros below); everything beyond it is don't-care.  Appending a string = one funnel shift per word +
// one store per word at compile-time offsets from that word; no byte stores, no per-byte address
// arithmetic, and one 32-bit counter to maintain per token (the message length is 64 * blocks + pos).
struct LeafStream {
    uint32_t* base;  // word 0 of this thread
    uint32_t pos;    // bytes in the buffer

    // X[0..NW): the string, left aligned, big-endian characters, zero padded; len <= 4*NW bytes
    template <int NW>
    __device__ __forceinline__ void append(const uint32_t (&X)[NW], uint32_t len) {
        uint32_t* wp = base + (pos >> 2) * LEAF_THREADS;
        const uint32_t sh = 8u * (pos & 3u);
        const uint32_t part = *wp;
        wp[0] = part | (X[0] >> sh);
#pragma unroll
        for (int k = 1; k < NW; k++) wp[k * LEAF_THREADS] = __funnelshift_r(X[k], X[k - 1], sh);
        wp[NW * LEAF_THREADS] = __funnelshift_r(0u, X[NW - 1], sh);
        pos += len;
    }
};

constexpr uint32_t pack4(const char* s, int n, int off) {
    uint32_t w = 0;
    for (int i = 0; i < 4; i++) w = (w << 8) | (uint32_t)(off + i < n ? (unsigned char)s[off + i] : 0);
    return w;
}
template <int N>
__device__ __forceinline__ void put_lit(LeafStream& st, const char (&s)[N]) {
    constexpr int len = N - 1;
    static_assert(len >= 1 && len <= 16, "literals are appended in pieces of at most 16 bytes");
    constexpr int NW = (len + 3) / 4;
    uint32_t X[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) X[k] = pack4(s, len, 4 * k);
    st.append<NW>(X, (uint32_t)len);
}

// decimal digits of v (< 2^64), most significant first, no leading zeros; zero prints "0" (or
// nothing when zero_empty): ark-ff Display of a field element
__device__ __forceinline__ void put_decimal(LeafStream& st, uint64_t v, int zero_empty, const uint32_t* __restrict__ dec4) {
    // five 4-digit groups: v = ((((g4)*10^4 + g3)*10^4 + g2)*10^4 + g1)*10^4 + g0
    const uint64_t q1 = v / 100000000ULL;                       // < 1.85e11
    const uint32_t r1 = (uint32_t)(v - q1 * 100000000ULL);      // < 1e8
    const uint32_t g4 = (uint32_t)(q1 / 100000000ULL);          // < 1845
    const uint32_t r2 = (uint32_t)(q1 - (uint64_t)g4 * 100000000ULL);
    const uint32_t g3 = r2 / 10000u, g2 = r2 - g3 * 10000u;
    const uint32_t g1 = r1 / 10000u, g0 = r1 - g1 * 10000u;
    uint32_t A[5] = {__ldg(&dec4[g4]), __ldg(&dec4[g3]), __ldg(&dec4[g2]), __ldg(&dec4[g1]), __ldg(&dec4[g0])};
    // leading zero characters
    uint32_t skip;
    const uint32_t z0 = A[0] ^ 0x30303030u;
    if (z0) {
        skip = (uint32_t)__clz(z0) >> 3;  // 0..3: the common case for 64-bit field elements
    } else {
        skip = 4;
#pragma unroll
        for (int j = 1; j < 5; j++) {
            const uint32_t z = A[j] ^ 0x30303030u;
            if (skip == 4u * j) skip += z ? ((uint32_t)__clz(z) >> 3) : 4u;
        }
        if (skip == 20) skip = zero_empty ? 20 : 19;  // v == 0
        // drop whole leading words (rare path; values below 10^16)
        for (uint32_t wsk = skip >> 2; wsk; wsk--) {
            A[0] = A[1]; A[1] = A[2]; A[2] = A[3]; A[3] = A[4]; A[4] = 0;
        }
    }
    const uint32_t sb = 8u * (skip & 3u);
    uint32_t X[5];
#pragma unroll
    for (int k = 0; k < 4; k++) X[k] = __funnelshift_l(A[k + 1], A[k], sb);
    X[4] = A[4] << sb;
    st.append<5>(X, 20u - skip);
}

// One "token" of the Display string of element coordinates.
template <int DEG>
struct Tokens;
template <>
struct Tokens<1> {
    static constexpr int PER_ELEM = 1;
};
template <>
struct Tokens<2> {
    static constexpr int PER_ELEM = 5;  // "QuadExtField(" c0 " + " c1 " * u)"
};
template <>
struct Tokens<4> {
    static constexpr int PER_ELEM = 11;  // "QuadExtField(" "QuadExtField(" a " + " b " * u) + " "QuadExtField(" c " + " d " * u)" " * u)"
};

// data layout: coordinate d of the element at (row, col) is data[(col*DEG + d)*stride + row];
// flat element f = row*width + col (the reference's row-major Matrix / Vec order).
// GATHER: `data` is instead a device table of width*DEG column pointers (coordinate plane p of this
// row range starts at table[p]); the planes may live in the memory of peer GPUs (NVLink loads), which
// is how a rank hashes its row range of a column-sharded LDE without a separate exchange pass.
template <class F, int DEG, bool GATHER>
__global__ void __launch_bounds__(LEAF_THREADS)
k_leaf_hash(const typename F::T* __restrict__ data, uint64_t stride, uint64_t width, uint64_t lpn, uint64_t n_groups,
            int zero_empty, const uint32_t* __restrict__ dec4, uint32_t* __restrict__ nodes) {
    using T = typename F::T;
    const T* const* __restrict__ planes = reinterpret_cast<const T* const*>(data);
    auto load = [&](uint64_t plane, uint64_t r) -> T { return GATHER ? planes[plane][r] : data[plane * stride + r]; };
    __shared__ uint32_t buf[LEAF_WORDS * LEAF_THREADS];
    const uint64_t g = (uint64_t)blockIdx.x * LEAF_THREADS + threadIdx.x;
    const bool live = g < n_groups;
    LeafStream st;
    st.base = buf + threadIdx.x;
    st.pos = 0;
    st.base[0] = 0;
    uint32_t h[8];
    sha256_init(h);
    // 32-bit token / column counters (the host side bounds lpn and width); only the row index is 64-bit
    const uint32_t ntok = live ? (uint32_t)lpn * Tokens<DEG>::PER_ELEM : 0;
    const uint32_t wid = (uint32_t)width;
    uint32_t tok = 0;
    const uint64_t f = g * lpn;    // flat index of the first element
    uint64_t row = live ? f / width : 0;
    uint32_t col = live ? (uint32_t)(f % width) : 0;
    int sub = 0;                   // token index inside the current element
    bool padded = !live, finished = !live;
    uint32_t nblk = 0;             // blocks compressed so far
    // base-field leaves out of one matrix: the element pointer walks along the row (+stride per column, back to the
    // next row's first column at the end) instead of being recomputed from (col, row) per element
    const T* ep = data + (uint64_t)col * stride + row;
    const int64_t ep_wrap = 1 - (int64_t)((width - 1) * stride);
    // base-field leaves: the element of the next token is loaded one token ahead, so that the load
    // (HBM, or NVLink in GATHER mode) is in flight during the decimal conversion / compression
    // (GATHER only: measured 3 % slower on local HBM, where 32 resident warps already hide the latency)
    T ahead = (GATHER && DEG == 1 && ntok) ? load(col, row) : (T)0;
    while (__any_sync(0xffffffffu, !finished)) {
        // ---- fill: append tokens until a full block is pending
        while (st.pos < 64 && tok < ntok) {
            if (DEG == 1 && GATHER) {
                const T v = ahead;
                if (tok + 1 < ntok) {
                    const bool wrap = col + 1 == wid;
                    ahead = load(wrap ? 0 : col + 1, wrap ? row + 1 : row);
                }
                put_decimal(st, (uint64_t)v, zero_empty, dec4);
            } else if (DEG == 1) {
                put_decimal(st, (uint64_t)*ep, zero_empty, dec4);
                ep += (col + 1 == wid) ? ep_wrap : (int64_t)stride;
            } else if (DEG == 2) {
                switch (sub) {
                    case 0: put_lit(st, "QuadExtField("); break;
                    case 1: put_decimal(st, (uint64_t)load(((uint64_t)col * 2 + 0), row), zero_empty, dec4); break;
                    case 2: put_lit(st, " + "); break;
                    case 3: put_decimal(st, (uint64_t)load(((uint64_t)col * 2 + 1), row), zero_empty, dec4); break;
                    default: put_lit(st, " * u)"); break;
                }
            } else {
                switch (sub) {
                    case 0: put_lit(st, "QuadExtField("); break;
                    case 1: put_lit(st, "QuadExtField("); break;
                    case 2: put_decimal(st, (uint64_t)load(((uint64_t)col * 4 + 0), row), zero_empty, dec4); break;
                    case 3: put_lit(st, " + "); break;
                    case 4: put_decimal(st, (uint64_t)load(((uint64_t)col * 4 + 1), row), zero_empty, dec4); break;
                    case 5: put_lit(st, " * u) + "); break;
                    case 6: put_lit(st, "QuadExtField("); break;
                    case 7: put_decimal(st, (uint64_t)load(((uint64_t)col * 4 + 2), row), zero_empty, dec4); break;
                    case 8: put_lit(st, " + "); break;
                    case 9: put_decimal(st, (uint64_t)load(((uint64_t)col * 4 + 3), row), zero_empty, dec4); break;
                    default: put_lit(st, " * u) * u)"); break;
                }
            }
            tok++;
            if (++sub == Tokens<DEG>::PER_ELEM) {
                sub = 0;
                if (++col == wid) { col = 0; row++; }
            }
        }
        // ---- padding once the message is complete (FIPS 180-4 5.1.1)
        if (!padded && st.pos < 64 && tok >= ntok) {
            const uint64_t bits = ((uint64_t)nblk * 64 + st.pos) * 8;
            const uint32_t one[1] = {0x80000000u};
            st.append<1>(one, 1);
            const uint32_t nw = (st.pos + 3u) >> 2;  // words holding data
            const uint32_t end_words = nw <= 14 ? 16u : 32u;
            for (uint32_t i = nw; i < end_words - 2; i++) st.base[i * LEAF_THREADS] = 0;
            st.base[(end_words - 2) * LEAF_THREADS] = (uint32_t)(bits >> 32);
            st.base[(end_words - 1) * LEAF_THREADS] = (uint32_t)bits;
            st.pos = 4u * end_words;
            padded = true;
        }
        // ---- compress one pending block and slide the rest of the buffer down
        if (!finished && st.pos >= 64) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = st.base[i * LEAF_THREADS];
            sha256_compress(h, w);
            st.pos -= 64;
            nblk++;
            if (padded) {
                if (st.pos == 0) finished = true;
                else {
#pragma unroll
                    for (int i = 0; i < 16; i++) st.base[i * LEAF_THREADS] = st.base[(i + 16) * LEAF_THREADS];
                }
            } else {
                // at most 6 complete words plus the partial one were beyond the block
#pragma unroll
                for (int i = 0; i < 7; i++) st.base[i * LEAF_THREADS] = st.base[(i + 16) * LEAF_THREADS];
            }
        }
    }
    if (live) {
        uint4* o = reinterpret_cast<uint4*>(nodes + g * 8);
        o[0] = make_uint4(h[0], h[1], h[2], h[3]);
        o[1] = make_uint4(h[4], h[5], h[6], h[7]);
    }
}

// parent p = SHA256(child[p*k] .. child[p*k + k-1]); digests are 8 state words
template <int K>
__device__ __forceinline__ void node_hash_one(const uint32_t* children, uint32_t* parents, uint64_t p) {
    const uint4* src = reinterpret_cast<const uint4*>(children + p * 8 * K);
    uint32_t st[8];
    sha256_init(st);
#pragma unroll
    for (int blk = 0; blk < K / 2; blk++) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint4 v = src[blk * 4 + i];
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
        sha256_compress(st, w);
    }
    sha256_compress_padblock<32u * K * 8u>(st);
    uint4* o = reinterpret_cast<uint4*>(parents + p * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}
// The top of a tree in one launch: from a level of `lv` <= TOP_NODES digests down to `stop` digests, one CTA,
// levels separated by __syncthreads (the ~log_k(lv) launches it replaces cost more than the hashing).
// `nodes` points at the first digest of the starting level; the levels follow each other as in d_nodes.
constexpr int TOP_THREADS = 512;
constexpr uint64_t TOP_NODES = 512;  // (a 4096-digest start costs 16 sequential hash times in the one CTA, 65 us per tree; 512: 9, the three wider levels go through k_node_hash)
template <int K>
__global__ void __launch_bounds__(TOP_THREADS)
k_tree_top(uint32_t* nodes, uint64_t lv, uint64_t stop) {
    uint64_t src = 0;
    while (lv > stop) {
        const uint64_t np = lv / K;
        for (uint64_t p = threadIdx.x; p < np; p += TOP_THREADS) node_hash_one<K>(nodes + src * 8, nodes + (src + lv) * 8, p);
        src += lv;
        lv = np;
        __syncthreads();
    }
}
template <int K>
__global__ void __launch_bounds__(256)
k_node_hash(const uint32_t* __restrict__ children, uint32_t* __restrict__ parents, uint64_t n_parents) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_parents) return;
    const uint4* src = reinterpret_cast<const uint4*>(children + p * 8 * K);
    uint32_t st[8];
    sha256_init(st);
#pragma unroll
    for (int blk = 0; blk < K / 2; blk++) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint4 v = src[blk * 4 + i];
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
        sha256_compress(st, w);
    }
    sha256_compress_padblock<32u * K * 8u>(st);
    uint4* o = reinterpret_cast<uint4*>(parents + p * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}

inline uint64_t merkle_node_count(uint64_t n_groups, uint64_t k) {
    if (k < 2 || !is_pow2(k) || !is_pow2(n_groups)) return 0;
    int lgk = ilog2(k), lg1 = ilog2(n_groups);
    if (lg1 % lgk) return 0;
    uint64_t total = 0, lv = n_groups;
    while (true) {
        total += lv;
        if (lv == 1) break;
        lv /= k;
    }
    return total;
}

inline void digest_words_to_bytes(const uint32_t* w, uint8_t* out) {
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(w[i] >> 24);
        out[4 * i + 1] = (uint8_t)(w[i] >> 16);
        out[4 * i + 2] = (uint8_t)(w[i] >> 8);
        out[4 * i + 3] = (uint8_t)w[i];
    }
}

// levels above the first: d_nodes holds `n1` level-1 digests followed by room for every upper level;
// climbs until `stop` digests are left and returns their offset (in digests) through *top
inline int merkle_climb(Ctx* c, uint32_t* d_nodes, uint64_t n1, uint64_t k, uint64_t stop, uint64_t* top);
inline int merkle_upper_levels(Ctx* c, uint32_t* d_nodes, uint64_t n1, uint64_t k) {
    uint64_t top = 0;
    return merkle_climb(c, d_nodes, n1, k, 1, &top);
}
inline int merkle_climb(Ctx* c, uint32_t* d_nodes, uint64_t n1, uint64_t k, uint64_t stop, uint64_t* top) {
    uint64_t src = 0, lv = n1;
    while (lv > stop) {
        if (lv <= TOP_NODES) {  // the rest of the tree in one launch
            uint32_t* base = d_nodes + src * 8;
            prof_begin(c, "k_tree_top");
            switch (k) {
                case 2: k_tree_top<2><<<1, TOP_THREADS, 0, c->stream>>>(base, lv, stop); break;
                case 4: k_tree_top<4><<<1, TOP_THREADS, 0, c->stream>>>(base, lv, stop); break;
                case 8: k_tree_top<8><<<1, TOP_THREADS, 0, c->stream>>>(base, lv, stop); break;
                default: k_tree_top<16><<<1, TOP_THREADS, 0, c->stream>>>(base, lv, stop); break;
            }
            prof_end(c);
            MS_LAUNCH_CHECK(c);
            while (lv > stop) { src += lv; lv /= k; }
            break;
        }
        uint64_t np = lv / k;
        const uint32_t* ch = d_nodes + src * 8;
        uint32_t* pa = d_nodes + (src + lv) * 8;
        unsigned nb = (unsigned)((np + 255) / 256);
        prof_begin(c, "k_node_hash");
        switch (k) {
            case 2: k_node_hash<2><<<nb, 256, 0, c->stream>>>(ch, pa, np); break;
            case 4: k_node_hash<4><<<nb, 256, 0, c->stream>>>(ch, pa, np); break;
            case 8: k_node_hash<8><<<nb, 256, 0, c->stream>>>(ch, pa, np); break;
            default: k_node_hash<16><<<nb, 256, 0, c->stream>>>(ch, pa, np); break;
        }
        prof_end(c);
        MS_LAUNCH_CHECK(c);
        src += lv;
        lv = np;
    }
    *top = src;
    return MS_OK;
}

// join already hashed digests (e.g. the subtree roots gathered from the other GPUs) into one root
inline int merkle_reduce(Ctx* c, const uint32_t* d_digests, uint64_t n, uint64_t k, uint8_t* root32) {
    const uint64_t total = merkle_node_count(n, k);
    if (total == 0 || (k != 2 && k != 4 && k != 8 && k != 16) || !root32)
        return fail(c, MS_ERR_BAD_SHAPE, "merkle_reduce: %llu digests do not form a full %llu-ary tree", (unsigned long long)n, (unsigned long long)k);
    Scratch nodes(c);
    MS_TRY(nodes.alloc(total * 32));
    MS_CUDA(c, cudaMemcpyAsync(nodes.p, d_digests, n * 32, cudaMemcpyDeviceToDevice, c->stream));
    MS_TRY(merkle_upper_levels(c, nodes.as<uint32_t>(), n, k));
    uint32_t w[8];
    MS_CUDA(c, cudaMemcpyAsync(w, nodes.as<uint32_t>() + (total - 1) * 8, 32, cudaMemcpyDeviceToHost, c->stream));
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    digest_words_to_bytes(w, root32);
    return MS_OK;
}

inline int ensure_dec4(Ctx* c) {
    if (c->dec4) return MS_OK;
    uint32_t* t;
    MS_CUDA(c, cudaMalloc(&t, 10000 * sizeof(uint32_t)));
    k_build_dec4<<<(10000 + 255) / 256, 256, 0, c->stream>>>(t);
    MS_LAUNCH_CHECK(c);
    c->dec4 = t;
    return MS_OK;
}

// gather: d_data is a device table of width*deg plane pointers (see k_leaf_hash)
template <class F>
int merkle_leaf_level(Ctx* c, const typename F::T* d_data, uint64_t stride, uint64_t width, int deg, uint64_t lpn, uint64_t n1,
                      uint32_t* d_nodes, bool gather = false) {
    unsigned blocks = (unsigned)((n1 + LEAF_THREADS - 1) / LEAF_THREADS);
    if (lpn > (1ULL << 27) || width >= (1ULL << 32))  // the leaf kernel counts tokens and columns in 32 bits
        return fail(c, MS_ERR_UNSUPPORTED, "leaf groups of %llu elements / rows of %llu columns are beyond the leaf kernel's counters", (unsigned long long)lpn, (unsigned long long)width);
    prof_begin(c, "k_leaf_hash");
    MS_TRY(ensure_dec4(c));
    if (gather) {
        if (deg == 1) k_leaf_hash<F, 1, true><<<blocks, LEAF_THREADS, 0, c->stream>>>(d_data, stride, width, lpn, n1, c->zero_display_empty, c->dec4, d_nodes);
        else k_leaf_hash<F, F::D, true><<<blocks, LEAF_THREADS, 0, c->stream>>>(d_data, stride, width, lpn, n1, c->zero_display_empty, c->dec4, d_nodes);
    } else if (deg == 1) k_leaf_hash<F, 1, false><<<blocks, LEAF_THREADS, 0, c->stream>>>(d_data, stride, width, lpn, n1, c->zero_display_empty, c->dec4, d_nodes);
    else k_leaf_hash<F, F::D, false><<<blocks, LEAF_THREADS, 0, c->stream>>>(d_data, stride, width, lpn, n1, c->zero_display_empty, c->dec4, d_nodes);
    prof_end(c);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

// MerkleTree::new.  d_nodes: caller buffer for all nodes or nullptr (scratch); root32: host or nullptr.
template <class F>
int merkle_commit(Ctx* c, const typename F::T* d_data, uint64_t stride, uint64_t rows, uint64_t width, int deg,
                  uint64_t lpn, uint64_t k, uint32_t* d_nodes, uint8_t* root32) {
    const uint64_t n_elems = rows * width;
    if (lpn == 0 || n_elems == 0 || n_elems % lpn) return fail(c, MS_ERR_BAD_SHAPE, "leaf count %llu not divisible by leafs_per_node %llu (merkle.rs:99)", (unsigned long long)n_elems, (unsigned long long)lpn);
    const uint64_t n1 = n_elems / lpn;
    const uint64_t total = merkle_node_count(n1, k);
    if (total == 0 || (k != 2 && k != 4 && k != 8 && k != 16))
        return fail(c, MS_ERR_BAD_SHAPE, "Tree is not full! %llu leaf groups, inner_children %llu (merkle.rs:93-104)", (unsigned long long)n1, (unsigned long long)k);
    if (deg != 1 && deg != F::D) return fail(c, MS_ERR_BAD_SHAPE, "deg must be 1 or the extension degree");
    Scratch own(c);
    if (!d_nodes) {
        MS_TRY(own.alloc(total * 32));
        d_nodes = own.as<uint32_t>();
    }
    MS_TRY(merkle_leaf_level<F>(c, d_data, stride, width, deg, lpn, n1, d_nodes));
    MS_TRY(merkle_upper_levels(c, d_nodes, n1, k));
    if (root32) {
        uint32_t w[8];
        MS_CUDA(c, cudaMemcpyAsync(w, d_nodes + (total - 1) * 8, 32, cudaMemcpyDeviceToHost, c->stream));
        MS_CUDA(c, cudaStreamSynchronize(c->stream));
        digest_words_to_bytes(w, root32);
    }
    return MS_OK;
}

// A rank's share of a tree: hash the leaf groups of a contiguous, aligned range of
// rows and climb while whole groups of k digests remain.  Leaves *n_out < k digests (1 when the range
// is itself a full subtree) in d_out; the caller gathers every rank's digests and joins them with
// merkle_reduce.  The range must hold a power-of-two number of leaf groups.
template <class F>
int merkle_subtree(Ctx* c, const typename F::T* d_data, uint64_t stride, uint64_t rows, uint64_t width, int deg, uint64_t lpn,
                   uint64_t k, uint32_t* d_out, uint64_t* n_out, bool gather = false) {
    const uint64_t n_elems = rows * width;
    if (lpn == 0 || n_elems == 0 || n_elems % lpn) return fail(c, MS_ERR_BAD_SHAPE, "leaf count %llu not divisible by leafs_per_node %llu (merkle.rs:99)", (unsigned long long)n_elems, (unsigned long long)lpn);
    const uint64_t n1 = n_elems / lpn;
    if (!is_pow2(n1) || (k != 2 && k != 4 && k != 8 && k != 16)) return fail(c, MS_ERR_BAD_SHAPE, "subtree of %llu leaf groups, inner_children %llu", (unsigned long long)n1, (unsigned long long)k);
    if (deg != 1 && deg != F::D) return fail(c, MS_ERR_BAD_SHAPE, "deg must be 1 or the extension degree");
    uint64_t total = 0, lv = n1;
    while (true) {
        total += lv;
        if (lv == 1 || lv % k) break;
        lv /= k;
    }
    const uint64_t rem = lv, full = n1 / rem;  // `rem` independent full subtrees of `full` leaf groups
    Scratch nodes(c);
    MS_TRY(nodes.alloc(total * 32));
    uint32_t* d_nodes = nodes.as<uint32_t>();
    MS_TRY(merkle_leaf_level<F>(c, d_data, stride, width, deg, lpn, n1, d_nodes, gather));
    // the same climb as merkle_upper_levels, stopping at `rem` digests
    uint64_t src = 0;
    MS_TRY(merkle_climb(c, d_nodes, n1, k, rem, &src));
    (void)full;
    MS_CUDA(c, cudaMemcpyAsync(d_out, d_nodes + src * 8, rem * 32, cudaMemcpyDeviceToDevice, c->stream));
    if (n_out) *n_out = rem;
    return MS_OK;
}

}  // namespace ms
