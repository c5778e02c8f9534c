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

struct LeafRing {
    uint32_t* w;     // word 0 of this thread; word i at w[i * LEAF_THREADS]
    uint32_t base;   // byte offset of the first pending byte (0 or 64)
    uint32_t pos;    // pending bytes
    __device__ __forceinline__ void put(uint32_t byte) {
        uint32_t k = (base + pos) & 127u;
        reinterpret_cast<unsigned char*>(w + (k >> 2) * LEAF_THREADS)[3 - (k & 3u)] = (unsigned char)byte;
        pos++;
    }
};

template <int N>
__device__ __forceinline__ void put_lit(LeafRing& r, const char (&s)[N]) {
#pragma unroll
    for (int i = 0; i < N - 1; i++) r.put((unsigned char)s[i]);
}

// decimal digits of v (< 2^64), most significant first, no leading zeros; zero prints "0" (or
// nothing when zero_empty)
__device__ __forceinline__ void put_decimal(LeafRing& r, uint64_t v, int zero_empty) {
    // split into four 5-digit limbs: v = ((l3 * 10^5 + l2) * 10^5 + l1) * 10^5 + l0
    uint64_t hi = v / 10000000000ULL;          // < 1.85e9
    uint64_t lo = v - hi * 10000000000ULL;     // < 1e10
    uint32_t l3 = (uint32_t)hi / 100000u, l2 = (uint32_t)hi % 100000u;
    uint32_t l1 = (uint32_t)(lo / 100000ULL), l0 = (uint32_t)(lo - (uint64_t)l1 * 100000ULL);
    uint32_t limbs[4] = {l3, l2, l1, l0};
    unsigned char d[20];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t x = limbs[i];
#pragma unroll
        for (int k = 4; k >= 0; k--) {
            uint32_t qd = x / 10u;
            d[i * 5 + k] = (unsigned char)(x - qd * 10u);
            x = qd;
        }
    }
    int skip = 0;
    bool lead = true;
#pragma unroll
    for (int i = 0; i < 19; i++) {
        lead = lead && (d[i] == 0);
        skip += lead ? 1 : 0;
    }
    if (zero_empty && v == 0) skip = 20;
#pragma unroll
    for (int i = 0; i < 20; i++)
        if (i >= skip) r.put('0' + d[i]);
}

// One "token" of the Display string of element coordinates (tokens keep each burst <= 26 bytes).
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
    static constexpr int PER_ELEM = 9;  // "QuadExtField(QuadExtField(" a " + " b " * u) + QuadExtField(" c " + " d " * u) * u)"
};

// data layout: coordinate d of the element at (row, col) is data[(col*DEG + d)*stride + row];
// flat element f = row*width + col (the reference's row-major Matrix / Vec order).
template <class F, int DEG>
__global__ void __launch_bounds__(LEAF_THREADS)
k_leaf_hash(const typename F::T* __restrict__ data, uint64_t stride, uint64_t width, uint64_t lpn, uint64_t n_groups,
            int zero_empty, uint32_t* __restrict__ nodes) {
    __shared__ uint32_t ring[32 * LEAF_THREADS];
    const uint64_t g = (uint64_t)blockIdx.x * LEAF_THREADS + threadIdx.x;
    const bool live = g < n_groups;
    LeafRing r;
    r.w = ring + threadIdx.x;
    r.base = 0;
    r.pos = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) r.w[i * LEAF_THREADS] = 0;
    uint32_t st[8];
    sha256_init(st);
    const uint64_t ntok = live ? lpn * Tokens<DEG>::PER_ELEM : 0;
    uint64_t tok = 0, total_bytes = 0;
    uint64_t f = g * lpn;          // flat index of the current element
    uint64_t row = live ? f / width : 0, col = live ? f % width : 0;
    int sub = 0;                   // token index inside the current element
    bool padded = !live, finished = !live;
    while (__any_sync(0xffffffffu, !finished)) {
        // ---- fill: append tokens until a full block is pending
        while (r.pos < 64 && tok < ntok) {
            const uint32_t before = r.pos;
            if (DEG == 1) {
                put_decimal(r, (uint64_t)data[col * stride + row], zero_empty);
            } else if (DEG == 2) {
                switch (sub) {
                    case 0: put_lit(r, "QuadExtField("); break;
                    case 1: put_decimal(r, (uint64_t)data[(col * 2 + 0) * stride + row], zero_empty); break;
                    case 2: put_lit(r, " + "); break;
                    case 3: put_decimal(r, (uint64_t)data[(col * 2 + 1) * stride + row], zero_empty); break;
                    default: put_lit(r, " * u)"); break;
                }
            } else {
                switch (sub) {
                    case 0: put_lit(r, "QuadExtField(QuadExtField("); break;
                    case 1: put_decimal(r, (uint64_t)data[(col * 4 + 0) * stride + row], zero_empty); break;
                    case 2: put_lit(r, " + "); break;
                    case 3: put_decimal(r, (uint64_t)data[(col * 4 + 1) * stride + row], zero_empty); break;
                    case 4: put_lit(r, " * u) + QuadExtField("); break;
                    case 5: put_decimal(r, (uint64_t)data[(col * 4 + 2) * stride + row], zero_empty); break;
                    case 6: put_lit(r, " + "); break;
                    case 7: put_decimal(r, (uint64_t)data[(col * 4 + 3) * stride + row], zero_empty); break;
                    default: put_lit(r, " * u) * u)"); break;
                }
            }
            total_bytes += r.pos - before;
            tok++;
            if (++sub == Tokens<DEG>::PER_ELEM) {
                sub = 0;
                if (++col == width) { col = 0; row++; }
            }
        }
        // ---- padding once the message is complete (FIPS 180-4 5.1.1)
        if (!padded && r.pos < 64 && tok >= ntok) {
            r.put(0x80);
            uint32_t end = (r.pos <= 56) ? 64u : 128u;  // bytes beyond are already zero
            uint64_t bits = total_bytes * 8;
            r.pos = end - 8;
#pragma unroll
            for (int i = 7; i >= 0; i--) r.put((uint32_t)(bits >> (8 * i)) & 0xffu);
            padded = true;
        }
        // ---- compress one pending block
        if (!finished && r.pos >= 64) {
            uint32_t w[16];
            const uint32_t wb = r.base >> 2;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                w[i] = r.w[(wb + i) * LEAF_THREADS];
                r.w[(wb + i) * LEAF_THREADS] = 0;
            }
            sha256_compress(st, w);
            r.base ^= 64u;
            r.pos -= 64;
            if (padded && r.pos == 0) finished = true;
        }
    }
    if (live) {
        uint4* o = reinterpret_cast<uint4*>(nodes + g * 8);
        o[0] = make_uint4(st[0], st[1], st[2], st[3]);
        o[1] = make_uint4(st[4], st[5], st[6], st[7]);
    }
}

// parent p = SHA256(child[p*k] .. child[p*k + k-1]); digests are 8 state words
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
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
    w[0] = 0x80000000u;
    w[15] = 32u * K * 8u;
    sha256_compress(st, w);
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

// levels above the first: d_nodes holds `n1` level-1 digests followed by room for every upper level
inline int merkle_upper_levels(Ctx* c, uint32_t* d_nodes, uint64_t n1, uint64_t k) {
    uint64_t src = 0, lv = n1;
    while (lv > 1) {
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

template <class F>
int merkle_leaf_level(Ctx* c, const typename F::T* d_data, uint64_t stride, uint64_t width, int deg, uint64_t lpn, uint64_t n1,
                      uint32_t* d_nodes) {
    unsigned blocks = (unsigned)((n1 + LEAF_THREADS - 1) / LEAF_THREADS);
    prof_begin(c, "k_leaf_hash");
    if (deg == 1) k_leaf_hash<F, 1><<<blocks, LEAF_THREADS, 0, c->stream>>>(d_data, stride, width, lpn, n1, c->zero_display_empty, d_nodes);
    else k_leaf_hash<F, F::D><<<blocks, LEAF_THREADS, 0, c->stream>>>(d_data, stride, width, lpn, n1, c->zero_display_empty, d_nodes);
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
                   uint64_t k, uint32_t* d_out, uint64_t* n_out) {
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
    MS_TRY(merkle_leaf_level<F>(c, d_data, stride, width, deg, lpn, n1, d_nodes));
    // the same level-by-level climb as merkle_upper_levels, stopping at `rem` digests
    uint64_t src = 0;
    lv = n1;
    while (lv > rem) {
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
    (void)full;
    MS_CUDA(c, cudaMemcpyAsync(d_out, d_nodes + src * 8, rem * 32, cudaMemcpyDeviceToDevice, c->stream));
    if (n_out) *n_out = rem;
    return MS_OK;
}

}  // namespace ms
