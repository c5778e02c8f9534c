// ntt.cuh -- batched radix-2 (i)NTT / coset low-degree extension over column-major matrices.
//
// One primitive serves trace interpolation (src/air.rs:147-160), the LDE loop
// (src/starks.rs:82-91) and every FRI codeword (src/fri.rs:345-351):
//
//   lde_batch:  out[c][B*k + j] = sum_{m<N} in[c][m] * (s_j * w^k)^m ,   s_j = shift * w_L^j ,
//               k in [0,N), j in [0,B)          (B = 1, shift = 1, w = w_N^-1, scale = 1/N: iNTT)
//
// i.e. the size-L evaluation on the coset is B coset transforms of size N (the first log2 B stages
// of the zero padded size-L transform are pure replication and are never executed).
//
// Work decomposition.  N = n1 * n2 (n2 = 2^a, n1 = 2^b) is covered by at most two passes over
// HBM; both passes run the same kernel (k_ntt_fixed for the planner's shapes, k_ntt_tile otherwise) on a
// "tile" = [2^a' points] x [2^beta batch entries] of 8192 elements that lives in shared memory:
//
//   pass 1  transform along m2 (stride n1 in the input) for R consecutive m1 and all B cosets.
//           The coset shift is folded into the stage twiddles t1[j][2^s+q] = s_j^(N/2^(s+1)) w_(2^(s+1))^q
//           (no "distribute powers" pass), the B replicas are made on load, and the inter-pass
//           factor ft = scale * s_j^m1 * w_N^(m1 k2) is applied on store.
//   pass 2  transform along m1 for R2 consecutive k2 and all B cosets; plain twiddles; natural-order
//           output.  With R = 1 the intermediate has the index map of the output and pass 2 runs
//           in place (no temporary: the 2^24 x 64 LDE needs no second 32 GiB buffer).
//
// Inside a tile the log2(points) DIT stages are grouped in rounds of G <= 4 stages.  In a round a
// thread owns "slots" of 2^G elements whose positions differ in the round's G bits, keeps them in
// registers for the G stages, and only touches shared memory at round boundaries (XOR-swizzled, so
// both the contiguous and the strided access patterns are bank-conflict free).  The first round
// loads straight from HBM into registers (bit-reversed gather), the last one stores straight to
// HBM, so a tile costs ceil(a/4) - 1 shared-memory exchanges instead of one per stage.
//
// Arithmetic (Goldilocks): butterflies run on lazily reduced values.  mul() returns a canonical
// product, add()/sub() take (lazy, canonical) and return lazy values in [0, 2^64); carries are
// folded back with 2^64 = 2^32 - 1 (mod p) through PTX carry chains (see Fast<GL>).  Everything
// written to HBM is canonical.  In the fixed-shape kernel a Goldilocks round is "multiply element r
// by beta^brev(r), then a plain DFT-2^G whose twiddles are powers of two" (ShiftTw, gl_shift_dft,
// tw16_entry): one general product per element per round, shifts inside.  BabyBear keeps canonical
// data, Montgomery-form twiddles and radix-2 butterflies; so does the generic k_ntt_tile.
#pragma once
#include "common.cuh"
#include "field.cuh"

namespace ms {

constexpr int NTT_MAXLOG = 14;         // largest in-tile transform (plain twiddle table size)
constexpr int NTT_LOG_TILE_PREF = 13;  // tile: 8192 elements (64 KB Goldilocks, 3 CTAs / SM)
constexpr int NTT_THREADS = 256;
constexpr int NTT_MAXG = 4;            // stages per round: 16 elements in registers
#ifndef MS_NTT_MINB
#define MS_NTT_MINB 3  // CTAs per SM for 64 KB tiles (80 registers, no spills; +4% over 2 on B200)
#endif

#ifdef __CUDA_ARCH__
#define MS_LDG(p) __ldg(p)
#else
#define MS_LDG(p) (*(p))
#endif

MS_HD uint32_t brev_bits(uint32_t x, int bits) {
#ifdef __CUDA_ARCH__
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

// ------------------------------------------------------------------------------------------------
// Butterfly arithmetic.  Contract: mul(x: any representative, w: twiddle form) -> canonical;
// add/sub(a: lazy, t: canonical) -> lazy; canon(lazy) -> canonical; to_tw(canonical) -> twiddle form.
// ------------------------------------------------------------------------------------------------
template <class F>
struct Fast;

// 2^64 mod p = 2^32 - 1 kept in constant memory on purpose.  Measured on B200 (scratch/ubench/pipes.cu,
// SM clocks per warp instruction per scheduler): IADD3 / LOP3 / IMAD 2.0, IMAD.WIDE 2.6, IMAD.HI 5.4.
// With the literal 0xFFFFFFFF ptxas rewrites every "x * (2^32-1) + y" into IMAD.HI + IADD3 (7.4 clocks);
// with an operand it cannot see through it stays one IMAD.WIDE (2.6 clocks).
#ifdef __CUDACC__
__constant__ uint32_t GL_EPS_OPAQUE = 0xFFFFFFFFu;
__constant__ uint32_t BB_P_OPAQUE = 2013265921u;
#endif

template <>
struct Fast<GL> {
    using T = uint64_t;
    static MS_HD T to_tw(T w) { return w; }
    static MS_HD T mul(T a, T b) {
#ifdef __CUDA_ARCH__
        // 128-bit product c3..c0, then c0 + c1 2^32 + c2 (2^32 - 1) - c3 with every carry folded
        // back once; the last fold also subtracts p when the sum landed in [p, 2^64).
        uint32_t r0, r1;
        const uint32_t eps = GL_EPS_OPAQUE;
        asm("{\n\t"
            ".reg .u32 c0, c1, c2, c3, m, k, d;\n\t"
            ".reg .u64 pp;\n\t"
            "mul.wide.u32 pp, %2, %4;\n\t"
            "mov.b64 {c0, c1}, pp;\n\t"
            "mad.lo.cc.u32 c1, %2, %5, c1;\n\t"
            "madc.hi.u32 c2, %2, %5, 0;\n\t"
            "mad.lo.cc.u32 c1, %3, %4, c1;\n\t"
            "madc.hi.cc.u32 c2, %3, %4, c2;\n\t"
            "addc.u32 c3, 0, 0;\n\t"
            "mad.lo.cc.u32 c2, %3, %5, c2;\n\t"
            "madc.hi.u32 c3, %3, %5, c3;\n\t"
            "sub.cc.u32 c0, c0, c3;\n\t"
            "subc.cc.u32 c1, c1, 0;\n\t"
            "subc.u32 m, 0, 0;\n\t"
            "sub.cc.u32 c0, c0, m;\n\t"
            "subc.u32 c1, c1, 0;\n\t"
            "mad.lo.cc.u32 c0, c2, %6, c0;\n\t"
            "madc.hi.cc.u32 c1, c2, %6, c1;\n\t"
            "addc.u32 k, 0, 0;\n\t"
            "add.cc.u32 d, c0, 0xFFFFFFFF;\n\t"
            "addc.cc.u32 d, c1, 0;\n\t"
            "addc.u32 k, k, 0;\n\t"
            "mad.lo.cc.u32 %0, k, %6, c0;\n\t"
            "madc.hi.u32 %1, k, %6, c1;\n\t"
            "}"
            : "=r"(r0), "=r"(r1)
            : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)), "r"(eps));
        return ((uint64_t)r1 << 32) | r0;
#else
        return GL::mul(a % GL::P, b % GL::P);
#endif
    }
    static MS_HD T add(T a, T t) {
#ifdef __CUDA_ARCH__
        uint32_t r0, r1;
        const uint32_t eps = GL_EPS_OPAQUE;
        asm("{\n\t.reg .u32 k, s0, s1;\n\t"
            "add.cc.u32 s0, %2, %4;\n\t"
            "addc.cc.u32 s1, %3, %5;\n\t"
            "addc.u32 k, 0, 0;\n\t"
            "mad.lo.cc.u32 %0, k, %6, s0;\n\t"
            "madc.hi.u32 %1, k, %6, s1;\n\t}"
            : "=r"(r0), "=r"(r1)
            : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)t), "r"((uint32_t)(t >> 32)), "r"(eps));
        return ((uint64_t)r1 << 32) | r0;
#else
        T s = a + t;
        return s < a ? s + GL::EPS : s;
#endif
    }
    static MS_HD T sub(T a, T t) {
#ifdef __CUDA_ARCH__
        uint32_t r0, r1;
        asm("{\n\t.reg .u32 m, s0, s1;\n\t"
            "sub.cc.u32 s0, %2, %4;\n\t"
            "subc.cc.u32 s1, %3, %5;\n\t"
            "subc.u32 m, 0, 0;\n\t"
            "sub.cc.u32 %0, s0, m;\n\t"
            "subc.u32 %1, s1, 0;\n\t}"
            : "=r"(r0), "=r"(r1)
            : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)t), "r"((uint32_t)(t >> 32)));
        return ((uint64_t)r1 << 32) | r0;
#else
        T d = a - t;
        return a < t ? d - GL::EPS : d;
#endif
    }
    // x * 2^S, 0 < S < 96, x any representative, result canonical.  Every root of unity of order <= 64 is a
    // power of two in this field (2^96 = -1), so the butterflies inside a radix-16 block multiply by shifts:
    // with x << r = W2 2^64 + W1 2^32 + W0 the product is  W0 + W1 2^32 + W2 (2^32-1)           (S = r),
    // W1 2^32 + W2 (2^32-1) - W3 (S = 32 + r)  or  W2 (2^32-1) - W3 - W4 2^32 (S = 64 + r; always canonical),
    // 9 to 16 instructions against 22 for a general product (and a third of its register reads).
    template <int S>
    static MS_HD T mulpow2(T x) {
        static_assert(S > 0 && S < 96, "shift out of range");
#ifdef __CUDA_ARCH__
        constexpr int Q = S / 32, R = S % 32;
        const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), eps = GL_EPS_OPAQUE;
        uint32_t a, b, c;  // x << R as three words (R = 0: a = x0, b = x1, c = 0)
        if constexpr (R == 0) { a = x0; b = x1; c = 0; }
        else { a = x0 << R; b = __funnelshift_l(x0, x1, R); c = x1 >> ((32 - R) & 31); }
        uint32_t r0, r1;
        if constexpr (Q == 0) {
            asm("{\n\t.reg .u32 k, d, s0, s1;\n\t"
                "mad.lo.cc.u32 s0, %4, %5, %2;\n\t"
                "madc.hi.cc.u32 s1, %4, %5, %3;\n\t"
                "addc.u32 k, 0, 0;\n\t"
                "add.cc.u32 d, s0, 0xFFFFFFFF;\n\t"
                "addc.cc.u32 d, s1, 0;\n\t"
                "addc.u32 k, k, 0;\n\t"
                "mad.lo.cc.u32 %0, k, %5, s0;\n\t"
                "madc.hi.u32 %1, k, %5, s1;\n\t}"
                : "=r"(r0), "=r"(r1) : "r"(a), "r"(b), "r"(c), "r"(eps));
        } else if constexpr (Q == 1) {
            asm("{\n\t.reg .u32 k, m, d, s0, s1;\n\t"
                "sub.cc.u32 s0, 0, %4;\n\t"
                "subc.cc.u32 s1, %2, 0;\n\t"
                "subc.u32 m, 0, 0;\n\t"
                "sub.cc.u32 s0, s0, m;\n\t"
                "subc.u32 s1, s1, 0;\n\t"
                "mad.lo.cc.u32 s0, %3, %5, s0;\n\t"
                "madc.hi.cc.u32 s1, %3, %5, s1;\n\t"
                "addc.u32 k, 0, 0;\n\t"
                "add.cc.u32 d, s0, 0xFFFFFFFF;\n\t"
                "addc.cc.u32 d, s1, 0;\n\t"
                "addc.u32 k, k, 0;\n\t"
                "mad.lo.cc.u32 %0, k, %5, s0;\n\t"
                "madc.hi.u32 %1, k, %5, s1;\n\t}"
                : "=r"(r0), "=r"(r1) : "r"(a), "r"(b), "r"(c), "r"(eps));
        } else {
            asm("{\n\t.reg .u32 m, s0, s1;\n\t.reg .u64 pp;\n\t"
                "mul.wide.u32 pp, %2, %5;\n\t"
                "mov.b64 {s0, s1}, pp;\n\t"
                "sub.cc.u32 s0, s0, %3;\n\t"
                "subc.cc.u32 s1, s1, %4;\n\t"
                "subc.u32 m, 0, 0;\n\t"
                "sub.cc.u32 %0, s0, m;\n\t"
                "subc.u32 %1, s1, 0;\n\t}"
                : "=r"(r0), "=r"(r1) : "r"(a), "r"(b), "r"(c), "r"(eps));
        }
        return ((uint64_t)r1 << 32) | r0;
#else
        constexpr uint64_t pw = S < 64 ? (1ULL << (S & 63)) : (0xFFFFFFFFULL << ((S - 64) & 31));  // 2^S mod p
        return GL::mul(x % GL::P, pw);
#endif
    }
    static MS_HD T canon(T a) {
#ifdef __CUDA_ARCH__
        // a >= p  <=>  a + (2^32 - 1) carries out of 64 bits; then a - p = a + (2^32 - 1) mod 2^64
        uint32_t r0, r1;
        asm("{\n\t.reg .u32 k, d;\n\t"
            "add.cc.u32 d, %2, 0xFFFFFFFF;\n\t"
            "addc.cc.u32 d, %3, 0;\n\t"
            "addc.u32 k, 0, 0;\n\t"
            "mad.lo.cc.u32 %0, k, %4, %2;\n\t"
            "madc.hi.u32 %1, k, %4, %3;\n\t}"
            : "=r"(r0), "=r"(r1)
            : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"(GL_EPS_OPAQUE));
        return ((uint64_t)r1 << 32) | r0;
#else
        return a >= GL::P ? a - GL::P : a;
#endif
    }
};

template <>
struct Fast<BB> {
    using T = uint32_t;
    static constexpr uint32_t NPINV = 2013265919u;  // -p^-1 mod 2^32
    static MS_HD T to_tw(T w) { return (T)((((uint64_t)w) << 32) % BB::P); }  // Montgomery form, R = 2^32
    // REDC(x * w R): x any 32-bit value, result canonical.  (t + m p) >> 32 with m = -t p^-1 is one
    // IMAD.WIDE with a 64-bit addend; p comes from constant memory so that ptxas keeps that form instead of
    // an IMAD.HI (5.4 vs 2.6 clocks, see GL_EPS_OPAQUE).
    static MS_HD T mul(T x, T w) {
        const uint64_t t = (uint64_t)x * w;
        const uint32_t m = (uint32_t)t * NPINV;
#ifdef __CUDA_ARCH__
        const uint64_t s = (uint64_t)m * BB_P_OPAQUE + t;  // < 2^33 p: no overflow
#else
        const uint64_t s = (uint64_t)m * BB::P + t;
#endif
        const uint32_t u = (uint32_t)(s >> 32), v = u - BB::P;  // u in [0, 2p)
        return u < v ? u : v;
    }
    static MS_HD T add(T a, T t) {
        uint32_t s = a + t, v = s - BB::P;
        return s < v ? s : v;
    }
    static MS_HD T sub(T a, T t) {
        uint32_t d = a - t, v = d + BB::P;
        return d < v ? d : v;
    }
    static MS_HD T canon(T a) { return a; }
};

// ------------------------------------------------------------------------------------------------
// Radix-2^G block with power-of-two twiddles (Goldilocks).  The in-register part of a round is a plain
// DIT DFT of size 2^G <= 16 once every element r has been multiplied by beta^brev(r) (beta = the round's
// coset / position factor, see fixed_round): its own twiddles are the 2^(st+1)-th roots of unity, all
// powers of two: w_2 = 2^96, w_4 = 2^48, w_8 = 2^120, w_16 = 2^156 for ark's generator
// (inverse: 2^96, 2^144, 2^72, 2^36).  Exponents >= 96 use 2^96 = -1: the two butterfly outputs swap.
// ------------------------------------------------------------------------------------------------
template <class F>
struct ShiftTw {
    static constexpr bool ON = false;
};
template <>
struct ShiftTw<GL> {
    static constexpr bool ON = true;
    // log2 of w_(2^l) as a power of two, l = 0..4
    static constexpr int expo(int l, bool inv) {
        return l == 0 ? 0 : l == 1 ? 96 : l == 2 ? (inv ? 144 : 48) : l == 3 ? (inv ? 72 : 120) : (inv ? 36 : 156);
    }
};
// x arrives canonical in stage 0 (loaded from HBM or just multiplied), lazy afterwards
template <int E, bool X_CANON>
MS_HD void gl_shift_butterfly(uint64_t& y, uint64_t& x) {
    using A = Fast<GL>;
    constexpr int e = E % 96;
    uint64_t t;
    if constexpr (e == 0) t = X_CANON ? x : A::canon(x);
    else t = A::template mulpow2<e>(x);
    const uint64_t u = y;
    if constexpr (E < 96) { y = A::add(u, t); x = A::sub(u, t); }
    else { y = A::sub(u, t); x = A::add(u, t); }
}
template <int G, bool INV, int ST, int R>
MS_HD void gl_shift_dft_steps(uint64_t* v) {
    if constexpr (ST < G) {
        if constexpr (R < (1 << G)) {
            if constexpr (!(R & (1 << ST))) {
                constexpr int j = R & ((1 << ST) - 1);
                constexpr int E = (ShiftTw<GL>::expo(ST + 1, INV) * j) % 192;
                gl_shift_butterfly<E, ST == 0>(v[R], v[R | (1 << ST)]);
            }
            gl_shift_dft_steps<G, INV, ST, R + 1>(v);
        } else {
            gl_shift_dft_steps<G, INV, ST + 1, 0>(v);
        }
    }
}
// v[r], r odd, canonical on entry (stage 0 operands); every output lazy
template <int G, bool INV>
MS_HD void gl_shift_dft(uint64_t* v) {
    static_assert(G >= 0 && G <= 4, "w_32 and w_64 are powers of two as well, but blocks stop at 16 elements");
    gl_shift_dft_steps<G, INV, 0, 0>(v);
}

// ------------------------------------------------------------------------------------------------
// twiddle tables (built on the device with the canonical field ops, stored in twiddle form)
// ------------------------------------------------------------------------------------------------
// plain DIT twiddles W[2^s + q] = g_(2^(s+1))^q, s < NTT_MAXLOG, g = (inverse) root of unity
template <class F>
__global__ void k_build_wtab(typename F::T* w, const typename F::T* roots /*[NTT_MAXLOG]*/) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (1 << NTT_MAXLOG)) return;
    if (idx == 0) { w[0] = 0; return; }
    int u = 31 - __clz(idx);
    int q = idx - (1 << u);
    w[idx] = Fast<F>::to_tw(fpow<F>(roots[u], (uint64_t)q));
}

// t1[j*n2 + 2^s + q] = sbase[j*a + s] * W[2^s + q]     (sbase canonical, W in twiddle form)
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

// ft[(((tile << a) + k2) << beta) | rr << logB | j] = scale * s_j^m1 * wN^(m1*k2), m1 = tile*R + rr
template <class F>
__global__ void k_build_ft(typename F::T* ft, const typename F::T* shifts /*[B]*/, typename F::T wN, typename F::T scale,
                           int a, int b, int logB, int logR, int chunk_log) {
    using T = typename F::T;
    const uint64_t lanes = (1ULL << b) << logB;  // (m1, j)
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lane = gid % lanes, chunk = gid / lanes;
    uint64_t nchunks = (1ULL << a) >> chunk_log;
    if (chunk >= nchunks) return;
    const uint64_t m1 = lane >> logB;
    const int j = (int)(lane & ((1u << logB) - 1));
    const uint64_t tile = m1 >> logR, rr = m1 & ((1u << logR) - 1);
    const int beta = logR + logB;
    uint64_t k2 = chunk << chunk_log;
    T rho = fpow<F>(wN, m1);
    T v = F::mul(F::mul(scale, fpow<F>(shifts[j], m1)), fpow<F>(rho, k2));
    for (uint64_t i = 0; i < (1ULL << chunk_log); i++, k2++) {
        ft[((((tile << a) + k2) << beta) | (rr << logB)) + j] = Fast<F>::to_tw(v);
        v = F::mul(v, rho);
    }
}

// Block twiddles of the fixed-shape Goldilocks kernel (ShiftTw): the round that starts at stage U and
// spans G stages multiplies element r of the block at in-tile offset lo by beta^brev_G(r),
//   beta = sigma^(2^(A-U-G)) * w_(2^(U+G))^lo        (sigma = the coset factor of the 2^A-point transform),
// stored at tab[2^(U+G) + (n << U) + lo], n = brev_G(r) in [1, 2^G): the ranges of different rounds are
// disjoint because U + G grows from round to round, so one coset needs 2^(A+1) entries.
template <class F>
struct TwRoots {
    typename F::T w[NTT_MAXLOG + 1];  // w[l] = the (inverse) primitive 2^l-th root of unity
};
MS_HD int tile_rounds(int a);
MS_HD int tile_round_size(int a, int r);
template <class F>
MS_HD typename F::T tw16_entry(uint32_t e, int A, typename F::T sigma, const TwRoots<F>& roots) {
    using T = typename F::T;
    if (e < 2) return 0;
    int p = 0;
    while ((2u << p) <= e) p++;  // p = floor(log2 e)
    int U = 0, G = 0;
    bool ok = false;
    for (int r = 0; r < tile_rounds(A); r++) {
        G = tile_round_size(A, r);
        if (U + G == p) { ok = true; break; }
        U += G;
    }
    if (!ok) return 0;
    const uint32_t n = (e - (1u << p)) >> U, lo = e & ((1u << U) - 1);
    T beta = fpow<F>(roots.w[p], (uint64_t)lo);
    for (int i = 0; i < A - p; i++) sigma = F::mul(sigma, sigma);
    beta = F::mul(beta, sigma);
    return Fast<F>::to_tw(fpow<F>(beta, (uint64_t)n));
}
// sigma_j = shifts[j]^(2^log_n1) (pass 1 transforms x[m1 + n1 m2] s^(n1 m2) along m2); shifts == nullptr: 1
template <class F>
__global__ void k_build_tw16(typename F::T* tab, const typename F::T* shifts, TwRoots<F> roots, int A, int B, int log_n1) {
    using T = typename F::T;
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ((uint32_t)B << (A + 1))) return;
    const uint32_t j = idx >> (A + 1), e = idx & ((2u << A) - 1);
    T sigma = 1;
    if (shifts) {
        sigma = shifts[j];
        for (int i = 0; i < log_n1; i++) sigma = F::mul(sigma, sigma);
    }
    tab[idx] = tw16_entry<F>(e, A, sigma, roots);
}

// ------------------------------------------------------------------------------------------------
// the tile kernel
// ------------------------------------------------------------------------------------------------
template <class F>
struct NttTile {
    using T = typename F::T;
    const T* src;
    uint64_t src_stride;  // elements between columns
    T* dst;
    uint64_t dst_stride;
    const T* tw;          // twiddle of (stage s, index q, batch b): tw[(b & jmask) * jstride + 2^s + q]
    const T* ft;          // per-element factor indexed like dst (pass 1 of two), or nullptr
    T scale;              // twiddle form; applied on store when ft == nullptr and has_scale
    int has_scale;
    int a;                // log2(points of the in-tile transform)
    int beta;             // log2(batch entries per point)
    int logB;             // the low logB bits of a batch index are the coset j
    uint32_t jmask, jstride;
    int mode;             // 0: first (or only) pass, 1: second pass
    int plain;            // tw is the plain table (no coset factors): twiddles with q = 0 are exactly 1
    int inv;              // inverse transform (selects the power-of-two twiddles of the Goldilocks blocks)
    int bq;               // mode 0: log2(n1), the input stride of one transform step
    int cs;               // mode 0: coset-split bits: tile = (m1 << cs) | sub, coset j = (sub << beta) | b (then R = 1)
    int a1, beta1, logR1; // mode 1: geometry of the first pass (a1 = log2 n2)
    uint32_t tiles;       // tiles per column
    uint32_t cols;
};

// element index of (position P, batch b) of tile `tile` in the source / destination column
template <class F>
MS_HD uint64_t tile_src_index(const NttTile<F>& g, uint32_t tile, uint32_t P, uint32_t b) {
    if (g.mode == 0) {
        if (g.cs) return (uint64_t)(tile >> g.cs) + ((uint64_t)brev_bits(P, g.a) << g.bq);
        const int logR = g.beta - g.logB;
        return ((uint64_t)tile << logR) + (b >> g.logB) + ((uint64_t)brev_bits(P, g.a) << g.bq);
    }
    // second pass: the tile owns 2^beta consecutive lanes (k2, j) of the intermediate
    const uint64_t m1 = brev_bits(P, g.a);
    const uint64_t lane = ((uint64_t)tile << g.beta) + b;
    const uint64_t k2 = lane >> g.logB;
    const uint32_t j = (uint32_t)lane & ((1u << g.logB) - 1);
    return (((((m1 >> g.logR1) << g.a1) + k2) << g.beta1) | ((m1 & ((1u << g.logR1) - 1)) << g.logB)) + j;
}
template <class F>
MS_HD uint64_t tile_dst_index(const NttTile<F>& g, uint32_t tile, uint32_t P, uint32_t b) {
    if (g.mode == 0) {
        if (g.cs) {
            const uint32_t j = ((tile & ((1u << g.cs) - 1)) << g.beta) | b;
            return (((((uint64_t)(tile >> g.cs)) << g.a) + P) << g.logB) | j;
        }
        return ((((uint64_t)tile << g.a) + P) << g.beta) | b;
    }
    return ((uint64_t)P << (g.a1 + g.logB)) + ((uint64_t)tile << g.beta) + b;
}
// coset of batch entry b (selects the twiddle table of a first-pass tile)
template <class F>
MS_HD uint32_t tile_coset(const NttTile<F>& g, uint32_t tile, uint32_t b) {
    return g.cs ? (((tile & ((1u << g.cs) - 1)) << g.beta) | b) : b;
}

// shared-memory swizzle: bijection on every aligned group of 2^(2*SW) elements; makes the strided
// slot accesses of the u = 0 rounds and the contiguous accesses of the later rounds conflict free
template <class F>
MS_HD uint32_t tile_swz(uint32_t idx) {
    constexpr int SW = sizeof(typename F::T) == 8 ? 4 : 5;  // elements per 128-byte bank row
    return idx ^ ((idx >> SW) & ((1u << SW) - 1));
}

// rounds of a 2^a-point tile: ceil(a/4) rounds, sizes as even as possible, larger ones first
MS_HD int tile_rounds(int a) { return (a + NTT_MAXG - 1) / NTT_MAXG; }
MS_HD int tile_round_size(int a, int r) {
    int nr = tile_rounds(a);
    return a / nr + (r < a % nr ? 1 : 0);
}

// One slot of one round: stages [u, u+G) on the 2^G elements P0 + (r << u) of batch b.
template <class F, int G>
MS_HD void tile_slot(const NttTile<F>& g, typename F::T* S, const typename F::T* __restrict__ src, typename F::T* __restrict__ dst,
                     uint32_t tile, uint32_t slot, int u, bool first, bool last) {
    using T = typename F::T;
    using A = Fast<F>;
    constexpr int E = 1 << G;
    const uint32_t b = slot & ((1u << g.beta) - 1), t = slot >> g.beta;
    const uint32_t lo = t & ((1u << u) - 1), hi = t >> u;
    const uint32_t P0 = lo + (hi << (u + G));
    const T* __restrict__ twj = g.tw + (size_t)(tile_coset<F>(g, tile, b) & g.jmask) * g.jstride + lo;
    T w[E];  // w[2^s' + r'] = twiddle of local stage s', local index r'
#pragma unroll
    for (int s = 0; s < G; s++)
#pragma unroll
        for (int r = 0; r < (1 << s); r++) {
#ifdef __CUDA_ARCH__
            w[(1 << s) + r] = __ldg(&twj[(1u << (u + s)) + ((uint32_t)r << u)]);
#else
            w[(1 << s) + r] = twj[(1u << (u + s)) + ((uint32_t)r << u)];
#endif
        }
    T v[E];
    if (first) {
#pragma unroll
        for (int r = 0; r < E; r++) v[r] = src[tile_src_index<F>(g, tile, P0 + ((uint32_t)r << u), b)];
    } else {
#pragma unroll
        for (int r = 0; r < E; r++) v[r] = S[tile_swz<F>(((P0 + ((uint32_t)r << u)) << g.beta) | b)];
    }
#pragma unroll
    for (int s = 0; s < G; s++) {
        const int h = 1 << s;
#pragma unroll
        for (int r = 0; r < E; r++) {
            if (r & h) continue;
            T x = A::mul(v[r | h], w[h + (r & (h - 1))]);
            T y = v[r];
            v[r] = A::add(y, x);
            v[r | h] = A::sub(y, x);
        }
    }
    if (last) {
#pragma unroll
        for (int r = 0; r < E; r++) {
            const uint64_t o = tile_dst_index<F>(g, tile, P0 + ((uint32_t)r << u), b);
            T x = v[r];
            if (g.ft) {
#ifdef __CUDA_ARCH__
                x = A::mul(x, __ldg(&g.ft[o]));
#else
                x = A::mul(x, g.ft[o]);
#endif
            } else if (g.has_scale) x = A::mul(x, g.scale);
            else x = A::canon(x);
            dst[o] = x;
        }
    } else {
#pragma unroll
        for (int r = 0; r < E; r++) S[tile_swz<F>(((P0 + ((uint32_t)r << u)) << g.beta) | b)] = v[r];
    }
}

// all slots of round `round` that thread `tid` of `nthreads` owns
template <class F>
MS_HD void tile_round(const NttTile<F>& g, typename F::T* S, const typename F::T* src, typename F::T* dst, uint32_t tile,
                      int round, int u, int G, uint32_t tid, uint32_t nthreads) {
    const bool first = round == 0, last = round == tile_rounds(g.a) - 1 || g.a == 0;
    const uint32_t nslots = 1u << (g.a - G + g.beta);
    for (uint32_t s = tid; s < nslots; s += nthreads) {
        switch (G) {
            case 4: tile_slot<F, 4>(g, S, src, dst, tile, s, u, first, last); break;
            case 3: tile_slot<F, 3>(g, S, src, dst, tile, s, u, first, last); break;
            case 2: tile_slot<F, 2>(g, S, src, dst, tile, s, u, first, last); break;
            case 1: tile_slot<F, 1>(g, S, src, dst, tile, s, u, first, last); break;
            default: tile_slot<F, 0>(g, S, src, dst, tile, s, u, true, true); break;
        }
    }
}

// block -> (column, tile): 8 neighbouring tiles of every column run in the same wave, so the 32-byte
// sectors pass 1 touches (8 of 32 bytes per tile) and the ft rows are shared through L2
template <class F>
MS_HD void tile_of_block(const NttTile<F>& g, uint32_t bid, uint32_t* col, uint32_t* tile) {
    const uint32_t grp = g.tiles < 8 ? g.tiles : 8;
    const uint32_t tl = bid % grp, rest = bid / grp;
    *col = rest % g.cols;
    *tile = (rest / g.cols) * grp + tl;
}

template <class F>
__global__ void __launch_bounds__(NTT_THREADS, 2)
k_ntt_tile(const NttTile<F> g) {
    using T = typename F::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);
    uint32_t col, tile;
    tile_of_block<F>(g, blockIdx.x, &col, &tile);
    const T* src = g.src + (uint64_t)col * g.src_stride;
    T* dst = g.dst + (uint64_t)col * g.dst_stride;
    const int nr = tile_rounds(g.a);
    if (nr == 0) {
        tile_round<F>(g, S, src, dst, tile, 0, 0, 0, threadIdx.x, blockDim.x);
        return;
    }
    int u = 0;
    for (int r = 0; r < nr; r++) {
        const int G = tile_round_size(g.a, r);
        if (r) __syncthreads();
        tile_round<F>(g, S, src, dst, tile, r, u, G, threadIdx.x, blockDim.x);
        u += G;
    }
}

// ------------------------------------------------------------------------------------------------
// Specialised tile kernel: tile shape (A points, BETA batch bits) fixed at compile time, so the
// round structure is unrolled and every per-element index is base + constant (global: base +
// constant * stride).  Same algorithm and the same result as k_ntt_tile; the generic kernel stays
// as the path for unusual shapes.
// ------------------------------------------------------------------------------------------------
template <int A>
struct Rounds {
    static constexpr int NR = (A + NTT_MAXG - 1) / NTT_MAXG;
    static constexpr int size(int r) { return A / NR + (r < A % NR ? 1 : 0); }
    static constexpr int start(int r) {
        int u = 0;
        for (int i = 0; i < r; i++) u += size(i);
        return u;
    }
};
constexpr uint32_t cbrev(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

template <class F, int A, int BETA>
struct FixedSwz {
    static constexpr int G0 = Rounds<A>::size(0);
    static constexpr int BANKBITS = sizeof(typename F::T) == 8 ? 4 : 5;
    static constexpr int K0 = BANKBITS > BETA ? BANKBITS - BETA : 0;
    static constexpr int K1 = K0 < G0 ? K0 : G0;
    static constexpr int K = K1 < A - G0 ? K1 : (A - G0 > 0 ? A - G0 : 0);  // swizzle bits
    static constexpr uint32_t MASK = (1u << K) - 1;
};

template <class F, int A, int BETA, int R, int THREADS, bool INV>
MS_HD void fixed_round(const NttTile<F>& g, typename F::T* S, const typename F::T* __restrict__ src,
                       typename F::T* __restrict__ dst, uint32_t tile, uint32_t tid) {
    using T = typename F::T;
    using Ar = Fast<F>;
    using SW = FixedSwz<F, A, BETA>;
    constexpr bool SHIFT = ShiftTw<F>::ON;  // g.tw is then the block-twiddle table (tw16_entry)
    constexpr int G = Rounds<A>::size(R), U = Rounds<A>::start(R), E = 1 << G;
    constexpr bool FIRST = R == 0, LAST = R == Rounds<A>::NR - 1;
    constexpr uint32_t NSLOTS = 1u << (A - G + BETA);
    const int logR = BETA - g.logB;  // pass 1: log2 R1, pass 2: log2 R2
#pragma unroll 1
    for (uint32_t s = tid; s < NSLOTS; s += THREADS) {
        const uint32_t b = s & ((1u << BETA) - 1), t = s >> BETA;
        const uint32_t lo = t & ((1u << U) - 1), hi = t >> U;
        const uint32_t P0 = lo | (hi << (U + G));
        const T* __restrict__ twj = g.tw + (size_t)(tile_coset<F>(g, tile, b) & g.jmask) * g.jstride + lo;
        T w[E];
        if constexpr (SHIFT) {
            // w[n] = beta^n, the factor of the element with sub-index n = brev_G(r); none in a plain first round
            if (!(FIRST && g.plain)) {
#pragma unroll
                for (int n = 1; n < E; n++) w[n] = MS_LDG(&twj[(1u << (U + G)) + ((uint32_t)n << U)]);
            }
        } else {
#pragma unroll
            for (int st = 0; st < G; st++)
#pragma unroll
                for (int r = 0; r < (1 << st); r++) w[(1 << st) + r] = MS_LDG(&twj[(1u << (U + st)) + ((uint32_t)r << U)]);
        }
        T v[E];
        if (FIRST) {
            // element r sits at position P0 + r whose bit reversal is brev(P0) | brev_G(r) << (A - G)
            const uint32_t m0 = brev_bits(P0, A);
            const T* p;
            uint32_t stride;  // elements between brev_G(r) = 1 and 2
            if (g.mode == 0) {
                p = src + (g.cs ? (uint64_t)(tile >> g.cs) : ((uint64_t)tile << logR) + (b >> g.logB)) + ((uint64_t)m0 << g.bq);
                stride = 1u << (A - G + g.bq);
            } else {
                const uint64_t lane = ((uint64_t)tile << BETA) + b;
                const uint64_t k2 = lane >> g.logB;
                const uint32_t j = (uint32_t)lane & ((1u << g.logB) - 1);
                p = src + ((((((uint64_t)m0 >> g.logR1) << g.a1) + k2) << g.beta1) | ((m0 & ((1u << g.logR1) - 1)) << g.logB)) + j;
                stride = 1u << (A - G - g.logR1 + g.a1 + g.beta1);
            }
#pragma unroll
            for (int r = 0; r < E; r++) v[r] = p[(uint64_t)cbrev(r, G) * stride];
        } else {
            const uint32_t Pb = P0 ^ ((P0 >> SW::G0) & SW::MASK);
#pragma unroll
            for (int r = 0; r < E; r++) {
                const uint32_t cr = (((uint32_t)r << U) >> SW::G0) & SW::MASK;
                v[r] = S[((((Pb ^ cr) << BETA) | b)) + ((uint32_t)r << (U + BETA))];
            }
        }
        if constexpr (SHIFT) {
            if (!(FIRST && g.plain)) {
#pragma unroll
                for (int r = 1; r < E; r++) v[r] = Ar::mul(v[r], w[cbrev(r, G)]);
            }
            gl_shift_dft<G, INV>(v);
        } else if (FIRST && g.plain) {
            // U == 0 and plain twiddles: w[2^st + 0] = 1, so those butterflies need no product, only
            // a canonical operand (the loaded values are canonical already: stage 0 needs nothing)
#pragma unroll
            for (int st = 0; st < G; st++) {
                const int h = 1 << st;
#pragma unroll
                for (int r = 0; r < E; r++) {
                    if (r & h) continue;
                    T x = (r & (h - 1)) ? Ar::mul(v[r | h], w[h + (r & (h - 1))]) : (st ? Ar::canon(v[r | h]) : v[r | h]);
                    T y = v[r];
                    v[r] = Ar::add(y, x);
                    v[r | h] = Ar::sub(y, x);
                }
            }
        } else {
#pragma unroll
            for (int st = 0; st < G; st++) {
                const int h = 1 << st;
#pragma unroll
                for (int r = 0; r < E; r++) {
                    if (r & h) continue;
                    T x = Ar::mul(v[r | h], w[h + (r & (h - 1))]);
                    T y = v[r];
                    v[r] = Ar::add(y, x);
                    v[r | h] = Ar::sub(y, x);
                }
            }
        }
        if (LAST) {
            // positions P0 + (r << U) with U + G == A: consecutive lanes hold consecutive (b, lo)
            uint64_t o;
            uint32_t stride;
            if (g.mode == 0) {
                o = tile_dst_index<F>(g, tile, P0, b);
                stride = 1u << (U + (g.cs ? g.logB : BETA));
            } else {
                o = ((uint64_t)P0 << (g.a1 + g.logB)) + ((uint64_t)tile << BETA) + b;
                stride = 1u << (U + g.a1 + g.logB);
            }
            T* q = dst + o;
            if (g.ft) {
                const T* __restrict__ f = g.ft + o;
#pragma unroll
                for (int r = 0; r < E; r++) q[(uint64_t)r * stride] = Ar::mul(v[r], MS_LDG(&f[(uint64_t)r * stride]));
            } else if (g.has_scale) {
#pragma unroll
                for (int r = 0; r < E; r++) q[(uint64_t)r * stride] = Ar::mul(v[r], g.scale);
            } else {
#pragma unroll
                for (int r = 0; r < E; r++) q[(uint64_t)r * stride] = Ar::canon(v[r]);
            }
        } else if (FIRST) {
            // U == 0: the swizzle term depends on hi only and permutes the slot's E positions
            const uint32_t m = (P0 >> SW::G0) & SW::MASK;
#pragma unroll
            for (int r = 0; r < E; r++) S[((P0 | ((uint32_t)r ^ m)) << BETA) | b] = v[r];
        } else {
            const uint32_t Pb = P0 ^ ((P0 >> SW::G0) & SW::MASK);
#pragma unroll
            for (int r = 0; r < E; r++) {
                const uint32_t cr = (((uint32_t)r << U) >> SW::G0) & SW::MASK;
                S[(((Pb ^ cr) << BETA) | b) + ((uint32_t)r << (U + BETA))] = v[r];
            }
        }
    }
}

template <class F, int A, int BETA, int R, int THREADS, bool INV>
__device__ __forceinline__ void fixed_rounds_from(const NttTile<F>& g, typename F::T* S, const typename F::T* src,
                                                  typename F::T* dst, uint32_t tile) {
    if constexpr (R < Rounds<A>::NR) {
        if (R) __syncthreads();
        fixed_round<F, A, BETA, R, THREADS, INV>(g, S, src, dst, tile, threadIdx.x);
        fixed_rounds_from<F, A, BETA, R + 1, THREADS, INV>(g, S, src, dst, tile);
    }
}

template <class F, int A, int BETA, int THREADS, int MINB, bool INV>
__global__ void __launch_bounds__(THREADS, MINB)
k_ntt_fixed(const NttTile<F> g) {
    using T = typename F::T;
    static_assert(Rounds<A>::NR >= 2, "fixed tiles have at least two rounds");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);
    uint32_t col, tile;
    tile_of_block<F>(g, blockIdx.x, &col, &tile);
    fixed_rounds_from<F, A, BETA, 0, THREADS, INV>(g, S, g.src + (uint64_t)col * g.src_stride, g.dst + (uint64_t)col * g.dst_stride, tile);
}

template <class F>
__global__ void k_transpose(const typename F::T* __restrict__ in, typename F::T* __restrict__ out, uint64_t rows,
                            uint64_t width, uint64_t cm_stride, int to_colmajor) {
    // row-major [rows][width] <-> column-major [width][cm_stride >= rows]; 32x32 tiles through shared memory
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
            if (r < rows && c < width) out[c * cm_stride + r] = tile[threadIdx.x][i];
        }
    } else {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            uint64_t c = c0 + i, r = r0 + threadIdx.x;
            if (r < rows && c < width) tile[threadIdx.x][i] = in[c * cm_stride + r];
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            uint64_t r = r0 + i, c = c0 + threadIdx.x;
            if (r < rows && c < width) out[r * width + c] = tile[i][threadIdx.x];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#ifndef MS_NTT_NO_HOST
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
#endif

struct NttPlan {
    int a, b;          // N = 2^(a+b): pass 1 transforms 2^a points (stride 2^b), pass 2 2^b points
    int beta1, beta2;  // batch bits of a pass-1 / pass-2 tile
    int logR1;         // pass 1: consecutive m1 per tile (beta1 = logR1 + logB) ...
    int cs1;           // ... or, when 2^a points x B cosets exceed a tile, cosets split over 2^cs1 tiles
};
// Tiles hold 2^NTT_LOG_TILE_PREF elements.  Pass 1 batches extra m1 when the cosets do not fill a tile and
// splits the cosets over several tiles when they overflow it; pass 2 takes whatever number of consecutive
// (k2, coset) lanes fills a tile (fewer than B lanes for the largest transforms).
inline bool ntt_plan(int logN, int logB, NttPlan* p, int T = NTT_LOG_TILE_PREF) {
    if (logN + logB <= T) {
        *p = {logN, 0, logB, 0, 0, 0};
        return true;
    }
    int a = (logN + 1) / 2, b = logN - a;
    if (a > T || b > T) return false;
    NttPlan pl{};
    pl.a = a;
    pl.b = b;
    if (a + logB <= T) {
        int r1 = T - a - logB;
        if (r1 > b) r1 = b;
        pl.logR1 = r1;
        pl.beta1 = r1 + logB;
        pl.cs1 = 0;
    } else {
        pl.cs1 = a + logB - T;
        if (pl.cs1 > logB) return false;
        pl.logR1 = 0;
        pl.beta1 = logB - pl.cs1;
    }
    pl.beta2 = T - b;
    if (pl.beta2 > a + logB) pl.beta2 = a + logB;
    *p = pl;
    return true;
}

// The plan lde_batch uses: the context's tile size, except that a 4-byte field whose cosets would have to be split over
// pass-1 tiles takes 2^14-element tiles instead (64 KB, the byte size of a Goldilocks tile): no split, no second read of the
// input, and pass-2 runs of 16-32 contiguous bytes per point instead of 8 (BabyBear from 2^23 rows at blowup 4; the
// transform stays in place: such plans have logR1 = 0).  *tile_log returns the tile size the plan was made for.
template <class F>
inline bool ntt_plan_for(int logN, int logB, int pref_tile, NttPlan* p, int* tile_log) {
    *tile_log = pref_tile;
    if (!ntt_plan(logN, logB, p, pref_tile)) return false;
    if (sizeof(typename F::T) == 4 && pref_tile == NTT_LOG_TILE_PREF && p->cs1 > 0) {
        NttPlan q;
        if (ntt_plan(logN, logB, &q, NTT_LOG_TILE_PREF + 1)) {
            *p = q;
            *tile_log = NTT_LOG_TILE_PREF + 1;
        }
    }
    return true;
}

// tile shapes with a compile-time specialisation
// (A + BETA = 13: 8192-element tiles, 256 threads, 3 CTAs / SM; A + BETA = 12: 4096-element tiles, 128 threads, 6 CTAs / SM --
// the same warps per SM in CTAs half the size: barriers couple fewer warps and load / compute phases of more CTAs interleave)
#define MS_NTT_FIXED_SHAPES(X) X(8, 5) X(9, 4) X(10, 3) X(11, 2) X(12, 1) X(13, 0) X(8, 4) X(9, 3) X(10, 2) X(11, 1) X(12, 0)
// A + BETA = 14, 4-byte fields only (ntt_plan_for): the 64 KB tiles of large BabyBear transforms
#define MS_NTT_FIXED_SHAPES_W4(X) X(11, 3) X(12, 2) X(13, 1)
// does launch_tile run this tile through k_ntt_fixed (and, for Goldilocks, with block twiddles)?
template <class F>
inline bool tile_is_fixed(const NttTile<F>& g) {
    if (!(g.mode == 0 || g.logR1 <= g.a - tile_round_size(g.a, 0))) return false;
#define X(A_, B_) if (g.a == A_ && g.beta == B_) return true;
    MS_NTT_FIXED_SHAPES(X)
    if (sizeof(typename F::T) == 4) { MS_NTT_FIXED_SHAPES_W4(X) }
#undef X
    return false;
}

#ifndef MS_NTT_NO_HOST
template <class F, int A, int BETA, bool INV>
int launch_fixed_dir(Ctx* c, const NttTile<F>& g, const char* name) {
    using T = typename F::T;
    constexpr size_t smem = sizeof(T) << (A + BETA);
    constexpr bool big = smem > 100 * 1024;  // one CTA per SM: give it 512 threads
    constexpr bool half = A + BETA == NTT_LOG_TILE_PREF - 1;
    constexpr int threads = big ? 512 : (half ? NTT_THREADS / 2 : NTT_THREADS);
    auto kern = k_ntt_fixed<F, A, BETA, threads, big ? 1 : (half ? 2 * MS_NTT_MINB : MS_NTT_MINB), INV>;
    static uint64_t attr_done = 0;  // per instantiation: devices whose function attribute is set (once, not per launch)
    if (!(attr_done >> (c->device & 63) & 1)) {
        MS_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done |= 1ULL << (c->device & 63);
    }
    prof_begin(c, name);
    kern<<<g.cols * g.tiles, threads, smem, c->stream>>>(g);
    prof_end(c);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}
template <class F, int A, int BETA>
int launch_fixed(Ctx* c, const NttTile<F>& g, const char* name) {
    // only the Goldilocks blocks depend on the direction (power-of-two twiddles are compiled in)
    if constexpr (ShiftTw<F>::ON) {
        if (g.inv) return launch_fixed_dir<F, A, BETA, true>(c, g, name);
    }
    return launch_fixed_dir<F, A, BETA, false>(c, g, name);
}

// block-twiddle table (tw16_entry) of a 2^a-point transform: per coset when `d_shifts` is given (then the
// caller owns `own`), else the plain table of this direction, built once per context
template <class F>
int get_tw16(Ctx* c, int a, bool inverse, const typename F::T* d_shifts, int B, int log_n1, Scratch* own, const typename F::T** out) {
    using T = typename F::T;
    TwRoots<F> roots{};
    for (int l = 0; l <= NTT_MAXLOG; l++) {
        T g = (l <= F::TWO_ADICITY) ? root_of_unity<F>(l) : (T)1;
        roots.w[l] = inverse ? finv<F>(g) : g;
    }
    T* tab;
    if (d_shifts) {
        MS_TRY(own->alloc(((size_t)B << (a + 1)) * sizeof(T)));
        tab = own->template as<T>();
    } else {
        void*& slot = c->tw16_plain[inverse ? 1 : 0][a];
        if (slot) { *out = reinterpret_cast<const T*>(slot); return MS_OK; }
        MS_CUDA(c, cudaMalloc(&slot, ((size_t)2 << a) * sizeof(T)));
        tab = reinterpret_cast<T*>(slot);
        B = 1;
    }
    const unsigned n = (unsigned)B << (a + 1);
    prof_begin(c, "k_build_tw16");
    k_build_tw16<F><<<(n + 127) / 128, 128, 0, c->stream>>>(tab, d_shifts, roots, a, B, log_n1);
    prof_end(c);
    MS_LAUNCH_CHECK(c);
    *out = tab;
    return MS_OK;
}

template <class F>
int launch_tile(Ctx* c, const NttTile<F>& g, const char* name) {
    using T = typename F::T;
    if (tile_is_fixed<F>(g)) {
#define X(A_, B_) if (g.a == A_ && g.beta == B_) return launch_fixed<F, A_, B_>(c, g, name);
        MS_NTT_FIXED_SHAPES(X)
        if constexpr (sizeof(T) == 4) { MS_NTT_FIXED_SHAPES_W4(X) }
#undef X
    }
    const size_t smem = tile_rounds(g.a) > 1 ? (sizeof(T) << (g.a + g.beta)) : 0;
    if (smem > 48 * 1024)
        MS_CUDA(c, cudaFuncSetAttribute(k_ntt_tile<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t slots = 1ULL << (g.a + g.beta - (g.a ? tile_round_size(g.a, 0) : 0));
    unsigned threads = NTT_THREADS;
    while (threads > 32 && threads / 2 >= slots) threads /= 2;
    prof_begin(c, name);
    k_ntt_tile<F><<<g.cols * g.tiles, threads, smem, c->stream>>>(g);
    prof_end(c);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}

// Cached tables of two-pass transforms (Ctx::ntt_tables): look-up by key, and a fresh entry after evicting the least
// recently used ones beyond the byte budget.  Everything runs on the context's one stream, so a table handed back to the
// block cache is only reused behind the launches that still read it.
inline Ctx::NttTables* ntt_tables_find(Ctx* c, int logN, int logB, uint64_t shift, int inverse, int field, int tile_log) {
    for (auto& t : c->ntt_tables)
        if (t.ft && t.logN == logN && t.logB == logB && t.shift == shift && t.inverse == inverse && t.tile == tile_log && t.field == field) {
            t.stamp = ++c->ntt_tables_clock;
            return &t;
        }
    return nullptr;
}
inline Ctx::NttTables* ntt_tables_new(Ctx* c, size_t need_bytes) {
    auto& v = c->ntt_tables;
    auto total = [&] { size_t b = 0; for (auto& t : v) b += t.ft_bytes + t.tw_bytes; return b; };
    while (!v.empty() && (v.size() >= 48 || total() + need_bytes > c->ntt_tables_budget)) {
        size_t lru = 0;
        for (size_t i = 1; i < v.size(); i++) if (v[i].stamp < v[lru].stamp) lru = i;
        block_free(c, v[lru].ft, v[lru].ft_bytes);
        block_free(c, v[lru].tw, v[lru].tw_bytes);
        v.erase(v.begin() + lru);
    }
    v.emplace_back();
    v.back().stamp = ++c->ntt_tables_clock;
    return &v.back();
}

// See the header comment.  d_in and d_out must not alias.
template <class F>
int lde_batch(Ctx* c, const typename F::T* d_in, uint64_t in_stride, uint64_t cols, int logN, int logB,
              typename F::T shift, bool inverse, typename F::T* d_out, uint64_t out_stride) {
    using T = typename F::T;
    if (cols == 0) return MS_OK;
    if (logN + logB > F::TWO_ADICITY) return fail(c, MS_ERR_BAD_SHAPE, "domain 2^%d exceeds the field's two-adicity", logN + logB);
    if (inverse && logB != 0) return fail(c, MS_ERR_UNSUPPORTED, "inverse transform with blowup");
    if (cols >= (1ULL << 20)) return fail(c, MS_ERR_UNSUPPORTED, "too many columns");
    NttPlan pl;
    int tile_log = c->ntt_log_tile;
    if (!ntt_plan_for<F>(logN, logB, c->ntt_log_tile, &pl, &tile_log)) return fail(c, MS_ERR_UNSUPPORTED, "transform 2^%d x blowup 2^%d too large", logN, logB);
    MS_TRY(ensure_wtab<F>(c, inverse ? 1 : 0));
    const T* wtab = reinterpret_cast<const T*>(c->wtab[inverse ? 1 : 0]);
    const int B = 1 << logB;
    const uint64_t N = 1ULL << logN;
    T wN = root_of_unity<F>(logN);
    if (inverse) wN = finv<F>(wN);
    const T scale = inverse ? finv<F>((T)(N % (uint64_t)F::P)) : (T)1;
    const T wL = root_of_unity<F>(logN + logB);
    const bool two = pl.b > 0;
    // host-side per-coset constants
    const int na = pl.a ? pl.a : 1;
    std::vector<T> hbuf((size_t)B * na + B);
    T* sbase = hbuf.data();
    T* shifts = hbuf.data() + (size_t)B * na;
    T sj = shift;
    bool plain = true;  // every coset's stage factors are 1: use the plain table
    for (int j = 0; j < B; j++) {
        shifts[j] = sj;
        for (int u = 0; u < pl.a; u++) {
            sbase[j * na + u] = fpow<F>(sj, N >> (u + 1));
            plain = plain && sbase[j * na + u] == 1;
        }
        sj = F::mul(sj, wL);
    }
    Scratch consts(c), t1(c), tmp(c);
    // The coset block-twiddle table and the inter-pass factor table only depend on (n, blowup, shift, direction, tile size):
    // they are kept for the next call with the same key (the prover extends its columns in several calls per proof, a
    // benchmark repeats one call), built on this stream and only ever used on it.
    Ctx::NttTables tc_none;
    Ctx::NttTables* tcp = two ? ntt_tables_find(c, logN, logB, (uint64_t)shift, inverse ? 1 : 0, F::ID, tile_log) : nullptr;
    const bool tc_hit = tcp != nullptr;
    if (two && !tc_hit) tcp = ntt_tables_new(c, (((size_t)N << logB) + ((size_t)B << (pl.a + 1))) * sizeof(T));
    Ctx::NttTables& tc = tcp ? *tcp : tc_none;
    // geometry of the first pass decides which kernel runs it, and with it the twiddle-table format
    NttTile<F> g1{};
    g1.a = pl.a;
    g1.beta = pl.beta1;
    g1.cs = pl.cs1;
    g1.logB = logB;
    g1.mode = 0;
    g1.inv = inverse ? 1 : 0;
    const bool blocks1 = ShiftTw<F>::ON && tile_is_fixed<F>(g1);  // Goldilocks fixed-shape kernel: block twiddles
    const T* d_tw1 = wtab;
    if (!plain || two) {  // per-coset constants (the shifts are also needed by k_build_ft)
        MS_TRY(consts.alloc(hbuf.size() * sizeof(T)));
        MS_CUDA(c, cudaMemcpyAsync(consts.p, hbuf.data(), hbuf.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    }
    const T* d_shifts = consts.p ? consts.as<T>() + (size_t)B * na : nullptr;
    uint32_t jstride1 = plain ? 0u : (1u << pl.a);
    if (blocks1 && !plain && two) {
        // cached with the factor table below (same key)
        if (!tc_hit) {
            block_free(c, tc.tw, tc.tw_bytes);  // same stream: earlier launches that read it are ahead of its next user
            tc.tw = nullptr;
            MS_TRY(block_alloc(c, ((size_t)B << (pl.a + 1)) * sizeof(T), &tc.tw, &tc.tw_bytes));
            TwRoots<F> roots{};
            for (int l = 0; l <= NTT_MAXLOG; l++) {
                T g = (l <= F::TWO_ADICITY) ? root_of_unity<F>(l) : (T)1;
                roots.w[l] = inverse ? finv<F>(g) : g;
            }
            const unsigned nn = (unsigned)B << (pl.a + 1);
            prof_begin(c, "k_build_tw16");
            k_build_tw16<F><<<(nn + 127) / 128, 128, 0, c->stream>>>(reinterpret_cast<T*>(tc.tw), d_shifts, roots, pl.a, B, pl.b);
            prof_end(c);
            MS_LAUNCH_CHECK(c);
        }
        d_tw1 = reinterpret_cast<const T*>(tc.tw);
        jstride1 = 2u << pl.a;
    } else if (blocks1) {
        MS_TRY(get_tw16<F>(c, pl.a, inverse, plain ? nullptr : d_shifts, B, pl.b, &t1, &d_tw1));
        jstride1 = plain ? 0u : (2u << pl.a);
    } else if (!plain) {
        MS_TRY(t1.alloc(((size_t)B << pl.a) * sizeof(T)));
        int n = B << pl.a;
        prof_begin(c, "k_build_t1");
        k_build_t1<F><<<(n + 255) / 256, 256, 0, c->stream>>>(t1.as<T>(), wtab, consts.as<T>(), pl.a, B);
        prof_end(c);
        MS_LAUNCH_CHECK(c);
        d_tw1 = t1.as<T>();
    }
    const bool inplace = two && pl.logR1 == 0 && tile_rounds(pl.b) >= 2;
    const int lay1 = pl.logR1 + logB;  // low index bits (m1 offset, coset) of the intermediate layout
    if (two) {
        if (!tc_hit) {
            block_free(c, tc.ft, tc.ft_bytes);
            tc.ft = nullptr;
            tc.logN = -1;
            MS_TRY(block_alloc(c, ((size_t)N << logB) * sizeof(T), &tc.ft, &tc.ft_bytes));
            int chunk_log = pl.a < 6 ? pl.a : 6;
            uint64_t threads = (((uint64_t)1 << pl.b) << logB) * ((1ULL << pl.a) >> chunk_log);
            prof_begin(c, "k_build_ft");
            k_build_ft<F><<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<T*>(tc.ft), d_shifts, wN, scale, pl.a, pl.b,
                                                                                   logB, pl.logR1, chunk_log);
            prof_end(c);
            MS_LAUNCH_CHECK(c);
            tc.logN = logN; tc.logB = logB; tc.shift = (uint64_t)shift; tc.inverse = inverse ? 1 : 0; tc.tile = tile_log; tc.field = F::ID;
        }
        if (!inplace) MS_TRY(tmp.alloc(cols * ((size_t)N << logB) * sizeof(T)));
    }
    g1.src = d_in;
    g1.src_stride = in_stride;
    g1.dst = two ? (inplace ? d_out : tmp.as<T>()) : d_out;
    g1.dst_stride = two ? (inplace ? out_stride : ((uint64_t)N << logB)) : out_stride;
    g1.tw = d_tw1;
    g1.ft = two ? reinterpret_cast<const T*>(tc.ft) : nullptr;
    g1.scale = Fast<F>::to_tw(scale);
    g1.has_scale = (!two && scale != 1) ? 1 : 0;
    g1.a = pl.a;
    g1.beta = pl.beta1;
    g1.cs = pl.cs1;
    g1.logB = logB;
    g1.jmask = plain ? 0u : (uint32_t)(B - 1);
    g1.jstride = jstride1;
    g1.plain = plain ? 1 : 0;
    g1.bq = pl.b;
    g1.tiles = (uint32_t)(((1ULL << pl.b) >> pl.logR1) << pl.cs1);
    g1.cols = (uint32_t)cols;
    MS_TRY(launch_tile<F>(c, g1, "k_ntt_tile/pass1"));
    if (two) {
        NttTile<F> g2{};
        g2.src = g1.dst;
        g2.src_stride = g1.dst_stride;
        g2.dst = d_out;
        g2.dst_stride = out_stride;
        g2.ft = nullptr;
        g2.scale = 0;
        g2.has_scale = 0;
        g2.a = pl.b;
        g2.beta = pl.beta2;
        g2.logB = logB;
        g2.jmask = 0;
        g2.jstride = 0;
        g2.mode = 1;
        g2.plain = 1;
        g2.a1 = pl.a;
        g2.beta1 = lay1;
        g2.logR1 = pl.logR1;
        g2.tiles = (uint32_t)(((1ULL << pl.a) << logB) >> pl.beta2);
        g2.cols = (uint32_t)cols;
        g2.inv = inverse ? 1 : 0;
        g2.tw = wtab;
        if (ShiftTw<F>::ON && tile_is_fixed<F>(g2)) MS_TRY(get_tw16<F>(c, pl.b, inverse, nullptr, 1, 0, nullptr, &g2.tw));
        MS_TRY(launch_tile<F>(c, g2, "k_ntt_tile/pass2"));
    }
    return MS_OK;
}

// Device self-test of the butterfly arithmetic against the canonical host ops (edge values first,
// then xorshift-random operands): returns the number of mismatching (a, b) pairs.
constexpr int SELFTEST_SHIFTS[] = {1, 12, 24, 31, 32, 36, 48, 60, 63, 64, 72, 84, 95};
constexpr int SELFTEST_NSHIFT = sizeof(SELFTEST_SHIFTS) / sizeof(int);
constexpr int SELFTEST_OUT = 4 + SELFTEST_NSHIFT;
template <class F, int I>
__device__ void selftest_shifts(typename F::T a, typename F::T* out) {
    if constexpr (ShiftTw<F>::ON && I < SELFTEST_NSHIFT) {
        out[4 + I] = Fast<F>::template mulpow2<SELFTEST_SHIFTS[I]>(a);
        selftest_shifts<F, I + 1>(a, out);
    }
}
template <class F>
__global__ void k_selftest_ops(const typename F::T* a, const typename F::T* b, typename F::T* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    using A = Fast<F>;
    typename F::T m = A::mul(a[i], A::to_tw(b[i]));
    out[SELFTEST_OUT * i] = m;
    out[SELFTEST_OUT * i + 1] = A::add(a[i], m);
    out[SELFTEST_OUT * i + 2] = A::sub(a[i], m);
    out[SELFTEST_OUT * i + 3] = A::canon(a[i]);
    selftest_shifts<F, 0>(a[i], out + SELFTEST_OUT * i);  // Goldilocks: a * 2^S for the shifts above
}
template <class F>
int selftest_ops(Ctx* c, uint64_t n_random, uint64_t* n_bad) {
    using T = typename F::T;
    const uint64_t P = (uint64_t)F::P;
    const bool lazy = sizeof(T) == 8;  // Goldilocks accepts any 64-bit representative; BabyBear operands are canonical
    std::vector<uint64_t> edge = {0, 1, 2, P - 2, P - 1};
    if (lazy) {
        const uint64_t more[] = {P, P + 1, ~0ULL, ~0ULL - 1, 0xFFFFFFFFULL, 0x100000000ULL, 0xFFFFFFFF00000000ULL, 0xFFFFFFFEFFFFFFFFULL, 0x8000000000000000ULL};
        for (uint64_t e : more) edge.push_back(e);
    }
    std::vector<T> a, b;
    for (auto x : edge) for (auto y : edge) { a.push_back((T)x); b.push_back((T)(y % P)); }
    uint64_t s = 88172645463325252ULL;
    for (uint64_t i = 0; i < n_random; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint64_t x = s;
        s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint64_t y = s;
        if (i % 7 == 0) x |= 0xFFFFFFFF00000000ULL;
        if (i % 13 == 0) x &= 0xFFFFFFFFULL;
        a.push_back(lazy ? (T)x : (T)(x % P));
        b.push_back((T)(y % P));
    }
    const int n = (int)a.size();
    Scratch da(c), db(c), dout(c);
    MS_TRY(da.alloc(n * sizeof(T)));
    MS_TRY(db.alloc(n * sizeof(T)));
    MS_TRY(dout.alloc(SELFTEST_OUT * (size_t)n * sizeof(T)));
    MS_CUDA(c, cudaMemcpyAsync(da.p, a.data(), n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    MS_CUDA(c, cudaMemcpyAsync(db.p, b.data(), n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    k_selftest_ops<F><<<(n + 255) / 256, 256, 0, c->stream>>>(da.as<T>(), db.as<T>(), dout.as<T>(), n);
    MS_LAUNCH_CHECK(c);
    std::vector<T> out(SELFTEST_OUT * (size_t)n);
    MS_CUDA(c, cudaMemcpyAsync(out.data(), dout.p, out.size() * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    MS_CUDA(c, cudaStreamSynchronize(c->stream));
    uint64_t bad = 0;
    for (int i = 0; i < n; i++) {
        const uint64_t av = (uint64_t)a[i] % P, bv = (uint64_t)b[i];
        const uint64_t m = (uint64_t)(((unsigned __int128)av * bv) % P);
        const uint64_t ad = (uint64_t)(((unsigned __int128)av + m) % P), sb = (uint64_t)(((unsigned __int128)av + P - m) % P);
        const T* o = &out[(size_t)SELFTEST_OUT * i];
        bool ok = (uint64_t)o[0] == m && (uint64_t)o[1] % P == ad && (uint64_t)o[2] % P == sb && (uint64_t)o[3] == av;
        if (!lazy) ok = ok && (uint64_t)o[1] < P && (uint64_t)o[2] < P;
        if (ShiftTw<F>::ON) {
            for (int k = 0; k < SELFTEST_NSHIFT; k++) {  // canonical a * 2^S, by repeated doubling
                uint64_t want = av;
                for (int d = 0; d < SELFTEST_SHIFTS[k]; d++) want = (uint64_t)(((unsigned __int128)want * 2) % P);
                ok = ok && (uint64_t)o[4 + k] == want;
            }
        }
        bad += ok ? 0 : 1;
    }
    if (n_bad) *n_bad = bad;
    return MS_OK;
}

// cm_stride = elements between columns of the column-major side (0: rows)
template <class F>
int transpose(Ctx* c, const typename F::T* d_in, typename F::T* d_out, uint64_t rows, uint64_t width, bool to_colmajor,
              uint64_t cm_stride = 0) {
    if (rows == 0 || width == 0) return MS_OK;
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((width + 31) / 32));
    k_transpose<F><<<grid, dim3(32, 8), 0, c->stream>>>(d_in, d_out, rows, width, cm_stride ? cm_stride : rows, to_colmajor ? 1 : 0);
    MS_LAUNCH_CHECK(c);
    return MS_OK;
}
#endif  // MS_NTT_NO_HOST

}  // namespace ms
