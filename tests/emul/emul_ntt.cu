// Host emulation of k_ntt_tile (tests/test_ntt_emulation.py): the per-thread round body of
// csrc/ntt.cuh is __host__ __device__, so the index maps, round structure, swizzle and twiddle
// tables can be checked on a CPU-only box against a naive evaluation.  Test infrastructure only.
#define MS_NTT_NO_HOST
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ministark_b200/csrc/ntt.cuh"
using namespace ms;

template <class F>
static void run_tiles(const NttTile<F>& g, unsigned threads) {
    using T = typename F::T;
    std::vector<T> S((size_t)1 << (g.a + g.beta));
    for (uint32_t bid = 0; bid < g.cols * g.tiles; bid++) {
        uint32_t col, tile;
        tile_of_block<F>(g, bid, &col, &tile);
        const T* src = g.src + (uint64_t)col * g.src_stride;
        T* dst = g.dst + (uint64_t)col * g.dst_stride;
        int nr = tile_rounds(g.a);
        if (nr == 0) {
            for (unsigned t = 0; t < threads; t++) tile_round<F>(g, S.data(), src, dst, tile, 0, 0, 0, t, threads);
            continue;
        }
        int u = 0;
        for (int r = 0; r < nr; r++) {
            int G = tile_round_size(g.a, r);
            for (unsigned t = 0; t < threads; t++) tile_round<F>(g, S.data(), src, dst, tile, r, u, G, t, threads);
            u += G;
        }
    }
}

template <class F, int A, int BETA, int R, bool INV>
static void emu_fixed_rounds(const NttTile<F>& g, typename F::T* S, const typename F::T* src, typename F::T* dst, uint32_t tile) {
    if constexpr (R < Rounds<A>::NR) {
        for (uint32_t t = 0; t < 256; t++) fixed_round<F, A, BETA, R, 256, INV>(g, S, src, dst, tile, t);
        emu_fixed_rounds<F, A, BETA, R + 1, INV>(g, S, src, dst, tile);
    }
}
template <class F, int A, int BETA>
static void run_fixed(const NttTile<F>& g) {
    using T = typename F::T;
    std::vector<T> S((size_t)1 << (A + BETA));
    for (uint32_t bid = 0; bid < g.cols * g.tiles; bid++) {
        uint32_t col, tile;
        tile_of_block<F>(g, bid, &col, &tile);
        const T* src = g.src + (uint64_t)col * g.src_stride;
        T* dst = g.dst + (uint64_t)col * g.dst_stride;
        if (g.inv) emu_fixed_rounds<F, A, BETA, 0, true>(g, S.data(), src, dst, tile);
        else emu_fixed_rounds<F, A, BETA, 0, false>(g, S.data(), src, dst, tile);
    }
}
// the shapes the emulation runs through the fixed-shape body: the product's (MS_NTT_FIXED_SHAPES) plus small ones
#define EMU_FIXED_SHAPES(X) X(8, 5) X(9, 4) X(10, 3) X(11, 2) X(12, 1) X(13, 0) X(8, 4) X(9, 3) X(10, 2) X(11, 1) X(12, 0) X(5, 2) X(6, 0) X(7, 3) X(5, 1) X(7, 1)
template <class F>
static bool emu_is_fixed(const NttTile<F>& g) {
    if (!(g.mode == 0 || g.logR1 <= g.a - tile_round_size(g.a, 0))) return false;
#define X(A_, B_) if (g.a == A_ && g.beta == B_) return true;
    EMU_FIXED_SHAPES(X)
    if (sizeof(typename F::T) == 4) { MS_NTT_FIXED_SHAPES_W4(X) }  // the 2^14-element tiles of large 4-byte transforms
#undef X
    return false;
}
static int g_fixed_used = 0;
static int g_tile_log = NTT_LOG_TILE_PREF;  // the planner's tile size (Ctx::ntt_log_tile in the product)
template <class F>
static void run_any(const NttTile<F>& g) {
    if (emu_is_fixed<F>(g)) {
#define X(A_, B_) if (g.a == A_ && g.beta == B_) { g_fixed_used++; return run_fixed<F, A_, B_>(g); }
        EMU_FIXED_SHAPES(X)
        if constexpr (sizeof(typename F::T) == 4) { MS_NTT_FIXED_SHAPES_W4(X) }
#undef X
    }
    run_tiles<F>(g, 64);
}
// host copy of get_tw16 / k_build_tw16 (csrc/ntt.cuh)
template <class F>
static std::vector<typename F::T> host_tw16(int a, bool inverse, const std::vector<typename F::T>* shifts, int B, int log_n1) {
    using T = typename F::T;
    TwRoots<F> roots{};
    for (int l = 0; l <= NTT_MAXLOG; l++) {
        T g = (l <= F::TWO_ADICITY) ? root_of_unity<F>(l) : (T)1;
        roots.w[l] = inverse ? finv<F>(g) : g;
    }
    if (!shifts) B = 1;
    std::vector<T> tab((size_t)B << (a + 1));
    for (int j = 0; j < B; j++) {
        T sigma = 1;
        if (shifts) {
            sigma = (*shifts)[j];
            for (int i = 0; i < log_n1; i++) sigma = F::mul(sigma, sigma);
        }
        for (uint32_t e = 0; e < (2u << a); e++) tab[((size_t)j << (a + 1)) + e] = tw16_entry<F>(e, a, sigma, roots);
    }
    return tab;
}

// mirrors lde_batch (csrc/ntt.cuh) with host tables
template <class F>
static std::vector<typename F::T> emu_lde(const std::vector<typename F::T>& in, uint64_t cols, int logN, int logB,
                                          typename F::T shift, bool inverse, int force_a) {
    using T = typename F::T;
    NttPlan pl;
    int tile_log = 0;
    if (!ntt_plan_for<F>(logN, logB, g_tile_log, &pl, &tile_log)) { printf("plan failed\n"); exit(2); }
    if (force_a >= 0) {  // exercise two-pass geometry on small sizes
        pl.a = force_a; pl.b = logN - force_a;
        pl.logR1 = 0; pl.cs1 = 0; pl.beta1 = logB; pl.beta2 = logB;
        if (logB == 0 && pl.b >= 2) { pl.logR1 = 2; pl.beta1 = 2; }
        if (logB == 0 && pl.a >= 1) pl.beta2 = 1;
        if (logB == 2 && (force_a & 1)) { pl.cs1 = 1; pl.beta1 = 1; pl.beta2 = 1; }   // cosets split over two tiles, half-width pass 2
        if (logB == 2 && force_a == 3) pl.beta2 = 0;
    }
    const int B = 1 << logB;
    const uint64_t N = 1ULL << logN;
    std::vector<T> wtab((size_t)1 << NTT_MAXLOG, 0);
    for (int u = 0; u < NTT_MAXLOG; u++) {
        T g = (u + 1 <= F::TWO_ADICITY) ? root_of_unity<F>(u + 1) : (T)1;
        if (inverse) g = finv<F>(g);
        for (int q = 0; q < (1 << u); q++) wtab[(1 << u) + q] = Fast<F>::to_tw(fpow<F>(g, q));
    }
    T wN = root_of_unity<F>(logN);
    if (inverse) wN = finv<F>(wN);
    const T scale = inverse ? finv<F>((T)(N % (uint64_t)F::P)) : (T)1;
    const T wL = root_of_unity<F>(logN + logB);
    const bool two = pl.b > 0;
    std::vector<T> shifts(B), t1((size_t)B << pl.a, 0);
    T sj = shift;
    for (int j = 0; j < B; j++) {
        shifts[j] = sj;
        for (int u = 0; u < pl.a; u++) {
            T sb = fpow<F>(sj, N >> (u + 1));
            for (int q = 0; q < (1 << u); q++) t1[((size_t)j << pl.a) + (1 << u) + q] = F::mul(sb, wtab[(1 << u) + q]);
        }
        sj = F::mul(sj, wL);
    }
    std::vector<T> ft, tmp, out(cols * (N << logB), 0);
    const bool inplace = two && pl.logR1 == 0 && tile_rounds(pl.b) >= 2;
    if (two) {
        ft.resize(N << logB);
        const int beta = pl.logR1 + logB;
        for (uint64_t m1 = 0; m1 < (1ULL << pl.b); m1++)
            for (int j = 0; j < B; j++)
                for (uint64_t k2 = 0; k2 < (1ULL << pl.a); k2++) {
                    uint64_t tile = m1 >> pl.logR1, rr = m1 & ((1u << pl.logR1) - 1);
                    T v = F::mul(F::mul(scale, fpow<F>(shifts[j], m1)), fpow<F>(wN, m1 * k2));
                    ft[((((tile << pl.a) + k2) << beta) | (rr << logB)) + j] = Fast<F>::to_tw(v);
                }
        if (!inplace) tmp.resize(cols * (N << logB));
    }
    NttTile<F> g1{};
    g1.src = in.data(); g1.src_stride = N;
    g1.dst = two ? (inplace ? out.data() : tmp.data()) : out.data();
    g1.dst_stride = N << logB;
    g1.tw = t1.data(); g1.ft = two ? ft.data() : nullptr;
    g1.scale = Fast<F>::to_tw(scale); g1.has_scale = (!two && scale != 1) ? 1 : 0;
    g1.a = pl.a; g1.beta = pl.beta1; g1.cs = pl.cs1; g1.logB = logB;
    g1.jmask = B - 1; g1.jstride = 1u << pl.a; g1.mode = 0; g1.bq = pl.b; g1.inv = inverse ? 1 : 0;
    g1.tiles = (uint32_t)(((1ULL << pl.b) >> pl.logR1) << pl.cs1); g1.cols = (uint32_t)cols;
    std::vector<T> tw16_1, tw16_2;
    if (ShiftTw<F>::ON && emu_is_fixed<F>(g1)) {  // Goldilocks fixed-shape body: block twiddles (as lde_batch)
        tw16_1 = host_tw16<F>(pl.a, inverse, &shifts, B, pl.b);
        g1.tw = tw16_1.data(); g1.jstride = 2u << pl.a;
    }
    run_any<F>(g1);
    if (two) {
        NttTile<F> g2{};
        g2.src = g1.dst; g2.src_stride = g1.dst_stride; g2.dst = out.data(); g2.dst_stride = N << logB;
        g2.tw = wtab.data(); g2.a = pl.b; g2.beta = pl.beta2; g2.logB = logB; g2.mode = 1; g2.plain = 1;
        g2.a1 = pl.a; g2.beta1 = pl.logR1 + logB; g2.logR1 = pl.logR1;
        g2.tiles = (uint32_t)(((1ULL << pl.a) << logB) >> pl.beta2); g2.cols = (uint32_t)cols; g2.inv = inverse ? 1 : 0;
        if (ShiftTw<F>::ON && emu_is_fixed<F>(g2)) {
            tw16_2 = host_tw16<F>(pl.b, inverse, nullptr, 1, 0);
            g2.tw = tw16_2.data();
        }
        run_any<F>(g2);
    }
    return out;
}

template <class F>
static int check(int logN, int logB, bool inverse, int force_a, uint64_t cols) {
    using T = typename F::T;
    const uint64_t N = 1ULL << logN, L = N << logB;
    std::vector<T> in(cols * N);
    uint64_t s = 0x9E3779B97F4A7C15ULL * (logN * 131 + logB * 7 + inverse + 1);
    for (auto& v : in) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (T)(s % (uint64_t)F::P); }
    T shift = inverse ? (T)1 : (T)(0x1234567ULL % (uint64_t)F::P);
    auto out = emu_lde<F>(in, cols, logN, logB, shift, inverse, force_a);
    // naive: out[c][i] = sum in[c][m] (shift w_L^i)^m  (inverse: w_N^-1, scaled 1/N)
    T wL = root_of_unity<F>(logN + logB);
    if (inverse) wL = finv<F>(wL);
    T scale = inverse ? finv<F>((T)(N % (uint64_t)F::P)) : (T)1;
    int bad = 0;
    for (uint64_t c = 0; c < cols; c++)
        for (uint64_t i = 0; i < L; i += (L > 256 ? (L > 4096 ? L / 24 + 1 : 37) : 1)) {
            T x = F::mul(shift, fpow<F>(wL, i)), acc = 0;
            for (uint64_t m = N; m-- > 0;) acc = F::add(F::mul(acc, x), in[c * N + m]);
            acc = F::mul(acc, scale);
            if (acc != out[c * L + i]) bad++;
        }
    printf("field %d logN %d logB %d inv %d force_a %d: %s\n", F::ID, logN, logB, (int)inverse, force_a, bad ? "MISMATCH" : "ok");
    return bad;
}

// The register block on its own: gl_shift_dft<G, INV> (power-of-two twiddles, lazy inputs in every slot but the
// stage-0 operands) against the defining sum over ark's root of unity, and x * 2^S against repeated doubling.
template <int G, bool INV>
static int check_shift_block(uint64_t seed) {
    constexpr int E = 1 << G;
    uint64_t v[E], in[E];
    uint64_t s = seed;
    for (int r = 0; r < E; r++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        // odd slots are stage-0 operands (canonical by contract); even slots may be any 64-bit representative
        in[r] = (r & 1) ? s % GL::P : (r % 4 == 0 ? s | 0xFFFFFFFF00000000ULL : s);
        v[r] = in[r];
    }
    gl_shift_dft<G, INV>(v);
    uint64_t w = root_of_unity<GL>(G);
    if (INV) w = finv<GL>(w);
    int bad = 0;
    for (int k = 0; k < E; k++) {
        uint64_t acc = 0;
        for (int r = 0; r < E; r++)  // DIT: slot r holds sub-index brev(r)
            acc = GL::add(acc, GL::mul(in[r] % GL::P, fpow<GL>(w, (uint64_t)cbrev(r, G) * k)));
        if (v[k] % GL::P != acc) bad++;
    }
    printf("shift block G=%d inv=%d: %s\n", G, (int)INV, bad ? "MISMATCH" : "ok");
    return bad;
}
template <int S>
static int check_mulpow2_from() {
    if constexpr (S < 96) {
        int bad = 0;
        const uint64_t xs[] = {0, 1, GL::P - 1, GL::P, ~0ULL, 0xFFFFFFFFULL, 0x100000000ULL, 0x123456789ABCDEF1ULL, 0xFFFFFFFF00000001ULL};
        for (uint64_t x : xs) {
            uint64_t want = x % GL::P;
            for (int d = 0; d < S; d++) want = GL::add(want, want);
            if (Fast<GL>::mulpow2<S>(x) != want) bad++;
        }
        if (bad) printf("mulpow2<%d>: MISMATCH\n", S);
        return bad + check_mulpow2_from<S + 1>();
    } else {
        return 0;
    }
}

int main() {
    int bad = 0;
    bad += check_shift_block<1, false>(11) + check_shift_block<2, false>(12) + check_shift_block<3, false>(13) + check_shift_block<4, false>(14);
    bad += check_shift_block<1, true>(21) + check_shift_block<2, true>(22) + check_shift_block<3, true>(23) + check_shift_block<4, true>(24);
    bad += check_mulpow2_from<1>();
    for (int logN = 0; logN <= 11; logN++)
        for (int logB = 0; logB <= 3; logB++) {
            bad += check<GL>(logN, logB, false, -1, 2);
            if (logB == 0) bad += check<GL>(logN, 0, true, -1, 2);
        }
    for (int logN = 4; logN <= 10; logN += 3)
        for (int fa = 1; fa < logN; fa += 2)
            for (int logB = 0; logB <= 2; logB += 2) {
                bad += check<GL>(logN, logB, false, fa, 3);
                bad += check<BB>(logN, logB, false, fa, 3);
                if (logB == 0) bad += check<GL>(logN, 0, true, fa, 3);
            }
    for (int logN = 0; logN <= 9; logN += 3) bad += check<BB>(logN, 1, false, -1, 2) + check<BB>(logN, 0, true, -1, 1);
    bad += check<GL>(11, 2, false, -1, 1);   // single pass, fixed (11,2)
    bad += check<BB>(11, 2, false, -1, 1);
    bad += check<GL>(10, 3, false, -1, 1);   // fixed (10,3)
    bad += check<GL>(16, 2, false, -1, 1);   // two-pass (8,5)/(8,5)
    bad += check<BB>(16, 2, false, -1, 1);
    bad += check<GL>(17, 2, false, -1, 1);   // (9,4) + (8,5)
    bad += check<GL>(16, 0, true, -1, 2);    // iNTT two-pass with R1 = R2 = 32
    bad += check<GL>(18, 0, true, -1, 1);
    bad += check<BB>(18, 0, true, -1, 1);
    bad += check<GL>(7, 2, false, 5, 1);     // (5,2) pass 1 fixed
    bad += check<GL>(12, 0, false, 6, 1);    // (6,0): 4-bit swizzle
    bad += check<BB>(12, 0, false, 6, 1);
    bad += check<GL>(10, 3, false, 7, 1);    // (7,3)
    bad += check<GL>(12, 2, false, -1, 1);   // real two-pass plan
    bad += check<GL>(12, 2, false, 7, 1);    // (7,1): cosets split over two tiles + (5,1) second pass
    bad += check<BB>(12, 2, false, 7, 1);
    bad += check<GL>(23, 2, false, -1, 1);   // planner's own coset split: (12,1) + (11,2)
    bad += check<GL>(14, 0, true, -1, 1);
    bad += check<BB>(23, 2, false, -1, 1);   // 4-byte field, cosets would split: 2^14-element tiles, (12,2) + (11,3)
    bad += check<BB>(22, 3, false, -1, 1);   // (11,3) + (11,3)
    // ((12,2) as a second pass only occurs at 2^24 rows: covered on the GPU by test_coset_lde_largest_shapes_and_widest_batches)
    // half-size tiles (MINISTARK_NTT_TILE=12): 4096-element tiles, the headline shape becomes (11,1) + (11,1) with the cosets
    // split over two pass-1 tiles
    g_tile_log = NTT_LOG_TILE_PREF - 1;
    bad += check<GL>(22, 2, false, -1, 1);
    bad += check<GL>(16, 2, false, -1, 1);
    bad += check<BB>(16, 2, false, -1, 1);
    bad += check<GL>(17, 3, false, -1, 1);
    bad += check<GL>(11, 2, false, -1, 1);
    bad += check<GL>(12, 0, false, -1, 2);
    bad += check<GL>(18, 0, true, -1, 1);
    bad += check<BB>(20, 1, false, -1, 1);
    bad += check<GL>(24, 2, false, -1, 1);
    g_tile_log = NTT_LOG_TILE_PREF;
    printf("fixed-shape tiles used: %d\n", g_fixed_used);
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad ? 1 : 0;
}
