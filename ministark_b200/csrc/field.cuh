// field.cuh -- device (and host) arithmetic for the two STARK fields of the reference and their
// extension towers (reference src/field.rs:43-109).  Everything is exact modular integer
// arithmetic on canonical representatives: the reference keeps ark-ff Montgomery form internally
// but only canonical values are observable (Display, serialisation, challenges), so the device is
// free to use plain Goldilocks reduction (2^64 = 2^32 - 1 mod p, built from 32-bit IMAD chains by
// nvcc) and plain / Montgomery BabyBear.  Values in HBM are ALWAYS canonical.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define MS_HD __host__ __device__ __forceinline__

namespace ms {

// ------------------------------------------------------------------------------------------
// Goldilocks  p = 2^64 - 2^32 + 1   (field.rs:43-47; two-adic root 7^((p-1)/2^32))
// ------------------------------------------------------------------------------------------
struct GL {
    using T = uint64_t;
    static constexpr int ID = 0;
    static constexpr int D = 2;  // extension degree: Fp2 = Fp[u]/(u^2 - 7), field.rs:50-62
    static constexpr T P = 0xFFFFFFFF00000001ULL;
    static constexpr T EPS = 0xFFFFFFFFULL;  // 2^64 mod p
    static constexpr T ROOT = 1753635133440165772ULL;
    static constexpr int TWO_ADICITY = 32;
    static constexpr int BITS = 64;
    static constexpr T NONRES = 7;

    static MS_HD T add(T a, T b) {
        T s = a + b;
        if (s < a) return s + EPS;  // wrapped: true sum - p = s + 2^64 - p
        return s >= P ? s - P : s;
    }
    static MS_HD T sub(T a, T b) {
        T d = a - b;
        return a < b ? d - EPS : d;  // wrapped: a - b + p = d - 2^64 + p
    }
    static MS_HD T neg(T a) { return a ? P - a : 0; }
    // reduce hi*2^64 + lo; 2^64 = EPS, 2^96 = -1 (mod p)
    static MS_HD T reduce128(T lo, T hi) {
        T hh = hi >> 32, hl = hi & EPS;
        T t0 = lo - hh;
        if (lo < hh) t0 -= EPS;
        T t1 = (hl << 32) - hl;  // hl * EPS
        T r = t0 + t1;
        if (r < t1) r += EPS;
        return r >= P ? r - P : r;
    }
    static MS_HD T mul(T a, T b) {
#ifdef __CUDA_ARCH__
        return reduce128(a * b, __umul64hi(a, b));
#else
        unsigned __int128 x = (unsigned __int128)a * b;
        return reduce128((T)x, (T)(x >> 64));
#endif
    }
    static MS_HD T mul_small(T a, uint32_t c) {  // c < 2^32
#ifdef __CUDA_ARCH__
        return reduce128(a * c, __umul64hi(a, (T)c));
#else
        unsigned __int128 x = (unsigned __int128)a * c;
        return reduce128((T)x, (T)(x >> 64));
#endif
    }
};

// ------------------------------------------------------------------------------------------
// BabyBear  p = 2013265921 = 15 * 2^27 + 1   (field.rs:72-76; two-adic root 440564289^15)
// ------------------------------------------------------------------------------------------
struct BB {
    using T = uint32_t;
    static constexpr int ID = 1;
    static constexpr int D = 4;  // Fp4 = Fp2[v]/(v^2 - u) (ark-ff's effective tower, see ext_mul), Fp2 = Fp[u]/(u^2 - 11)
    static constexpr T P = 2013265921u;
    static constexpr T ROOT = 291241980u;
    static constexpr int TWO_ADICITY = 27;
    static constexpr int BITS = 31;
    static constexpr T NONRES = 11;

    static MS_HD T add(T a, T b) {
        T s = a + b;
        return s >= P ? s - P : s;
    }
    static MS_HD T sub(T a, T b) { return a >= b ? a - b : a + P - b; }
    static MS_HD T neg(T a) { return a ? P - a : 0; }
    // Montgomery reduction t 2^-32 mod p for t < p 2^32, result canonical (-p^-1 mod 2^32 = 2013265919)
    static MS_HD T redc(uint64_t t) {
        const uint32_t m = (uint32_t)t * 2013265919u;
        const uint32_t u = (uint32_t)(((uint64_t)m * P + t) >> 32), v = u - P;  // u in [0, 2p)
        return u < v ? u : v;
    }
    static MS_HD T mul(T a, T b) {
#ifdef __CUDA_ARCH__
        // canonical in, canonical out without the 64-bit division: REDC(REDC(a b) 2^64) -- two IMAD.WIDE pairs
        // (8 instructions) against 13 with two IMAD.HI chains for "% P"
        constexpr uint32_t R2 = (uint32_t)((((uint64_t)1 << 32) % P) * (((uint64_t)1 << 32) % P) % P);
        return redc((uint64_t)redc((uint64_t)a * b) * R2);
#else
        return (T)(((uint64_t)a * b) % P);
#endif
    }
    static MS_HD T mul_small(T a, uint32_t c) { return mul(a, c % P); }
};

template <class F>
MS_HD typename F::T fpow(typename F::T b, uint64_t e) {
    typename F::T r = 1;
    while (e) {
        if (e & 1) r = F::mul(r, b);
        b = F::mul(b, b);
        e >>= 1;
    }
    return r;
}
template <class F>
MS_HD typename F::T finv(typename F::T a) {
    return fpow<F>(a, (uint64_t)F::P - 2);
}
// ark-poly Radix2EvaluationDomain group generator for size 2^log_n
template <class F>
MS_HD typename F::T root_of_unity(int log_n) {
    typename F::T r = F::ROOT;
    for (int i = log_n; i < F::TWO_ADICITY; i++) r = F::mul(r, r);
    return r;
}

// ------------------------------------------------------------------------------------------
// Extension elements in ark's tower order (c0, c1 / c0.c0, c0.c1, c1.c0, c1.c1)
// ------------------------------------------------------------------------------------------
template <class F>
struct Ext {
    typename F::T c[F::D];
};

template <class F>
MS_HD Ext<F> ext_zero() {
    Ext<F> r;
#pragma unroll
    for (int i = 0; i < F::D; i++) r.c[i] = 0;
    return r;
}
template <class F>
MS_HD Ext<F> ext_from_base(typename F::T b) {
    Ext<F> r = ext_zero<F>();
    r.c[0] = b;
    return r;
}
template <class F>
MS_HD Ext<F> ext_add(const Ext<F>& a, const Ext<F>& b) {
    Ext<F> r;
#pragma unroll
    for (int i = 0; i < F::D; i++) r.c[i] = F::add(a.c[i], b.c[i]);
    return r;
}
template <class F>
MS_HD Ext<F> ext_sub(const Ext<F>& a, const Ext<F>& b) {
    Ext<F> r;
#pragma unroll
    for (int i = 0; i < F::D; i++) r.c[i] = F::sub(a.c[i], b.c[i]);
    return r;
}
template <class F>
MS_HD Ext<F> ext_mul_base(const Ext<F>& a, typename F::T b) {
    Ext<F> r;
#pragma unroll
    for (int i = 0; i < F::D; i++) r.c[i] = F::mul(a.c[i], b);
    return r;
}
template <class F>
MS_HD bool ext_eq(const Ext<F>& a, const Ext<F>& b) {
    bool e = true;
#pragma unroll
    for (int i = 0; i < F::D; i++) e = e && (a.c[i] == b.c[i]);
    return e;
}
template <class F>
MS_HD bool ext_is_zero(const Ext<F>& a) {
    bool e = true;
#pragma unroll
    for (int i = 0; i < F::D; i++) e = e && (a.c[i] == 0);
    return e;
}

// quadratic step: (a0 + a1 t)(b0 + b1 t) with t^2 = nr (a base-field non-residue)
template <class F>
MS_HD void fp2_mul(const typename F::T* a, const typename F::T* b, typename F::T* r) {
    using T = typename F::T;
    T r0 = F::add(F::mul(a[0], b[0]), F::mul_small(F::mul(a[1], b[1]), (uint32_t)F::NONRES));
    T r1 = F::add(F::mul(a[0], b[1]), F::mul(a[1], b[0]));
    r[0] = r0;
    r[1] = r1;
}

MS_HD Ext<GL> ext_mul(const Ext<GL>& a, const Ext<GL>& b) {
    Ext<GL> r;
    fp2_mul<GL>(a.c, b.c, r.c);
    return r;
}
// BabyBearFp4 (field.rs:93-109).  The tower ark-ff 0.5.0 actually computes in is v^2 = u: QuadExtField's
// mul/square/inverse call Fp4Config::mul_fp2_by_nonresidue_in_place, whose default body is
// (c0, c1) -> (Fp2::NONRESIDUE * c1, c0), i.e. multiplication by u, and the reference does not override it
// (the trait doc says NONRESIDUE "must equal (0, 1)").  The declared constant (2013265910, 1) = u - 11 of
// field.rs:96 is never read on the prover path; the declared Frobenius coefficients 11^((q^i-1)/4)
// (field.rs:98-107) are those of v^4 = 11, i.e. of v^2 = u.  [ark-ff source is not in this image: recalled.]
MS_HD Ext<BB> ext_mul(const Ext<BB>& a, const Ext<BB>& b) {
    using T = BB::T;
    T a0b0[2], a1b1[2], a0b1[2], a1b0[2];
    fp2_mul<BB>(&a.c[0], &b.c[0], a0b0);
    fp2_mul<BB>(&a.c[2], &b.c[2], a1b1);
    fp2_mul<BB>(&a.c[0], &b.c[2], a0b1);
    fp2_mul<BB>(&a.c[2], &b.c[0], a1b0);
    Ext<BB> r;
    r.c[0] = BB::add(a0b0[0], BB::mul_small(a1b1[1], (uint32_t)BB::NONRES));  // u * (x + y u) = 11 y + x u
    r.c[1] = BB::add(a0b0[1], a1b1[0]);
    r.c[2] = BB::add(a0b1[0], a1b0[0]);
    r.c[3] = BB::add(a0b1[1], a1b0[1]);
    return r;
}
template <class F>
MS_HD Ext<F> ext_pow(Ext<F> b, uint64_t e) {
    Ext<F> r = ext_from_base<F>(1);
    while (e) {
        if (e & 1) r = ext_mul(r, b);
        b = ext_mul(b, b);
        e >>= 1;
    }
    return r;
}

}  // namespace ms
