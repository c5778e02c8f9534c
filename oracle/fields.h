/*
 * oracle/fields.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Prime fields and extension towers of the reference, restated in plain C on canonical
 * integers.  The reference keeps everything in ark-ff 0.5.0 `MontBackend` (source not in
 * /root/reference); only canonical values are observable, so plain modular arithmetic on
 * canonical representatives is an exact restatement.
 *
 *   Goldilocks  p = 2^64 - 2^32 + 1, generator 7          reference src/field.rs:43-47
 *   Fp2 = Fp[u]/(u^2 - 7)                                   reference src/field.rs:50-62
 *   BabyBear    p = 2013265921, "generator" 440564289      reference src/field.rs:72-76
 *   Fp2 = Fp[u]/(u^2 - 11)                                  reference src/field.rs:78-91
 *   Fp4 = Fp2[v]/(v^2 - (u + 2013265910))                   reference src/field.rs:93-109
 *
 * Two-adic roots follow ark-ff's MontConfig derive: ROOT = GENERATOR^((p-1)/2^s)
 * (SURVEY.md App. A item 1): Goldilocks s=32, BabyBear s=27.
 *
 * Every base-field element crosses the oracle API as a canonical uint64_t (BabyBear
 * values simply stay below 2^31); an extension element is D consecutive uint64_t in
 * ark's tower order (c0, c1 / c0.c0, c0.c1, c1.c0, c1.c1).
 */
#ifndef ORACLE_FIELDS_H
#define ORACLE_FIELDS_H
#include <stdint.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ Goldilocks */
#define GL_P 0xFFFFFFFF00000001ULL
#define GL_ROOT 1753635133440165772ULL /* 7^((p-1)/2^32) */
#define GL_TWO_ADICITY 32
#define GL_BITS 64

static inline u64 gl_add(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
}
static inline u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : a + (GL_P - b); }
/* 128-bit product reduced with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p); same value as (a*b) % p */
static inline u64 gl_mul(u64 a, u64 b) {
    u128 x = (u128)a * b;
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= 0xFFFFFFFFULL;
    u64 t1 = (hl << 32) - hl;
    u64 r = t0 + t1;
    if (r < t1) r += 0xFFFFFFFFULL;
    return r >= GL_P ? r - GL_P : r;
}
static inline u64 gl_mul_slow(u64 a, u64 b) { return (u64)(((u128)a * b) % GL_P); }

/* ------------------------------------------------------------------ BabyBear */
#define BB_P 2013265921ULL
#define BB_ROOT 291241980ULL /* 440564289^15 */
#define BB_TWO_ADICITY 27
#define BB_BITS 31

static inline u64 bb_add(u64 a, u64 b) {
    u64 s = a + b;
    return s >= BB_P ? s - BB_P : s;
}
static inline u64 bb_sub(u64 a, u64 b) { return a >= b ? a - b : a + BB_P - b; }
static inline u64 bb_mul(u64 a, u64 b) { return (a * b) % BB_P; }

#endif
