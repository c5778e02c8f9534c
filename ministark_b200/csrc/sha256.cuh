// sha256.cuh -- FIPS 180-4 SHA-256 compression for device and host.
// The reference hashes with sha2::Sha256 through the `digest` traits (src/merkle.rs:162-177); digests
// live in device memory as the 8 big-endian state words H0..H7 held in native uint32_t (the byte
// string of the reference is the big-endian serialisation of those words, done on the host).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ms {

__host__ __device__ __forceinline__ uint32_t sha_rotr(uint32_t x, int n) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

// SHA-256 is all rotates, 3-input logic and adds: on sm_100a every one of those issues on the ALU
// pipe (SHF / LOP3 / IADD3, one warp instruction per two cycles per scheduler) while the FMA pipe
// idles; ncu shows the hashing kernels at ~90 % ALU-pipe utilisation.  A rotate is also the two halves
// of the 64-bit product x * 2^(32-n): one IMAD.WIDE on the FMA pipe, and since the halves have
// disjoint bits they can go straight into the 3-input XORs.  Per Sigma / sigma that trades
// (3 SHF + 1 LOP3) for (2 IMAD.WIDE + 1 SHF + 2 LOP3): one ALU instruction less.  The multiplier is
// read from constant memory so that ptxas cannot strength-reduce it back into shifts.
// MS_SHA_FMA_MASK selects where this is applied: 1 Sigma1, 2 Sigma0, 4 sigma0, 8 sigma1.
#ifndef MS_SHA_FMA_MASK
#define MS_SHA_FMA_MASK 0
#endif
#ifdef __CUDACC__
__constant__ uint32_t SHA_ROT_MUL[32] = {
    0u,        1u << 31, 1u << 30, 1u << 29, 1u << 28, 1u << 27, 1u << 26, 1u << 25, 1u << 24, 1u << 23, 1u << 22,
    1u << 21,  1u << 20, 1u << 19, 1u << 18, 1u << 17, 1u << 16, 1u << 15, 1u << 14, 1u << 13, 1u << 12, 1u << 11,
    1u << 10,  1u << 9,  1u << 8,  1u << 7,  1u << 6,  1u << 5,  1u << 4,  1u << 3,  1u << 2,  1u << 1};
#endif
// Adds on the FMA pipe (experiment, OFF).  ncu (profiles/r01_d_ncu_merkle.txt) shows the hashing kernels
// ALU-pipe bound (88-95 % ALU, 6-10 % FMA): every SHF / LOP3 / IADD3 issues on the ALU pipe.  a + b is
// also a * 1 + b, an IMAD on the idle FMA pipe; the multiplier comes from constant memory so that ptxas
// cannot fold it back into an IADD3.  MS_SHA_IMAD_ADD selects where: 1 the t1 chain, 2 e = d + t1,
// 4 t2 / a, 8 the message schedule, 16 the final state update.  Measured on B200 with every add moved
// (ALU instructions per binary node 2108 -> 1656, FMA 178 -> 992): LDE tree 16.68 -> 16.79 ms, i.e. no
// gain.  profiles/r01_d_pipes.txt explains it: two pipes only overlap while the register-file operand
// bandwidth lasts (~1.7 register reads per clock per scheduler), ptxas keeps the multiplier in a vector
// register (IMAD R, R, R, R: three reads), and the extra reads cost what the freed ALU slots gain.
// r02 sweep of the mask on B200 (profiles/r02_d_sha_pipe_ab.txt; headline LDE tree, ms): 0: 16.68, 2: 16.67, 8: 16.27,
// 10: 16.36, 24: 16.42, 9: 16.79, 11: 17.17, 14: 15.67, **12: 15.50**.  The message-schedule adds and the t2 / a adds pay
// (the multiplier sits in a uniform register: two vector-register reads per IMAD), the t1 chain does not: 12 is the default.
#ifndef MS_SHA_IMAD_ADD
#define MS_SHA_IMAD_ADD 12
#endif
#ifdef __CUDACC__
__constant__ uint32_t SHA_ONE_OPAQUE = 1u;
#endif
template <int WHICH>
__host__ __device__ __forceinline__ uint32_t sha_add(uint32_t a, uint32_t b, uint32_t one) {
#ifdef __CUDA_ARCH__
    if (MS_SHA_IMAD_ADD & WHICH) {
        uint32_t d;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
        return d;
    }
#endif
    (void)one;
    return a + b;
}
__host__ __device__ __forceinline__ uint32_t sha_one() {
#ifdef __CUDA_ARCH__
    return SHA_ONE_OPAQUE;
#else
    return 1u;
#endif
}

// xor of rotr(x, n1), rotr(x, n2) and `third` (a rotate or a shift computed by the caller)
template <int WHICH>
__host__ __device__ __forceinline__ uint32_t sha_xor3(uint32_t x, int n1, int n2, uint32_t third) {
#ifdef __CUDA_ARCH__
    if (MS_SHA_FMA_MASK & WHICH) {
        const uint64_t p1 = (uint64_t)x * SHA_ROT_MUL[n1], p2 = (uint64_t)x * SHA_ROT_MUL[n2];
        return ((uint32_t)p1 ^ (uint32_t)(p1 >> 32) ^ (uint32_t)p2) ^ ((uint32_t)(p2 >> 32) ^ third);
    }
#endif
    return sha_rotr(x, n1) ^ sha_rotr(x, n2) ^ third;
}

__host__ __device__ __forceinline__ void sha256_init(uint32_t st[8]) {
    st[0] = 0x6a09e667; st[1] = 0xbb67ae85; st[2] = 0x3c6ef372; st[3] = 0xa54ff53a;
    st[4] = 0x510e527f; st[5] = 0x9b05688c; st[6] = 0x1f83d9ab; st[7] = 0x5be0cd19;
}

// One compression: st <- st + F(st, w).  w[16] is the big-endian-decoded block and is clobbered
// (rolling message schedule kept in the same 16 registers).
__host__ __device__ __forceinline__ void sha256_compress(uint32_t st[8], uint32_t w[16], uint32_t one = sha_one()) {
    constexpr uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t wi;
        if (i < 16) {
            wi = w[i];
        } else {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = sha_xor3<4>(w15, 7, 18, w15 >> 3);
            uint32_t s1 = sha_xor3<8>(w2, 17, 19, w2 >> 10);
            wi = sha_add<8>(sha_add<8>(w[i & 15], s0, one), sha_add<8>(w[(i + 9) & 15], s1, one), one);
            w[i & 15] = wi;
        }
        uint32_t S1 = sha_xor3<1>(e, 6, 11, sha_rotr(e, 25));
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = sha_add<1>(sha_add<1>(sha_add<1>(wi, K[i], one), h, one), sha_add<1>(S1, ch, one), one);
        uint32_t S0 = sha_xor3<2>(a, 2, 13, sha_rotr(a, 22));
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = sha_add<4>(S0, mj, one);
        h = g; g = f; f = e; e = sha_add<2>(d, t1, one); d = c; c = b; b = a; a = sha_add<4>(t1, t2, one);
    }
    st[0] = sha_add<16>(st[0], a, one); st[1] = sha_add<16>(st[1], b, one); st[2] = sha_add<16>(st[2], c, one);
    st[3] = sha_add<16>(st[3], d, one); st[4] = sha_add<16>(st[4], e, one); st[5] = sha_add<16>(st[5], f, one);
    st[6] = sha_add<16>(st[6], g, one); st[7] = sha_add<16>(st[7], h, one);
}

// The last block of a message whose length is a multiple of 64 bytes is a constant: 0x80, zeros, the
// bit length.  Its message schedule folds into the round constants at compile time, so that compression
// is 64 rounds with immediates and no schedule arithmetic (inner Merkle nodes hash k*32 bytes: one such
// block per node, a third of a binary node's work).
struct ShaPadKW {
    uint32_t v[64];
};
constexpr uint32_t c_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
constexpr ShaPadKW sha_pad_kw(uint32_t bits) {
    constexpr uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64] = {};
    w[0] = 0x80000000u;
    w[15] = bits;
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = c_rotr(w[i - 15], 7) ^ c_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = c_rotr(w[i - 2], 17) ^ c_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    ShaPadKW r{};
    for (int i = 0; i < 64; i++) r.v[i] = K[i] + w[i];
    return r;
}
template <uint32_t BITS>
__host__ __device__ __forceinline__ void sha256_compress_padblock(uint32_t st[8], uint32_t one = sha_one()) {
    constexpr ShaPadKW KW = sha_pad_kw(BITS);
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = sha_add<1>(sha_add<1>(h, KW.v[i], one), sha_add<1>(S1, ch, one), one);
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = sha_add<4>(S0, mj, one);
        h = g; g = f; f = e; e = sha_add<2>(d, t1, one); d = c; c = b; b = a; a = sha_add<4>(t1, t2, one);
    }
    st[0] = sha_add<16>(st[0], a, one); st[1] = sha_add<16>(st[1], b, one); st[2] = sha_add<16>(st[2], c, one);
    st[3] = sha_add<16>(st[3], d, one); st[4] = sha_add<16>(st[4], e, one); st[5] = sha_add<16>(st[5], f, one);
    st[6] = sha_add<16>(st[6], g, one); st[7] = sha_add<16>(st[7], h, one);
}

}  // namespace ms
