// transcript.hpp -- the serial host side of the prover: nimue's Fiat-Shamir transcript as the reference
// uses it (src/fiatshamir.rs:48-64, 96-116; nimue @ 0e584985, `DigestBridge<Sha256>`), and the
// parameter derivation of StarkConfig::new (src/starks.rs:268-332, src/util.rs:30-44).
//
// nimue's source is not in /root/reference: the sponge construction below follows SURVEY.md App. A
// items 5-9 and is PARITY UNPINNED (no golden transcript exists).  Two details are Ctx switches
// (ms_set_transcript_option): the three domain-separation bytes, and the handling of digest bytes left
// over from the previous squeeze call -- nimue's leftovers branch, as published at 0e584985 (recalled),
// reads `self.leftovers[..len].copy_from_slice(&output[..len])`: the copy runs the wrong way, the
// leftovers are consumed but the caller's buffer keeps its old bytes.  That behaviour is the default
// (leftover_as_published); the intended one (leftovers reach the output) is the alternative.
// Everything else in the prover is independent of this file (challenges are inputs to the device stages).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "common.cuh"
#include "sha256.cuh"

namespace ms {

// ---------------------------------------------------------------------------------- host SHA-256
struct Sha256Host {
    uint32_t st[8];
    uint8_t buf[64];
    uint64_t len = 0;
    unsigned fill = 0;
    Sha256Host() { reset(); }
    void reset() {
        sha256_init(st);
        len = 0;
        fill = 0;
    }
    void block(const uint8_t* p) {
        uint32_t w[16];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        sha256_compress(st, w);
    }
    void update(const void* data, size_t n) {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        len += n;
        while (n) {
            size_t take = 64 - fill;
            if (take > n) take = n;
            memcpy(buf + fill, p, take);
            fill += (unsigned)take;
            p += take;
            n -= take;
            if (fill == 64) {
                block(buf);
                fill = 0;
            }
        }
    }
    // finalize a COPY (the hasher itself keeps streaming, like Digest::finalize on a clone)
    void digest(uint8_t out[32]) const {
        Sha256Host c = *this;
        uint64_t bits = c.len * 8;
        uint8_t pad = 0x80, z = 0;
        c.update(&pad, 1);
        while (c.fill != 56) c.update(&z, 1);
        uint8_t lb[8];
        for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        c.update(lb, 8);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)(c.st[i] >> 24);
            out[4 * i + 1] = (uint8_t)(c.st[i] >> 16);
            out[4 * i + 2] = (uint8_t)(c.st[i] >> 8);
            out[4 * i + 3] = (uint8_t)c.st[i];
        }
    }
};

// ---------------------------------------------------------------------------------- keccak-f[1600]
inline void keccak_f1600(uint64_t st[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int r = 0; r < 24; r++) {
        uint64_t bc[5], t;
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t b = bc[(i + 1) % 5];
            t = bc[(i + 4) % 5] ^ ((b << 1) | (b >> 63));
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = PIL[i];
            uint64_t b = st[j];
            st[j] = (t << ROT[i]) | (t >> (64 - ROT[i]));
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= RC[r];
    }
}

// Safe::new -> generate_tag: nimue's Keccak duplex (rate 136, overwrite-mode absorb, no padding,
// zero IV) absorbs the IO-pattern bytes and squeezes 32 (App. A item 7).
inline void nimue_tag(const std::string& io, uint8_t tag[32]) {
    const size_t R = 136;
    uint64_t lanes[25] = {0};
    uint8_t* state = reinterpret_cast<uint8_t*>(lanes);  // little-endian host
    size_t pos = 0, off = 0, n = io.size();
    while (off < n) {
        if (pos == R) {
            keccak_f1600(lanes);
            pos = 0;
        } else {
            size_t take = n - off < R - pos ? n - off : R - pos;
            memcpy(state + pos, io.data() + off, take);
            pos += take;
            off += take;
        }
    }
    keccak_f1600(lanes);
    memcpy(tag, state, 32);
}

// ---------------------------------------------------------------------------------- DigestBridge<Sha256>
struct DigestBridge {
    enum Mode { START, ABSORB, SQUEEZE };
    Sha256Host hasher;
    uint8_t cv[32];
    Mode mode = START;
    uint64_t squeeze_i = 0;
    std::vector<uint8_t> leftovers;
    uint8_t mask_absorb = 0x00, mask_squeeze = 0x01, mask_squeeze_end = 0x02;
    bool leftover_as_published = true;

    void init(const uint8_t tag[32], const uint8_t masks[3], bool as_published = true) {
        leftover_as_published = as_published;
        hasher.reset();
        memset(cv, 0, 32);
        mode = START;
        squeeze_i = 0;
        leftovers.clear();
        mask_absorb = masks[0];
        mask_squeeze = masks[1];
        mask_squeeze_end = masks[2];
        hasher.update(tag, 32);
    }
    static void be64(uint64_t v, uint8_t out[8]) {
        for (int i = 0; i < 8; i++) out[i] = (uint8_t)(v >> (56 - 8 * i));
    }
    void mask_block(uint8_t m, Sha256Host& h) {
        uint8_t blk[64] = {0};
        blk[0] = m;
        h.update(blk, 64);
    }
    void squeeze_end() {
        if (mode != SQUEEZE) return;
        hasher.reset();
        uint64_t byte_count = squeeze_i * 32 - leftovers.size();
        Sha256Host h;
        mask_block(mask_squeeze_end, h);
        h.update(cv, 32);
        uint8_t b[8];
        be64(byte_count, b);
        h.update(b, 8);
        h.digest(cv);
        mode = START;
        leftovers.clear();
    }
    void absorb(const uint8_t* data, size_t n) {
        squeeze_end();
        if (mode == START) {
            mode = ABSORB;
            mask_block(mask_absorb, hasher);
            hasher.update(cv, 32);
        }
        hasher.update(data, n);
    }
    void ratchet() {
        squeeze_end();
        uint8_t d1[32];
        hasher.digest(d1);
        hasher.reset();
        Sha256Host h;
        h.update(d1, 32);
        h.digest(cv);
        leftovers.clear();
        mode = START;
    }
    // squeeze_unchecked(output): `out` arrives with the caller's previous contents (see the header)
    void squeeze(uint8_t* out, size_t n) {
        size_t got = 0;
        for (;;) {
            if (mode == START) {
                mode = SQUEEZE;
                squeeze_i = 0;
                mask_block(mask_squeeze, hasher);
                hasher.update(cv, 32);
            } else if (mode == ABSORB) {
                ratchet();
            } else if (got == n) {
                return;
            } else if (!leftovers.empty()) {
                size_t take = n - got < leftovers.size() ? n - got : leftovers.size();
                if (!leftover_as_published) memcpy(out + got, leftovers.data(), take);
                leftovers.erase(leftovers.begin(), leftovers.begin() + take);
                got += take;
            } else {
                Sha256Host h = hasher;
                uint8_t b[8];
                be64(squeeze_i, b);
                h.update(b, 8);
                uint8_t d[32];
                h.digest(d);
                size_t take = n - got < 32 ? n - got : 32;
                memcpy(out + got, d, take);
                leftovers.insert(leftovers.end(), d + take, d + 32);
                got += take;
                squeeze_i++;
            }
        }
    }
};

// ---------------------------------------------------------------------------------- IO pattern + Merlin
struct IOPattern {
    std::string io;
    std::deque<std::pair<char, size_t>> ops;  // IOPattern::finalize: consecutive same-kind ops merged
    explicit IOPattern(const std::string& domsep) : io(domsep) {}
    void op(char kind, size_t count, const char* label) {
        io.push_back('\0');
        io.push_back(kind);
        io += std::to_string(count);
        io += label;
        if (!ops.empty() && ops.back().first == kind) ops.back().second += count;
        else ops.emplace_back(kind, count);
    }
    void add_bytes(size_t n, const char* label) { op('A', n, label); }
    void challenge_bytes(size_t n, const char* label) { op('S', n, label); }
};

// StarkIOPattern::new_stark + FriIOPattern::add_fri (src/fiatshamir.rs:48-64, 100-116); scalar counts
// become byte counts as in nimue's ark plugin: challenges (bits+128)/8 bytes per prime-field
// coordinate, absorbed scalars ceil(bits/8) bytes per coordinate (App. A items 5, 6).
inline IOPattern stark_iopattern(int bits, int ext_degree, size_t rounds, size_t constrain_queries, size_t fri_queries) {
    const size_t cb = (bits + 128) / 8, sb = (bits + 7) / 8, D = ext_degree;
    IOPattern io("\xF0\x9F\x90\xBA");  // domsep, src/starks.rs:307
    io.add_bytes(32, "commit to original trace");
    io.challenge_bytes(cb, "ZK: pick random shift of domain");
    io.add_bytes(32, "commit to quotients");
    io.challenge_bytes(cb, "batching: retrieve random scalar r");
    io.challenge_bytes(constrain_queries * D * cb, "number of queries in DEEP ALI");
    for (size_t i = 0; i + 1 < rounds; i++) {
        io.challenge_bytes(D * cb, "(DEEP) FRI: pick random z");
        io.add_bytes(2 * D * sb, "(DEEP) FRI: degree one B polynomial");
        io.challenge_bytes(D * cb, "FRI COMMIT Phase: random scalar challenge");
        io.add_bytes(32, "FRI COMMIT Phase: commit to folded codeword");
    }
    io.challenge_bytes(8 * fri_queries, "FRI QUERY Phase: choose a random element in the domain");
    return io;
}

struct Merlin {
    std::deque<std::pair<char, size_t>> stack;
    DigestBridge sponge;
    std::vector<uint8_t> transcript;  // absorbed bytes = StarkProof.arthur (src/starks.rs:160)
    bool ok = true;
    Merlin(const IOPattern& io, const uint8_t masks[3], bool leftover_as_published = true) : stack(io.ops) {
        uint8_t tag[32];
        nimue_tag(io.io, tag);
        sponge.init(tag, masks, leftover_as_published);
    }
    bool expect(char kind, size_t n) {
        if (stack.empty() || stack.front().first != kind || stack.front().second < n) {
            stack.clear();
            ok = false;
            return false;
        }
        if (stack.front().second == n) stack.pop_front();
        else stack.front().second -= n;
        return true;
    }
    bool add_bytes(const uint8_t* data, size_t n) {
        if (!expect('A', n)) return false;
        sponge.absorb(data, n);
        transcript.insert(transcript.end(), data, data + n);
        return true;
    }
    // Merlin::fill_challenge_bytes: `out` keeps its previous contents where the sponge does not write
    bool challenge_bytes(uint8_t* out, size_t n) {
        if (!expect('S', n)) return false;
        sponge.squeeze(out, n);
        return true;
    }
    // FieldChallenges::fill_challenge_scalars of nimue's ark plugin: ONE zero-initialised buffer of degree * cb
    // bytes per call, refilled for every output scalar; coordinates big-endian mod p in tower order
    // (`challenge_scalars::<1>()` is the same call with count = 1).  out: count * degree coordinates.
    bool challenge_scalars(int bits, uint64_t p, int degree, size_t count, uint64_t* out);
};

// from_be_bytes_mod_order of a (bits+128)/8-byte challenge (App. A item 6)
inline uint64_t be_bytes_mod(const uint8_t* b, size_t n, uint64_t p) {
    unsigned __int128 acc = 0;
    for (size_t i = 0; i < n; i++) acc = ((acc << 8) | b[i]) % p;
    return (uint64_t)acc;
}

inline bool Merlin::challenge_scalars(int bits, uint64_t p, int degree, size_t count, uint64_t* out) {
    const size_t cb = (size_t)(bits + 128) / 8;
    uint8_t buf[4 * 32] = {0};
    for (size_t i = 0; i < count; i++) {
        if (!challenge_bytes(buf, cb * degree)) return false;
        for (int d = 0; d < degree; d++) out[i * degree + d] = be_bytes_mod(buf + d * cb, cb, p);
    }
    return true;
}

// ---------------------------------------------------------------------------------- StarkConfig::new
inline bool is_power_of_two_u(uint64_t v) { return (v & (v - 1)) == 0; }  // util.rs:4-14 (0 counts)
inline uint64_t ceil_log2_k(uint64_t number, uint64_t base) {               // util.rs:30-44
    if (number == 1) return 1;
    unsigned log2_base = __builtin_ctzll(base), log2_number = __builtin_ctzll(number);
    if (is_power_of_two_u(number) && log2_number % log2_base == 0) return log2_number;
    unsigned next_power_2 = 64 - __builtin_clzll(number);
    return (uint64_t)((next_power_2 + log2_base - 1) / log2_base) * log2_base;
}

struct StarkDerived {
    uint64_t rounds, constrain_queries, fri_queries;
};

inline int stark_derive(int field, const ms_stark_params& p, StarkDerived* d) {
    const unsigned bits = field == MS_FIELD_GOLDILOCKS ? 64 : 31;
    if (p.security_bits < 20 || p.steps == 0 || p.blowup_factor < 2 || !is_pow2(p.blowup_factor)) return MS_ERR_BAD_SHAPE;  // starks.rs:317-320
    uint64_t log_steps = ceil_log2_k(p.steps, 2);
    if (log_steps >= bits) return MS_ERR_BAD_SHAPE;
    d->constrain_queries = (p.security_bits + (bits - log_steps) - 1) / (bits - log_steps);  // :321-323
    uint64_t rounds_q = ceil_log2_k(p.steps * p.blowup_factor, 2);                            // :325
    double rho = 1.0 / (double)p.blowup_factor;
    double denominator = std::log2(2.0 / (1.0 + rho));
    double total = (double)p.security_bits / denominator;
    d->fri_queries = (uint64_t)std::ceil(total / (double)rounds_q);  // :326-329
    d->rounds = ceil_log2_k(p.steps * p.blowup_factor + 1, 2);       // :277
    return MS_OK;
}

}  // namespace ms
