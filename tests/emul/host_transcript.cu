// Host run of the PRODUCT's Fiat-Shamir transcript (csrc/transcript.hpp: nimue IOPattern / Merlin / DigestBridge<Sha256> as the
// prover in libministark.so uses them) on a CPU-only box, for tests/test_transcript_product.py: the walk the prover makes through
// the STARK IO pattern (src/starks.rs:59-169, src/fri.rs:64-189) with scripted absorb data, every challenge printed, so that it can
// be compared with the oracle's two restatements (oracle/pyref.py, oracle/transcript.inc).  Test infrastructure only.
//   host_transcript FIELD ROUNDS CQ FQ PUBLISHED      FIELD 0 = Goldilocks, 1 = BabyBear
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../ministark_b200/csrc/transcript.hpp"
using namespace ms;

static void hex(const char* tag, const uint8_t* p, size_t n) {
    std::printf("%s ", tag);
    for (size_t i = 0; i < n; i++) std::printf("%02x", p[i]);
    std::printf("\n");
}

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const int field = std::atoi(argv[1]);
    const size_t R = std::atoi(argv[2]), CQ = std::atoi(argv[3]), FQ = std::atoi(argv[4]);
    const bool published = std::atoi(argv[5]) != 0;
    const int bits = field == 0 ? 64 : 31, D = field == 0 ? 2 : 4;
    const uint64_t p = field == 0 ? 18446744069414584321ULL : 2013265921ULL;
    const size_t sb = (bits + 7) / 8;
    StarkDerived der;
    ms_stark_params sp{100, 4, 1023, 8, 2};
    if (stark_derive(field, sp, &der) != MS_OK) return 3;
    std::printf("derived %llu %llu %llu\n", (unsigned long long)der.rounds, (unsigned long long)der.constrain_queries, (unsigned long long)der.fri_queries);
    IOPattern io = stark_iopattern(bits, D, R, CQ, FQ);
    hex("io", reinterpret_cast<const uint8_t*>(io.io.data()), io.io.size());
    uint8_t tag[32];
    nimue_tag(io.io, tag);
    hex("tag", tag, 32);
    const uint8_t masks[3] = {0x00, 0x01, 0x02};
    Merlin m(io, masks, published);
    auto absorb = [&](uint8_t fill, size_t n, bool ramp) {
        std::vector<uint8_t> b(n);
        for (size_t i = 0; i < n; i++) b[i] = ramp ? (uint8_t)(fill + i) : fill;
        if (!m.add_bytes(b.data(), n)) { std::printf("pattern violated\n"); std::exit(4); }
    };
    auto scalars = [&](const char* tagname, int degree, size_t count) {
        std::vector<uint64_t> v(count * degree);
        if (!m.challenge_scalars(bits, p, degree, count, v.data())) { std::printf("pattern violated\n"); std::exit(4); }
        std::printf("%s", tagname);
        for (uint64_t x : v) std::printf(" %llu", (unsigned long long)x);
        std::printf("\n");
    };
    absorb(0, 32, true);                 // trace root                          starks.rs:73
    scalars("shift", 1, 1);              // starks.rs:81
    absorb(32, 32, true);                // LDE root                            starks.rs:95
    scalars("r", 1, 1);                  // starks.rs:108
    scalars("queries", D, CQ);           // ONE fill_challenge_scalars call     starks.rs:124-125
    for (size_t i = 0; i + 1 < R; i++) {
        scalars("z", D, 1);              // fri.rs:89
        absorb((uint8_t)i, 2 * D * sb, false);   // the two deep values          fri.rs:94
        scalars("alpha", D, 1);          // fri.rs:96
        absorb((uint8_t)(0x40 + i), 32, false);  // round root                   fri.rs:107-108
    }
    std::vector<uint8_t> braw(8 * FQ, 0);
    if (!m.challenge_bytes(braw.data(), braw.size())) return 4;  // fri.rs:121-122
    hex("betas", braw.data(), braw.size());
    hex("arthur", m.transcript.data(), m.transcript.size());
    // the op queue is exhausted: one more operation must be refused (nimue Safe)
    uint8_t one = 0;
    std::printf("extra_squeeze_refused %d\n", m.challenge_bytes(&one, 1) ? 0 : 1);
    return 0;
}
