// e2e_fibonacci.cpp -- the reference's own end-to-end tests (tests/e2e_goldilocks.rs, tests/e2e_babybear.rs) on the C++ host
// mirror (include/ministark.hpp): same claim, witness, trace, constraints and parameters, the prover behind Stark::prove on the
// GPU through the C ABI.
//
//   g++ -std=c++17 -O2 -Iinclude examples/e2e_fibonacci.cpp -Lministark_b200 -lministark -Wl,-rpath,$PWD/ministark_b200 -o e2e_fibonacci
//   ./e2e_fibonacci host                     host-side part only (no GPU): padding value, affine form, derived parameters
//   ./e2e_fibonacci prove <outdir>           both fields: prove, verify, reject a corrupted proof; writes <outdir>/<field>.proof
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

#include "ministark.hpp"

using namespace ministark;

struct Witness {
    u64 secret_b;
};

// tests/e2e_goldilocks.rs:11-63
struct FibonacciClaim : Provable<Witness> {
    const StarkField& F;
    size_t step;  // nth fibonacci number
    u64 output;   // FIXME in the reference too: the output is not used in the proof
    FibonacciClaim(const StarkField& f, size_t s, u64 o) : F(f), step(s), output(o) {}

    TraceTable trace(const Witness& witness) const override {
        const size_t trace_width = 3;
        TraceTable trace(F, step, trace_width);
        // initial state
        u64 a = 1, b = witness.secret_b % F.p, c = F.add(a, b);
        // set initial state constrains
        trace.add_boundary_constrain(0, 0);
        trace.add_boundary_constrain(0, 1);
        trace.add_boundary_constrain(0, 2);
        // trace
        for (size_t i = 0; i < trace.step_number(); i++) {
            trace.add_row(i, {a, b, c});
            a = b;
            b = c;
            c = F.add(a, b);
        }
        // set output constrains
        trace.add_boundary_constrain(step - 1, 2);
        // add transition constrains
        const DensePolynomial omega = DensePolynomial::from_coefficients_vec(F, {trace.omega});
        // a[1] == b[0]
        trace.add_transition_constrain([omega](const std::vector<DensePolynomial>& trace_polys) { return trace_polys[0].clone() * omega - trace_polys[1].clone(); });
        // b[1] == c[0]
        trace.add_transition_constrain([omega](const std::vector<DensePolynomial>& trace_polys) { return trace_polys[0].clone() * omega - trace_polys[1].clone(); });
        trace.add_transition_constrain([](const std::vector<DensePolynomial>& trace_polys) { return trace_polys[2].clone() - trace_polys[0].clone() - trace_polys[1].clone(); });
        return trace;
    }
};

static void host_part(const StarkField& F, size_t steps) {
    FibonacciClaim claim(F, steps, 13);
    const TraceTable trace = claim.trace(Witness{2});
    std::vector<u64> m, c;
    trace.affine_form(&m, &c);
    StarkConfig config(F, 20, 2, trace.step_number(), trace.constrain_number());
    std::printf("%s padding_value %llu length %zu width %zu omega %llu constrain_number %zu rounds %llu constrain_queries %llu fri_queries %llu matrix",
                F.name, (unsigned long long)padding_value(F), trace.length(), trace.width(), (unsigned long long)trace.omega, trace.constrain_number(),
                (unsigned long long)config.rounds, (unsigned long long)config.constrain_queries, (unsigned long long)config.fri_queries);
    for (u64 v : m) std::printf(" %llu", (unsigned long long)v);
    std::printf(" constants");
    for (u64 v : c) std::printf(" %llu", (unsigned long long)v);
    std::printf(" last_row");
    for (size_t j = 0; j < trace.width(); j++) std::printf(" %llu", (unsigned long long)trace.data()[(steps - 1) * trace.width() + j]);
    std::printf("\n");
}

// test_fibonacci_air_constrains + test_stark_prover
static int prove_part(const StarkField& F, size_t steps, const std::string& outdir) {
    FibonacciClaim claim(F, steps, 13);
    const Witness witness{2};
    const TraceTable trace = claim.trace(witness);
    const size_t blowup_factor = 2, columns = trace.constrain_number();
    StarkConfig config = StarkConfig::new_(F, 20, blowup_factor, trace.step_number(), columns);
    Stark proof_system = Stark::new_(config);
    const Constrains constrains = trace.derive_constrains(proof_system.gpu());

    // check output constrain: constrain polynomials times the vanishing polynomial vanish on the trace domain (e2e_goldilocks.rs:79-95)
    const size_t N = trace.length();
    for (size_t k : {size_t(2), size_t(3)}) {
        DensePolynomial poly(F, constrains.get_constrain_poly(k));
        std::vector<u64> van(N + 1, 0);
        van[0] = F.p - 1;
        van[N] = 1;
        const DensePolynomial prod = poly * DensePolynomial(F, van);
        for (size_t i = 0; i + 1 < trace.step_number(); i++) {
            const u64 w_i = F.pow(trace.omega, i);
            u64 acc = 0;
            for (size_t d = prod.coeffs.size(); d-- > 0;) acc = F.add(F.mul(acc, w_i), prod.coeffs[d]);
            if (acc != 0) { std::printf("%s: constraint %zu does not vanish at w^%zu\n", F.name, k, i); return 1; }
        }
    }
    // the third trace polynomial interpolates the c register: f_2(w^i) = c_i
    {
        const std::vector<u64> f2 = constrains.get_constrain_poly(2);
        for (size_t i = 0; i < trace.step_number(); i++) {
            const u64 w_i = F.pow(trace.omega, i);
            u64 acc = 0;
            for (size_t d = f2.size(); d-- > 0;) acc = F.add(F.mul(acc, w_i), f2[d]);
            if (acc != trace.data()[i * trace.width() + 2]) { std::printf("%s: trace polynomial 2 misses row %zu\n", F.name, i); return 1; }
        }
    }

    const StarkProof proof = proof_system.prove(claim, witness);
    const bool is_alright = proof_system.verify(constrains, proof);
    const bool strict_ok = proof_system.verify(constrains, proof, true);
    // a corrupted opening must be rejected like the reference's assert!s would (std::logic_error naming the line)
    bool rejected = false;
    {
        StarkProof bad = proof;
        bad.raw[24 + bad.arthur.size() + 64 + 16 + 3] ^= 1;  // first constrain_queries scalar
        try { proof_system.verify(constrains, bad); } catch (const std::logic_error&) { rejected = true; } catch (const Error&) { rejected = true; }
    }
    std::ofstream(outdir + "/" + F.name + ".proof", std::ios::binary).write(reinterpret_cast<const char*>(proof.raw.data()), (std::streamsize)proof.raw.size());
    std::printf("%s proof_len %zu arthur_len %zu verify %d strict %d corrupted_rejected %d\n", F.name, proof.raw.size(), proof.arthur.size(), (int)is_alright,
                (int)strict_ok, (int)rejected);
    return (is_alright && strict_ok && rejected) ? 0 : 1;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "host";
    try {
        if (mode == "host") {
            host_part(Goldilocks(), 9);
            host_part(BabyBear(), 7);
            return 0;
        }
        const std::string outdir = argc > 2 ? argv[2] : ".";
        int rc = prove_part(Goldilocks(), 9, outdir);   // tests/e2e_goldilocks.rs:65-77: step 9
        rc |= prove_part(BabyBear(), 7, outdir);        // tests/e2e_babybear.rs: step 7
        return rc;
    } catch (const Error& e) {
        std::fprintf(stderr, "ministark::Error %d: %s\n", (int)e.code, e.what());
        return 3;
    } catch (const std::logic_error& e) {
        std::fprintf(stderr, "panic: %s\n", e.what());
        return 4;
    }
}
