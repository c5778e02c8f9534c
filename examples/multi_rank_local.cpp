// multi_rank_local.cpp -- the multi-GPU prover of libministark.so driven from compiled host code, no Python and no torch:
// `world` contexts bound into one communicator with ms_comm_init_local, one host thread per rank, every rank calling
// ms_stark_prove_multi with the same arguments (include/ministark.h, "multi-GPU inside the library").  The contexts go round
// robin over the visible GPUs (MINISTARK_EXAMPLE_GPUS, default 1: "virtual ranks" sharing one GPU, which runs every sharded
// code path).  The sharded proof must equal the single-context proof byte for byte.
//
//   g++ -std=c++17 -O2 -pthread -Iinclude examples/multi_rank_local.cpp -Lministark_b200 -lministark -Wl,-rpath,$PWD/ministark_b200 -o multi_rank_local
//   ./multi_rank_local [world = 2] [log_rows = 14] [trace columns = 8]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "ministark.hpp"

using namespace ministark;

int main(int argc, char** argv) {
    const int world = argc > 1 ? std::atoi(argv[1]) : 2;
    const int log_n = argc > 2 ? std::atoi(argv[2]) : 14;
    const size_t W = argc > 3 ? (size_t)std::atoi(argv[3]) : 8;
    const int gpus = std::getenv("MINISTARK_EXAMPLE_GPUS") ? std::atoi(std::getenv("MINISTARK_EXAMPLE_GPUS")) : 1;
    const StarkField& F = Goldilocks();
    const u64 n = 1ULL << log_n;
    try {
        // bidiagonal transition constraints f_{W+t} = a_t f_t + b_t f_{t+1}: linear, so provable for any trace
        std::vector<u64> matrix(W * W, 0);
        for (size_t t = 0; t < W; t++) {
            matrix[t * W + t] = F.pow(3, 1000 + t);
            matrix[t * W + (t + 1) % W] = F.p - F.pow(5, 77 + t);
        }
        ms_stark_params params{100, 4, n - 1, 2 * W, 2};
        const u64 bound = ms_stark_proof_bound(F.id, &params, n, 2 * W);
        if (!bound) { std::fprintf(stderr, "bad parameters\n"); return 2; }

        // single context: the proof every rank count has to reproduce
        std::vector<uint8_t> single(bound);
        u64 single_len = bound;
        {
            Gpu g(F, 0);
            DeviceBuffer trace(g, W * n * 8);
            g.check(ms_trace_synth(g.ctx(), 0x5EED, n, W, trace.ptr()), "ms_trace_synth");
            g.check(ms_stark_prove_device(g.ctx(), &params, trace.ptr(), n, W, matrix.data(), W, single.data(), &single_len), "ms_stark_prove_device");
        }

        // `world` ranks: one context and one thread each
        std::vector<std::unique_ptr<Gpu>> ranks;
        std::vector<ms_ctx*> ctxs;
        for (int r = 0; r < world; r++) {
            ranks.emplace_back(new Gpu(F, r % gpus));
            ctxs.push_back(ranks.back()->ctx());
        }
        ranks[0]->check(ms_comm_init_local(ctxs.data(), world), "ms_comm_init_local");
        std::vector<std::unique_ptr<DeviceBuffer>> traces;
        for (int r = 0; r < world; r++) {  // every rank holds the trace (generated on its device: no upload)
            traces.emplace_back(new DeviceBuffer(*ranks[r], W * n * 8));
            ranks[r]->check(ms_trace_synth(ranks[r]->ctx(), 0x5EED, n, W, traces[r]->ptr()), "ms_trace_synth");
            ranks[r]->check(ms_sync(ranks[r]->ctx()), "ms_sync");
        }
        std::vector<uint8_t> proof(bound);
        std::vector<u64> lens(world, 0);
        std::vector<int32_t> rcs(world, 0);
        std::vector<std::string> errs(world);
        int32_t rank = -1, size = -1;
        const char* backend = "";
        ms_comm_info(ranks[world - 1]->ctx(), &rank, &size, &backend);
        std::vector<std::thread> threads;
        for (int r = 0; r < world; r++)
            threads.emplace_back([&, r] {
                u64 len = r == 0 ? bound : 0;  // rank 0 receives the proof bytes
                rcs[r] = ms_stark_prove_multi(ranks[r]->ctx(), &params, traces[r]->ptr(), n, W, matrix.data(), W, r == 0 ? proof.data() : nullptr, &len, 0);
                lens[r] = len;
                if (rcs[r] != MS_OK) errs[r] = ms_last_error(ranks[r]->ctx());
                // collective (the ranks leave the group together), so it runs on the rank's own thread, never in a loop on one
                // thread.  A rank that failed skips it: the library fails on all ranks or on none, and ms_ctx_destroy below
                // tears a communicator down without waiting for anybody.
                if (rcs[r] == MS_OK) ms_comm_destroy(ranks[r]->ctx());
            });
        for (auto& t : threads) t.join();
        for (int r = 0; r < world; r++)
            if (rcs[r] != MS_OK) {
                std::fprintf(stderr, "rank %d: %s (code %d)\n", r, errs[r].c_str(), (int)rcs[r]);
                return 1;
            }
        const bool same = lens[0] == single_len && std::memcmp(proof.data(), single.data(), single_len) == 0;
        std::printf("world %d backend %s last_rank %d proof_len %llu single_len %llu identical %d\n", size, backend, rank, (unsigned long long)lens[0],
                    (unsigned long long)single_len, (int)same);
        return same ? 0 : 1;
    } catch (const Error& e) {
        std::fprintf(stderr, "ministark::Error %d: %s\n", (int)e.code, e.what());
        return 3;
    } catch (const std::logic_error& e) {
        std::fprintf(stderr, "panic: %s\n", e.what());
        return 4;
    }
}
