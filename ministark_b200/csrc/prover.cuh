#pragma once
#include <utility>
#include "common.cuh"
#include "field.cuh"
namespace ms {
struct StarkDerived { uint64_t rounds, constrain_queries, fri_queries; };
struct ProverState { std::vector<std::pair<const char*, float>> timings; };
inline int stark_derive(int, const ms_stark_params&, StarkDerived*) { return MS_ERR_UNSUPPORTED; }
template <class F> int stark_prove(Ctx* c, ProverState*, const ms_stark_params&, const void*, const void*, uint64_t, uint64_t, const typename F::T*, uint64_t, uint8_t*, uint64_t*) { return fail(c, MS_ERR_UNSUPPORTED, "not built yet"); }
}
