#pragma once
#include "common.cuh"
#include "field.cuh"
namespace ms {
template <class F> int fri_commit(Ctx* c, const typename F::T*, uint64_t, uint64_t, uint64_t, typename F::T*, uint64_t, uint32_t*, uint8_t*) { return fail(c, MS_ERR_UNSUPPORTED, "not built yet"); }
template <class F> int fri_deep_coeffs(Ctx* c, const typename F::T*, uint64_t, uint64_t, const typename F::T*, typename F::T*) { return fail(c, MS_ERR_UNSUPPORTED, "not built yet"); }
template <class F> int fri_fold(Ctx* c, const typename F::T*, uint64_t, uint64_t, const typename F::T*, const typename F::T*, const typename F::T*, typename F::T*, uint64_t) { return fail(c, MS_ERR_UNSUPPORTED, "not built yet"); }
}
