#define MS_NTT_NO_HOST
#include "../../ministark_b200/csrc/ntt.cuh"
using namespace ms;
extern "C" __global__ void bf(uint64_t* x, const uint64_t* w) {
    uint64_t v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = x[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 8; i++) { uint64_t t = Fast<GL>::mul(v[i + 8], w[i]); uint64_t A = v[i]; v[i] = Fast<GL>::add(A, t); v[i + 8] = Fast<GL>::sub(A, t); }
#pragma unroll
    for (int i = 0; i < 16; i++) x[threadIdx.x + 32 * i] = v[i];
}
