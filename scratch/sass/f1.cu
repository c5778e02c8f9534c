#include <cstdint>
#include "../../ministark_b200/csrc/field.cuh"
using namespace ms;
extern "C" __global__ void k_mul(uint64_t* a, const uint64_t* b) {
    int i = threadIdx.x;
    a[i] = GL::mul(a[i], b[i]);
}
extern "C" __global__ void k_add(uint64_t* a, const uint64_t* b) {
    int i = threadIdx.x;
    a[i] = GL::add(a[i], b[i]);
}
extern "C" __global__ void k_sub(uint64_t* a, const uint64_t* b) {
    int i = threadIdx.x;
    a[i] = GL::sub(a[i], b[i]);
}
