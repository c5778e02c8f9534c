#include <cstdint>
typedef uint32_t u32;
__device__ __forceinline__ u32 rotr_fma(u32 x, int n) {
    u32 r, c = 1u << (32 - n);
    asm("{\n\t.reg .u32 t;\n\tmul.hi.u32 t, %1, %2;\n\tmad.lo.u32 %0, %1, %2, t;\n\t}" : "=r"(r) : "r"(x), "r"(c));
    return r;
}
extern "C" __global__ void kr(u32* x) {
    u32 a = x[threadIdx.x];
    u32 s = rotr_fma(a, 6) ^ rotr_fma(a, 11) ^ rotr_fma(a, 25);
    x[threadIdx.x] = s;
}
