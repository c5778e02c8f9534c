#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
__device__ __forceinline__ u64 pack(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
// sub with FMA-pipe fix
__device__ __forceinline__ u64 subF(u64 a, u64 b) {
    u64 r;
    asm("{\n\t.reg .u32 m, s0, s1; .reg .u64 d;\n\t"
        "sub.cc.u32 s0, %1, %3;\n\t"
        "subc.cc.u32 s1, %2, %4;\n\t"
        "subc.u32 m, 0, 0;\n\t"          // m = -borrow
        "add.u32 s1, s1, m;\n\t"         // -borrow * 2^32
        "mov.b64 d, {s0, s1};\n\t"
        "mad.wide.s32 %0, m, -1, d;\n\t" // + borrow
        "}" : "=l"(r) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return r;
}
__device__ __forceinline__ u64 addF(u64 a, u64 b) {
    u64 r;
    asm("{\n\t.reg .u32 k, s0, s1; .reg .u64 d;\n\t"
        "add.cc.u32 s0, %1, %3;\n\t"
        "addc.cc.u32 s1, %2, %4;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "mov.b64 d, {s0, s1};\n\t"
        "mad.wide.u32 %0, k, 0xFFFFFFFF, d;\n\t"
        "}" : "=l"(r) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return r;
}
__device__ __forceinline__ u64 canonF(u64 a) {
    u64 r;
    asm("{\n\t.reg .u32 k, d;\n\t"
        "add.cc.u32 d, %1, 0xFFFFFFFF;\n\t"
        "addc.cc.u32 d, %2, 0;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "mad.wide.u32 %0, k, 0xFFFFFFFF, %3;\n\t"
        "}" : "=l"(r) : "r"((u32)a), "r"((u32)(a >> 32)), "l"(a));
    return r;
}
extern "C" __global__ void kS(u64* x, const u64* w) {
    u64 v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = x[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 8; i++) { u64 A = v[i], t = v[i+8]; v[i] = addF(A, t); v[i+8] = subF(A, t); }
#pragma unroll
    for (int i = 0; i < 16; i++) x[threadIdx.x + 32 * i] = canonF(v[i]);
}
