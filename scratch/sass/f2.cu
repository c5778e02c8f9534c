#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
#define EPS 0xFFFFFFFFu
// ---- variant A: C with halves
__device__ __forceinline__ u64 mulA(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u64 p00 = (u64)a0 * b0;
    u64 t = (u64)a0 * b1 + (p00 >> 32);
    u64 u = (u64)a1 * b0 + (u32)t;
    u64 v = (u64)a1 * b1 + (t >> 32) + (u >> 32);
    u64 lo = (u64)(u32)p00 | (u << 32);
    u32 c2 = (u32)v, c3 = (u32)(v >> 32);
    u64 x = (u64)c2 * EPS;
    u64 s = lo + x;
    if (s < x) s += EPS;
    u64 r = s - c3;
    if (s < c3) r -= EPS;
    return r;
}
// ---- variant B: PTX carry chains
__device__ __forceinline__ u64 mulB(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 c0, c1, c2, c3;
    asm("{\n\t"
        ".reg .u32 t1, t2;\n\t"
        "mul.lo.u32 %0, %4, %6;\n\t"
        "mul.hi.u32 %1, %4, %6;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.u32 %2, %4, %7, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, 0, 0;\n\t"
        "mad.lo.cc.u32 %2, %5, %7, %2;\n\t"
        "madc.hi.u32 %3, %5, %7, %3;\n\t"
        "}" : "=&r"(c0), "=&r"(c1), "=&r"(c2), "=&r"(c3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    // r = (c1:c0) + c2*EPS - c3
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 m, k;\n\t"
        "sub.cc.u32 %0, %2, %5;\n\t"     // lo - c3
        "subc.cc.u32 %1, %3, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"          // m = borrow ? 0xFFFFFFFF : 0
        "sub.cc.u32 %0, %0, m;\n\t"      // -= EPS if borrow
        "subc.u32 %1, %1, 0;\n\t"
        "sub.cc.u32 %0, %0, %4;\n\t"     // - c2
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        "add.cc.u32 %1, %1, %4;\n\t"     // + c2<<32
        "addc.u32 k, 0, 0;\n\t"
        "sub.u32 k, 0, k;\n\t"          // k = carry ? EPS : 0
        "add.cc.u32 %0, %0, k;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}" : "=&r"(r0), "=&r"(r1) : "r"(c0), "r"(c1), "r"(c2), "r"(c3));
    return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ u64 addL(u64 a, u64 b) { u64 s = a + b; if (s < a) s += EPS; return s; }
__device__ __forceinline__ u64 subL(u64 a, u64 b) { u64 d = a - b; if (a < b) d -= EPS; return d; }
__device__ __forceinline__ u64 addP(u64 a, u64 b) {
    u32 r0, r1;
    asm("{\n\t.reg .u32 k;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "sub.u32 k, 0, k;\n\t"
        "add.cc.u32 %0, %0, k;\n\t"
        "addc.u32 %1, %1, 0;\n\t}"
        : "=&r"(r0), "=&r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ u64 subP(u64 a, u64 b) {
    u32 r0, r1;
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t}"
        : "=&r"(r0), "=&r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return ((u64)r1 << 32) | r0;
}
#define KERN(name, MUL, ADD, SUB) \
extern "C" __global__ void name(u64* x, const u64* w) { \
    u64 v[16]; \
    _Pragma("unroll") for (int i = 0; i < 16; i++) v[i] = x[threadIdx.x + 32 * i]; \
    _Pragma("unroll") for (int i = 0; i < 8; i++) { \
        u64 t = MUL(v[i + 8], w[i]); \
        u64 A = v[i]; \
        v[i] = ADD(A, t); v[i + 8] = SUB(A, t); } \
    _Pragma("unroll") for (int i = 0; i < 16; i++) x[threadIdx.x + 32 * i] = v[i]; \
}
KERN(bfA, mulA, addL, subL)
KERN(bfB, mulB, addP, subP)
KERN(bfAP, mulA, addP, subP)
