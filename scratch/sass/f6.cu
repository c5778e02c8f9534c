#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
__device__ __forceinline__ u64 pack(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
__device__ __forceinline__ u64 mulM(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 c0, c1, c2, c3, m, d;\n\t"
        ".reg .u64 pp;\n\t"
        "mul.wide.u32 pp, %2, %4;\n\t"
        "mov.b64 {c0, c1}, pp;\n\t"
        "mad.lo.cc.u32 c1, %2, %5, c1;\n\t"
        "madc.hi.u32 c2, %2, %5, 0;\n\t"
        "mad.lo.cc.u32 c1, %3, %4, c1;\n\t"
        "madc.hi.cc.u32 c2, %3, %4, c2;\n\t"
        "addc.u32 c3, 0, 0;\n\t"
        "mad.lo.cc.u32 c2, %3, %5, c2;\n\t"
        "madc.hi.u32 c3, %3, %5, c3;\n\t"
        // t = (c1:c0) - c3, borrow folded: t -= EPS
        "sub.cc.u32 c0, c0, c3;\n\t"
        "subc.cc.u32 c1, c1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 c0, c0, m;\n\t"
        "subc.u32 c1, c1, 0;\n\t"
        // r = t + c2*EPS, carry -> mask
        "mad.lo.cc.u32 c0, c2, 0xFFFFFFFF, c0;\n\t"
        "madc.hi.cc.u32 c1, c2, 0xFFFFFFFF, c1;\n\t"
        "subc.u32 m, 0, 0;\n\t"           // m = -carry
        // carry of r + EPS (r >= p), accumulate into mask:  m |= -(carry')
        "add.cc.u32 d, c0, 0xFFFFFFFF;\n\t"
        "addc.cc.u32 d, c1, 0;\n\t"
        "subc.u32 d, 0, 0;\n\t"
        "or.b32 m, m, d;\n\t"
        "add.cc.u32 %0, c0, m;\n\t"
        "addc.u32 %1, c1, 0;\n\t"
        "}" : "=r"(r0), "=r"(r1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack(r0, r1);
}
__device__ __forceinline__ u64 addM(u64 a, u64 b) {
    u32 r0, r1;
    asm("{\n\t.reg .u32 m, s0, s1;\n\t"
        "add.cc.u32 s0, %2, %4;\n\t"
        "addc.cc.u32 s1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "add.cc.u32 %0, s0, m;\n\t"
        "addc.u32 %1, s1, 0;\n\t}"
        : "=r"(r0), "=r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return pack(r0, r1);
}
__device__ __forceinline__ u64 subM(u64 a, u64 b) {
    u32 r0, r1;
    asm("{\n\t.reg .u32 m, s0, s1;\n\t"
        "sub.cc.u32 s0, %2, %4;\n\t"
        "subc.cc.u32 s1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, s0, m;\n\t"
        "subc.u32 %1, s1, 0;\n\t}"
        : "=r"(r0), "=r"(r1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return pack(r0, r1);
}
extern "C" __global__ void bfM(u64* x, const u64* w) {
    u64 v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = x[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 8; i++) { u64 t = mulM(v[i + 8], w[i]); u64 A = v[i]; v[i] = addM(A, t); v[i + 8] = subM(A, t); }
#pragma unroll
    for (int i = 0; i < 16; i++) x[threadIdx.x + 32 * i] = v[i];
}
extern "C" __global__ void test_ops(const u64* a, const u64* b, u64* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[3 * i] = mulM(a[i], b[i]); out[3 * i + 1] = addM(a[i], out[3 * i]); out[3 * i + 2] = subM(a[i], out[3 * i]);
}
