// GPU check of Fast<GL>/Fast<BB> (csrc/ntt.cuh) against host 128-bit arithmetic on random and edge inputs.
#define MS_NTT_NO_HOST
#include <cstdio>
#include <vector>
#include "../../ministark_b200/csrc/ntt.cuh"
using namespace ms;
__global__ void k_ops(const uint64_t* a, const uint64_t* b, uint64_t* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t m = Fast<GL>::mul(a[i], b[i]);        // canonical
    out[4 * i] = m;
    out[4 * i + 1] = Fast<GL>::add(a[i], m);       // lazy + canonical
    out[4 * i + 2] = Fast<GL>::sub(a[i], m);
    out[4 * i + 3] = Fast<GL>::canon(a[i]);
}
int main() {
    const uint64_t P = GL::P;
    std::vector<uint64_t> edge = {0, 1, 2, P - 1, P, P + 1, ~0ULL, ~0ULL - 1, 0xFFFFFFFFULL, 0x100000000ULL, 0xFFFFFFFF00000000ULL,
                                  0xFFFFFFFEFFFFFFFFULL, 0x8000000000000000ULL, P - 2, 0xFFFFFFFF00000002ULL};
    std::vector<uint64_t> a, b;
    for (auto x : edge) for (auto y : edge) { a.push_back(x); b.push_back(y); }
    uint64_t s = 88172645463325252ULL;
    for (int i = 0; i < 1 << 20; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint64_t x = s;
        s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint64_t y = s;
        if (i % 7 == 0) x |= 0xFFFFFFFF00000000ULL;
        if (i % 11 == 0) y |= 0xFFFFFFFF00000000ULL;
        if (i % 13 == 0) x &= 0xFFFFFFFFULL;
        a.push_back(x); b.push_back(y);
    }
    int n = (int)a.size();
    uint64_t *da, *db, *dout;
    cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dout, n * 32);
    cudaMemcpy(da, a.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
    k_ops<<<(n + 255) / 256, 256>>>(da, db, dout, n);
    std::vector<uint64_t> out(4 * (size_t)n);
    if (cudaMemcpy(out.data(), dout, n * 32, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 2; }
    int bad = 0;
    for (int i = 0; i < n; i++) {
        unsigned __int128 pr = (unsigned __int128)a[i] * b[i];
        uint64_t m = (uint64_t)(pr % P);
        uint64_t ad = (uint64_t)(((unsigned __int128)a[i] + m) % P), sb = (uint64_t)(((unsigned __int128)a[i] % P + P - m) % P);
        bool ok = out[4 * i] == m && out[4 * i + 1] % P == ad && out[4 * i + 2] % P == sb && out[4 * i + 3] == a[i] % P;
        if (!ok && bad++ < 8)
            printf("MISMATCH a=%016llx b=%016llx mul %016llx want %016llx add %016llx want %016llx sub %016llx want %016llx canon %016llx\n",
                   (unsigned long long)a[i], (unsigned long long)b[i], (unsigned long long)out[4 * i], (unsigned long long)m,
                   (unsigned long long)out[4 * i + 1], (unsigned long long)ad, (unsigned long long)out[4 * i + 2], (unsigned long long)sb,
                   (unsigned long long)out[4 * i + 3]);
    }
    printf("%s: %d cases, %d bad (MS_GL_MASKFIX=%d)\n", bad ? "FAILED" : "OK", n, bad, MS_GL_MASKFIX);
    return bad ? 1 : 0;
}
