// Pure-register version of one Goldilocks NTT round (16 general products + the power-of-two DFT-16 of csrc/ntt.cuh), no
// loads, stores or barriers inside the loop: what the instruction stream of k_ntt_fixed can reach on the integer pipes when
// memory phases are taken away.  Compare its ALU-pipe utilisation (ncu) and time per element-round with the real kernel's.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -o ntt_math ntt_math.cu
//   ./ntt_math            (or under ncu --metrics sm__pipe_alu_cycles_active...,smsp__issue_active...,smsp__inst_executed.sum)
#define MS_NTT_NO_HOST
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../ministark_b200/csrc/ntt.cuh"
using namespace ms;

#ifndef ITERS
#define ITERS 512
#endif

// MODE 0: products + DFT-16 (one full round), 1: DFT-16 only, 2: products only, 3: add/sub butterflies only (no shifts)
template <int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) k_round(uint64_t* out, const uint64_t* __restrict__ tw, uint64_t seed) {
    using A = Fast<GL>;
    uint64_t v[16], w[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        v[r] = (seed * (threadIdx.x + 1 + 256ull * blockIdx.x) + 0x9E3779B97F4A7C15ull * (r + 1)) % GL::P;
        w[r] = tw[(threadIdx.x & 15) * 16 + r];
    }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int r = 1; r < 16; r++) v[r] = A::mul(v[r], w[r]);
        }
        if (MODE == 0 || MODE == 1) gl_shift_dft<4, false>(v);
        if (MODE == 3) {
#pragma unroll
            for (int st = 0; st < 4; st++)
#pragma unroll
                for (int r = 0; r < 16; r++) {
                    if (r & (1 << st)) continue;
                    const uint64_t x = A::canon(v[r | (1 << st)]), y = v[r];
                    v[r] = A::add(y, x);
                    v[r | (1 << st)] = A::sub(y, x);
                }
        }
        if (MODE == 1 || MODE == 3) {  // keep the stage-0 operands canonical as the real round does (they come out of a product)
#pragma unroll
            for (int r = 1; r < 16; r += 2) v[r] = A::canon(v[r]);
        }
    }
    uint64_t acc = 0;
#pragma unroll
    for (int r = 0; r < 16; r++) acc ^= v[r];
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int MODE, int MINB>
static void run(const char* name, int ctas_per_sm) {
    const int blocks = 148 * ctas_per_sm;
    uint64_t *out, *tw;
    cudaMalloc(&out, (size_t)blocks * 256 * 8);
    cudaMalloc(&tw, 256 * 8);
    uint64_t h[256];
    for (int i = 0; i < 256; i++) h[i] = (0x123456789ABCDEFull * (i + 3)) % GL::P;
    cudaMemcpy(tw, h, sizeof h, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_round<MODE, MINB><<<blocks, 256>>>(out, tw, 12345 + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double elem_rounds = (double)blocks * 256 * 16 * ITERS;
    printf("%-44s %d CTA/SM  %.3f ms  %.2f ns per 1000 element-rounds  (%.3e element-rounds/s)\n", name, ctas_per_sm, best,
           best * 1e6 / (elem_rounds / 1000), elem_rounds / (best * 1e-3));
    cudaFree(out);
    cudaFree(tw);
}

int main() {
    run<0, 3>("products + DFT-16 (one round), MINB 3", 3);
    run<0, 2>("products + DFT-16 (one round), MINB 2", 2);
    run<0, 1>("products + DFT-16 (one round), MINB 1", 1);
    run<1, 3>("DFT-16 only", 3);
    run<2, 3>("15 general products only", 3);
    run<3, 3>("add/sub butterflies + canon only", 3);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
