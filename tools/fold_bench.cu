// Register-resident throughput of the fold step lo + r (hi - lo) on the two pipes (development aid, run under gpurun):
//   int   : fold2 on the integer multiplier (fr_mul_const, constants in the constant bank)
//   f64u  : fold2_f64, constants re-read through the uniform datapath every iteration (what the round kernels do)
//   f64s  : fold2_f64, constants in shared memory (LDS broadcast)
//   f64r  : fold2_f64, constants hoisted into registers by the compiler (one CTA per SM)
//   mix   : one int fold and one f64u fold per iteration (independent chains)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gkr_b200/csrc -o build/fold_bench tools/fold_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include "fr_f64.cuh"

constexpr int kThreads = 256;
template <class KT>
__device__ __forceinline__ Fr fold2(const Fr &lo, const Fr &hi, const KT &r) { return fr_add(lo, fr_mul_const(fr_sub(hi, lo), r)); }

struct SmemK { const double (*c)[12]; };

template <int MODE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_fold(Fr *out, int iters, const __grid_constant__ FrConstMul r,
                                                         const __grid_constant__ FrFoldF64 rf) {
    __shared__ FrFoldF64 srf;
    if (MODE == 2 || MODE == 5 || MODE == 6) {
        for (int i = threadIdx.x; i < 11 * 12; i += blockDim.x) (&srf.c[0][0])[i] = (&rf.c[0][0])[i];
        __syncthreads();
    }
    Fr x = fr_one(), y = fr_one(), u = fr_one(), v = fr_one();
    x.l[0] ^= threadIdx.x * 2654435761u; y.l[1] ^= blockIdx.x + 3u; u.l[2] ^= threadIdx.x + 11u; v.l[3] ^= blockIdx.x * 7u + 1u;
    x = fr_mul(x, y); y = fr_mul(y, x); u = fr_mul(u, x); v = fr_mul(v, y);
    for (int it = 0; it < iters; ++it) {
        const FrFoldF64 &rfz = *reinterpret_cast<const FrFoldF64 *>(reinterpret_cast<const char *>(&rf) + (size_t)(it >> 30) * 16);
        if (MODE == 0) { x = fold2(x, y, r); y = fold2(y, x, r); }
        if (MODE == 1) { x = fold2_f64(x, y, rfz); y = fold2_f64(y, x, rfz); }
        if (MODE == 2) { const FoldKSmem ks{(uint32_t)__cvta_generic_to_shared(&srf)}; x = fold2_f64_src(x, y, ks); y = fold2_f64_src(y, x, ks); }
        if (MODE == 5) { const FoldKSmem ks{(uint32_t)__cvta_generic_to_shared(&srf)}; x = fold2(x, y, r); u = fold2_f64_src(u, v, ks); y = fold2(y, x, r); v = fold2_f64_src(v, u, ks); }
        if (MODE == 6) { const FoldKSmem ks{(uint32_t)__cvta_generic_to_shared(&srf)}; x = fold2(x, y, r); u = fold2_f64_src(u, v, ks); y = fold2(y, x, r); v = fold2(v, u, r); }
        if (MODE == 3) { x = fold2_f64(x, y, rf); y = fold2_f64(y, x, rf); }
        if (MODE == 4) { x = fold2(x, y, r); u = fold2_f64(u, v, rfz); y = fold2(y, x, r); v = fold2_f64(v, u, rfz); }
    }
    Fr acc = fr_add(fr_add(x, y), fr_add(u, v));
    if (acc.l[7] == 0xffffffffu) out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = acc;
}

int main() {
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    Fr *out;
    cudaMalloc(&out, sizeof(Fr) * 256 * 148 * 8);
    FrConstMul r;
    FrFoldF64 rf;
    memset(&rf, 0, sizeof rf);
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 8; ++i) r.c[j][i] = 0x9e3779b9u * (8 * j + i + 1) & (i == 7 ? 0x0fffffffu : 0xffffffffu);
    for (int i = 0; i < 11; ++i) for (int j = 0; j < 11; ++j) rf.c[i][j] = (double)((int)((0x9e3779b9u * (11 * i + j + 1)) >> 9) - (1 << 22)) * (j == 10 ? 1.0 / 2048 : 1.0);
    for (int i = 0; i < 11; ++i) rf.c[i][10] = (double)(long)(rf.c[i][10]);
    const int iters = 2000;
    auto run = [&](const char *name, auto kern, int ctas, int folds_per_iter) {
        const int grid = n_sm * ctas;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        kern<<<grid, kThreads>>>(out, 50, r, rf);
        cudaEventRecord(e0);
        kern<<<grid, kThreads>>>(out, iters, r, rf);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"fold\":\"%s\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"gfolds_per_s\":%.2f,\"err\":\"%s\"}\n", name, ctas, ms,
               (double)grid * kThreads * iters * folds_per_iter / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    run("int", k_fold<0, 2>, 2, 2);
    run("int", k_fold<0, 4>, 4, 2);
    run("f64u", k_fold<1, 2>, 2, 2);
    run("f64u", k_fold<1, 4>, 4, 2);
    run("f64s", k_fold<2, 2>, 2, 2);
    run("f64s", k_fold<2, 4>, 4, 2);
    run("f64r", k_fold<3, 1>, 1, 2);
    run("mix(int+f64u)", k_fold<4, 2>, 2, 4);
    run("mix(2 int + 2 f64s)", k_fold<5, 2>, 2, 4);
    run("mix(3 int + 1 f64s)", k_fold<6, 2>, 2, 4);
    return 0;
}
