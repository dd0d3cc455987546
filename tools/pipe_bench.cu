// Pipe micro-benchmark for sm_100a.  Question: is the FP64 pipe a second multiplier that runs
// beside the half-rate 32x32->64 integer multiplier (IMAD.WIDE, FMA-heavy pipe), and what does the ALU pipe add?
// Workload per thread and iteration: NMUL Montgomery products (fr.cuh: 137 IMAD.WIDE each, two independent chains)
// + NDFMA DFMAs (8 independent chains) + NIADD IADD3s (4 independent chains) + NCVT I2F.F64.U32, all in one loop body
// so that the warp schedulers see them side by side.  Times are clock64() cycles per CTA (independent of DVFS).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gkr_b200/csrc -o build/pipe_bench tools/pipe_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "fr.cuh"

constexpr int kThreads = 256;

template <int NMUL, int NDFMA, int NIADD, int NCVT>
__global__ void __launch_bounds__(kThreads) k_pipe(Fr *out, int iters, uint32_t seed, long long *cycles) {
    Fr x0 = fr_one(), x1 = fr_one(), y = fr_one();
    x0.l[0] ^= threadIdx.x * 2654435761u; x1.l[1] ^= seed + threadIdx.x; y.l[2] ^= blockIdx.x + 77u;
    double f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 1.0 + 0.125 * j + 1e-3 * threadIdx.x;
    const double fm = 1.0000001 + 1e-9 * seed, fa = 0.5;
    uint32_t i0 = seed + threadIdx.x, i1 = seed * threadIdx.x + 1, i2 = seed + 2 * threadIdx.x, i3 = seed ^ threadIdx.x, a = (seed | 1u) + 3 * threadIdx.x, b = threadIdx.x * 7 + 1;
    uint32_t c0 = b + 5, c1 = b + 9;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < (NMUL > 0 ? NMUL : 1); ++u) {
            if (NMUL > 0) { if (u & 1) x1 = fr_mul(x1, y); else x0 = fr_mul(x0, y); }
            constexpr int per = NMUL > 0 ? NDFMA / NMUL : NDFMA;
#pragma unroll
            for (int d = 0; d < per; ++d)
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[d & 7]) : "d"(fm), "d"(fa));
            constexpr int peri = NMUL > 0 ? NIADD / NMUL : NIADD;
#pragma unroll
            for (int d = 0; d < peri; ++d) {
                // alternating add / xor on four thread-dependent chains: cannot be merged into 3-input adds
                if ((d & 7) == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(i0) : "r"(a));
                if ((d & 7) == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(i1) : "r"(b));
                if ((d & 7) == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(i2) : "r"(a));
                if ((d & 7) == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(i3) : "r"(b));
                if ((d & 7) == 4) asm volatile("xor.b32 %0, %0, %1;" : "+r"(i0) : "r"(b));
                if ((d & 7) == 5) asm volatile("xor.b32 %0, %0, %1;" : "+r"(i1) : "r"(a));
                if ((d & 7) == 6) asm volatile("xor.b32 %0, %0, %1;" : "+r"(i2) : "r"(b));
                if ((d & 7) == 7) asm volatile("xor.b32 %0, %0, %1;" : "+r"(i3) : "r"(a));
            }
            constexpr int perc = NMUL > 0 ? NCVT / NMUL : NCVT;
#pragma unroll
            for (int d = 0; d < perc; ++d) {
                double g;
                if (d & 1) { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(g) : "r"(c1)); c1 += (uint32_t)__double2hiint(g); }
                else { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(g) : "r"(c0)); c0 += (uint32_t)__double2hiint(g); }
            }
        }
    }
    const long long t1 = clock64();
    Fr r = fr_add(x0, x1);
    double fs = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) fs += f[j];
    r.l[0] ^= i0 ^ i1 ^ i2 ^ i3 ^ c0 ^ c1 ^ (uint32_t)__double2loint(fs);
    if (r.l[7] == 0xffffffffu) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int g_sm = 0;
static Fr *g_out;
static long long *g_cyc;
template <int NMUL, int NDFMA, int NIADD, int NCVT>
static void run(const char *name, int ctas_per_sm) {
    const int grid = g_sm * ctas_per_sm, iters = 1500;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_pipe<NMUL, NDFMA, NIADD, NCVT><<<grid, kThreads>>>(g_out, 32, 1u, g_cyc);
    cudaEventRecord(e0);
    k_pipe<NMUL, NDFMA, NIADD, NCVT><<<grid, kThreads>>>(g_out, iters, 1u, g_cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    static long long cyc[8192];
    cudaMemcpy(cyc, g_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += (double)cyc[i];
    avg /= grid;
    const double thr = (double)kThreads * ctas_per_sm * iters / avg;      // thread-iterations per clk per SM
    printf("{\"mix\":\"%s\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"cycles\":%.0f,\"per_clk_sm\":{\"imad_wide\":%.1f,\"dfma\":%.1f,"
           "\"alu\":%.1f,\"i2f64\":%.1f},\"cycles_per_iter_per_warp_sched\":%.1f,\"eff_ghz\":%.3f}\n",
           name, ctas_per_sm, ms, avg, 137.0 * NMUL * thr, (double)NDFMA * thr, (double)NIADD * thr, (double)NCVT * thr,
           avg / iters / (kThreads / 32 * ctas_per_sm / 4.0), avg / (ms * 1e6));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    cudaDeviceGetAttribute(&g_sm, cudaDevAttrMultiProcessorCount, 0);
    cudaMalloc(&g_out, sizeof(Fr) * 256 * 148 * 8);
    cudaMalloc(&g_cyc, sizeof(long long) * 8192);
    for (int c : {2, 4}) {
        run<2, 0, 0, 0>("fr_mul", c);
        run<0, 256, 0, 0>("dfma", c);
        run<0, 0, 256, 0>("alu", c);
        run<0, 0, 0, 64>("i2f64", c);
        run<2, 128, 0, 0>("fr_mul+dfma(64/mul)", c);
        run<2, 274, 0, 0>("fr_mul+dfma(137/mul)", c);
        run<2, 548, 0, 0>("fr_mul+dfma(274/mul)", c);
        run<2, 0, 274, 0>("fr_mul+alu(137/mul)", c);
        run<2, 0, 548, 0>("fr_mul+alu(274/mul)", c);
        run<2, 274, 274, 0>("fr_mul+dfma(137)+alu(137)", c);
        run<2, 548, 548, 0>("fr_mul+dfma(274)+alu(274)", c);
        run<0, 256, 256, 0>("dfma+alu", c);
        run<0, 256, 0, 32>("dfma+i2f64(1/8)", c);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
