// Shared device helpers of the kernel translation units (kernels.cu, kernels_prod3.cu): 128-bit table access,
// the fold step, command-block polling, reductions, grid sizing.  Internal to the library.
#pragma once
#include <cuda_runtime.h>

#include "fr_f64.cuh"
#include "kernels.cuh"

namespace gkr {


#ifndef GKR_FOLD_PREFETCH
#define GKR_FOLD_PREFETCH 0       // experiment: L2 prefetch of the next iteration in the fused degree-2 rounds too
#endif
#ifndef GKR_LOAD_AHEAD
#define GKR_LOAD_AHEAD 0          // experiment: issue the W and H loads of a fused degree-2 pair before any arithmetic
#endif
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxWarps = 16;     // capacity of the per-warp shared arrays of the reductions (CTAs of up to 512 threads)

// ------------------------------------------------------------------------------------------------
// 128-bit table access
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fr ld_fr(const Fr *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
    r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
// coherent (L2) load for data written earlier in the same kernel by other CTAs
__device__ __forceinline__ Fr ld_fr_cg(const Fr *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
    r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
__device__ __forceinline__ void st_fr(Fr *p, const Fr &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// pull the cache lines of a future iteration towards L2 (no register cost)
__device__ __forceinline__ void prefetch_l2(const Fr *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// lo + r * (hi - lo), r given by its constant-multiplier table (kernel parameter => constant bank operands)
template <class KT>
__device__ __forceinline__ Fr fold2(const Fr &lo, const Fr &hi, const KT &r) {
    return fr_add(lo, fr_mul_const(fr_sub(hi, lo), r));
}

// the same fold on the pipe chosen at compile time: F64 = FP64 pipe (fr_f64.cuh), else the integer multiplier
template <bool F64, class KT>
__device__ __forceinline__ Fr fold_sel(const Fr &lo, const Fr &hi, const KT &r, const FrFoldF64 &rf) {
    if (F64) return fold2_f64(lo, hi, rf);
    return fold2(lo, hi, r);
}

// constant-multiplier table received through a HostCmd (shared memory copy of its 80 raw words)
struct CmdConst {
    const uint32_t *raw;
    __device__ __forceinline__ uint32_t get(int j, int i) const {
        const int p = 8 * j + i;
        return raw[(p / 15) * 16 + (p % 15)];
    }
};
__device__ __forceinline__ uint32_t ld_sys(const volatile uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Warp 0 polls the command block until all five line tags equal `tag` (or an abort tag / timeout shows up),
// then leaves the 80 raw words in shared memory.  Returns false on abort or timeout.
__device__ __forceinline__ bool wait_cmd(const HostCmd *cmd, uint32_t tag, uint32_t *raw_smem, int *ok_smem) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int ok = 0;
        for (uint32_t spin = 0; spin < (1u << 21); ++spin) {       // ~4 s, then give up: the host retries without pre-launching
            const uint32_t v0 = ld_sys(&cmd->w[lane]);
            const uint32_t v1 = ld_sys(&cmd->w[32 + lane]);
            const uint32_t v2 = lane < 16 ? ld_sys(&cmd->w[64 + lane]) : tag;
            const bool is_tag_lane = (lane & 15) == 15;
            const bool good = !is_tag_lane || (v0 == tag && v1 == tag && v2 == tag);
            const bool abort = is_tag_lane && (v0 == kCmdAbort || v1 == kCmdAbort || v2 == kCmdAbort);
            if (__any_sync(0xffffffffu, abort)) break;
            if (__all_sync(0xffffffffu, good)) {
                raw_smem[lane] = v0;
                raw_smem[32 + lane] = v1;
                if (lane < 16) raw_smem[64 + lane] = v2;
                ok = 1;
                break;
            }
            __nanosleep(200);
        }
        if (lane == 0) *ok_smem = ok;
    }
    __syncthreads();
    return *ok_smem != 0;
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fr warp_sum(Fr v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], off);
        v = fr_add(v, o);
    }
    return v;
}

// Sum K accumulators over the CTA; result valid in thread 0.
template <int K>
__device__ __forceinline__ void block_sum(Fr (&acc)[K], Fr (*smem)[kMaxWarps]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        Fr w = warp_sum(acc[j]);
        if (lane == 0) smem[j][warp] = w;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            Fr v = lane < nw ? smem[j][lane] : fr_zero();
            acc[j] = warp_sum(v);
        }
    }
    __syncthreads();
}

// Grid-wide sum of K accumulators; the last CTA to arrive publishes the canonical totals.
// Multi-GPU (xa): the totals go to this rank's entry of the shared exchange row (xa.out; every rank's host adds the
// entries up), or -- fallback exchange -- to xa.dev_out for an NCCL all-gather; the host slot is not written then.
// second half of grid_sum_publish: thread 0 of every CTA holds the CTA totals in acc
template <int K>
__device__ __forceinline__ void grid_publish_cta_totals(Fr (&acc)[K], Fr (*red)[kMaxWarps], Fr *partials, unsigned int *counter,
                                                        HostSlot *slot, uint32_t seq, uint32_t aux0, XchgArg xa = XchgArg{}) {
    __shared__ bool is_last;
    if (gridDim.x > 1) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int j = 0; j < K; ++j) st_fr(&partials[(size_t)blockIdx.x * K + j], acc[j]);
            __threadfence();
            unsigned int ticket = atomicAdd(counter, 1u);
            is_last = (ticket == gridDim.x - 1);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] = fr_zero();
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
            for (int j = 0; j < K; ++j) acc[j] = fr_add(acc[j], ld_fr_cg(&partials[(size_t)b * K + j]));
        }
        block_sum<K>(acc, red);
        if (threadIdx.x == 0) *counter = 0;
    }
    // single-CTA launches (small tables) skip the partials / ticket round trip entirely
    if (threadIdx.x == 0 && xa.dev_out != nullptr) {
#pragma unroll
        for (int j = 0; j < K; ++j) st_fr(&xa.dev_out[j], acc[j]);
    }
    if (xa.dev_out != nullptr) return;
    if (xa.out != nullptr) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int j = 0; j < K; ++j) st_fr(&xa.out->v[j], acc[j]);
            xa.out->aux[0] = aux0;
            __threadfence_system();
            xa.out->flag = xa.seq;
        }
        return;
    }
    if (threadIdx.x == 0) {
        // totals are published in Montgomery form (the host shares the representation); aux[1] = non-zero mask
        uint32_t nz = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            nz |= fr_is_zero(acc[j]) ? 0u : (1u << j);
            st_fr(&slot->v[j], acc[j]);
        }
        slot->aux[0] = aux0;
        slot->aux[1] = nz;
        slot->aux[2] = 0;
        __threadfence_system();
        slot->seq = seq;
    }
}
template <int K>
__device__ __forceinline__ void grid_sum_publish(Fr (&acc)[K], Fr *partials, unsigned int *counter,
                                                 HostSlot *slot, uint32_t seq, uint32_t aux0, XchgArg xa = XchgArg{}) {
    __shared__ Fr red[K][kMaxWarps];
    block_sum<K>(acc, red);
    grid_publish_cta_totals<K>(acc, red, partials, counter, slot, seq, aux0, xa);
}

static inline int grid_for(uint64_t work_items, int max_blocks) {
    uint64_t b = (work_items + kThreads - 1) / kThreads;
    if (b < 1) b = 1;
    if (b > (uint64_t)max_blocks) b = (uint64_t)max_blocks;
    return (int)b;
}

static inline int stream_grid(uint64_t work_items) { return grid_for(work_items, device_sm_count() * 8); }

// lazy accumulation pays once a thread sees several pairs: fewer, fatter CTAs (2 resident per SM)
static inline bool use_lazy(uint64_t pairs) { return pairs >= ((uint64_t)1 << 20); }
static inline int round_grid(uint64_t pairs, const ReduceWs &ws) {
    const int cap = use_lazy(pairs) ? device_sm_count() * 2 : ws.max_blocks;
    return grid_for(pairs, cap < ws.max_blocks ? cap : ws.max_blocks);
}

}  // namespace gkr
