// Exclusive prefix sum int32 -> int64 used to turn set sizes into SpG row pointers
// (replaces the serial ncumsum loop of subg_acc/subg_acc.c:848-851 and scipy's
// COO->CSR row counting in sampler/random_walks.py:79).
//   out[j] = carry + sum_{i<j} in[i],  j = 0..n      (n+1 outputs)
// Three small launches: per-block totals, scan of the totals by one block,
// per-block rescan with the block offset.  4096 inputs per block.
#pragma once
#include "common.cuh"

namespace subg {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ long long block_excl_scan(long long v, long long *total, long long *warp_sums) {
    // inclusive scan inside the warp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long long t = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += t;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
        long long s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long t = __shfl_up_sync(FULL, s, d);
            if (lane >= d) s += t;
        }
        warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    const long long before = w ? warp_sums[w - 1] : 0;
    if (total) *total = warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
    return before + x - v;
}

static __global__ void scan_block_totals(const int32_t *in, int64_t n, long long *block_tot) {
    __shared__ long long ws[32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    long long s = 0;
    for (int k = 0; k < kScanItems; k++) {
        const int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (i < n) s += in[i];
    }
    long long tot;
    block_excl_scan(s, &tot, ws);
    if (threadIdx.x == 0) block_tot[blockIdx.x] = tot;
}

static __global__ void scan_totals_inplace(long long *block_tot, int nblocks, long long carry) {
    __shared__ long long ws[32];
    __shared__ long long run;
    if (threadIdx.x == 0) run = carry;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        const long long v = b < nblocks ? block_tot[b] : 0;
        long long tot;
        const long long ex = block_excl_scan(v, &tot, ws);
        if (b < nblocks) block_tot[b] = run + ex;
        __syncthreads();
        if (threadIdx.x == 0) run += tot;
        __syncthreads();
    }
}

static __global__ void scan_apply(const int32_t *in, int64_t n, const long long *block_off, long long *out) {
    __shared__ long long ws[32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    // thread owns kScanItems consecutive inputs so the per-thread partials stay in registers
    const int64_t first = base + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    long long s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        const int64_t i = first + k;
        v[k] = i < n ? in[i] : 0;
        s += v[k];
    }
    long long ex = block_excl_scan(s, nullptr, ws) + block_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        const int64_t i = first + k;
        if (i < n) out[i] = ex;
        ex += v[k];
        if (i == n - 1) out[n] = ex;
    }
}

// scratch: at least ceil(n / kScanTile) int64
inline int scan_num_blocks(int64_t n) { return (int)((n + kScanTile - 1) / kScanTile); }

inline cudaError_t exclusive_scan_i32_i64(const int32_t *in, long long *out, int64_t n, long long carry,
                                          long long *scratch, cudaStream_t st) {
    if (n <= 0) {
        return cudaMemcpyAsync(out, &carry, sizeof(long long), cudaMemcpyHostToDevice, st);
    }
    const int nb = scan_num_blocks(n);
    scan_block_totals<<<nb, kScanThreads, 0, st>>>(in, n, scratch);
    scan_totals_inplace<<<1, 1024, 0, st>>>(scratch, nb, carry);
    scan_apply<<<nb, kScanThreads, 0, st>>>(in, n, scratch, out);
    return cudaGetLastError();
}

}  // namespace subg
