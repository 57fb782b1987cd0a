// Edge list -> CSR graph on the device.
//
// Reference behaviour reproduced (file:line relative to /root/reference):
//   edge2csr: csr_matrix((ones(bool), (row, col)), shape=(max+1, max+1))      subg_acc/test/test.py:15-19
//             -> duplicate edges coalesced, columns ascending per row (scipy canonical CSR)
//   symmetrisation of the training graph (to_undirected / G + G.T)             dataloader.py:119-129
//
// One 64-bit key (row << 32 | col) per directed edge, CUB radix sort over the bits in use, unique, then
// the row pointer is read off the sorted keys (every key fills the pointers of the rows that start at
// it).  The row pointer is 64-bit when the coalesced graph has >= 2^31 entries (twitter-2010 shape).
#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"
#include "scan.cuh"

namespace subg {

namespace {

__global__ void edge_range_kernel(const long long *row, const long long *col, int64_t E, long long *minmax) {
    long long lo = INT64_MAX, hi = INT64_MIN;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
        const long long r = row[i], c = col[i];
        lo = min(lo, min(r, c));
        hi = max(hi, max(r, c));
    }
    for (int d = 16; d; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(FULL, lo, d));
        hi = max(hi, __shfl_xor_sync(FULL, hi, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(minmax, lo);
        atomicMax(minmax + 1, hi);
    }
}

// dropped edges (self loops when asked) become the sentinel key (row = N), which sorts behind every row
__global__ void edge_keys_kernel(const long long *row, const long long *col, int64_t E, int64_t N, int sym, int drop_self,
                                 unsigned long long *keys) {
    const unsigned long long sentinel = (unsigned long long)N << 32;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long r = (unsigned long long)row[i], c = (unsigned long long)col[i];
        const bool drop = drop_self && r == c;
        keys[i] = drop ? sentinel : (r << 32 | c);
        if (sym) keys[E + i] = drop ? sentinel : (c << 32 | r);
    }
}

// Unique over the sorted keys, sentinel dropped; 64-bit item counts (tile = 256 threads x 16 consecutive keys).
// Pass 1 counts the run heads of every tile, the tile counts are scanned, pass 2 writes.
constexpr int kUqItems = 16;
template <bool WRITE>
__global__ void __launch_bounds__(256) unique_keys_kernel(const unsigned long long *keys, int64_t K, int64_t N, int32_t *tile_count,
                                                          const long long *tile_off, unsigned long long *out) {
    __shared__ long long ws[32];
    const int64_t first = ((int64_t)blockIdx.x * 256 + threadIdx.x) * kUqItems;
    unsigned long long k[kUqItems];
    unsigned long long prev = first > 0 && first <= K ? keys[first - 1] : ~0ull;
    int cnt = 0;
    uint32_t flags = 0;
#pragma unroll
    for (int j = 0; j < kUqItems; j++) {
        const int64_t i = first + j;
        k[j] = i < K ? keys[i] : ~0ull;
        const bool head = i < K && (i == 0 || k[j] != prev) && (int64_t)(k[j] >> 32) < N;
        prev = k[j];
        flags |= head ? 1u << j : 0u;
        cnt += head ? 1 : 0;
    }
    long long total;
    const long long ex = block_excl_scan((long long)cnt, &total, ws);
    if (!WRITE) {
        if (threadIdx.x == 0) tile_count[blockIdx.x] = (int32_t)total;
        return;
    }
    long long o = tile_off[blockIdx.x] + ex;
#pragma unroll
    for (int j = 0; j < kUqItems; j++)
        if (flags >> j & 1u) out[o++] = k[j];
}

template <typename P>
__global__ void csr_from_keys_kernel(const unsigned long long *keys, int64_t Eu, int64_t N, P *rowptr, int32_t *col) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= Eu; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i < Eu ? (int64_t)(keys[i] >> 32) : N;       // row that starts (or continues) at i
        const int64_t rprev = i ? (int64_t)(keys[i - 1] >> 32) : -1;
        for (int64_t rr = rprev + 1; rr <= r; rr++) rowptr[rr] = (P)i;  // rows rprev+1..r begin at i (all but r are empty)
        if (i < Eu) col[i] = (int32_t)(uint32_t)keys[i];
    }
}

template <typename P>
__global__ void rowinfo_from_rowptr_kernel(const P *rowptr, unsigned long long *out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long start = (unsigned long long)rowptr[i];
        unsigned long long d = (unsigned long long)(rowptr[i + 1] - rowptr[i]);
        if (d > 0xFFFFFFull) d = 0xFFFFFFull;
        out[i] = (d << 40) | start;
    }
}

__global__ void widen_i32_kernel(const int32_t *in, long long *out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace

int graph_from_edges_impl(const int64_t *row_hd, const int64_t *col_hd, int64_t E, int64_t N_in, int symmetrize,
                          int drop_self_loops, int device, cudaStream_t st, Graph **out) {
    if (!out || E < 0 || (E > 0 && (!row_hd || !col_hd))) return fail(SUBG_ERR_ARG, "Input parsing error.");
    DeviceGuard guard(device);
    Graph *g = new Graph();
    g->tag.last = st;
    g->device = device;
    cudaDeviceGetAttribute(&g->num_sms, cudaDevAttrMultiProcessorCount, device);
    long long *d_row = nullptr, *d_col = nullptr, *d_minmax = nullptr, *d_tile_off = nullptr, *d_scratch = nullptr;
    int32_t *d_tile_count = nullptr;
    unsigned long long *k_a = nullptr, *k_b = nullptr;
    void *d_tmp = nullptr;
    int rc = SUBG_OK;
    cudaError_t e = cudaSuccess;
    const int64_t K = E * (symmetrize ? 2 : 1);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((E + 255) / 256, 16 * (int64_t)g->num_sms));
#define IK(call)                                                                                   \
    do {                                                                                           \
        e = (call);                                                                                \
        if (e != cudaSuccess) {                                                                    \
            rc = fail(e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA,               \
                      std::string(#call) + ": " + cudaGetErrorString(e));                          \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)
    {
        const long long *row = (const long long *)row_hd, *col = (const long long *)col_hd;
        if (E > 0 && !is_device_ptr(row_hd)) {
            IK(dmalloc(&d_row, (size_t)E, st));
            IK(cudaMemcpyAsync(d_row, row_hd, (size_t)E * 8, cudaMemcpyHostToDevice, st));
            row = d_row;
        }
        if (E > 0 && !is_device_ptr(col_hd)) {
            IK(dmalloc(&d_col, (size_t)E, st));
            IK(cudaMemcpyAsync(d_col, col_hd, (size_t)E * 8, cudaMemcpyHostToDevice, st));
            col = d_col;
        }
        long long h_minmax[2] = {INT64_MAX, INT64_MIN};
        IK(dmalloc(&d_minmax, 2, st));
        IK(cudaMemcpyAsync(d_minmax, h_minmax, sizeof(h_minmax), cudaMemcpyHostToDevice, st));
        if (E > 0) {
            edge_range_kernel<<<blocks, 256, 0, st>>>(row, col, E, d_minmax);
            count_launch();
            IK(cudaGetLastError());
        }
        IK(cudaMemcpyAsync(h_minmax, d_minmax, sizeof(h_minmax), cudaMemcpyDeviceToHost, st));
        IK(cudaStreamSynchronize(st));
        if (E > 0 && h_minmax[0] < 0) { rc = fail(SUBG_ERR_ARG, "edge list holds a negative node id"); goto done; }
        int64_t N = N_in;
        if (N < 0) N = E > 0 ? h_minmax[1] + 1 : 0;  // nmax + 1 (test.py:18-19)
        if (E > 0 && h_minmax[1] >= N) { rc = fail(SUBG_ERR_ARG, "edge list holds a node id >= num_nodes"); goto done; }
        if (N >= (1ll << 31) - 1) { rc = fail(SUBG_ERR_ARG, "node ids must fit int32"); goto done; }
        g->N = N;

        int64_t Eu = 0;
        if (K > 0) {
            IK(dmalloc(&k_a, (size_t)K, st));
            IK(dmalloc(&k_b, (size_t)K + 1, st));
            edge_keys_kernel<<<blocks, 256, 0, st>>>(row, col, E, N, symmetrize ? 1 : 0, drop_self_loops ? 1 : 0, k_a);
            count_launch();
            IK(cudaGetLastError());
            dfree(d_row, st); dfree(d_col, st);
            d_row = d_col = nullptr;
            int nbits = 1;
            while ((N >> nbits) != 0) nbits++;  // bits that hold N (the sentinel row)
            cub::DoubleBuffer<unsigned long long> db(k_a, k_b);
            size_t tmp_bytes = 0;
            IK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, K, 0, 32 + nbits, st));
            IK(cudaMallocAsync(&d_tmp, std::max<size_t>(tmp_bytes, 16), st));
            IK(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, db, K, 0, 32 + nbits, st));
            unsigned long long *sorted = db.Current(), *other = db.Alternate();
            const int64_t tiles = (K + 256 * kUqItems - 1) / (256 * kUqItems);
            IK(dmalloc(&d_tile_count, (size_t)tiles, st));
            IK(dmalloc(&d_tile_off, (size_t)tiles + 1, st));
            IK(dmalloc(&d_scratch, (size_t)std::max(1, scan_num_blocks(tiles)), st));
            unique_keys_kernel<false><<<(unsigned)tiles, 256, 0, st>>>(sorted, K, N, d_tile_count, nullptr, nullptr);
            IK(cudaGetLastError());
            IK(exclusive_scan_i32_i64(d_tile_count, d_tile_off, tiles, 0, d_scratch, st));
            unique_keys_kernel<true><<<(unsigned)tiles, 256, 0, st>>>(sorted, K, N, nullptr, d_tile_off, other);
            IK(cudaGetLastError());
            count_launch(10);
            long long h_count = 0;
            IK(cudaMemcpyAsync(&h_count, d_tile_off + tiles, sizeof(h_count), cudaMemcpyDeviceToHost, st));
            IK(cudaStreamSynchronize(st));
            Eu = h_count;
            if (Eu >= (1ll << 40)) { rc = fail(SUBG_ERR_ARG, "graphs of 2^40 entries or more are not supported"); goto done; }
            g->E = Eu;
            g->rowptr64 = Eu >= (1ll << 31);
            IK(cudaMallocAsync(&g->rowptr, ((size_t)N + 2) * (g->rowptr64 ? 8 : 4), st));
            IK(cudaMallocAsync((void **)&g->col, ((size_t)Eu + 16) * 4, st));
            IK(cudaMallocAsync(&g->rowinfo, ((size_t)N + 1) * 8, st));
            const unsigned cb = (unsigned)std::max<int64_t>(1, std::min<int64_t>((Eu + 256) / 256, 16 * (int64_t)g->num_sms));
            if (g->rowptr64) csr_from_keys_kernel<long long><<<cb, 256, 0, st>>>(other, Eu, N, (long long *)g->rowptr, g->col);
            else csr_from_keys_kernel<int32_t><<<cb, 256, 0, st>>>(other, Eu, N, (int32_t *)g->rowptr, g->col);
            count_launch();
            IK(cudaGetLastError());
        } else {
            g->E = 0;
            g->rowptr64 = false;
            IK(cudaMallocAsync(&g->rowptr, ((size_t)N + 2) * 4, st));
            IK(cudaMallocAsync((void **)&g->col, 16 * 4, st));
            IK(cudaMallocAsync(&g->rowinfo, ((size_t)N + 1) * 8, st));
            IK(cudaMemsetAsync(g->rowptr, 0, ((size_t)N + 2) * 4, st));
        }
        if (N > 0) {
            const unsigned rb = (unsigned)std::min<int64_t>((N + 255) / 256, 4 * (int64_t)g->num_sms);
            if (g->rowptr64) rowinfo_from_rowptr_kernel<long long><<<rb, 256, 0, st>>>((const long long *)g->rowptr, (unsigned long long *)g->rowinfo, N);
            else rowinfo_from_rowptr_kernel<int32_t><<<rb, 256, 0, st>>>((const int32_t *)g->rowptr, (unsigned long long *)g->rowinfo, N);
            count_launch();
            IK(cudaGetLastError());
        }
        g->sorted_state = 1;
        if (symmetrize) g->sym_state = 1;
        IK(cudaStreamSynchronize(st));
    }
done:
#undef IK
    dfree(d_row, st); dfree(d_col, st); dfree(d_minmax, st); dfree(d_tile_count, st); dfree(d_tile_off, st); dfree(d_scratch, st); dfree(k_a, st); dfree(k_b, st); dfree(d_tmp, st);
    if (rc != SUBG_OK) {
        if (g->rowptr) cudaFreeAsync(g->rowptr, st);
        if (g->col) cudaFreeAsync(g->col, st);
        if (g->rowinfo) cudaFreeAsync(g->rowinfo, st);
        delete g;
        return rc;
    }
    *out = g;
    return SUBG_OK;
}

// CSR back to the caller: rowptr as int64[N+1] (whatever the width in HBM), col int32[E]; host or device.
int graph_export_impl(const Graph *g, int64_t *rowptr_hd, int32_t *col_hd, cudaStream_t st) {
    if (!g) return fail(SUBG_ERR_ARG, "null graph");
    DeviceGuard guard(g->device);
    g->tag.use_on(st);
    long long *wide = nullptr;
    if (rowptr_hd) {
        const void *src = g->rowptr;
        if (!g->rowptr64) {
            SUBG_CUDA(dmalloc(&wide, (size_t)g->N + 1, st));
            const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((g->N + 256) / 256, 4 * (int64_t)g->num_sms));
            widen_i32_kernel<<<blocks, 256, 0, st>>>((const int32_t *)g->rowptr, wide, g->N + 1);
            count_launch();
            src = wide;
        }
        SUBG_CUDA(cudaMemcpyAsync(rowptr_hd, src, ((size_t)g->N + 1) * 8, cudaMemcpyDefault, st));
    }
    if (col_hd && g->E > 0) SUBG_CUDA(cudaMemcpyAsync(col_hd, g->col, (size_t)g->E * 4, cudaMemcpyDefault, st));
    SUBG_CUDA(cudaStreamSynchronize(st));
    dfree(wide, st);
    return SUBG_OK;
}

}  // namespace subg
