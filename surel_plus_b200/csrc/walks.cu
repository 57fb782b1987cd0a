// SUREL-v1 walk sampler + relative-position encoder on the device.
//
// Reference behaviour reproduced (file:line relative to /root/reference):
//   random_walk      every hop uniform with replacement                subg_acc/subg_acc.c:144-181
//   random_walk_wo   first hop without replacement, then uniform       subg_acc/subg_acc.c:183-248
//   rpe_encoder      per seed: unique nodes in first-visit order of the step-major / walk-minor scan,
//                    rpe[node][step] = walks at that node after `step` hops, rpe[root][0] = M
//                                                                     subg_acc/subg_acc.c:250-314
//   walk_sampler     argument handling and the returned pair           subg_acc/subg_acc.c:316-389
//
// Two kernels.  walk_sample_kernel: one warp per seed, lanes own walks (four at a time so that each
// hop has 4 x 32 gathers in flight), draws from Philox4x32-10 or from the reference's single-thread
// rand_r stream replayed by LCG jump-ahead (call offsets = prefix sum of the per-seed call counts).
// rpe_kernel: one CTA per seed sorts the (node, visit order) keys of the seed's walks in shared memory;
// equal nodes become runs, a bitmap over the visit orders of the run heads ranks the nodes in
// first-visit order, and each head counts its run per step.  It runs twice: sizes, scan, then rows.
#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"
#include "scan.cuh"

namespace subg {

struct WalkSet {
    StreamTag tag;
    int device = 0;
    int64_t n = 0, T = 0;
    int M = 0, m = 0;
    uint32_t status = 0;
    int32_t *walks = nullptr;   // [n, M, m+1]
    long long *off = nullptr;   // [n+1]
    int32_t *ids = nullptr;     // [T]
    int32_t *rpe = nullptr;     // [T, m+1]
};

namespace {

constexpr int kWalkWarps = 4;
constexpr int kWalkGroup = 4;
constexpr int kRpeThreads = 256;
constexpr int kMaxRpeKeys = 16384;  // keys sorted in shared memory per seed (128 KB)
constexpr int kMaxFyWalks = 4096;

struct WalkArgs {
    const unsigned long long *rowinfo;
    const void *rowptr;
    int rowptr64;
    const int32_t *col;
    const int32_t *seeds;
    int64_t n, N;
    int M, m, without, replay;
    uint32_t rng_lo, rng_hi;
    const long long *call_base;
    int32_t *walks;
    uint32_t *status;  // [0] status bits, [1] bad seed
    int fy_cap, smem_per_warp;
};

__device__ __forceinline__ void row_of(const WalkArgs &a, uint32_t v, int64_t &start, uint32_t &deg) {
    const unsigned long long q = __ldg(a.rowinfo + v);
    start = (int64_t)(q & 0xFFFFFFFFFFull);
    deg = (uint32_t)(q >> 40);
    if (deg == 0xFFFFFFu) {
        const int64_t e = a.rowptr64 ? (int64_t)__ldg((const long long *)a.rowptr + v + 1)
                                     : (int64_t)__ldg((const int *)a.rowptr + v + 1);
        deg = (uint32_t)min(e - start, (int64_t)0xffffffffll);
    }
}

__global__ void __launch_bounds__(kWalkWarps * 32) walk_sample_kernel(const WalkArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    int32_t *fy_pick = (int32_t *)(smem_raw + (size_t)wib * a.smem_per_warp);
    int32_t *fy_dense = fy_pick + a.M;
    int32_t *fy_key = fy_dense + a.M;
    int32_t *fy_val = fy_key + a.fy_cap;
    const int M = a.M, m = a.m, ncol = a.m + 1;

    for (int64_t i = (int64_t)blockIdx.x * wpb + wib; i < a.n; i += (int64_t)gridDim.x * wpb) {
        const int32_t u = __ldg(a.seeds + i);
        if ((uint64_t)(int64_t)u >= (uint64_t)a.N) {
            if (lane == 0) atomicOr(a.status + 1, 1u);
            continue;
        }
        int64_t rp0;
        uint32_t d;
        row_of(a, (uint32_t)u, rp0, d);
        int64_t calls0 = a.replay ? __ldg(a.call_base + i) : 0;
        const uint32_t i_lo = (uint32_t)i, i_hi = (uint32_t)((uint64_t)i >> 32);
        const bool fy = a.without && d > (uint32_t)M;

        if (fy) {  // partial Fisher-Yates over the neighbour positions (subg_acc.c:201-214)
            for (int k = lane; k < M; k += 32) {
                uint32_t pick;
                if (a.replay) {
                    uint32_t st = lcg_jump(a.rng_lo, 3u * (uint32_t)(calls0 + k));
                    pick = rand_r_dev(st) % (d - (uint32_t)k) + (uint32_t)k;
                } else {
                    const uint4 r4 = philox4x32_10(make_uint4(i_lo, i_hi, (uint32_t)k, 0x46597331u), make_uint2(a.rng_lo, a.rng_hi));
                    pick = (uint32_t)k + __umulhi(r4.x, d - (uint32_t)k);
                }
                fy_pick[k] = (int32_t)pick;
                fy_dense[k] = k;
            }
            for (int h = lane; h < a.fy_cap; h += 32) fy_key[h] = -1;
            __syncwarp();
            if (lane == 0) {  // the swaps are sequential by definition; positions >= M live in a sparse map
                const int hm = a.fy_cap - 1;
                for (int k = 0; k < M; k++) {
                    const int s = fy_pick[k];
                    const int vk = fy_dense[k];
                    if (s < M) {
                        const int vs = fy_dense[s];
                        fy_dense[s] = vk;
                        fy_dense[k] = vs;
                    } else {
                        int p = (int)(mix32((uint32_t)s) & (uint32_t)hm);
                        while (fy_key[p] != -1 && fy_key[p] != s) p = (p + 1) & hm;
                        const int vs = (fy_key[p] == s) ? fy_val[p] : s;
                        fy_key[p] = s;
                        fy_val[p] = vk;
                        fy_dense[k] = vs;
                    }
                }
            }
            __syncwarp();
            calls0 += M;
        }

        const int draws_per_walk = a.without ? m - 1 : m;  // rand_r calls of one walk (no dead ends)
        int32_t *out = a.walks + i * (int64_t)M * ncol;
        for (int g = 0; g * 32 < M; g += kWalkGroup) {
            uint32_t cur[kWalkGroup], rst[kWalkGroup];
#pragma unroll
            for (int t = 0; t < kWalkGroup; t++) {
                const int w = lane + 32 * (g + t);
                cur[t] = (uint32_t)u;
                rst[t] = 0;
                if (w < M) {
                    out[(int64_t)w * ncol] = u;
                    if (a.replay) rst[t] = lcg_jump(a.rng_lo, 3u * (uint32_t)(calls0 + (int64_t)w * draws_per_walk));
                }
            }
            for (int s = 0; s < m; s++) {
                int64_t rp[kWalkGroup];
                uint32_t dn[kWalkGroup];
#pragma unroll
                for (int t = 0; t < kWalkGroup; t++) {
                    if (s == 0) { rp[t] = rp0; dn[t] = d; }
                    else row_of(a, cur[t], rp[t], dn[t]);
                }
#pragma unroll
                for (int t = 0; t < kWalkGroup; t++) {
                    const int w = lane + 32 * (g + t);
                    if (w >= M) continue;
                    if (dn[t] > 0) {
                        uint32_t off;
                        if (a.without && s == 0) {
                            off = fy ? (uint32_t)fy_dense[w] : (uint32_t)w % dn[t];
                        } else if (a.replay) {
                            off = rand_r_dev(rst[t]) % dn[t];
                        } else {
                            const uint4 r4 = philox4x32_10(make_uint4(i_lo, i_hi, (uint32_t)w, 0x57414c4bu ^ (uint32_t)s),
                                                           make_uint2(a.rng_lo, a.rng_hi));
                            off = __umulhi(r4.x, dn[t]);
                        }
                        cur[t] = (uint32_t)__ldg(a.col + rp[t] + off);
                    } else if (a.replay && d > 0) {
                        atomicOr(a.status, SUBG_STATUS_DEAD_END);
                    }
                    out[(int64_t)w * ncol + s + 1] = (int32_t)cur[t];
                }
            }
        }
        __syncwarp();
    }
}

// rand_r calls a seed consumes in the reference's single stream (no dead ends assumed)
__global__ void walk_calls_kernel(const unsigned long long *rowinfo, const void *rowptr, int rowptr64, const int32_t *seeds,
                                  int64_t n, int64_t N, int M, int m, int without, int32_t *calls) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = seeds[i];
        if (u < 0 || u >= N) { calls[i] = 0; continue; }
        const int64_t d = rowptr64 ? ((const long long *)rowptr)[u + 1] - ((const long long *)rowptr)[u]
                                   : (int64_t)((const int *)rowptr)[u + 1] - ((const int *)rowptr)[u];
        if (without) calls[i] = (int32_t)((d > M ? M : 0) + (d > 0 ? (int64_t)M * (m - 1) : 0));
        else calls[i] = (int32_t)(d > 0 ? (int64_t)M * m : 0);
    }
}

// ---------------------------------------------------------------- relative-position encoder
// key = node << 32 | order, order = 0 for the root entry and 1 + (step-1)*M + walk for the node a walk is
// at after `step` hops: ascending order is the scan order of subg_acc.c:263-277.
template <bool EMIT>
__global__ void __launch_bounds__(kRpeThreads) rpe_kernel(const int32_t *walks, int64_t n, int M, int m, int P, int nbw,
                                                          int32_t *nsize, const long long *off, int32_t *ids, int32_t *rpe) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = (unsigned long long *)smem_raw;
    uint32_t *bitmap = (uint32_t *)(keys + P);
    uint32_t *bprefix = bitmap + nbw;
    __shared__ int s_heads;
    const int Kt = M * m + 1, ncol = m + 1;
    const int tid = threadIdx.x;
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const int32_t *wk = walks + i * (int64_t)M * ncol;
        for (int j = tid; j < P; j += kRpeThreads) {
            unsigned long long k = ~0ull;
            if (j == 0) k = (unsigned long long)(uint32_t)__ldg(wk) << 32;
            else if (j < Kt) {
                const int e = j - 1, s = e / M, w = e - s * M;  // step s+1, walk w
                k = ((unsigned long long)(uint32_t)__ldg(wk + (int64_t)w * ncol + s + 1) << 32) | (uint32_t)j;
            }
            keys[j] = k;
        }
        if (EMIT)
            for (int b = tid; b < nbw; b += kRpeThreads) bitmap[b] = 0u;
        if (tid == 0) s_heads = 0;
        __syncthreads();
        // bitonic sort, ascending
        for (int k2 = 2; k2 <= P; k2 <<= 1)
            for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                for (int t = tid; t < (P >> 1); t += kRpeThreads) {
                    const int lo = ((t / j2) * (j2 << 1)) + (t % j2);
                    const int hi = lo + j2;
                    const bool up = (lo & k2) == 0;
                    const unsigned long long x = keys[lo], y = keys[hi];
                    if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
                }
                __syncthreads();
            }
        // run heads
        int mine = 0;
        for (int j = tid; j < Kt; j += kRpeThreads) {
            const uint32_t node = (uint32_t)(keys[j] >> 32);
            const bool head = j == 0 || (uint32_t)(keys[j - 1] >> 32) != node;
            if (head) {
                mine++;
                if (EMIT) {
                    const uint32_t ord = (uint32_t)keys[j];
                    atomicOr(&bitmap[ord >> 5], 1u << (ord & 31));
                }
            }
        }
        if (!EMIT) {
            if (mine) atomicAdd(&s_heads, mine);
            __syncthreads();
            if (tid == 0) nsize[i] = s_heads;
            __syncthreads();
            continue;
        }
        __syncthreads();
        if (tid < 32) {  // exclusive popcount prefix over the bitmap words
            uint32_t running = 0;
            for (int b0 = 0; b0 < nbw; b0 += 32) {
                const int b = b0 + tid;
                const uint32_t cnt = b < nbw ? __popc(bitmap[b]) : 0u;
                const uint32_t inc = warp_incl_scan(cnt);
                if (b < nbw) bprefix[b] = running + inc - cnt;
                running += __shfl_sync(FULL, inc, 31);
            }
        }
        __syncthreads();
        const long long base = off[i];
        for (int j = tid; j < Kt; j += kRpeThreads) {
            const uint32_t node = (uint32_t)(keys[j] >> 32);
            if (j != 0 && (uint32_t)(keys[j - 1] >> 32) == node) continue;
            const uint32_t ord = (uint32_t)keys[j];
            const uint32_t rank = bprefix[ord >> 5] + __popc(bitmap[ord >> 5] & ((1u << (ord & 31)) - 1u));
            ids[base + rank] = (int32_t)node;
            int32_t *row = rpe + (base + rank) * ncol;
            int p = j;
            if (ord == 0) { row[0] = M; p++; }  // Coarr1[0] = num_walks (subg_acc.c:294)
            else row[0] = 0;
            for (int c = 1; c <= m; c++) {
                const uint32_t lim = (uint32_t)c * (uint32_t)M;  // orders of step c are (c-1)*M+1 .. c*M
                int cnt = 0;
                while (p < Kt && (uint32_t)(keys[p] >> 32) == node && (uint32_t)keys[p] <= lim) { cnt++; p++; }
                row[c] = cnt;
            }
        }
        __syncthreads();
    }
}

void free_walkset_arrays(WalkSet *w, cudaStream_t st) {
    dfree(w->walks, st); dfree(w->off, st); dfree(w->ids, st); dfree(w->rpe, st);
    w->walks = nullptr; w->off = nullptr; w->ids = nullptr; w->rpe = nullptr;
}

}  // namespace

int walk_sample_impl(const Graph *g, const int32_t *seeds_hd, int64_t n, int M, int m, uint64_t seed, int rng_mode,
                     int without, cudaStream_t st, WalkSet **out) {
    if (!g || !out || n < 0 || (n > 0 && !seeds_hd)) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (M < 1 || m < 1) return fail(SUBG_ERR_ARG, "num_walks and num_steps must be >= 1");
    if (rng_mode != SUBG_RNG_PHILOX && rng_mode != SUBG_RNG_RAND_R)
        return fail(SUBG_ERR_ARG, "walk_sampler draws from SUBG_RNG_PHILOX or SUBG_RNG_RAND_R");
    if ((int64_t)M * m + 1 > kMaxRpeKeys)
        return fail(SUBG_ERR_UNSUPPORTED, "walk_sampler: num_walks * num_steps + 1 must be <= 16384 (keys sorted in shared memory)");
    if (without > 0 && M > kMaxFyWalks)
        return fail(SUBG_ERR_UNSUPPORTED, "walk_sampler without replacement: num_walks must be <= 4096");
    DeviceGuard guard(g->device);
    g->tag.use_on(st);
    WalkSet *w = new WalkSet();
    w->tag.last = st;
    w->device = g->device; w->n = n; w->M = M; w->m = m;
    const int ncol = m + 1;
    int32_t *d_seeds = nullptr, *d_calls = nullptr, *d_nsize = nullptr;
    long long *call_base = nullptr, *scratch = nullptr;
    uint32_t *d_status = nullptr;
    int rc = SUBG_OK;
    cudaError_t e = cudaSuccess;
#define WK(call)                                                                                   \
    do {                                                                                           \
        e = (call);                                                                                \
        if (e != cudaSuccess) {                                                                    \
            rc = fail(e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA,               \
                      std::string(#call) + ": " + cudaGetErrorString(e));                          \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)
    {
        WK(dmalloc(&w->off, (size_t)n + 1, st));
        WK(dmalloc(&w->walks, (size_t)n * M * ncol, st));
        WK(dmalloc(&d_nsize, (size_t)n, st));
        WK(dmalloc(&d_status, 2, st));
        WK(dmalloc(&scratch, (size_t)std::max(1, scan_num_blocks(n)), st));
        WK(cudaMemsetAsync(d_status, 0, 2 * sizeof(uint32_t), st));
        const int32_t *seeds = seeds_hd;
        if (n > 0 && !is_device_ptr(seeds_hd)) {
            WK(dmalloc(&d_seeds, (size_t)n, st));
            WK(cudaMemcpyAsync(d_seeds, seeds_hd, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            seeds = d_seeds;
        }
        const unsigned gen_blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 8 * (int64_t)g->num_sms));
        const int replay = rng_mode == SUBG_RNG_RAND_R;
        if (replay && n > 0) {
            WK(dmalloc(&d_calls, (size_t)n, st));
            WK(dmalloc(&call_base, (size_t)n + 1, st));
            walk_calls_kernel<<<gen_blocks, 256, 0, st>>>((const unsigned long long *)g->rowinfo, g->rowptr, g->rowptr64 ? 1 : 0,
                                                          seeds, n, g->N, M, m, without > 0, d_calls);
            count_launch();
            WK(cudaGetLastError());
            WK(exclusive_scan_i32_i64(d_calls, call_base, n, 0, scratch, st));
            count_launch(3);
        }
        if (n > 0) {
            WalkArgs a{};
            a.rowinfo = (const unsigned long long *)g->rowinfo; a.rowptr = g->rowptr; a.rowptr64 = g->rowptr64 ? 1 : 0;
            a.col = g->col; a.seeds = seeds; a.n = n; a.N = g->N; a.M = M; a.m = m;
            a.without = without > 0; a.replay = replay;
            a.rng_lo = (uint32_t)seed; a.rng_hi = (uint32_t)(seed >> 32);
            a.call_base = call_base; a.walks = w->walks; a.status = d_status;
            int cap = 16;
            while (cap < 2 * M) cap <<= 1;
            a.fy_cap = cap;
            a.smem_per_warp = a.without ? (int)sizeof(int32_t) * (2 * M + 2 * cap) : 0;
            int warps = kWalkWarps;
            while (warps > 1 && (size_t)warps * a.smem_per_warp > 200 * 1024) warps >>= 1;
            const int smem = warps * a.smem_per_warp;
            WK(cudaFuncSetAttribute(walk_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(smem, 1024)));
            const int64_t want = (n + warps - 1) / warps;
            const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, 16 * (int64_t)g->num_sms));
            timing_begin(SUBG_TIMING_SAMPLER, st);
            walk_sample_kernel<<<blocks, warps * 32, smem, st>>>(a);
            timing_end(SUBG_TIMING_SAMPLER, st);
            count_launch();
            WK(cudaGetLastError());

            // relative-position encoder: sizes, scan, rows
            const int Kt = M * m + 1;
            int P = 64;
            while (P < Kt) P <<= 1;
            const int nbw = (Kt + 31) / 32;
            const int rsmem = P * (int)sizeof(unsigned long long) + 2 * nbw * (int)sizeof(uint32_t);
            WK(cudaFuncSetAttribute(rpe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rsmem));
            WK(cudaFuncSetAttribute(rpe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rsmem));
            const unsigned rblocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(n, 8 * (int64_t)g->num_sms));
            rpe_kernel<false><<<rblocks, kRpeThreads, rsmem, st>>>(w->walks, n, M, m, P, nbw, d_nsize, nullptr, nullptr, nullptr);
            count_launch();
            WK(cudaGetLastError());
            WK(exclusive_scan_i32_i64(d_nsize, w->off, n, 0, scratch, st));
            count_launch(3);
            long long T = 0;
            uint32_t h_status[2] = {0, 0};
            WK(cudaMemcpyAsync(&T, w->off + n, sizeof(long long), cudaMemcpyDeviceToHost, st));
            WK(cudaMemcpyAsync(h_status, d_status, sizeof(h_status), cudaMemcpyDeviceToHost, st));
            WK(cudaStreamSynchronize(st));
            if (h_status[1]) {
                rc = fail(SUBG_ERR_ARG, "query holds a node id outside [0, N)");
                goto done;
            }
            w->status = h_status[0];
            w->T = T;
            WK(dmalloc(&w->ids, (size_t)T, st));
            WK(dmalloc(&w->rpe, (size_t)T * ncol, st));
            rpe_kernel<true><<<rblocks, kRpeThreads, rsmem, st>>>(w->walks, n, M, m, P, nbw, nullptr, w->off, w->ids, w->rpe);
            count_launch();
            WK(cudaGetLastError());
        } else {
            WK(cudaMemsetAsync(w->off, 0, sizeof(long long), st));
        }
    }
done:
#undef WK
    dfree(d_seeds, st); dfree(d_calls, st); dfree(d_nsize, st); dfree(call_base, st); dfree(scratch, st); dfree(d_status, st);
    if (rc != SUBG_OK) {
        free_walkset_arrays(w, st);
        delete w;
        return rc;
    }
    *out = w;
    return SUBG_OK;
}

int walkset_export_impl(const WalkSet *w, int32_t *walks_hd, int64_t *off_hd, int32_t *ids_hd, int32_t *rpe_hd, cudaStream_t st) {
    if (!w) return fail(SUBG_ERR_ARG, "null walk set");
    DeviceGuard guard(w->device);
    w->tag.use_on(st);
    const size_t ncol = (size_t)w->m + 1;
    if (walks_hd && w->n > 0)
        SUBG_CUDA(cudaMemcpyAsync(walks_hd, w->walks, (size_t)w->n * w->M * ncol * sizeof(int32_t), cudaMemcpyDefault, st));
    if (off_hd) SUBG_CUDA(cudaMemcpyAsync(off_hd, w->off, ((size_t)w->n + 1) * sizeof(int64_t), cudaMemcpyDefault, st));
    if (ids_hd && w->T > 0) SUBG_CUDA(cudaMemcpyAsync(ids_hd, w->ids, (size_t)w->T * sizeof(int32_t), cudaMemcpyDefault, st));
    if (rpe_hd && w->T > 0) SUBG_CUDA(cudaMemcpyAsync(rpe_hd, w->rpe, (size_t)w->T * ncol * sizeof(int32_t), cudaMemcpyDefault, st));
    SUBG_CUDA(cudaStreamSynchronize(st));
    return SUBG_OK;
}

// ---------------------------------------------------------------- walk_join (SUREL v1), subg_acc.c:509-647
// For every query (u, v) and every position j of the rows of u and v in `walks`: the 1-based position of the visited
// node in the concatenated key sets, looked up in the set of u and in the set of v (0 = not a member).  The
// reference builds a two-level hash (root -> row, (row, node) -> running index); here both levels are sorted arrays
// searched by bisection: (row << 32 | node) -> index for the sets, root node -> row for the roots.
namespace {

__global__ void join_keys_kernel(const long long *off, const int32_t *ids, int64_t n, unsigned long long *keys, int32_t *vals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nw)
        for (long long j = off[i] + lane; j < off[i + 1]; j += 32) {
            keys[j] = ((unsigned long long)i << 32) | (uint32_t)ids[j];
            vals[j] = (int32_t)(j + 1);  // idx starts at 1 and runs over all rows (subg_acc.c:563,583-586)
        }
}
__global__ void join_roots_kernel(const int32_t *walks, int64_t n, int64_t stride, unsigned long long *keys) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = ((unsigned long long)(uint32_t)walks[i * stride] << 32) | (uint32_t)i;  // root node -> row (subg_acc.c:572-575)
}
__device__ __forceinline__ int32_t find_row(const unsigned long long *roots, int64_t n, int32_t node) {
    int64_t lo = 0, hi = n;
    const unsigned long long want = (unsigned long long)(uint32_t)node << 32;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (roots[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n && (roots[lo] >> 32) == (uint32_t)node) ? (int32_t)(uint32_t)roots[lo] : -1;
}
__global__ void join_query_rows_kernel(const unsigned long long *roots, int64_t n, const int32_t *query, int64_t Q2, int32_t *xq) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Q2; i += (int64_t)gridDim.x * blockDim.x)
        xq[i] = find_row(roots, n, query[i]);
}
__device__ __forceinline__ int32_t find_member(const unsigned long long *keys, const int32_t *vals, const long long *off,
                                               int32_t row, int32_t node) {
    if (row < 0) return -1;  // find_idx: unknown main key (subg_acc.c:96-99)
    long long lo = off[row], hi = off[row + 1];
    const long long end = hi;
    const unsigned long long want = ((unsigned long long)(uint32_t)row << 32) | (uint32_t)node;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    return (lo < end && keys[lo] == want) ? vals[lo] : 0;
}
__global__ void walk_join_kernel(const int32_t *walks, int64_t stride, const unsigned long long *keys, const int32_t *vals,
                                 const long long *off, const int32_t *xq, int64_t Q, int32_t *out) {
    const int64_t total = Q * stride, half = Q * 2 * stride;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t x = t / stride, j = t - x * stride;
        const int32_t r1 = xq[2 * x], r2 = xq[2 * x + 1];
        const int64_t o = 2 * x * stride + 2 * j;
        // a query node that is no root has no walks: its entries are -1 (the reference reads out of bounds there)
        const int32_t w1 = r1 >= 0 ? walks[(int64_t)r1 * stride + j] : -1;
        const int32_t w2 = r2 >= 0 ? walks[(int64_t)r2 * stride + j] : -1;
        out[o] = r1 >= 0 ? find_member(keys, vals, off, r1, w1) : -1;
        out[o + 1] = r1 >= 0 ? find_member(keys, vals, off, r2, w1) : -1;
        out[half + o] = r2 >= 0 ? find_member(keys, vals, off, r1, w2) : -1;
        out[half + o + 1] = r2 >= 0 ? find_member(keys, vals, off, r2, w2) : -1;
    }
}

}  // namespace

int walk_join_impl(const int32_t *walks_hd, int64_t n, int64_t stride, const int64_t *key_off_hd, const int32_t *key_ids_hd,
                   const int32_t *query_hd, int64_t Q, int32_t *out_hd, int32_t *xq_hd, int device, cudaStream_t st) {
    if (n < 0 || stride < 1 || Q < 0 || (n > 0 && (!walks_hd || !key_off_hd)) || (Q > 0 && (!query_hd || !out_hd)))
        return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (n >= (1ll << 31)) return fail(SUBG_ERR_ARG, "too many rows");
    DeviceGuard guard(device);
    int num_sms = 148;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device);
    int32_t *d_walks = nullptr, *d_ids = nullptr, *d_query = nullptr, *d_out = nullptr, *d_xq = nullptr, *v_a = nullptr, *v_b = nullptr;
    long long *d_off = nullptr;
    unsigned long long *k_a = nullptr, *k_b = nullptr, *r_a = nullptr, *r_b = nullptr;
    void *tmp = nullptr;
    int rc = SUBG_OK;
    cudaError_t e = cudaSuccess;
#define JK(call)                                                                                   \
    do {                                                                                           \
        e = (call);                                                                                \
        if (e != cudaSuccess) {                                                                    \
            rc = fail(e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA,               \
                      std::string(#call) + ": " + cudaGetErrorString(e));                          \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)
    {
        long long T = 0;
        const long long *off = (const long long *)key_off_hd;
        if (n > 0) {
            if (!is_device_ptr(key_off_hd)) {
                T = key_off_hd[n];
                JK(dmalloc(&d_off, (size_t)n + 1, st));
                JK(cudaMemcpyAsync(d_off, key_off_hd, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
                off = d_off;
            } else {
                JK(cudaMemcpyAsync(&T, key_off_hd + n, 8, cudaMemcpyDeviceToHost, st));
                JK(cudaStreamSynchronize(st));
            }
        }
        if (T < 0 || (T > 0 && !key_ids_hd)) { rc = fail(SUBG_ERR_ARG, "Input parsing error."); goto done; }
        const int32_t *walks = walks_hd, *ids = key_ids_hd, *query = query_hd;
        if (n > 0 && !is_device_ptr(walks_hd)) {
            JK(dmalloc(&d_walks, (size_t)n * stride, st));
            JK(cudaMemcpyAsync(d_walks, walks_hd, (size_t)n * stride * 4, cudaMemcpyHostToDevice, st));
            walks = d_walks;
        }
        if (T > 0 && !is_device_ptr(key_ids_hd)) {
            JK(dmalloc(&d_ids, (size_t)T, st));
            JK(cudaMemcpyAsync(d_ids, key_ids_hd, (size_t)T * 4, cudaMemcpyHostToDevice, st));
            ids = d_ids;
        }
        if (Q > 0 && !is_device_ptr(query_hd)) {
            JK(dmalloc(&d_query, (size_t)2 * Q, st));
            JK(cudaMemcpyAsync(d_query, query_hd, (size_t)2 * Q * 4, cudaMemcpyHostToDevice, st));
            query = d_query;
        }
        int32_t *out = out_hd, *xq = xq_hd;
        if (Q > 0 && !is_device_ptr(out_hd)) { JK(dmalloc(&d_out, (size_t)4 * Q * stride, st)); out = d_out; }
        if (Q > 0 && (!xq_hd || !is_device_ptr(xq_hd))) { JK(dmalloc(&d_xq, (size_t)2 * Q, st)); xq = d_xq; }
        JK(dmalloc(&k_a, (size_t)T + 1, st)); JK(dmalloc(&k_b, (size_t)T + 1, st));
        JK(dmalloc(&v_a, (size_t)T + 1, st)); JK(dmalloc(&v_b, (size_t)T + 1, st));
        JK(dmalloc(&r_a, (size_t)n + 1, st)); JK(dmalloc(&r_b, (size_t)n + 1, st));
        const unsigned gb = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n * 32 + 255) / 256, 8 * (int64_t)num_sms));
        cub::DoubleBuffer<unsigned long long> dk(k_a, k_b), dr(r_a, r_b);
        cub::DoubleBuffer<int32_t> dv(v_a, v_b);
        if (n > 0) {
            join_keys_kernel<<<gb, 256, 0, st>>>(off, ids, n, k_a, v_a);
            join_roots_kernel<<<gb, 256, 0, st>>>(walks, n, stride, r_a);
            int nbits = 1;
            while ((n >> nbits) != 0) nbits++;
            size_t b1 = 0, b2 = 0;
            JK(cub::DeviceRadixSort::SortPairs(nullptr, b1, dk, dv, (int64_t)T, 0, 32 + nbits, st));
            JK(cub::DeviceRadixSort::SortKeys(nullptr, b2, dr, (int64_t)n, 0, 64, st));
            JK(cudaMallocAsync(&tmp, std::max<size_t>(std::max(b1, b2), 16), st));
            if (T > 0) JK(cub::DeviceRadixSort::SortPairs(tmp, b1, dk, dv, (int64_t)T, 0, 32 + nbits, st));
            JK(cub::DeviceRadixSort::SortKeys(tmp, b2, dr, (int64_t)n, 0, 64, st));
            count_launch(8);
        }
        if (Q > 0) {
            const unsigned qb = (unsigned)std::max<int64_t>(1, std::min<int64_t>((Q * stride + 255) / 256, 16 * (int64_t)num_sms));
            join_query_rows_kernel<<<std::max(1u, std::min(qb, (unsigned)((2 * Q + 255) / 256))), 256, 0, st>>>(dr.Current(), n, query, 2 * Q, xq);
            walk_join_kernel<<<qb, 256, 0, st>>>(walks, stride, dk.Current(), dv.Current(), off, xq, Q, out);
            count_launch(2);
            JK(cudaGetLastError());
            if (out != out_hd) JK(cudaMemcpyAsync(out_hd, out, (size_t)4 * Q * stride * 4, cudaMemcpyDeviceToHost, st));
            if (xq_hd && xq != xq_hd) JK(cudaMemcpyAsync(xq_hd, xq, (size_t)2 * Q * 4, cudaMemcpyDeviceToHost, st));
        }
        JK(cudaStreamSynchronize(st));
    }
done:
#undef JK
    dfree(d_walks, st); dfree(d_ids, st); dfree(d_query, st); dfree(d_out, st); dfree(d_xq, st); dfree(d_off, st);
    dfree(k_a, st); dfree(k_b, st); dfree(v_a, st); dfree(v_b, st); dfree(r_a, st); dfree(r_b, st); dfree(tmp, st);
    return rc;
}

// ------------------------------------------------------------------ batch_sampler (subg_acc.c:391-507)
// The reference's serial mini-batch node sampler: ONE rand_r stream, and every seed's early exit depends on the running
// number of distinct nodes over all seeds before it, so the walk order is part of the result.  One warp replays it: lane 0
// owns the stream and the walks (a chain of dependent draws), all lanes initialise the Fisher-Yates index array of a seed
// with more than num_walks neighbours.  The batch is a bitmap over the nodes plus the list of distinct nodes in insertion
// order (what uthash's iteration returns, subg_acc.c:484-490).  No caller in the reference and no parallelism to speak of:
// provided for drop-in completeness, bit-exact given the stream's start (seed + pid).
namespace {
__global__ void batch_sample_kernel(const void *rowptr, int rowptr64, const int32_t *col, const int32_t *seeds, int64_t n, int64_t N,
                                    int M, int m, int thld, uint32_t state, uint32_t *seen, int32_t *rseq, int32_t *out,
                                    int64_t cap, long long *result) {
    const int lane = threadIdx.x;
    long long count = 0;
    bool bad = false;
    auto row = [&](int64_t v) -> int64_t {
        return rowptr64 ? (int64_t)((const long long *)rowptr)[v] : (int64_t)((const int32_t *)rowptr)[v];
    };
    auto add = [&](int32_t v) {   // lane 0 only
        const uint32_t bit = 1u << (v & 31);
        if (!(seen[v >> 5] & bit)) {
            seen[v >> 5] |= bit;
            if (count < cap) out[count] = v;
            count++;
        }
    };
    for (int64_t i = 0; i < n; i++) {
        const int32_t u = seeds[i];
        if ((uint64_t)(int64_t)u >= (uint64_t)N) { bad = true; break; }
        const int64_t r0 = row(u);
        const int64_t hop1 = row((int64_t)u + 1) - r0;
        if (hop1 > M) {   // partial Fisher-Yates, num_walks draws (subg_acc.c:430-441)
            for (int64_t j = lane; j < hop1; j += 32) rseq[j] = (int32_t)j;
            __syncwarp();
            if (lane == 0) {
                for (int k = 0; k < M; k++) {
                    const int64_t sidx = (int64_t)(rand_r_dev(state) % (uint32_t)(hop1 - k)) + k;
                    const int32_t t = rseq[k];
                    rseq[k] = rseq[sidx];
                    rseq[sidx] = t;
                }
            }
        }
        if (lane == 0) {
            add(u);
            for (int walk = 0; walk < M; walk++) {
                if (hop1 < 1) break;
                int32_t curr = hop1 <= M ? col[r0 + walk % hop1] : col[r0 + rseq[walk]];
                add(curr);
                for (int step = 1; step < m; step++) {
                    const int64_t c0 = row(curr);
                    const int64_t nn = row((int64_t)curr + 1) - c0;
                    if (nn > 0) {
                        curr = col[c0 + (int64_t)(rand_r_dev(state) % (uint32_t)nn)];
                        add(curr);
                    }
                }
                if ((int)count >= (int)((i + 1) * (int64_t)thld / n)) break;   // subg_acc.c:472 (int arithmetic)
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        result[0] = count;
        result[1] = bad ? 1 : 0;
    }
}
__global__ void max_degree_kernel(const void *rowptr, int rowptr64, const int32_t *seeds, int64_t n, int64_t N, unsigned long long *out) {
    unsigned long long mx = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = seeds[i];
        if (u < 0 || u >= N) continue;
        const int64_t d = rowptr64 ? ((const long long *)rowptr)[u + 1] - ((const long long *)rowptr)[u]
                                   : (int64_t)((const int32_t *)rowptr)[u + 1] - ((const int32_t *)rowptr)[u];
        mx = max(mx, (unsigned long long)d);
    }
    if (mx) atomicMax(out, mx);
}
}  // namespace

// out_hd: host or device int32[cap]; *count_out = distinct nodes (may exceed cap: then SUBG_ERR_MEM and nothing useful in out)
int batch_sample_impl(const Graph *g, const int32_t *seeds_hd, int64_t n, int M, int m, int thld, uint32_t state, int32_t *out_hd,
                      int64_t cap, int64_t *count_out, cudaStream_t st) {
    if (!g || n < 0 || (n > 0 && !seeds_hd) || M < 1 || m < 1 || cap < 0 || (cap > 0 && !out_hd) || !count_out)
        return fail(SUBG_ERR_ARG, "Input parsing error.");
    DeviceGuard guard(g->device);
    g->tag.use_on(st);
    *count_out = 0;
    if (n == 0) return SUBG_OK;
    int32_t *d_seeds = nullptr, *d_rseq = nullptr, *d_out = nullptr;
    uint32_t *d_seen = nullptr;
    long long *d_res = nullptr;
    unsigned long long *d_mx = nullptr;
    int rc = SUBG_OK;
    cudaError_t e = cudaSuccess;
    long long hres[2] = {0, 0};
    unsigned long long hmx = 0;
#define BK(call)                                                                                   \
    do {                                                                                           \
        e = (call);                                                                                \
        if (e != cudaSuccess) {                                                                    \
            rc = fail(e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA,               \
                      std::string(#call) + ": " + cudaGetErrorString(e));                          \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)
    BK(dmalloc(&d_seeds, (size_t)n, st));
    BK(cudaMemcpyAsync(d_seeds, seeds_hd, (size_t)n * 4, cudaMemcpyDefault, st));
    BK(dmalloc(&d_mx, 1, st));
    BK(cudaMemsetAsync(d_mx, 0, 8, st));
    max_degree_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, st>>>(g->rowptr, g->rowptr64 ? 1 : 0, d_seeds, n, g->N, d_mx);
    BK(cudaMemcpyAsync(&hmx, d_mx, 8, cudaMemcpyDeviceToHost, st));
    BK(cudaStreamSynchronize(st));
    BK(dmalloc(&d_rseq, (size_t)std::max<unsigned long long>(hmx, 1), st));
    BK(dmalloc(&d_seen, (size_t)(g->N / 32 + 1), st));
    BK(cudaMemsetAsync(d_seen, 0, (size_t)(g->N / 32 + 1) * 4, st));
    BK(dmalloc(&d_out, (size_t)std::max<int64_t>(cap, 1), st));
    BK(dmalloc(&d_res, 2, st));
    batch_sample_kernel<<<1, 32, 0, st>>>(g->rowptr, g->rowptr64 ? 1 : 0, g->col, d_seeds, n, g->N, M, m, thld, state, d_seen, d_rseq,
                                          d_out, cap, d_res);
    BK(cudaGetLastError());
    count_launch(2);
    BK(cudaMemcpyAsync(hres, d_res, 16, cudaMemcpyDeviceToHost, st));
    BK(cudaStreamSynchronize(st));
    if (hres[1]) { rc = fail(SUBG_ERR_ARG, "query contains node ids outside [0, N)"); goto done; }
    *count_out = hres[0];
    if (hres[0] > cap) { rc = fail(SUBG_ERR_MEM, "batch_sampler: output capacity too small"); goto done; }
    if (hres[0] > 0) {
        BK(cudaMemcpyAsync(out_hd, d_out, (size_t)hres[0] * 4, cudaMemcpyDefault, st));
        BK(cudaStreamSynchronize(st));
    }
done:
#undef BK
    dfree(d_seeds, st); dfree(d_rseq, st); dfree(d_seen, st); dfree(d_out, st); dfree(d_res, st); dfree(d_mx, st);
    return rc;
}

int walkset_info_impl(const WalkSet *w, int64_t *n, int64_t *T, int32_t *M, int32_t *ncol, uint32_t *status) {
    if (!w) return fail(SUBG_ERR_ARG, "null walk set");
    if (n) *n = w->n;
    if (T) *T = w->T;
    if (M) *M = w->M;
    if (ncol) *ncol = w->m + 1;
    if (status) *status = w->status;
    return SUBG_OK;
}

int walkset_views_impl(const WalkSet *w, const int32_t **walks, const int64_t **off, const int32_t **ids, const int32_t **rpe) {
    if (!w) return fail(SUBG_ERR_ARG, "null walk set");
    if (walks) *walks = w->walks;
    if (off) *off = (const int64_t *)w->off;
    if (ids) *ids = w->ids;
    if (rpe) *rpe = w->rpe;
    return SUBG_OK;
}

void walkset_free_impl(WalkSet *w) {
    if (!w) return;
    DeviceGuard guard(w->device);
    free_walkset_arrays(w, w->tag.free_stream());
    delete w;
}

}  // namespace subg
