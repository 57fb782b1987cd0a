// Walk-based set sampler + LP encoder: one warp per seed, everything between the
// CSR gathers and the staged set rows stays in registers / shared memory.
//
// Reference behaviour reproduced (file:line relative to /root/reference):
//   first hop without replacement, later hops uniform    subg_acc/subg_acc.c:763-809
//   per-seed dedup, first-visit slot order, LP counts    subg_acc/subg_acc.c:784-844
//   bucket overflow drops late nodes, walk continues      subg_acc/subg_acc.c:814-828
//   64-bit LP key, LEAD bit on the root                   subg_acc/subg_acc.c:900-955
//   first-occurrence ids of unique LP rows                subg_acc/subg_acc.c:957-978
//
// B200 design: a seed's M*m visits are packed as (node << OB | order) keys, held
// EPL per lane, sorted by a register/shuffle bitonic network (no shared-memory
// traffic, no atomics), so equal nodes become runs: run head = first visit,
// run population per step = landing counts, and the set comes out already in
// ascending node order, which is the order the SpG CSR needs.
#pragma once
#include "common.cuh"

namespace subg {

constexpr int kWarpsPerBlock = 4;
constexpr uint64_t kEmptyKey = ~0ull;
constexpr uint32_t kStatusTableFull = 1u << 31;  // internal status bit
constexpr int kFirstHopCap = 1000000;  // NEBMAX, subg_acc.c:13,750

struct SamplerArgs {
    const void *rowptr;
    int rowptr64;
    const int32_t *col;
    const int32_t *seeds;  // chunk-local [n_chunk]
    int64_t n_chunk;
    int64_t seed_base;     // global index of seeds[0]
    int M, m, stride;      // stride = reference's bucket stride (cap on set size)
    int OB;                // bits of the order field
    int SHIFT;             // 32 - clz(M), bits per LP column in the key
    int rng_mode;
    uint32_t rng_lo, rng_hi;
    const int64_t *call_base;  // RAND_R: exclusive prefix of rand_r calls, global seed index
    const int32_t *walks;      // TRACE: chunk-local [n_chunk, M, m]
    // staging rows (chunk-local), row pitch S_pad
    int32_t *st_node;
    int32_t *st_prov;
    uint16_t *st_rank;
    int S_pad;
    int32_t *nsize;  // chunk-local
    // LP-key intern table (global, L2 resident)
    unsigned long long *tab_key;
    unsigned long long *tab_pos;
    uint32_t tab_mask;
    uint32_t *tab_count;
    uint32_t *status;
    // shared memory carve-up (per warp)
    int rec_cap;  // records (multiple of 8)
    int nbw;      // bitmap words
    int fy_cap;   // Fisher-Yates overflow map capacity (power of two)
    int smem_per_warp;
};

__device__ __forceinline__ int64_t load_rowptr(const SamplerArgs &a, int64_t i) {
    return a.rowptr64 ? __ldg((const long long *)a.rowptr + i) : (int64_t)__ldg((const int *)a.rowptr + i);
}

// ---------------------------------------------------------------- register bitonic sort
template <typename K>
__device__ __forceinline__ void cswap(K &a, K &b) {
    K lo = a < b ? a : b;
    K hi = a < b ? b : a;
    a = lo;
    b = hi;
}

// Sorts the 32*EPL keys held by a warp (lane L register r = element L*EPL + r) ascending.
// "Flip" formulation of the bitonic network: every comparator keeps the minimum at the lower index.
template <typename K, int EPL>
__device__ __forceinline__ void warp_sort(K (&k)[EPL], int lane) {
    constexpr int NT = 32 * EPL;
#pragma unroll
    for (int size = 2; size <= NT; size <<= 1) {
        if (size <= EPL) {
#pragma unroll
            for (int r = 0; r < EPL; r++) {
                const int p = r ^ (size - 1);
                if (p > r) cswap(k[r], k[p]);
            }
        } else {
            const int lm = size / EPL - 1;
            const bool keep_min = (lane & (size / (2 * EPL))) == 0;
            K nk[EPL];
#pragma unroll
            for (int r = 0; r < EPL; r++) {
                const K o = __shfl_xor_sync(FULL, k[EPL - 1 - r], lm);
                const K lo = k[r] < o ? k[r] : o;
                const K hi = k[r] < o ? o : k[r];
                nk[r] = keep_min ? lo : hi;
            }
#pragma unroll
            for (int r = 0; r < EPL; r++) k[r] = nk[r];
        }
#pragma unroll
        for (int j = size >> 2; j > 0; j >>= 1) {
            if (j < EPL) {
#pragma unroll
                for (int r = 0; r < EPL; r++)
                    if ((r & j) == 0) cswap(k[r], k[r | j]);
            } else {
                const int lm = j / EPL;
                const bool keep_min = (lane & lm) == 0;
#pragma unroll
                for (int r = 0; r < EPL; r++) {
                    const K o = __shfl_xor_sync(FULL, k[r], lm);
                    const K lo = k[r] < o ? k[r] : o;
                    const K hi = k[r] < o ? o : k[r];
                    k[r] = keep_min ? lo : hi;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- LP-key interning
// Open-addressing table in global memory (a few MB, L2 resident).  Returns the slot of
// `key`; tab_pos[slot] keeps the smallest stream position at which the key occurs, which
// later yields the reference's first-occurrence ids (subg_acc.c:957-978).
__device__ __forceinline__ uint32_t intern_key(const SamplerArgs &a, unsigned long long key,
                                               unsigned long long pos) {
    uint32_t h = (uint32_t)mix64(key) & a.tab_mask;
    for (uint32_t probes = 0;; probes++) {
        if (probes > a.tab_mask) {  // table full: the host sees tab_count > cap/2 and reruns with a larger one
            atomicOr(a.status, kStatusTableFull);
            return 0;
        }
        unsigned long long cur = a.tab_key[h];
        if (cur == kEmptyKey) {
            cur = atomicCAS(&a.tab_key[h], kEmptyKey, key);
            if (cur == kEmptyKey) {
                atomicAdd(a.tab_count, 1u);
                cur = key;
            }
        }
        if (cur == key) break;
        h = (h + 1) & a.tab_mask;
    }
    if (pos < a.tab_pos[h]) atomicMin(&a.tab_pos[h], pos);
    return h;
}

// ---------------------------------------------------------------- the sampler kernel
// WT walks per lane, MS step slots per walk (power of two >= m); EPL = WT*MS keys per lane.
template <typename K, int WT, int MS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) gset_sample_kernel(const SamplerArgs a) {
    constexpr int EPL = WT * MS;
    constexpr int LS = (MS == 1) ? 0 : (MS == 2 ? 1 : 2);
    constexpr K SENT = ~(K)0;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    unsigned char *wsm = smem_raw + (size_t)wib * a.smem_per_warp;
    unsigned long long *rec_cnt = (unsigned long long *)wsm;
    int32_t *rec_node = (int32_t *)(rec_cnt + a.rec_cap);
    uint16_t *rec_ord = (uint16_t *)(rec_node + a.rec_cap);
    uint32_t *bitmap = (uint32_t *)(rec_ord + a.rec_cap);
    uint32_t *bprefix = bitmap + a.nbw;
    // Fisher-Yates scratch aliases the record area (used strictly before it)
    int32_t *fy_pick = (int32_t *)wsm;
    int32_t *fy_dense = fy_pick + a.M;
    int32_t *fy_key = fy_dense + a.M;
    int32_t *fy_val = fy_key + a.fy_cap;

    const int M = a.M, m = a.m;
    const uint32_t ord_mask = (1u << a.OB) - 1u;
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;

    for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + wib; i < a.n_chunk; i += nwarps) {
        const int64_t gi = a.seed_base + i;
        const int32_t u = __ldg(a.seeds + i);
        K key[EPL];
#pragma unroll
        for (int r = 0; r < EPL; r++) key[r] = SENT;

        if (a.rng_mode == SUBG_RNG_TRACE) {
            const int32_t *wk = a.walks + i * (int64_t)M * m;
#pragma unroll
            for (int s = 0; s < MS; s++) {
                if (s < m) {
#pragma unroll
                    for (int t = 0; t < WT; t++) {
                        const int w = lane + 32 * t;
                        if (w < M) {
                            const uint32_t v = (uint32_t)__ldg(wk + (int64_t)w * m + s);
                            key[s * WT + t] = ((K)v << a.OB) | (K)(1u + ((uint32_t)w << LS) + s);
                        }
                    }
                }
            }
        } else {
            const int64_t rp0 = load_rowptr(a, u);
            const int64_t dfull = load_rowptr(a, (int64_t)u + 1) - rp0;
            const int d = dfull > kFirstHopCap ? kFirstHopCap : (int)dfull;
            const bool replay = a.rng_mode == SUBG_RNG_RAND_R;
            const uint32_t gi_lo = (uint32_t)gi, gi_hi = (uint32_t)((uint64_t)gi >> 32);
            int64_t calls0 = 0;
            if (replay) calls0 = __ldg((const long long *)a.call_base + gi);

            // ---- first hop without replacement (subg_acc.c:763-776, 790-800)
            if (d > M) {
                for (int k = lane; k < M; k += 32) {
                    uint32_t pick;
                    if (replay) {
                        uint32_t st = lcg_jump(a.rng_lo, 3u * (uint32_t)(calls0 + k));
                        pick = rand_r_dev(st) % (uint32_t)(d - k) + k;
                    } else {
                        const uint4 r4 = philox4x32_10(make_uint4(gi_lo, gi_hi, (uint32_t)k, 0x46597331u),
                                                       make_uint2(a.rng_lo, a.rng_hi));
                        pick = k + __umulhi(r4.x, (uint32_t)(d - k));
                    }
                    fy_pick[k] = (int32_t)pick;
                    fy_dense[k] = k;
                }
                for (int h = lane; h < a.fy_cap; h += 32) fy_key[h] = -1;
                __syncwarp();
                if (lane == 0) {
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

            // Walks are advanced in groups of GW per lane: all loads of one hop of a group are
            // issued back to back (GW x 32 gathers in flight per warp), while only the group's
            // RNG / row state is live in registers.
            constexpr int GW = WT < 4 ? WT : 4;
#pragma unroll
            for (int g = 0; g < WT; g += GW) {
                uint32_t cur[GW];
#pragma unroll
                for (int tt = 0; tt < GW; tt++) {
                    const int t = g + tt;
                    const int w = lane + 32 * t;
                    cur[tt] = (uint32_t)u;
                    if (w < M && d > 0) {
                        const int off = (d <= M) ? (w % d) : fy_dense[w];
                        cur[tt] = (uint32_t)__ldg(a.col + rp0 + off);
                    }
                    if (w < M) key[t] = ((K)cur[tt] << a.OB) | (K)(1u + ((uint32_t)w << LS));
                }
                // ---- later hops, uniform with replacement (subg_acc.c:802-809)
                uint32_t rst[GW];                  // RAND_R state per walk
                uint32_t rx[GW], ry[GW], rz[GW];   // Philox draw per walk (steps 1..3)
                if (m > 1) {
#pragma unroll
                    for (int tt = 0; tt < GW; tt++) {
                        const int w = lane + 32 * (g + tt);
                        if (replay) {
                            rst[tt] = lcg_jump(a.rng_lo, 3u * (uint32_t)(calls0 + (int64_t)w * (m - 1)));
                        } else {
                            const uint4 r4 = philox4x32_10(make_uint4(gi_lo, gi_hi, (uint32_t)w, 0x57414c4bu),
                                                           make_uint2(a.rng_lo, a.rng_hi));
                            rx[tt] = r4.x; ry[tt] = r4.y; rz[tt] = r4.z;
                        }
                    }
                }
#pragma unroll
                for (int s = 1; s < MS; s++) {
                    if (s < m) {
                        int64_t rp[GW];
                        uint32_t dn[GW];
#pragma unroll
                        for (int tt = 0; tt < GW; tt++) {
                            rp[tt] = load_rowptr(a, cur[tt]);
                            dn[tt] = (uint32_t)(load_rowptr(a, (int64_t)cur[tt] + 1) - rp[tt]);
                        }
#pragma unroll
                        for (int tt = 0; tt < GW; tt++) {
                            const int t = g + tt;
                            const int w = lane + 32 * t;
                            if (w < M) {
                                if (dn[tt] > 0) {
                                    uint32_t off;
                                    if (replay) {
                                        off = rand_r_dev(rst[tt]) % dn[tt];
                                    } else {
                                        const uint32_t r = s == 1 ? rx[tt] : (s == 2 ? ry[tt] : rz[tt]);
                                        off = __umulhi(r, dn[tt]);
                                    }
                                    cur[tt] = (uint32_t)__ldg(a.col + rp[tt] + off);
                                } else if (replay && d > 0) {
                                    atomicOr(a.status, SUBG_STATUS_DEAD_END);
                                }
                                key[s * WT + t] = ((K)cur[tt] << a.OB) | (K)(1u + ((uint32_t)w << LS) + s);
                            }
                        }
                    }
                }
            }
            __syncwarp();  // fy_dense reads done before the record area is reused
        }
        // root: order 0, parked in the last register of lane 31 (free by template choice)
        if (lane == 31) key[EPL - 1] = (K)(uint32_t)u << a.OB;

        warp_sort<K, EPL>(key, lane);

        // ---- runs of equal node = one set member each
        const K prev_last = __shfl_up_sync(FULL, key[EPL - 1], 1);
        uint32_t headmask = 0;  // EPL <= 32 bits per word; EPL == 64 uses two words
        uint32_t headmask_hi = 0;
#pragma unroll
        for (int r = 0; r < EPL; r++) {
            const K pk = r ? key[r - 1] : prev_last;
            const bool valid = key[r] != SENT;
            const bool head = valid && ((r == 0 && lane == 0) || ((pk >> a.OB) != (key[r] >> a.OB)));
            if (head) {
                if (r < 32) headmask |= 1u << (r & 31);
                else headmask_hi |= 1u << (r & 31);
            }
        }
        const uint32_t nhead = __popc(headmask) + __popc(headmask_hi);
        const uint32_t incl = warp_incl_scan(nhead);
        const int s_total = (int)__shfl_sync(FULL, incl, 31);
        int idx = (int)(incl - nhead) - 1;

        for (int t = lane; t < s_total; t += 32) rec_cnt[t] = 0ull;
        for (int b = lane; b < a.nbw; b += 32) bitmap[b] = 0u;
        __syncwarp();

        {
            unsigned long long acc = 0ull;
#pragma unroll
            for (int r = 0; r < EPL; r++) {
                const bool valid = key[r] != SENT;
                const bool head = r < 32 ? ((headmask >> (r & 31)) & 1u) : ((headmask_hi >> (r & 31)) & 1u);
                if (valid) {
                    const uint32_t ord = (uint32_t)key[r] & ord_mask;
                    if (head) {
                        if (acc) atomicAdd(&rec_cnt[idx], acc);
                        acc = 0ull;
                        idx++;
                        rec_node[idx] = (int32_t)(key[r] >> a.OB);
                        rec_ord[idx] = (uint16_t)ord;
                        atomicOr(&bitmap[ord >> 5], 1u << (ord & 31));
                    }
                    if (ord) acc += 1ull << (16 * ((ord - 1u) & (uint32_t)(MS - 1)));
                }
            }
            if (acc) atomicAdd(&rec_cnt[idx], acc);
        }
        __syncwarp();

        // ---- first-visit rank of every member = popcount prefix over the order bitmap
        {
            uint32_t running = 0;
            for (int b0 = 0; b0 < a.nbw; b0 += 32) {
                const int b = b0 + lane;
                const uint32_t cnt = b < a.nbw ? __popc(bitmap[b]) : 0u;
                const uint32_t inc = warp_incl_scan(cnt);
                if (b < a.nbw) bprefix[b] = running + inc - cnt;
                running += __shfl_sync(FULL, inc, 31);
            }
        }
        __syncwarp();

        // ---- emit the set: ascending node id, provisional LP id, first-visit rank
        int kept = 0;
        const int64_t row = i * (int64_t)a.S_pad;
        for (int t0 = 0; t0 < s_total; t0 += 32) {
            const int t = t0 + lane;
            const bool act = t < s_total;
            uint32_t ord = 0, rank = 0;
            if (act) {
                ord = rec_ord[t];
                rank = bprefix[ord >> 5] + __popc(bitmap[ord >> 5] & ((1u << (ord & 31)) - 1u));
            }
            const bool keep = act && (int)rank < a.stride;
            const uint32_t km = __ballot_sync(FULL, keep);
            if (keep) {
                const int o = kept + __popc(km & ((1u << lane) - 1u));
                const unsigned long long cnt = rec_cnt[t];
                unsigned long long lp = 0ull;
                for (int j = 0; j < m; j++) lp = (lp << a.SHIFT) | ((cnt >> (16 * j)) & 0xffffull);
                if (ord == 0) lp |= 1ull << (m * a.SHIFT);
                const uint32_t prov = intern_key(a, lp, ((unsigned long long)gi << 16) | rank);
                a.st_node[row + o] = rec_node[t];
                a.st_prov[row + o] = (int32_t)prov;
                a.st_rank[row + o] = (uint16_t)rank;
            }
            kept += __popc(km);
        }
        if (lane == 0) {
            a.nsize[i] = kept;
            if (kept < s_total) atomicOr(a.status, SUBG_STATUS_BUCKET_OVERFLOW);
        }
        __syncwarp();  // record area is reused by the next seed
    }
}

// rand_r calls consumed per seed in the reference's single stream:
// M for the Fisher-Yates draw if deg > M, plus M*(m-1) later hops if deg > 0.
static __global__ void rand_r_calls_kernel(const void *rowptr, int rowptr64, const int32_t *seeds, int64_t n,
                                    int M, int m, int32_t *calls) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = seeds[i];
        int64_t d = rowptr64 ? ((const long long *)rowptr)[u + 1] - ((const long long *)rowptr)[u]
                             : (int64_t)((const int *)rowptr)[u + 1] - ((const int *)rowptr)[u];
        if (d > kFirstHopCap) d = kFirstHopCap;
        calls[i] = (d > M ? M : 0) + (d > 0 ? M * (m - 1) : 0);
    }
}

}  // namespace subg
