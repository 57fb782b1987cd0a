// SpG construction on the device: sampler launch, set-size scan, row compaction,
// first-occurrence ranking of the unique LP rows, export in the reference's layout.
//
//   dense compaction            subg_acc/subg_acc.c:848-872  -> rows are written once at a cursor by the sampler;
//                                                               compact_rows_kernel only for CSR views / chunked runs
//   unique ids / enc table      subg_acc/subg_acc.c:957-1000 -> collect/finalize kernels + remap
//   CSR-of-sets (sorted cols)   sampler/random_walks.py:79-80 -> rows leave the sampler sorted
//   return list                 subg_acc/subg_acc.c:1017-1024 -> export kernels
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <cstdlib>
#include <string>
#include <vector>

#include "sampler.cuh"
#include "scan.cuh"

namespace subg {

// implemented in sampler_k32*.cu / sampler_k64*.cu (EPL = keys per lane of the warp sort)
#define SUBG_DECL(n) cudaError_t n(const SamplerArgs &a, int EPL, int num_sms, cudaStream_t st);
SUBG_DECL(launch_gset_sample_k32a) SUBG_DECL(launch_gset_sample_k32b) SUBG_DECL(launch_gset_sample_k32c)
SUBG_DECL(launch_gset_sample_k32d) SUBG_DECL(launch_gset_sample_k64a) SUBG_DECL(launch_gset_sample_k64b)
SUBG_DECL(launch_gset_sample_k64c) SUBG_DECL(launch_gset_sample_k64d)
#undef SUBG_DECL
static cudaError_t launch_gset_sample_k32(const SamplerArgs &a, int EPL, int num_sms, cudaStream_t st) {
    if (EPL <= 13) return launch_gset_sample_k32a(a, EPL, num_sms, st);
    if (EPL <= 21) return launch_gset_sample_k32b(a, EPL, num_sms, st);
    if (EPL <= 33) return launch_gset_sample_k32c(a, EPL, num_sms, st);
    return launch_gset_sample_k32d(a, EPL, num_sms, st);
}
static cudaError_t launch_gset_sample_k64(const SamplerArgs &a, int EPL, int num_sms, cudaStream_t st) {
    if (EPL <= 13) return launch_gset_sample_k64a(a, EPL, num_sms, st);
    if (EPL <= 21) return launch_gset_sample_k64b(a, EPL, num_sms, st);
    if (EPL <= 33) return launch_gset_sample_k64c(a, EPL, num_sms, st);
    return launch_gset_sample_k64d(a, EPL, num_sms, st);
}
static const int kEplList[] = {3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 25, 29, 33, 41, 49, 63};
// keys per lane that are a power of two sort with the register bitonic network instead of the merge path; taken when the
// seed's keys fit 128 / 256 slots (SUBG_SAMPLER_POW2=0: measurement only).  Measured (profiles/r2_sampler_sweeps.txt): dblp
// shape, 201 keys: EPL 8 bitonic 4.15 ms against EPL 7 merge path 4.89 ms; collab shape, 401 keys: EPL 16 bitonic 1.22 ms
// against EPL 13 merge path 1.02 ms (the padding to 512 slots eats the saving), so 16 is not offered.
static const int kEplPow2[] = {4, 8};

static int ceil_log2(uint64_t x) {
    int b = 0;
    while ((1ull << b) < x) b++;
    return b;
}
static int64_t env_i64(const char *name, int64_t dflt) {
    const char *v = getenv(name);
    return v ? atoll(v) : dflt;
}

__global__ void check_seeds_kernel(const int32_t *seeds, int64_t n, int64_t N, uint32_t *bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (seeds[i] < 0 || seeds[i] >= N) atomicOr(bad, 1u);
}

__global__ void pack_col3_kernel(const int32_t *col, int64_t E, unsigned long long *out) {
    const int64_t W = (E + 2) / 3;
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < W; w += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long v = 0ull;
        for (int q = 0; q < 3; q++) {
            const int64_t e = 3 * w + q;
            if (e < E) v |= (unsigned long long)(uint32_t)col[e] << (21 * q);
        }
        out[w] = v;
    }
}
// decides once per graph whether the packed column array pays off, and builds it
static const unsigned long long *graph_col3(const Graph *g, cudaStream_t st) {
    if (g->col3_state < 0) {
        const int64_t csr_bytes = 4 * g->E + 8 * g->N;
        bool want = g->N <= (1ll << 21) && g->E < (1ll << 32) && g->E > 0 && csr_bytes > (96ll << 20);
        const int64_t force = env_i64("SUBG_COL_PACK", -1);
        if (force == 0) want = false;
        if (force == 1) want = g->N <= (1ll << 21) && g->E < (1ll << 32) && g->E > 0;
        g->col3_state = 0;
        if (want) {
            const int64_t W = (g->E + 2) / 3;
            if (cudaMallocAsync((void **)&g->col3, (size_t)(W + 2) * 8, st) == cudaSuccess) {
                pack_col3_kernel<<<8 * g->num_sms, 256, 0, st>>>(g->col, g->E, g->col3);
                g->col3_state = 1;
                count_launch(1);
            } else {
                cudaGetLastError();
                g->col3 = nullptr;
            }
        }
    }
    return g->col3_state == 1 ? g->col3 : nullptr;
}

__global__ void fill_u64_kernel(unsigned long long *p, int64_t n, unsigned long long v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// rows at rowbeg[i] (any order) -> dense rows at indptr[i]; one warp per row
__global__ void compact_rows_kernel(const int32_t *src_node, const int32_t *src_data, const uint16_t *src_slot,
                                    const long long *rowbeg, const int32_t *nsize, const long long *indptr,
                                    int64_t n_rows, int32_t *indices, int32_t *data, uint16_t *slot) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n_rows; i += nwarps) {
        const int s = nsize[i];
        const int64_t src = rowbeg[i], dst = indptr[i];
        for (int j = lane; j < s; j += 32) {
            indices[dst + j] = src_node[src + j];
            data[dst + j] = src_data[src + j];
            if (slot) slot[dst + j] = src_slot[src + j];
        }
    }
}

__global__ void collect_unique_kernel(const unsigned long long *tab_key, const unsigned long long *tab_pos,
                                      uint32_t cap, unsigned long long *u_pos, uint32_t *u_slot, uint32_t *cnt) {
    for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < cap; h += gridDim.x * blockDim.x)
        if (tab_key[h] != kEmptyKey) {
            const uint32_t j = atomicAdd(cnt, 1u);
            u_pos[j] = tab_pos[h];
            u_slot[j] = h;
        }
}

// sorted by first occurrence: id j <- table slot; decode the key back into an int16 LP row.
// c_dev (nullable): the number of valid entries lives on the device (the multi-GPU merge sorts an upper bound)
__global__ void finalize_unique_kernel(const uint32_t *sorted_slot, uint32_t c, const uint32_t *c_dev,
                                       const unsigned long long *tab_key, int M, int m, int SHIFT,
                                       int32_t *rank_of_slot, int16_t *enc, unsigned long long *lp_key) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c || (c_dev && j >= *c_dev)) return;
    const uint32_t h = sorted_slot[j];
    if (rank_of_slot) rank_of_slot[h] = (int32_t)j;
    const unsigned long long key = tab_key[h];
    if (lp_key) lp_key[j] = key;
    const int ncol = m + 1;
    const unsigned long long fm = (1ull << SHIFT) - 1ull;
    enc[(int64_t)j * ncol] = ((key >> (m * SHIFT)) & 1ull) ? (int16_t)M : (int16_t)0;
    for (int q = 1; q <= m; q++) enc[(int64_t)j * ncol + q] = (int16_t)((key >> (SHIFT * (m - q))) & fm);
}

// LP-key table -> ids in first-occurrence order (subg_acc.c:957-978 without the serial scan): the occupied slots are
// collected, sorted by the smallest stream position of their key, and numbered.  c_max bounds the number of keys (the
// exact count is left in *d_cnt).  Outputs: rank_of_slot[cap] (slot -> id), enc int16[c, m+1], lp_key[c], lp_pos[c]
// (ids ascending); entries beyond the count are undefined.  All launches on `st`, no synchronisation.
int rank_unique_keys(const unsigned long long *tab_key, const unsigned long long *tab_pos, uint32_t cap, uint32_t c_max,
                     bool count_exact, int M, int m, int SHIFT, int32_t *rank_of_slot, int16_t *enc,
                     unsigned long long *lp_key, unsigned long long *lp_pos, uint32_t *d_cnt, int num_sms, cudaStream_t st) {
    if (c_max == 0) return SUBG_OK;
    unsigned long long *u_pos = nullptr;
    uint32_t *u_slot = nullptr, *u_slot2 = nullptr;
    void *cub_tmp = nullptr;
    SUBG_CUDA(dmalloc(&u_pos, (size_t)c_max, st));
    SUBG_CUDA(dmalloc(&u_slot, (size_t)c_max, st));
    SUBG_CUDA(dmalloc(&u_slot2, (size_t)c_max, st));
    if (!count_exact) SUBG_CUDA(cudaMemsetAsync(u_pos, 0xff, (size_t)c_max * 8, st));  // padding sorts behind every key
    SUBG_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t), st));
    collect_unique_kernel<<<4 * num_sms, 256, 0, st>>>(tab_key, tab_pos, cap, u_pos, u_slot, d_cnt);
    size_t tmp_bytes = 0;
    SUBG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, u_pos, lp_pos, u_slot, u_slot2, (int)c_max, 0, 64, st));
    SUBG_CUDA(cudaMallocAsync(&cub_tmp, tmp_bytes ? tmp_bytes : 1, st));
    SUBG_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, tmp_bytes, u_pos, lp_pos, u_slot, u_slot2, (int)c_max, 0, 64, st));
    finalize_unique_kernel<<<(c_max + 255) / 256, 256, 0, st>>>(u_slot2, c_max, count_exact ? nullptr : d_cnt, tab_key, M, m, SHIFT,
                                                              rank_of_slot, enc, lp_key);
    SUBG_CUDA(cudaGetLastError());
    dfree(u_pos, st); dfree(u_slot, st); dfree(u_slot2, st); dfree(cub_tmp, st);
    count_launch(4);
    return SUBG_OK;
}

// provisional table slot -> LP-row id + 1, in place: one streaming pass (16-byte accesses; the slot -> id map is a few
// thousand hot entries of an L1/L2-resident array)
__global__ void remap_ids_kernel(int32_t *data, int64_t T, const int32_t *rank_of_slot) {
    const int64_t T4 = T >> 2;
    int4 *d4 = (int4 *)data;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T4; i += (int64_t)gridDim.x * blockDim.x) {
        int4 v = d4[i];
        v.x = __ldg(rank_of_slot + v.x) + 1;
        v.y = __ldg(rank_of_slot + v.y) + 1;
        v.z = __ldg(rank_of_slot + v.z) + 1;
        v.w = __ldg(rank_of_slot + v.w) + 1;
        d4[i] = v;
    }
    for (int64_t i = (T4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x)
        data[i] = __ldg(rank_of_slot + data[i]) + 1;
}

__global__ void relabel_ids_kernel(int32_t *data, int64_t T, const int32_t *id_map, int32_t c_old, int32_t c_new) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t d = data[i];
        if (d >= 1 && d <= c_old) {
            const int32_t g = id_map[d - 1];
            data[i] = (g >= 0 && g < c_new) ? g + 1 : 0;
        }
    }
}

// reference layout: entries of a set in first-visit order, ids without the +1.
// dstptr: exclusive scan of the set sizes (== indptr of the compact layout)
__global__ void export_remap_kernel(const long long *rowbeg, const int32_t *nsize, const long long *dstptr,
                                    const int32_t *indices, const int32_t *data, const uint16_t *slot, int64_t n,
                                    int64_t T, int32_t *remap, const int16_t *enc, int ncol, int16_t *raw) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t b = rowbeg[i], e = b + nsize[i], db = dstptr[i];
        for (int64_t j = b + lane; j < e; j += 32) {
            const int64_t d = db + slot[j];
            const int32_t id = data[j] - 1;
            remap[d] = indices[j];
            remap[T + d] = id;
            if (raw)
                for (int q = 0; q < ncol; q++) raw[d * ncol + q] = enc[(int64_t)id * ncol + q];
        }
    }
}

__global__ void row_sizes_kernel(const long long *indptr, int64_t n, int32_t *nsize) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        nsize[i] = (int32_t)(indptr[i + 1] - indptr[i]);
}

static void free_spg_arrays(SpG *s, cudaStream_t st) {
    if (s->rowbeg && (void *)s->rowbeg != (void *)s->indptr) dfree(s->rowbeg, st);
    dfree(s->indptr, st); dfree(s->slot, st);
    if (!s->borrowed) { dfree(s->indices, st); dfree(s->data, st); }
    dfree(s->enc, st); dfree(s->nsize, st); dfree(s->seeds, st); dfree(s->lp_key, st); dfree(s->lp_pos, st); dfree(s->walks, st);
    s->walks = nullptr;
    s->indptr = nullptr; s->rowbeg = nullptr; s->indices = nullptr; s->data = nullptr; s->slot = nullptr;
    s->enc = nullptr; s->nsize = nullptr; s->seeds = nullptr; s->lp_key = nullptr; s->lp_pos = nullptr;
}

// scattered rows -> compact CSR (rows back to back in seed order); frees the slack of the cursor layout
int spg_ensure_csr(SpG *s, cudaStream_t st) {
    if (!s) return fail(SUBG_ERR_ARG, "null SpG");
    if (s->indptr) return SUBG_OK;
    DeviceGuard guard(s->device);
    s->tag.use_on(st);
    const int64_t n = s->n;
    int64_t *indptr = nullptr;
    long long *scratch = nullptr;
    int32_t *ni = nullptr, *nd = nullptr;
    uint16_t *ns = nullptr;
    SUBG_CUDA(dmalloc(&indptr, (size_t)n + 1, st));
    SUBG_CUDA(dmalloc(&scratch, (size_t)std::max(1, scan_num_blocks(n)), st));
    SUBG_CUDA(exclusive_scan_i32_i64(s->nsize, (long long *)indptr, n, 0, scratch, st));
    SUBG_CUDA(dmalloc(&ni, (size_t)s->T + 16, st));
    SUBG_CUDA(dmalloc(&nd, (size_t)s->T + 16, st));
    if (s->slot) SUBG_CUDA(dmalloc(&ns, (size_t)s->T + 16, st));
    if (n > 0) {
        const int64_t cblocks = std::min<int64_t>((n * 32 + 255) / 256, 8 * (int64_t)s->num_sms);
        compact_rows_kernel<<<(unsigned)std::max<int64_t>(cblocks, 1), 256, 0, st>>>(
            s->indices, (const int32_t *)s->data, s->slot, (const long long *)s->rowbeg, s->nsize,
            (const long long *)indptr, n, ni, nd, ns);
        SUBG_CUDA(cudaGetLastError());
        count_launch(4);
    }
    SUBG_CUDA(cudaStreamSynchronize(st));
    dfree(scratch, st);
    if (!s->borrowed) { dfree(s->indices, st); dfree(s->data, st); }
    s->borrowed = false;   // the compact copy is owned (a linked SpG has just pulled its remote rows over NVLink)
    dfree(s->slot, st); dfree(s->rowbeg, st);
    s->indices = ni; s->data = nd; s->slot = ns;
    s->indptr = indptr; s->rowbeg = indptr;
    s->extent = s->T; s->cap = s->T + 16;
    return SUBG_OK;
}

// free device memory as of the first request (refresh = ask the driver again)
static int64_t free_memory_estimate(int device, bool refresh) {
    static std::mutex mu;
    static std::vector<int64_t> cache;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)cache.size() <= device) cache.resize(device + 1, -1);
    if (cache[device] < 0 || refresh) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
            cudaGetLastError();
            free_b = (size_t)8 << 30;
        }
        cache[device] = (int64_t)free_b;
    }
    return cache[device];
}

struct SamplePlan {
    int EPL, OB, LS, SHIFT, stride, Kt, rowcap, nbw, fy_cap, lp_off, bitmap_off, smem_per_warp;
    bool key64, lp64;
};

static int make_plan(const Graph *g, int M, int m, int bucket, SamplePlan *p) {
    if (M < 1 || M > 32767) return fail(SUBG_ERR_ARG, "num_walks must be in [1, 32767] (int16 landing counts)");
    if (m < 1) return fail(SUBG_ERR_ARG, "num_steps must be >= 1");
    if (bucket == 0) return fail(SUBG_ERR_ARG, "bucket must be >= 1 or negative");
    int shift = 0;
    while ((M >> shift) != 0) shift++;
    if ((int64_t)m * shift + 1 > 64)  // subg_acc.c:905-915
        return fail(SUBG_ERR_ASSERT, "Longer width of type for hasing key needed > INT64.");
    const int64_t Kt = (int64_t)M * m + 1;
    const int max_epl = kEplList[sizeof(kEplList) / sizeof(int) - 1];
    if (m > 16 || Kt > 32 * (int64_t)max_epl)
        return fail(SUBG_ERR_UNSUPPORTED, "this build supports num_steps <= 16 and num_walks * num_steps <= 2015");
    p->SHIFT = shift;
    p->Kt = (int)Kt;
    p->LS = ceil_log2((uint64_t)m);  // step slots per walk: 1, 2, 4, 8, 16
    p->EPL = max_epl;
    for (int e : kEplList)
        if (32 * e >= Kt) { p->EPL = e; break; }
    const uint32_t max_ord = ((uint32_t)M << p->LS) | (uint32_t)(m - 1);
    p->OB = ceil_log2((uint64_t)max_ord + 1);
    p->key64 = !((uint64_t)g->N <= (1ull << (32 - p->OB)) - 1ull);
    // 32-bit keys that fit 128 / 256 slots: a power-of-two EPL sorts with the register bitonic network (64-bit keys
    // cost two shuffles and a multi-instruction compare per step: twitter shape 168 -> 243 ms, so they keep the merge path)
    if (!p->key64 && Kt > 96 && env_i64("SUBG_SAMPLER_POW2", 1) != 0)
        for (int e : kEplPow2)
            if (32 * e >= Kt) { p->EPL = e; break; }
    const int ksz = p->key64 ? 8 : 4;
    p->stride = bucket < 0 ? (int)Kt : bucket;
    p->rowcap = (std::min(p->stride, (int)Kt) + 3) & ~3;
    p->nbw = (int)((max_ord + 1 + 31) / 32);
    int fc = 16;
    while (fc < M + M / 4) fc <<= 1;
    p->fy_cap = fc;
    // region 1: key buffer, later the member keys; Fisher-Yates picks + hash keys before the walk
    // region 2: member LP rows; Fisher-Yates permutation + hash values during the first hop
    p->lp64 = m * shift + 1 > 32;
    const int fy_half = 4 * M + 4 * fc;
    p->lp_off = (std::max(ksz * (32 * p->EPL + 64), fy_half) + 15) & ~15;  // keys + two merge sentinels per run
    p->bitmap_off = (p->lp_off + std::max((p->lp64 ? 8 : 4) * (int)Kt, fy_half) + 15) & ~15;
    p->smem_per_warp = (p->bitmap_off + 8 * p->nbw + 15) & ~15;
    return SUBG_OK;
}

// seeds_hd holds the whole query (n_all entries); the sets of the window [lo, hi) are sampled.  Seed
// indices stay global (Philox counters, rand_r call offsets, first-occurrence positions), so the
// shards of a range-partitioned query concatenate to exactly the single-call result.
int gset_sample_impl(const Graph *g, const int32_t *seeds_hd, int64_t n_all, int64_t lo, int64_t hi, int M, int m,
                     int bucket, uint64_t seed, int rng_mode, const int32_t *walks_hd, int flags, cudaStream_t st,
                     SpG **out) {
    if (!g || !out || n_all < 0 || (n_all > 0 && !seeds_hd) || lo < 0 || hi < lo || hi > n_all)
        return fail(SUBG_ERR_ARG, "Input parsing error.");
    const int64_t n = hi - lo;
    if (rng_mode < 0 || rng_mode > 2) return fail(SUBG_ERR_ARG, "unknown rng_mode");
    if (rng_mode == SUBG_RNG_TRACE && !walks_hd && n > 0) return fail(SUBG_ERR_ARG, "trace mode needs walks");
    SamplePlan pl;
    if (int rc = make_plan(g, M, m, bucket, &pl)) return rc;
    DeviceGuard guard(g->device);
    const bool want_slot = !(flags & SUBG_SAMPLE_NO_RANKS);
    const bool want_rank = want_slot || pl.stride < pl.Kt;
    if (!want_rank) {  // no first-visit ranks: the order bitmap is never touched, its shared memory buys a resident CTA on ppa
        pl.nbw = 0;
        pl.smem_per_warp = (pl.bitmap_off + 15) & ~15;
    }

    g->tag.use_on(st);
    SpG *s = new SpG();
    s->tag.last = st;
    s->device = g->device; s->n = n; s->ncol = m + 1; s->M = M; s->num_sms = g->num_sms; s->value_kind = 0;
    s->shift = pl.SHIFT;
    HostProf prof;

    // everything below that is not part of the SpG is scratch
    int32_t *d_walks = nullptr, *d_calls = nullptr, *rank_of_slot = nullptr, *d_all_seeds = nullptr;
    int32_t *c_node = nullptr, *c_prov = nullptr;  // chunk rows (chunked mode only)
    uint16_t *c_slot = nullptr;
    long long *call_base = nullptr, *scan_scratch = nullptr, *c_rowbeg = nullptr;
    unsigned long long *tab_key = nullptr, *tab_pos = nullptr, *d_ctr = nullptr;
    uint32_t *d_flags = nullptr;  // [0]=status [1]=tab_count [2]=bad seeds [3]=unique cnt
    int32_t *d_maxset = nullptr;
    bool walks_owned = false;
    int rc = SUBG_OK;
    int64_t cap = 0;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            rc = fail(_e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA,              \
                      std::string(#call) + ": " + cudaGetErrorString(_e));                         \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)

    {
        CK(dmalloc(&s->seeds, (size_t)n, st));
        CK(dmalloc(&s->nsize, (size_t)n, st));
        CK(dmalloc(&d_flags, 8, st));  // [0] status [1] tab_count [2] bad seeds [3] unique cnt [4] max set size
        d_maxset = (int32_t *)(d_flags + 4);
        CK(dmalloc(&d_ctr, kCtrWords, st));
        CK(cudaMemsetAsync(d_flags, 0, 8 * sizeof(uint32_t), st));
        if (n > 0) {
            CK(cudaMemcpyAsync(s->seeds, seeds_hd + lo, (size_t)n * sizeof(int32_t), cudaMemcpyDefault, st));
            check_seeds_kernel<<<std::min<int64_t>((n + 255) / 256, 4 * g->num_sms), 256, 0, st>>>(s->seeds, n, g->N, d_flags + 2);
        }
        const int nscan = std::max(1, scan_num_blocks(n));
        CK(dmalloc(&scan_scratch, (size_t)nscan, st));

        if (rng_mode == SUBG_RNG_RAND_R && n > 0) {
            // the single rand_r stream is consumed seed by seed: offsets are a prefix over the WHOLE query
            const int32_t *all_seeds = s->seeds;
            if (n_all != n) {
                CK(dmalloc(&d_all_seeds, (size_t)n_all, st));
                CK(cudaMemcpyAsync(d_all_seeds, seeds_hd, (size_t)n_all * sizeof(int32_t), cudaMemcpyDefault, st));
                check_seeds_kernel<<<std::min<int64_t>((n_all + 255) / 256, 4 * g->num_sms), 256, 0, st>>>(d_all_seeds, n_all, g->N, d_flags + 2);
                all_seeds = d_all_seeds;
            }
            long long *scratch_all = nullptr;
            CK(dmalloc(&d_calls, (size_t)n_all, st));
            CK(dmalloc(&call_base, (size_t)n_all + 1, st));
            CK(dmalloc(&scratch_all, (size_t)std::max(1, scan_num_blocks(n_all)), st));
            rand_r_calls_kernel<<<std::min<int64_t>((n_all + 255) / 256, 4 * g->num_sms), 256, 0, st>>>(
                g->rowptr, g->rowptr64 ? 1 : 0, all_seeds, n_all, g->N, M, m, d_calls);
            CK(exclusive_scan_i32_i64(d_calls, call_base, n_all, 0, scratch_all, st));
            dfree(scratch_all, st);
        }
        if (rng_mode == SUBG_RNG_TRACE && n > 0) {
            if (is_device_ptr(walks_hd)) d_walks = const_cast<int32_t *>(walks_hd);
            else {
                const size_t cnt = (size_t)n * M * m;
                CK(dmalloc(&d_walks, cnt, st));
                walks_owned = true;
                CK(cudaMemcpyAsync(d_walks, walks_hd, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            }
        }
        if (rng_mode == SUBG_RNG_RAND_R && n > 0) {
            // rand_r_calls_kernel indexes the row pointer with every seed: the range check has to be known first
            // (the sampler kernel itself skips out-of-range seeds and reports them through the status word)
            uint32_t bad = 0;
            CK(cudaMemcpyAsync(&bad, d_flags + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (bad) { rc = fail(SUBG_ERR_ARG, "query contains node ids outside [0, N)"); goto done; }
        }

        if ((flags & SUBG_SAMPLE_DUMP_WALKS) && rng_mode != SUBG_RNG_TRACE && n > 0)
            CK(dmalloc(&s->walks, (size_t)n * M * m, st));
        prof.mark("setup");
        // Rows are written once, at a cursor, into arrays sized for the worst case (rowcap entries per
        // seed).  If that does not fit the budget the seeds go through in chunks and every chunk is
        // compacted into the growing CSR (the layout of the reference's dense `encoding`, subg_acc.c:848-872).
        const int entry_bytes = want_slot ? 10 : 8;
        const int64_t worst = std::max<int64_t>(n, 1) * pl.rowcap * entry_bytes;
        int64_t chunk = 0, rows_cap = 0;
        bool chunked = false;
        for (int tries = 0;; tries++) {
            // cudaMemGetInfo is a driver round trip that can take milliseconds while other processes use
            // the GPUs: the free-memory figure is cached per device and refreshed only when an allocation fails
            int64_t free_b = (int64_t)8 << 30;
            if (worst > (1ll << 30)) free_b = free_memory_estimate(g->device, tries > 0);
            const int64_t budget = env_i64("SUBG_STAGING_BYTES", (int64_t)(free_b * 0.45));
            chunk = std::max<int64_t>(1, budget / ((int64_t)pl.rowcap * entry_bytes));
            chunk = std::min<int64_t>(chunk, std::max<int64_t>(n, 1));
            chunked = chunk < n;
            rows_cap = chunk * pl.rowcap + 16;
            cudaError_t e = cudaSuccess;
            if (chunked) {
                e = dmalloc(&c_node, (size_t)rows_cap, st);
                if (e == cudaSuccess) e = dmalloc(&c_prov, (size_t)rows_cap, st);
                if (e == cudaSuccess && want_slot) e = dmalloc(&c_slot, (size_t)rows_cap, st);
                if (e == cudaSuccess) e = dmalloc(&c_rowbeg, (size_t)chunk, st);
                if (e == cudaSuccess) e = dmalloc(&s->indptr, (size_t)n + 1, st);
            } else {
                e = dmalloc(&s->indices, (size_t)rows_cap, st);
                if (e == cudaSuccess) e = dmalloc((int32_t **)&s->data, (size_t)rows_cap, st);
                if (e == cudaSuccess && want_slot) e = dmalloc(&s->slot, (size_t)rows_cap, st);
                if (e == cudaSuccess) e = dmalloc(&s->rowbeg, (size_t)std::max<int64_t>(n, 1), st);
                cap = rows_cap;
            }
            if (e == cudaSuccess) break;
            if (e != cudaErrorMemoryAllocation || tries >= 1) CK(e);
            cudaGetLastError();
            dfree(c_node, st); dfree(c_prov, st); dfree(c_slot, st); dfree(c_rowbeg, st);
            dfree(s->indptr, st); dfree(s->indices, st); dfree(s->data, st); dfree(s->slot, st); dfree(s->rowbeg, st);
            c_node = c_prov = nullptr; c_slot = nullptr; c_rowbeg = nullptr;
            s->indptr = nullptr; s->indices = nullptr; s->data = nullptr; s->slot = nullptr; s->rowbeg = nullptr;
            cap = 0;
        }

        prof.mark("alloc_rows");
        int tab_log2 = (int)env_i64("SUBG_LP_TABLE_LOG2", 20);
        for (int attempt = 0;; attempt++) {
            const uint32_t tab_cap = 1u << tab_log2;
            CK(dmalloc(&tab_key, (size_t)tab_cap, st));
            CK(dmalloc(&tab_pos, (size_t)tab_cap, st));
            fill_u64_kernel<<<4 * g->num_sms, 256, 0, st>>>(tab_key, tab_cap, kEmptyKey);
            fill_u64_kernel<<<4 * g->num_sms, 256, 0, st>>>(tab_pos, tab_cap, ~0ull);
            CK(cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint32_t), st));
            CK(cudaMemsetAsync(d_maxset, 0, sizeof(int32_t), st));
            CK(cudaMemsetAsync(d_flags + 3, 0, sizeof(uint32_t), st));

            int64_t T = 0, extent = 0;
            bool table_full = false;
            uint32_t flags_h[5] = {0, 0, 0, 0, 0};
            for (int64_t base = 0; base < n; base += chunk) {
                const int64_t nc = std::min(chunk, n - base);
                CK(cudaMemsetAsync(d_ctr, 0, kCtrWords * sizeof(unsigned long long), st));
                SamplerArgs a{};
                a.rowinfo = (const unsigned long long *)g->rowinfo; a.rowptr = g->rowptr; a.rowptr64 = g->rowptr64 ? 1 : 0; a.col = g->col;
                a.col3 = graph_col3(g, st);
                a.seeds = s->seeds + base; a.n_chunk = nc; a.seed_base = lo + base;
                a.N = g->N; a.M = M; a.m = m; a.stride = pl.stride; a.Kt = pl.Kt; a.OB = pl.OB; a.LS = pl.LS; a.SHIFT = pl.SHIFT;
                a.rng_mode = rng_mode; a.rng_lo = (uint32_t)seed; a.rng_hi = (uint32_t)(seed >> 32);
                a.call_base = (const int64_t *)call_base;
                a.walks = d_walks ? d_walks + base * (int64_t)M * m : nullptr;
                a.dump_walks = s->walks ? s->walks + base * (int64_t)M * m : nullptr;
                a.out_node = chunked ? c_node : s->indices;
                a.out_prov = chunked ? c_prov : (int32_t *)s->data;
                a.out_slot = chunked ? c_slot : s->slot;
                a.rowbeg = chunked ? c_rowbeg : (long long *)s->rowbeg;
                a.nsize = s->nsize + base;
                a.ctr = d_ctr; a.max_set = d_maxset; a.want_rank = want_rank ? 1 : 0;
                a.stop_after = (int)env_i64("SUBG_SAMPLER_STOP", 0);
                {   // every gather in flight holds an L1 line: when the CSR does not fit the L2 the walk phase is
                    // bound by that count, so shared memory is capped to leave about 80 KB of the SM to L1 (sweep: profiles/r1_sampler_sweeps.txt)
                    const int64_t csr_bytes = 4 * g->E + 8 * g->N;
                    int cap_blocks = 0;
                    if (csr_bytes > (96ll << 20))
                        cap_blocks = std::max(2, (int)((152 << 10) / (kWarpsPerBlock * pl.smem_per_warp + 1024)));
                    a.blocks_per_sm = (int)env_i64("SUBG_SAMPLER_BLOCKS", cap_blocks);
                }
                a.tab_key = tab_key; a.tab_pos = tab_pos; a.tab_mask = tab_cap - 1;
                a.tab_count = d_flags + 1; a.status = d_flags;
                a.nbw = pl.nbw; a.fy_cap = pl.fy_cap; a.lp_off = pl.lp_off; a.lp64 = pl.lp64 ? 1 : 0; a.bitmap_off = pl.bitmap_off;
                a.smem_per_warp = pl.smem_per_warp;
                const int l2_persist = (int)env_i64("SUBG_L2_PERSIST", 0);
                if (l2_persist) {  // experiment knob: pin the row-info array in the persisting L2 carve-out
                    const size_t bytes = ((size_t)g->N + 1) * 8;
                    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, 64u << 20));
                    cudaStreamAttrValue av{};
                    av.accessPolicyWindow.base_ptr = g->rowinfo;
                    av.accessPolicyWindow.num_bytes = bytes;
                    av.accessPolicyWindow.hitRatio = 1.0f;
                    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
                }
                prof.mark("table_init");
                timing_begin(SUBG_TIMING_SAMPLER, st);
                CK(pl.key64 ? launch_gset_sample_k64(a, pl.EPL, g->num_sms, st)
                            : launch_gset_sample_k32(a, pl.EPL, g->num_sms, st));
                timing_end(SUBG_TIMING_SAMPLER, st);
                count_launch(1);
                if (l2_persist) {
                    cudaStreamAttrValue av{};
                    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
                }
                unsigned long long hctr[kCtrWords];
                CK(cudaMemcpyAsync(hctr, d_ctr, sizeof(hctr), cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(flags_h, d_flags, sizeof(flags_h), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                prof.mark("kernel+sync");
                if (flags_h[2] || (flags_h[0] & kStatusBadSeed)) { rc = fail(SUBG_ERR_ARG, "query contains node ids outside [0, N)"); goto done; }
                if (flags_h[1] > tab_cap / 2 || (flags_h[0] & kStatusTableFull)) { table_full = true; break; }
                if (!chunked) {
                    T = (int64_t)hctr[kCtrTotal];
                    extent = (int64_t)hctr[kCtrCursor];
                    break;
                }
                // ---- chunked mode: append the chunk's rows to the CSR
                timing_begin(SUBG_TIMING_BUILD, st);
                CK(exclusive_scan_i32_i64(s->nsize + base, (long long *)s->indptr + base, nc, T, scan_scratch, st));
                count_launch(3);
                const int64_t T_new = T + (int64_t)hctr[kCtrTotal];
                if (T_new > cap) {
                    int64_t want = T_new;
                    if (base + nc < n) want = std::max<int64_t>(T_new, (int64_t)((double)T_new * n / (base + nc) * 1.05) + 1024);
                    int32_t *ni = nullptr, *nd = nullptr; uint16_t *ns = nullptr;
                    CK(dmalloc(&ni, (size_t)want + 16, st));
                    CK(dmalloc(&nd, (size_t)want + 16, st));
                    if (want_slot) CK(dmalloc(&ns, (size_t)want + 16, st));
                    if (T > 0) {
                        CK(cudaMemcpyAsync(ni, s->indices, (size_t)T * 4, cudaMemcpyDeviceToDevice, st));
                        CK(cudaMemcpyAsync(nd, s->data, (size_t)T * 4, cudaMemcpyDeviceToDevice, st));
                        if (want_slot) CK(cudaMemcpyAsync(ns, s->slot, (size_t)T * 2, cudaMemcpyDeviceToDevice, st));
                    }
                    dfree(s->indices, st); dfree(s->data, st); dfree(s->slot, st);
                    s->indices = ni; s->data = nd; s->slot = ns;
                    cap = want + 16;
                }
                const int64_t cblocks = std::min<int64_t>((nc * 32 + 255) / 256, 8 * (int64_t)g->num_sms);
                compact_rows_kernel<<<(unsigned)std::max<int64_t>(cblocks, 1), 256, 0, st>>>(
                    c_node, c_prov, c_slot, c_rowbeg, s->nsize + base, (const long long *)s->indptr + base, nc,
                    s->indices, (int32_t *)s->data, s->slot);
                timing_end(SUBG_TIMING_BUILD, st);
                count_launch(1);
                T = T_new;
                extent = T_new;
            }
            if (table_full) {
                dfree(tab_key, st); dfree(tab_pos, st); tab_key = nullptr; tab_pos = nullptr;
                tab_log2 += 3;
                if (tab_log2 > 30) { rc = fail(SUBG_ERR_MEM, "LP-row table exceeds 2^30 entries"); goto done; }
                continue;
            }
            if (chunked && n > 0) s->rowbeg = s->indptr;
            s->T = T; s->extent = extent; s->cap = cap;

            // ---- unique LP rows in first-occurrence order
            // (status, unique count and max set size came back with the last chunk's counters)
            s->status = flags_h[0] & ~(kStatusTableFull | kStatusBadSeed);
            s->max_set = (int32_t)flags_h[4];
            const uint32_t c = flags_h[1];
            s->c = (int32_t)c;
            CK(dmalloc(&s->enc, (size_t)c * (m + 1), st));
            if (c > 0) {
                timing_begin(SUBG_TIMING_BUILD, st);
                CK(dmalloc(&s->lp_pos, (size_t)c, st));
                CK(dmalloc(&s->lp_key, (size_t)c, st));
                CK(dmalloc(&rank_of_slot, (size_t)tab_cap, st));
                CK(cudaMemsetAsync(rank_of_slot, 0, (size_t)tab_cap * 4, st));
                if (int urc = rank_unique_keys(tab_key, tab_pos, tab_cap, c, true, M, m, pl.SHIFT, rank_of_slot, s->enc, s->lp_key,
                                               s->lp_pos, d_flags + 3, g->num_sms, st)) { rc = urc; goto done; }
                if (extent > 0 && env_i64("SUBG_SAMPLER_STOP", 0) == 0) {  // a truncated measurement run leaves no valid rows
                    const int64_t rb = std::min<int64_t>((extent / 4 + 255) / 256 + 1, 16 * (int64_t)g->num_sms);
                    remap_ids_kernel<<<(unsigned)rb, 256, 0, st>>>((int32_t *)s->data, extent, rank_of_slot);
                }
                timing_end(SUBG_TIMING_BUILD, st);
                count_launch(1);
            }
            break;
        }
        if (!s->indices) {  // n == 0: keep valid (padded) arrays
            CK(dmalloc(&s->indices, 16, st)); CK(dmalloc((int32_t **)&s->data, 16, st));
            if (want_slot) CK(dmalloc(&s->slot, 16, st));
            cap = 16;
        }
        if (n == 0) {
            if (!s->indptr) CK(dmalloc(&s->indptr, 1, st));
            CK(cudaMemsetAsync(s->indptr, 0, sizeof(int64_t), st));
            if (s->rowbeg && (void *)s->rowbeg != (void *)s->indptr) dfree(s->rowbeg, st);
            s->rowbeg = s->indptr;
        }
        prof.mark("unique_launch");
        // No synchronisation here: the id remap is still in flight; everything that touches the SpG later
        // (SpJoin, export, views, free) is ordered behind it on the stream.
        // A mostly empty worst-case allocation is compacted only when the slack is a sizeable share of the device (default:
        // more than 15 % of its memory, SUBG_COMPACT_SLACK_PCT): SpJoin and the exchange read the scattered rows in place, and
        // the copy costs a pass over the SpG (dblp shape: 0.4 ms of a 5 ms pass for 2.3 GB of slack; twitter shape: 33 ms of
        // 255 ms for 24 GB of slack on a 180 GB device)
        static size_t total_by_device[64] = {};   // cudaMemGetInfo costs milliseconds once large blocks are cached: ask once
        size_t &mem_total = total_by_device[g->device & 63];
        if (mem_total == 0) {
            size_t mem_free = 0;
            cudaMemGetInfo(&mem_free, &mem_total);
        }
        const int64_t slack_bytes = (cap - s->extent) * (int64_t)(want_slot ? 10 : 8);
        const int64_t slack_limit = (int64_t)((double)mem_total * (double)env_i64("SUBG_COMPACT_SLACK_PCT", 15) / 100.0);
        if (!s->indptr && !(flags & SUBG_SAMPLE_NO_COMPACT) && cap > s->extent + s->extent / 4 + (16ll << 20) && slack_bytes > slack_limit) {
            timing_begin(SUBG_TIMING_BUILD, st);
            const int erc = spg_ensure_csr(s, st);
            timing_end(SUBG_TIMING_BUILD, st);
            if (erc != SUBG_OK) { rc = erc; goto done; }
        }
    }
done:
#undef CK
    prof.mark("tail");
    if (walks_owned) dfree(d_walks, st);
    dfree(d_calls, st); dfree(call_base, st); dfree(scan_scratch, st); dfree(d_all_seeds, st);
    dfree(c_node, st); dfree(c_prov, st); dfree(c_slot, st); dfree(c_rowbeg, st);
    dfree(tab_key, st); dfree(tab_pos, st); dfree(rank_of_slot, st);
    dfree(d_flags, st); dfree(d_ctr, st);
    if (rc != SUBG_OK) {
        free_spg_arrays(s, st);
        delete s;
        return rc;
    }
    *out = s;
    return SUBG_OK;
}

// Re-label the LP rows of a shard: data <- id_map[data - 1] + 1, enc <- the merged table.
// (multi-GPU: local first-occurrence ids -> ids of the table merged over all shards in rank order)
int spg_set_lp_table_impl(SpG *s, const int32_t *id_map_hd, const int16_t *enc_hd, int32_t c_new, int32_t ncol,
                          cudaStream_t st) {
    if (!s || c_new < 0 || (s->c > 0 && !id_map_hd) || (c_new > 0 && !enc_hd)) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (s->value_kind != 0) return fail(SUBG_ERR_ARG, "LP table of a value SpG");
    if (s->borrowed) return fail(SUBG_ERR_ARG, "a linked SpG's rows live in the exchange slabs and are not relabelled");
    if (ncol > 0) s->ncol = ncol;
    if (s->ncol < 1) return fail(SUBG_ERR_ARG, "LP table width unknown");
    DeviceGuard guard(s->device);
    s->tag.use_on(st);
    int32_t *d_map = nullptr;
    int16_t *d_enc = nullptr;
    SUBG_CUDA(dmalloc(&d_map, (size_t)s->c + 1, st));
    SUBG_CUDA(dmalloc(&d_enc, (size_t)c_new * s->ncol, st));
    if (s->c > 0) SUBG_CUDA(cudaMemcpyAsync(d_map, id_map_hd, (size_t)s->c * 4, cudaMemcpyDefault, st));
    if (c_new > 0) SUBG_CUDA(cudaMemcpyAsync(d_enc, enc_hd, (size_t)c_new * s->ncol * 2, cudaMemcpyDefault, st));
    if (s->extent > 0 && s->c > 0) {
        const int64_t rb = std::min<int64_t>((s->extent + 255) / 256, 16 * (int64_t)s->num_sms);
        relabel_ids_kernel<<<(unsigned)rb, 256, 0, st>>>((int32_t *)s->data, s->extent, d_map, s->c, c_new);
        SUBG_CUDA(cudaGetLastError());
        count_launch(1);
    }
    SUBG_CUDA(cudaStreamSynchronize(st));
    dfree(d_map, st);
    dfree(s->enc, st);
    s->enc = d_enc;
    s->c = c_new;
    return SUBG_OK;
}

// An empty LP SpG in the compact layout whose arrays the caller fills in place (the multi-GPU exchange lets NCCL
// write the gathered shards straight into them): nsize int32[n], indices int32[T], data int32[T].  spg_seal_impl then
// derives the row pointer and the largest set on the device.
__global__ void max_i32_kernel(const int32_t *v, int64_t n, int32_t *out) {
    int32_t m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, v[i]);
    for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(FULL, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

int spg_alloc_impl(int64_t n, int64_t T, int device, cudaStream_t st, SpG **out) {
    if (!out || n < 0 || T < 0) return fail(SUBG_ERR_ARG, "Input parsing error.");
    DeviceGuard guard(device);
    SpG *s = new SpG();
    s->tag.last = st;
    s->device = device; s->n = n; s->T = T; s->value_kind = 0;
    cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, device);
    cudaError_t e = dmalloc(&s->indptr, (size_t)n + 1, st);
    if (e == cudaSuccess) e = dmalloc(&s->nsize, (size_t)std::max<int64_t>(n, 1), st);
    if (e == cudaSuccess) e = dmalloc(&s->indices, (size_t)T + 16, st);
    if (e == cudaSuccess) e = dmalloc((int32_t **)&s->data, (size_t)T + 16, st);
    if (e != cudaSuccess) {
        free_spg_arrays(s, st);
        delete s;
        return fail(e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA, cudaGetErrorString(e));
    }
    s->rowbeg = s->indptr; s->extent = T; s->cap = T + 16;
    *out = s;
    return SUBG_OK;
}

int spg_seal_impl(SpG *s, cudaStream_t st) {
    if (!s || !s->indptr || !s->nsize) return fail(SUBG_ERR_ARG, "seal needs an SpG from subg_spg_alloc");
    DeviceGuard guard(s->device);
    s->tag.use_on(st);
    long long *scratch = nullptr;
    int32_t *d_max = nullptr;
    SUBG_CUDA(dmalloc(&scratch, (size_t)std::max(1, scan_num_blocks(s->n)), st));
    SUBG_CUDA(dmalloc(&d_max, 1, st));
    SUBG_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int32_t), st));
    SUBG_CUDA(exclusive_scan_i32_i64(s->nsize, (long long *)s->indptr, s->n, 0, scratch, st));
    if (s->n > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>((s->n + 255) / 256, 4 * (int64_t)s->num_sms);
        max_i32_kernel<<<blocks, 256, 0, st>>>(s->nsize, s->n, d_max);
    }
    count_launch(4);
    long long total = 0;
    int32_t mx = 0;
    SUBG_CUDA(cudaMemcpyAsync(&total, s->indptr + s->n, sizeof(long long), cudaMemcpyDeviceToHost, st));
    SUBG_CUDA(cudaMemcpyAsync(&mx, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SUBG_CUDA(cudaStreamSynchronize(st));
    dfree(scratch, st); dfree(d_max, st);
    if (total != s->T) return fail(SUBG_ERR_ARG, "set sizes do not add up to the number of entries");
    s->max_set = mx;
    return SUBG_OK;
}

// Rows in seed order -> one row per graph node (what subg_matrix's csr_matrix((data, (repeat(idx, nsize), nodes)), (N, N))
// gives for a query that is not arange(N), random_walks.py:79): row idx[i] = set i, every other row empty.  The entries stay
// where they are (scattered layout); only the row table is rebuilt.
__global__ void expand_rows_kernel(const int32_t *seeds, const long long *rowbeg, const int32_t *nsize, int64_t n,
                                   long long *out_rowbeg, int32_t *out_nsize) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t u = seeds[i];
        out_rowbeg[u] = rowbeg[i];
        out_nsize[u] = nsize[i];
    }
}
int spg_expand_rows_impl(SpG *s, int64_t num_nodes, cudaStream_t st) {
    if (!s || num_nodes < 0) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (!s->seeds || !s->nsize) return fail(SUBG_ERR_ARG, "expand needs a sampler-built SpG (rows in seed order)");
    DeviceGuard guard(s->device);
    s->tag.use_on(st);
    long long *rb = nullptr;
    int32_t *ns = nullptr;
    SUBG_CUDA(dmalloc(&rb, (size_t)std::max<int64_t>(num_nodes, 1), st));
    SUBG_CUDA(dmalloc(&ns, (size_t)std::max<int64_t>(num_nodes, 1), st));
    SUBG_CUDA(cudaMemsetAsync(rb, 0, (size_t)std::max<int64_t>(num_nodes, 1) * 8, st));
    SUBG_CUDA(cudaMemsetAsync(ns, 0, (size_t)std::max<int64_t>(num_nodes, 1) * 4, st));
    if (s->n > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>((s->n + 255) / 256, 4 * (int64_t)s->num_sms);
        expand_rows_kernel<<<blocks, 256, 0, st>>>(s->seeds, (const long long *)s->rowbeg, s->nsize, s->n, rb, ns);
        SUBG_CUDA(cudaGetLastError());
        count_launch(1);
    }
    if (s->indptr) dfree(s->indptr, st);                       // compact layout: rowbeg aliases indptr
    else dfree(s->rowbeg, st);
    dfree(s->nsize, st); dfree(s->seeds, st);
    s->indptr = nullptr; s->rowbeg = (int64_t *)rb; s->nsize = ns; s->seeds = nullptr;
    s->n = num_nodes;
    return SUBG_OK;
}

int spg_export_impl(const SpG *s, int32_t *nsize_hd, int32_t *remap_hd, int16_t *enc_hd, int16_t *raw_hd,
                    cudaStream_t st) {
    if (!s) return fail(SUBG_ERR_ARG, "null SpG");
    if (s->value_kind != 0 || !s->slot)
        return fail(SUBG_ERR_ARG, "export needs a sampler-built LP SpG with first-visit ranks (not SUBG_SAMPLE_NO_RANKS)");
    DeviceGuard guard(s->device);
    s->tag.use_on(st);
    const int64_t T = s->T, n = s->n;
    const int ncol = s->ncol;
    if (nsize_hd && n > 0)
        SUBG_CUDA(cudaMemcpyAsync(nsize_hd, s->nsize, (size_t)n * sizeof(int32_t), cudaMemcpyDefault, st));
    if (enc_hd && s->c > 0)
        SUBG_CUDA(cudaMemcpyAsync(enc_hd, s->enc, (size_t)s->c * ncol * sizeof(int16_t), cudaMemcpyDefault, st));
    if ((remap_hd || raw_hd) && T > 0) {
        int32_t *d_remap = nullptr;
        int16_t *d_raw = nullptr;
        const bool remap_dev = remap_hd && is_device_ptr(remap_hd);
        const bool raw_dev = raw_hd && is_device_ptr(raw_hd);
        if (remap_dev) d_remap = remap_hd;
        else SUBG_CUDA(dmalloc(&d_remap, (size_t)2 * T, st));
        if (raw_hd) {
            if (raw_dev) d_raw = raw_hd;
            else SUBG_CUDA(dmalloc(&d_raw, (size_t)T * ncol, st));
        }
        long long *dstptr = (long long *)s->indptr, *scratch = nullptr;
        if (!dstptr) {  // scattered rows: the reference's dense order is the scan of the set sizes
            SUBG_CUDA(dmalloc(&dstptr, (size_t)n + 1, st));
            SUBG_CUDA(dmalloc(&scratch, (size_t)std::max(1, scan_num_blocks(n)), st));
            SUBG_CUDA(exclusive_scan_i32_i64(s->nsize, dstptr, n, 0, scratch, st));
        }
        const int64_t blocks = std::min<int64_t>((n * 32 + 255) / 256, 8 * (int64_t)s->num_sms);
        export_remap_kernel<<<(unsigned)std::max<int64_t>(blocks, 1), 256, 0, st>>>(
            (const long long *)s->rowbeg, s->nsize, dstptr, s->indices, (const int32_t *)s->data, s->slot, n, T, d_remap,
            s->enc, ncol, d_raw);
        SUBG_CUDA(cudaGetLastError());
        if (!s->indptr) { dfree(dstptr, st); dfree(scratch, st); }
        if (remap_hd && !remap_dev)
            SUBG_CUDA(cudaMemcpyAsync(remap_hd, d_remap, (size_t)2 * T * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (raw_hd && !raw_dev)
            SUBG_CUDA(cudaMemcpyAsync(raw_hd, d_raw, (size_t)T * ncol * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
        SUBG_CUDA(cudaStreamSynchronize(st));
        if (!remap_dev) dfree(d_remap, st);
        if (raw_hd && !raw_dev) dfree(d_raw, st);
    } else {
        SUBG_CUDA(cudaStreamSynchronize(st));
    }
    return SUBG_OK;
}

int spg_from_csr_impl(const int64_t *indptr_hd, const int32_t *indices_hd, const void *data_hd, int value_kind,
                      int64_t n_rows, int64_t nnz, int device, cudaStream_t st, SpG **out) {
    if (!out || n_rows < 0 || nnz < 0 || !indptr_hd || (nnz > 0 && (!indices_hd || !data_hd)))
        return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (value_kind != 0 && value_kind != 1) return fail(SUBG_ERR_ARG, "value_kind must be 0 (int32) or 1 (float64)");
    DeviceGuard guard(device);
    SpG *s = new SpG();
    s->tag.last = st;
    s->device = device; s->n = n_rows; s->T = nnz; s->value_kind = value_kind;
    cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, device);
    const size_t vb = value_kind ? 8 : 4;
    cudaError_t e = dmalloc(&s->indptr, (size_t)n_rows + 1, st);
    if (e == cudaSuccess) e = dmalloc(&s->indices, (size_t)nnz + 16, st);
    if (e == cudaSuccess) e = dmalloc((unsigned char **)&s->data, ((size_t)nnz + 16) * vb, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->indptr, indptr_hd, ((size_t)n_rows + 1) * 8, cudaMemcpyDefault, st);
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(s->indices, indices_hd, (size_t)nnz * 4, cudaMemcpyDefault, st);
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(s->data, data_hd, (size_t)nnz * vb, cudaMemcpyDefault, st);
    // max set size (host side when the row pointer is host memory, else via a copy)
    std::vector<int64_t> hp;
    const int64_t *hptr = indptr_hd;
    if (e == cudaSuccess && is_device_ptr(indptr_hd)) {
        hp.resize((size_t)n_rows + 1);
        e = cudaMemcpyAsync(hp.data(), indptr_hd, ((size_t)n_rows + 1) * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        hptr = hp.data();
    }
    if (e != cudaSuccess) {
        free_spg_arrays(s, st);
        delete s;
        return fail(e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA, cudaGetErrorString(e));
    }
    int64_t mx = 0;
    for (int64_t i = 0; i < n_rows; i++) mx = std::max(mx, hptr[i + 1] - hptr[i]);
    if (hptr[0] != 0 || hptr[n_rows] != nnz) {
        free_spg_arrays(s, st);
        delete s;
        return fail(SUBG_ERR_ARG, "indptr does not describe nnz entries");
    }
    s->max_set = (int32_t)std::min<int64_t>(mx, INT32_MAX);
    s->rowbeg = s->indptr; s->extent = nnz; s->cap = nnz + 16;
    SUBG_CUDA(cudaStreamSynchronize(st));
    *out = s;
    return SUBG_OK;
}

void spg_free_impl(SpG *s) {
    if (!s) return;
    DeviceGuard guard(s->device);
    const cudaStream_t st = s->tag.free_stream();  // the stream of the last kernel that touched the arrays
    free_spg_arrays(s, st);
    dfree(s->join_sizes, st);
    dfree(s->join_tot, st);
    if (s->join_host) cudaFreeHost(s->join_host);
    delete s;
}

}  // namespace subg
